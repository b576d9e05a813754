for L in 1 2 3; do VBT_LANES=$L python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lanes $L', round(d['value']), round(d['ms_per_batch'],3))"; done
python scripts/net_probe.py lite0 64 20 mbconv | tail -18
