# usage: bash scripts/gpu_ab.sh "ENV1=.. ENV2=.." ["ENV=.."] ...   one short device-resident bench per environment
for E in "$@"; do
  env $E python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$E', round(d['value']), round(d['ms_per_batch'],3))"
done
