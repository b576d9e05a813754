# what the driver runs at round end, on one GPU: smoke(), the GPU tests, both bench arms
TAG=${1:-r2final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
r = json.load(open('gpurun_out/${TAG}_ref.json')); d = json.load(open('gpurun_out/${TAG}_bench.json'))
print('reference arm', round(r['value'], 2), 'frames/s', r['cpu_baseline']['cores'], 'cores')
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'], 2), 'ms/batch', round(d['ms_per_batch'], 3), 'launches', d['gpu_launches'])
print('roofline', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'e2e_frac')})
print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
print('ratio value', round(d['value'] / r['value']), 'e2e', round(d['e2e']['value'] / r['value']))
PY
