# usage: bash scripts/gpu_round.sh <tag> [pytest-k-expr]   (run under gpurun)
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --op-dump gpurun_out/${TAG}_ops.tsv > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'] if d.get('e2e') else None,'launches',d['gpu_launches'])
print('roofline',d['roofline'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],3), v['gbs'])
PY
