# usage: bash scripts/gpu_n2.sh <tag> <N>   (under gpurun --gpus N): the bench and the configs4 workload on N ranks
TAG=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/${TAG}_n${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload configs4 > gpurun_out/${TAG}_configs4_n${N}.json 2> gpurun_out/${TAG}_configs4_n${N}.err; echo "configs4 N=$N rc=$?"
tail -2 gpurun_out/${TAG}_configs4_n${N}.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_n${N}.json'))
    print('N', d['n_gpus'], 'value', round(d['value']), 'per gpu', round(d['value'] / d['n_gpus']), 'e2e', round(d['e2e']['value']), 'gather_ms', round(d['gather_ms'], 3), 'ms/batch', round(d['ms_per_batch'], 3), 'clock samples', d['clocks'].get('samples'))
except Exception as e:
    print('bench line missing', e)
try:
    d = json.load(open('gpurun_out/${TAG}_configs4_n${N}.json'))
    print('configs4 N', d['n_gpus'], round(d['value']), d['parity'], 'lpt bound', round(d['config']['lpt_bound_efficiency'], 4), 'eff vs single', d['config']['efficiency_vs_single_rank'])
except Exception as e:
    print('configs4 line missing', e)
PY
