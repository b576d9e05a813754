# usage: bash scripts/gpu_bench3.sh <tag>   (under gpurun): bench at 1 / 4 videos per GPU + configs4
TAG=${1:-r2}
mkdir -p gpurun_out
for V in 1 4; do
  timeout 600 python bench.py --steps 8 --warmup 2 --videos $V --no-cpu-baseline > gpurun_out/${TAG}_bench_v$V.json 2> gpurun_out/${TAG}_bench_v$V.err; echo "bench V=$V rc=$?"
  tail -3 gpurun_out/${TAG}_bench_v$V.err
done
timeout 600 python bench.py --workload configs4 > gpurun_out/${TAG}_configs4.json 2> gpurun_out/${TAG}_configs4.err; echo "configs4 rc=$?"; tail -3 gpurun_out/${TAG}_configs4.err
python - <<PY
import json
for V in (1, 4):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_v%d.json' % V))
    except Exception as e:
        print(V, 'no line', e); continue
    print('V', V, 'value', round(d['value']), 'ms/batch', round(d['ms_per_batch'], 3), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None,
          'gather_ms', d['gather_ms'], 'e2e_frac', round(d['roofline']['e2e_frac'], 4), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'], 4),
          d['roofline']['dominant_by_time'], 'clock samples', d['clocks'].get('samples'))
    for k, v in d['kernels'].items():
        print('    ', k, round(v['ms_per_step'], 3), v['gbs'] and round(v['gbs']), v['bound'])
try:
    d = json.load(open('gpurun_out/${TAG}_configs4.json'))
    print('configs4', round(d['value']), d['parity'], d['config']['single_rank_frames_per_s_same_job'])
except Exception as e:
    print('configs4: no line', e)
PY
