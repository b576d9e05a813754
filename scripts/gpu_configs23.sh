# usage: bash scripts/gpu_configs23.sh <tag>  (under gpurun): BASELINE configs[2] (Lite1 b256) and configs[3] (Lite2 b256, int8 vs bf16 heads)
TAG=${1:-r2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --steps 4 --warmup 2 --cpu-sample 2 "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/${TAG}_$name.err; }
run config2_lite1_b256 --variant lite1 --batch 256
run config3_lite2_b256_int8_heads --variant lite2 --batch 256
run config3_lite2_b256_bf16_heads --variant lite2 --batch 256 --head-dtype bf16
python - <<PY
import json
for n in ('config2_lite1_b256', 'config3_lite2_b256_int8_heads', 'config3_lite2_b256_bf16_heads'):
    try:
        d = json.load(open('gpurun_out/${TAG}_%s.json' % n))
        print(n, 'value', round(d['value']), 'ms/batch', round(d['ms_per_batch'], 2), 'e2e', round(d['e2e']['value']), 'dets/frame', round(d['config']['detections_per_frame'], 2), 'live', d['config']['live_tracks'],
              {k: round(v['ms_per_step'], 2) for k, v in list(d['kernels'].items())[:5]}, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2))
    except Exception as e:
        print(n, 'no line', e)
PY
