# usage: bash scripts/gpu_sweep.sh <tag> <ENVVAR> "<v1> <v2> ..."   (under gpurun): short bench per value
TAG=$1; VAR=$2
mkdir -p gpurun_out
for V in $3; do
  env $VAR=$V timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --clip-frames 256 \
     --op-dump gpurun_out/${TAG}_${VAR}_${V}_ops.tsv > gpurun_out/${TAG}_${VAR}_${V}.json 2> gpurun_out/${TAG}_${VAR}_${V}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_${VAR}_${V}.json'))
print('$VAR=$V value',round(d['value']),'ms/step',round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('pw','dw3','dw5','stem','fuse_add','K7_tracker')})
PY
done
