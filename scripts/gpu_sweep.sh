# usage: bash scripts/gpu_sweep.sh <tag> <ENVVAR> "<v1> <v2> ..." [extra bench args]  (under gpurun)
TAG=$1; VAR=$2; EXTRA=${4:-"--steps 40 --warmup 4 --no-e2e --no-cpu-baseline"}
mkdir -p gpurun_out
for V in $3; do
  env $VAR=$V timeout 300 python bench.py $EXTRA \
     --op-dump gpurun_out/${TAG}_${VAR}_${V}_ops.tsv > gpurun_out/${TAG}_${VAR}_${V}.json 2> gpurun_out/${TAG}_${VAR}_${V}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_${VAR}_${V}.json'))
print('$VAR=$V value',round(d['value']),'ms/step',round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('pw','dw3','dw5','stem','fuse_add','K7_tracker')})
PY
done
