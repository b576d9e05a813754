# usage: bash scripts/gpu_profile.sh <tag> "<kernel-regex> [<kernel-regex> ...]" [count]  (under gpurun; one GPU)
set -x
TAG=${1:-r1}; KRES=${2:-pw_umma}; CNT=${3:-4}
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 256"
mkdir -p gpurun_out
timeout 300 python bench.py $ARGS --op-dump gpurun_out/${TAG}_ops.tsv > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
for K in $KRES; do
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$K" -c $CNT \
      -o gpurun_out/${TAG}_prof_$K python bench.py $ARGS > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_$K.log
done
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv 30
