# usage: bash scripts/gpu_launchlist.sh <tag>   (under gpurun): per-launch device times of 2 steps
set -x
TAG=${1:-r1}
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 256"
mkdir -p gpurun_out
export VBT_GRAPH=0 VBT_LANES=1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv 30
