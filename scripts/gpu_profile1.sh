# usage: bash scripts/gpu_profile1.sh <tag> <kernel-regex> [count] [skip]  (under gpurun; one ncu --set full pass)
set -x
TAG=${1:-r1}; K=${2:-pw_umma}; CNT=${3:-4}; SKIP=${4:-0}
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 256"
mkdir -p gpurun_out
export VBT_GRAPH=0 VBT_LANES=1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT \
    -o gpurun_out/${TAG}_prof_$K python bench.py $ARGS > gpurun_out/${TAG}_ncu_$K.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_$K.log
