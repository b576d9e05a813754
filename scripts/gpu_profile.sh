# usage: bash scripts/gpu_profile.sh <tag> <kernel-regex>   (run under gpurun; one GPU)
set -x
TAG=${1:-r1}; KRE=${2:-pw_umma}
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 256"
mkdir -p gpurun_out
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$KRE" -c 6 \
    -o gpurun_out/${TAG}_prof python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log
