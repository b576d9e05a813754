# usage: bash scripts/gpu_profile_r2.sh <tag>  (under gpurun; one GPU)
# (1) plain run of the profiled command; (2) launch list + DRAM traffic of EVERY launch of two batches
# (3 metrics, CSV); (3) ncu --set full of the 16 fused MBConv launches of one batch (summarised here, the
# .ncu-rep stays in gpurun_out/).
set -x
TAG=${1:-r2}
ARGS="--steps 1 --warmup 2 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 128"
mkdir -p gpurun_out
export VBT_GRAPH=0 VBT_LANES=1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mbconv -s 0 -c 16 \
    -o gpurun_out/${TAG}_prof_mbconv python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/ | tail -5
