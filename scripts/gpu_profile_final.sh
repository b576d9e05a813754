# usage: bash scripts/gpu_profile_final.sh <tag>  (under gpurun; one GPU)
# (1) launch list of two steps; (2) DRAM traffic + time of EVERY launch of one step (3 metrics,
# CSV only); (3) a full-set capture of the three most expensive pointwise launches (pw_persist: b2.0.expand, b2.0.project, b2.1.expand), summarised to
# text on the box (the .ncu-rep itself stays small: 3 launches).
set -x
TAG=${1:-r1}
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profiler-range --clip-frames 256"
mkdir -p gpurun_out
export VBT_GRAPH=0 VBT_LANES=1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pw_persist -s 1 -c 3 \
    -o gpurun_out/${TAG}_prof_pw_persist python bench.py $ARGS > gpurun_out/${TAG}_ncu_pw.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_pw.log
ls -la gpurun_out/ | tail -8
# (4) the fused BiFPN-node / head-stage kernel: first four launches (5x5 ... 40x40 nodes of cell 0)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:node_umma -s 0 -c 4 \
    -o gpurun_out/${TAG}_prof_node_umma python bench.py $ARGS > gpurun_out/${TAG}_ncu_node.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_node.log
