# usage: bash scripts/gpu_n2b.sh <tag> <N>: one long clip chunk-sharded over N ranks (bench.py --workload onevideo)
TAG=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload onevideo > gpurun_out/${TAG}_onevideo_n${N}.json 2> gpurun_out/${TAG}_onevideo_n${N}.err; echo "onevideo N=$N rc=$?"
tail -3 gpurun_out/${TAG}_onevideo_n${N}.err
cat gpurun_out/${TAG}_onevideo_n${N}.json | cut -c1-1200
