# usage: bash scripts/gpu_round.sh <tag> [pytest-k-expr]   (run under gpurun)
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | head -c 3500
