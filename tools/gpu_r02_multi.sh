#!/bin/bash
# multi-GPU evidence: N = $1 GPUs on one box
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus_n$N.txt
if [ "$N" = "2" ]; then
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs > gpurun_out/r02_pytest_multi_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_multi_gpu.log ); tail -4 gpurun_out/r02_pytest_multi_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_2gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r02_smoke_2gpu.txt ); tail -3 gpurun_out/r02_smoke_2gpu.txt
fi
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "exit $?" >> gpurun_out/r02_bench_n$N.err ); tail -3 gpurun_out/r02_bench_n$N.err; cut -c1-400 gpurun_out/r02_bench_n$N.json
