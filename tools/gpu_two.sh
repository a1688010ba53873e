#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest_multi_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_gpu.log )
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload config5 --no-e2e > gpurun_out/bench_2gpu_config5.json 2> gpurun_out/bench_2gpu_config5.err )
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_config3.json 2> gpurun_out/bench_2gpu_config3.err )
tail -3 gpurun_out/pytest_multi_gpu.log; cut -c1-260 gpurun_out/bench_2gpu_config5.json; cut -c1-260 gpurun_out/bench_2gpu_config3.json
