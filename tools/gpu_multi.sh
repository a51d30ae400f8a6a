#!/usr/bin/env bash
# bench.py under torchrun on N GPUs of one box (records: profiles/r02_bench_line_nN.json):
#   gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
SECONDS=0
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/multi_bench_n$N.err)
tail -2 gpurun_out/multi_bench_n$N.err
(timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout=250 > gpurun_out/multi_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi_pytest_n$N.log)
tail -3 gpurun_out/multi_pytest_n$N.log
