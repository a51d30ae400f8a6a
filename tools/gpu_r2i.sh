#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
(timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout=500 > gpurun_out/r2i_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_multi.log)
SECONDS=0
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/r2i_bench_n2.err)
SECONDS=0
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2i_ref_n2.json 2> gpurun_out/r2i_ref_n2.err; echo "ref rc=$? wall=${SECONDS}s" >> gpurun_out/r2i_ref_n2.err)
tail -4 gpurun_out/r2i_pytest_multi.log; tail -3 gpurun_out/r2i_bench_n2.err; tail -2 gpurun_out/r2i_ref_n2.err; head -c 600 gpurun_out/r2i_bench_n2.json
