#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
SECONDS=0
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/r2p_bench_n8.err)
tail -2 gpurun_out/r2p_bench_n8.err
SECONDS=0
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/rowshard_check.py > gpurun_out/r2p_rowshard_n8.log 2>&1; echo "rowshard rc=$? wall=${SECONDS}s" >> gpurun_out/r2p_rowshard_n8.log)
tail -12 gpurun_out/r2p_rowshard_n8.log
