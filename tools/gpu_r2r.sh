#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256,512,1024 --reps 20 > gpurun_out/r2r_sweep_bal_$i.json 2>&1)
(LSLB_BALANCED_TILES=0 timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256,512,1024 --reps 20 > gpurun_out/r2r_sweep_old_$i.json 2>&1)
done
(timeout 200 python tools/bench_configs.py --configs s3 --reps 20 > gpurun_out/r2r_s3_bal.json 2>&1)
(LSLB_BALANCED_TILES=0 timeout 200 python tools/bench_configs.py --configs s3 --reps 20 > gpurun_out/r2r_s3_old.json 2>&1)
cat gpurun_out/r2r_sweep_*.json gpurun_out/r2r_s3_*.json
