#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
(timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/r2s_bench.err)
tail -2 gpurun_out/r2s_bench.err
SECONDS=0
(timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; echo "ref rc=$? wall=${SECONDS}s" >> gpurun_out/r2s_ref.err)
tail -1 gpurun_out/r2s_ref.err
(timeout 300 ncu -k "regex:banded_x2|finalize|gather_small" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 10 --warmup 3 --no-ess --no-secondary --no-cpu-baseline > gpurun_out/r2s_ncu_launches.log 2>&1)
(timeout 300 ncu -k "regex:banded_x2" --set full --import-source on --clock-control none -s 4 -c 2 -o gpurun_out/r2s_x2_full -f python bench.py --steps 6 --warmup 3 --no-ess --no-secondary --no-cpu-baseline > gpurun_out/r2s_ncu_full.log 2>&1)
(timeout 300 python tools/bench_configs.py > gpurun_out/r2s_cfg.jsonl 2>&1)
cat gpurun_out/r2s_cfg.jsonl
