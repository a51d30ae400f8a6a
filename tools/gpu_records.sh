#!/usr/bin/env bash
# One gpurun call that regenerates the records behind profiles/r02_*: default bench line (ESS run included),
# reference arm, ncu launch list and full capture of the headline kernel, per-config times, IWLS sampler.
# Afterwards: python tools/ncu_summary.py launches gpurun_out/rec_launches.csv profiles/r02_launches.md
#             python tools/ncu_summary.py full gpurun_out/rec_x2_full.ncu-rep profiles/r02_banded_x2_full.md
set -u
mkdir -p gpurun_out
SECONDS=0
(timeout 900 python bench.py > gpurun_out/rec_bench.json 2> gpurun_out/rec_bench.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/rec_bench.err)
tail -2 gpurun_out/rec_bench.err
SECONDS=0
(timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/rec_ref.json 2> gpurun_out/rec_ref.err; echo "ref rc=$? wall=${SECONDS}s" >> gpurun_out/rec_ref.err)
tail -1 gpurun_out/rec_ref.err
(timeout 300 ncu -k "regex:banded_x2|finalize|gather_small" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/rec_launches.csv python bench.py --steps 10 --warmup 3 --no-ess --no-secondary --no-cpu-baseline > gpurun_out/rec_ncu_launches.log 2>&1)
(timeout 300 ncu -k "regex:banded_x2" --set full --import-source on --clock-control none -s 4 -c 2 -o gpurun_out/rec_x2_full -f python bench.py --steps 6 --warmup 3 --no-ess --no-secondary --no-cpu-baseline > gpurun_out/rec_ncu_full.log 2>&1)
(timeout 300 python tools/bench_configs.py > gpurun_out/rec_cfg.jsonl 2>&1)
cat gpurun_out/rec_cfg.jsonl
(timeout 300 python tools/ess_bench.py --sampler iwls --chains 256 --warmup 100 --samples 100 > gpurun_out/rec_iwls_sampler.json 2> gpurun_out/rec_iwls_sampler.err)
cat gpurun_out/rec_iwls_sampler.json; tail -2 gpurun_out/rec_iwls_sampler.err
(timeout 600 python tools/config_samplers.py --configs s1,s2,s3 > gpurun_out/rec_cfg_samplers.jsonl 2> gpurun_out/rec_cfg_samplers.err)
(timeout 400 python tools/config_samplers.py --configs s5 --s5-chains 256 --s5-warmup 300 --s5-samples 300 >> gpurun_out/rec_cfg_samplers.jsonl 2>> gpurun_out/rec_cfg_samplers.err)
cat gpurun_out/rec_cfg_samplers.jsonl; tail -2 gpurun_out/rec_cfg_samplers.err
