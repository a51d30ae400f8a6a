#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "parity or s2 or families or reproducible or nan_inf or center or lower_order or interface" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log)
(timeout 300 python tools/bench_configs.py --configs s2 > gpurun_out/r2l_cfg.json 2>&1)
(LSLB_NO_MSORT=1 timeout 300 python tools/bench_configs.py --configs s2 > gpurun_out/r2l_cfg_old.json 2>&1)
(timeout 300 ncu -k "regex:msort|finalize|gather_small" --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/r2l_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2l_s2_ncu.log 2>&1)
(timeout 300 ncu -k "regex:msort_x2" --set full --import-source on --clock-control none -c 1 -o gpurun_out/r2l_msort -f python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2l_msort_ncu.log 2>&1)
tail -5 gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_cfg.json gpurun_out/r2l_cfg_old.json
