#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log)
tail -6 gpurun_out/r2o_pytest.log
(timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2o_smoke.log)
tail -4 gpurun_out/r2o_smoke.log
SECONDS=0
(timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/r2o_bench.err)
tail -2 gpurun_out/r2o_bench.err
(timeout 600 python tools/ess_diag.py --async --chains 1024 --warmup 650 --term 500 --samples 1000 > gpurun_out/r2o_async1024_full.json 2> gpurun_out/r2o_async1024_full.err)
cat gpurun_out/r2o_async1024_full.json; tail -3 gpurun_out/r2o_async1024_full.err
