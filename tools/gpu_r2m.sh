#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "async" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log)
tail -15 gpurun_out/r2m_pytest.log
(timeout 500 python tools/ess_diag.py --async --chains 1024 --warmup 500 --term 250 --samples 150 > gpurun_out/r2m_async1024.json 2> gpurun_out/r2m_async1024.err)
cat gpurun_out/r2m_async1024.json; tail -3 gpurun_out/r2m_async1024.err
(timeout 300 python tools/ess_diag.py --async --chains 256 --warmup 300 --term 150 --samples 150 > gpurun_out/r2m_async256.json 2> gpurun_out/r2m_async256.err)
cat gpurun_out/r2m_async256.json; tail -3 gpurun_out/r2m_async256.err
