#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "async" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log)
tail -25 gpurun_out/r2n_pytest.log
