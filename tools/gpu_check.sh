#!/usr/bin/env bash
# One gpurun call: GPU test suite, smoke, bench line (each under its own timeout); logs in gpurun_out/.
set -u
tag=${1:-run}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log)
(timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log)
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err)
(timeout 120 python tools/bench_configs.py --chain-sweep 48,64,96,128,192,256,512,1024 > gpurun_out/${tag}_sweep.json 2>&1)
(LSLB_X2_MIN_CHAINS=129 timeout 120 python tools/bench_configs.py --chain-sweep 48,64,96,128 > gpurun_out/${tag}_sweep_old.json 2>&1)
tail -15 gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_bench.err
