#!/bin/bash
# round 2, final evidence (1 GPU): full GPU suite, smoke, bench line, stage table at 512^3.  (The launch list, the full
# ncu capture of the headline kernels, the reference arm, the BASELINE configs and the 1024^3 stage table were taken
# with the same script earlier in the round: profiles/README.md.)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 300 python tools/bench_all.py 512 > gpurun_out/r2z_all_512.txt 2> gpurun_out/r2z_all_512.err
