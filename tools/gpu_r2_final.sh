#!/bin/bash
# round 2, final evidence (1 GPU): full GPU suite, bench line, reference arm, ncu launch list + full capture of the
# headline kernels, BASELINE configs[1..3], every hot-path row at 1024^3 / 512^3
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-one-gpu --no-e2e > gpurun_out/r2z_ncu_launch.log 2>&1; echo "ncu launch rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_rows_inv|k_cols_tma|k_x_c2r" -s 9 -c 3 -o gpurun_out/prof_r2_final python bench.py --steps 1 --warmup 3 --no-cpu --no-one-gpu --no-e2e > gpurun_out/r2z_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
for c in lognormal_rsd_512 filter_beam_poles_1024 halos_cross_1024; do
timeout 400 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2z_cfg_$c.json 2> gpurun_out/r2z_cfg_$c.err; echo "cfg $c rc=$?"
done
timeout 400 python tools/bench_all.py 1024 > gpurun_out/r2z_all_1024.txt 2> gpurun_out/r2z_all_1024.err
timeout 300 python tools/bench_all.py 512 > gpurun_out/r2z_all_512.txt 2> gpurun_out/r2z_all_512.err
timeout 600 ncu --set full --clock-control none -k regex:"k_beam|k_rsd" -c 6 -o gpurun_out/prof_r2_beam_rsd python tools/ncu_beam.py 1024 1 > gpurun_out/r2z_ncu_beam.log 2>&1; echo "ncu beam rc=$?"
