#!/bin/bash
# round 2, call L (1 GPU): final single-GPU record - whole suite, smoke, bench line, ncu launch list of the bench command,
# ncu --set full of the FP64 LEAD / INNER kernels and of the final tcgen05 TF32 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -16 gpurun_out/r2l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; tail -2 gpurun_out/r2l_smoke.log
timeout 1200 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err; cut -c1-250 gpurun_out/r2l_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2l_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^mttkrp_(lead|inner)_kernel' -c 3 -o gpurun_out/r2l_mttkrp_fp64_full python tools/perf_probe.py 4096 4096 64 8192 64 1 > gpurun_out/r2l_ncu_fp64.log 2>&1
echo "ncu fp64 rc=$?"
PROBE_PREC=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^mttkrp_tc_kernel' -c 3 -o gpurun_out/r2l_tc_tf32_tmem_full python tools/perf_probe.py 4096 4096 64 8192 64 1 > gpurun_out/r2l_ncu_tc.log 2>&1
echo "ncu tc rc=$?"; ls -la gpurun_out/r2l_*.ncu-rep
