#!/bin/bash
# round 2, call F (1 GPU): LEAD kernel after the per-slab fold (C2 + a C3 slab), ncu --set full of em_kernel and of the tcgen05 kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mttkrp or config2 or dimension_tree or golden" > gpurun_out/r2f_pytest_mttkrp.log 2>&1
echo "mttkrp rc=$?"; tail -3 gpurun_out/r2f_pytest_mttkrp.log
timeout 200 python tools/perf_probe.py 1000 1000 1000 5000 32 5 > gpurun_out/r2f_probe_c2_fp64.log 2>&1; grep mttkrp gpurun_out/r2f_probe_c2_fp64.log
timeout 200 python tools/perf_probe.py 4096 4096 128 8192 64 5 > gpurun_out/r2f_probe_c3slab_fp64.log 2>&1; grep mttkrp gpurun_out/r2f_probe_c3slab_fp64.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_kernel -c 2 -o gpurun_out/r2f_em_full python tools/em_probe.py 512 32 > gpurun_out/r2f_em_ncu.log 2>&1
echo "ncu em rc=$?"
PROBE_PREC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mttkrp_tc_kernel -c 3 -o gpurun_out/r2f_tc_tf32_full python tools/perf_probe.py 4096 4096 64 8192 64 2 > gpurun_out/r2f_tc_ncu.log 2>&1
echo "ncu tc rc=$?"
ls -la gpurun_out/*.ncu-rep
