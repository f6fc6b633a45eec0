#!/bin/bash
# round 2, call 3E (1 GPU): prep_system with the 2-D mapping; ncu --set full of the cooperative ADMM kernel and of prep_system
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or script6 or linear_coupl or chol or unconstrained or ridge" > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3e_pytest.log
timeout 600 python tools/perf_probe.py 4096 4096 128 8192 64 10 > gpurun_out/r3e_probe_c3slab.log 2>&1; tail -4 gpurun_out/r3e_probe_c3slab.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:admm_tile_kernel|prep_system_kernel' -s 12 -c 6 -o gpurun_out/r3e_admm_prep_full -f python tools/perf_probe.py 4096 4096 128 8192 64 2 > gpurun_out/r3e_ncu.log 2>&1; echo "ncu rc=$?"
