#!/bin/bash
# round 2, call 3H (1 GPU): finer-sampled ncu source profile of the cooperative ADMM kernel; parity subset for reduce_jobs
mkdir -p gpurun_out
timeout 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:admm_tile_kernel' -s 8 -c 4 -o gpurun_out/r3h_admm_full -f python tools/perf_probe.py 4096 4096 128 8192 64 2 > gpurun_out/r3h_ncu.log 2>&1; echo "ncu rc=$?"
