#!/bin/bash
# round 2, call 3D (1 GPU): launch list (warm caches: --cache-control none) of the per-GPU slab of the 8-GPU bench
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r3d_launches_c3slab.csv python tools/perf_probe.py 4096 4096 128 8192 64 4 > gpurun_out/r3d_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r3d_ncu.log
