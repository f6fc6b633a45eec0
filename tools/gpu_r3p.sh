#!/bin/bash
# round 2, call 3P (1 GPU): nvecs probe repeated (timing stability after the shared-memory Jacobi / split-K products)
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r3p_nvecs_probe_$i.log 2>&1; cut -c1-100 gpurun_out/r3p_nvecs_probe_$i.log; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3p_nvecs_launches.csv python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r3p_ncu.log 2>&1; echo "ncu rc=$?"
