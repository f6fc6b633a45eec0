#!/bin/bash
# round 2, call V (1 GPU): ncu --set full of the current EM kernel
mkdir -p gpurun_out
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r2v_em_probe.log 2>&1; cat gpurun_out/r2v_em_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^em_kernel$' -c 2 -o gpurun_out/r2v_em_full -f python tools/em_probe.py 512 32 > gpurun_out/r2v_em_ncu.log 2>&1; echo "ncu rc=$?"
