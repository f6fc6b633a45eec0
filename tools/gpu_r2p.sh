#!/bin/bash
# round 2, call P (1 GPU): current launch list of C4 (after the leaner Jacobi / warp-parallel K-sums), EM probe
mkdir -p gpurun_out
timeout 300 python tools/c4_probe.py > gpurun_out/r2p_c4_plain.log 2>&1; echo "c4 rc=$?"; tail -1 gpurun_out/r2p_c4_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches_c4.csv python tools/c4_probe.py > gpurun_out/r2p_c4_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r2p_em_probe.log 2>&1; cat gpurun_out/r2p_em_probe.log
