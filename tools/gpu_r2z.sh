#!/bin/bash
# round 2, call Z (1 GPU): pipelined EM kernel - parity, probe, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "em_imputation or znorm or handles_release" > gpurun_out/r2z_pytest_em.log 2>&1
echo "em rc=$?"; tail -5 gpurun_out/r2z_pytest_em.log
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r2z_em_probe.log 2>&1; cat gpurun_out/r2z_em_probe.log
timeout 300 python tools/em_probe.py 512 64 > gpurun_out/r2z_em_probe_r64.log 2>&1; cat gpurun_out/r2z_em_probe_r64.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:em_pipe_kernel' -c 2 -o gpurun_out/r2z_em_pipe_full -f python tools/em_probe.py 512 32 > gpurun_out/r2z_em_ncu.log 2>&1; echo "ncu rc=$?"
