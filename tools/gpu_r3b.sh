#!/bin/bash
# round 2, call 3B (1 GPU): pipelined EM kernel, 8 GEMM warps of 8 x 32 + 8 comparison warps (one of them the producer)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "em_imputation or znorm" > gpurun_out/r3b_pytest_em.log 2>&1
echo "em rc=$?"; tail -3 gpurun_out/r3b_pytest_em.log
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r3b_em_probe.log 2>&1; cat gpurun_out/r3b_em_probe.log
timeout 300 python tools/em_probe.py 512 64 > gpurun_out/r3b_em_probe_r64.log 2>&1; cat gpurun_out/r3b_em_probe_r64.log
