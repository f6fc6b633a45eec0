#!/bin/bash
# round 2, call D (1 GPU): tcgen05 TF32/BF16 MTTKRP first light (guarded by short timeouts), EM probe
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reduced_precision" > gpurun_out/r2d_pytest_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/r2d_pytest_tc.log; tail -40 gpurun_out/r2d_pytest_tc.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 200 python tools/em_probe.py > gpurun_out/r2d_em_probe.log 2>&1; tail -3 gpurun_out/r2d_em_probe.log
PROBE_PREC=1 timeout 120 python tools/perf_probe.py 1000 1000 1000 5000 32 5 > gpurun_out/r2d_probe_c2_tf32.log 2>&1; tail -8 gpurun_out/r2d_probe_c2_tf32.log
PROBE_PREC=2 timeout 120 python tools/perf_probe.py 1000 1000 1000 5000 32 5 > gpurun_out/r2d_probe_c2_bf16.log 2>&1; tail -8 gpurun_out/r2d_probe_c2_bf16.log
PROBE_PREC=3 timeout 120 python tools/perf_probe.py 1000 1000 1000 5000 32 5 > gpurun_out/r2d_probe_c2_tf32_legacy.log 2>&1; tail -8 gpurun_out/r2d_probe_c2_tf32_legacy.log
