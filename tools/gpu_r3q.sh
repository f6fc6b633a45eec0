#!/bin/bash
# round 2, call 3Q (1 GPU): Jacobi in shared memory + split-K block products of the eigen-iteration: every caller, nvecs probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nvecs or linear_coupl or ortho or quadratic or script14 or script15 or script1a or front_end or handles_release" > gpurun_out/r3q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3q_pytest.log
timeout 600 python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r3q_nvecs_probe.log 2>&1; cut -c1-130 gpurun_out/r3q_nvecs_probe.log
