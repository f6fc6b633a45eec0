#!/bin/bash
# round 2, call 3K (1 GPU): PARAFAC2 ranks above 64 and slice-size variants; whole PARAFAC2 family
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "parafac2 or par2 or script1a or script2 or script14 or script11 or config4 or degenerate or em_imputation_cp_coupled or front_end" > gpurun_out/r3k_pytest_par2.log 2>&1
echo "par2 rc=$?"; tail -30 gpurun_out/r3k_pytest_par2.log
