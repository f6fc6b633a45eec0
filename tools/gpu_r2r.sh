#!/bin/bash
# round 2, call R (1 GPU): PARAFAC2 parity with the warp-per-slice polar-factor kernel, C4 rate and launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "parafac2 or par2 or script1a or script2 or script14 or script11 or config4 or degenerate or em_imputation_cp_coupled or front_end" > gpurun_out/r2r_pytest_par2.log 2>&1
echo "par2 rc=$?"; tail -4 gpurun_out/r2r_pytest_par2.log
timeout 600 python tools/bench_configs.py c4 --iters 20 > gpurun_out/r2r_bench_c4.jsonl 2>&1; cut -c1-330 gpurun_out/r2r_bench_c4.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches_c4.csv python tools/c4_probe.py > gpurun_out/r2r_c4_ncu.log 2>&1; echo "ncu rc=$?"
