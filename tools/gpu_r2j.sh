#!/bin/bash
# round 2, call J (1 GPU): PARAFAC2 parity after the warp-parallel K-sums, C4 / C5 rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "parafac2 or par2 or script1a or script2 or script14 or script11 or config4 or degenerate or release" > gpurun_out/r2j_pytest_par2.log 2>&1
echo "par2 rc=$?"; tail -4 gpurun_out/r2j_pytest_par2.log
timeout 600 python tools/bench_configs.py c4 c5 --iters 20 > gpurun_out/r2j_bench_configs.jsonl 2>&1; cut -c1-330 gpurun_out/r2j_bench_configs.jsonl
