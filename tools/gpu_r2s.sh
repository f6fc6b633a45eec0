#!/bin/bash
# round 2, call S (1 GPU): TMA-fed unfolding Gram kernel - nvecs parity, Gram kernel times at 4096x4096x64; C4 rate with
# the shorter rotation chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nvecs or front_end or handles_release" > gpurun_out/r2s_pytest_nvecs.log 2>&1
echo "nvecs rc=$?"; tail -4 gpurun_out/r2s_pytest_nvecs.log
timeout 600 python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r2s_nvecs_probe.log 2>&1; cut -c1-200 gpurun_out/r2s_nvecs_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gram -c 40 --csv --log-file gpurun_out/r2s_nvecs_launches.csv python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r2s_nvecs_ncu.log 2>&1; echo "ncu rc=$?"
grep -i "gram" gpurun_out/r2s_nvecs_launches.csv | cut -d, -f5,10- | head -12
timeout 600 python tools/bench_configs.py c4 --iters 20 > gpurun_out/r2s_bench_c4.jsonl 2>&1; cut -c1-330 gpurun_out/r2s_bench_c4.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "parafac2 or par2 or script2 or script14 or config4" > gpurun_out/r2s_pytest_par2.log 2>&1; echo "par2 rc=$?"; tail -2 gpurun_out/r2s_pytest_par2.log
