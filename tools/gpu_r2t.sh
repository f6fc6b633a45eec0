#!/bin/bash
# round 2, call T (1 GPU): Gram kernel with makespan splits; full launch list of one nvecs call
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nvecs" > gpurun_out/r2t_pytest_nvecs.log 2>&1
echo "nvecs rc=$?"; tail -2 gpurun_out/r2t_pytest_nvecs.log
timeout 600 python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r2t_nvecs_probe.log 2>&1; cut -c1-120 gpurun_out/r2t_nvecs_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2t_nvecs_launches.csv python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r2t_nvecs_ncu.log 2>&1; echo "ncu rc=$?"
