#!/bin/bash
# round 2, call A (1 GPU): whole GPU test-suite incl. the full-size parity tests, then the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_gpu.txt
free -g | head -2 >> gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
