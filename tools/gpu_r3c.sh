#!/bin/bash
# round 2, call 3C (1 GPU): whole suite after the EM / Gram / PARAFAC2 / prep_system / reduction changes; per-GPU-slab probe
# (non-MTTKRP time of a step, the 8-GPU efficiency limiter); C4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r3c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3c_pytest.log; tail -12 gpurun_out/r3c_pytest.log
timeout 600 python tools/perf_probe.py 4096 4096 128 8192 64 10 > gpurun_out/r3c_probe_c3slab.log 2>&1; cat gpurun_out/r3c_probe_c3slab.log
timeout 600 python tools/bench_configs.py c4 --iters 20 > gpurun_out/r3c_bench_c4.jsonl 2>&1; cut -c1-330 gpurun_out/r3c_bench_c4.jsonl
