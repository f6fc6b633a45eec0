#!/bin/bash
# round 2, call 3F (1 GPU): generation barrier in the cooperative ADMM kernel, batched loads in prep_system: parity subset,
# slab probe, warm-cache launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not nvecs and not em_imputation and not prox_" > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3f_pytest.log
timeout 600 python tools/perf_probe.py 4096 4096 128 8192 64 10 > gpurun_out/r3f_probe_c3slab.log 2>&1; tail -4 gpurun_out/r3f_probe_c3slab.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r3f_launches_c3slab.csv python tools/perf_probe.py 4096 4096 128 8192 64 4 > gpurun_out/r3f_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/bench_configs.py c1 > gpurun_out/r3f_bench_c1.jsonl 2>&1; cut -c1-300 gpurun_out/r3f_bench_c1.jsonl
