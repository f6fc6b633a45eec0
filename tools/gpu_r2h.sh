#!/bin/bash
# round 2, call H (1 GPU): whole GPU suite after the LEAD / prep_system / tcgen05 changes, ncu --set full of em_kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -22 gpurun_out/r2h_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^em_kernel$' -c 2 -o gpurun_out/r2h_em_full python tools/em_probe.py 512 32 > gpurun_out/r2h_em_ncu.log 2>&1
echo "ncu em rc=$?"; ls -la gpurun_out/r2h_em_full.ncu-rep
timeout 300 python tools/perf_probe.py 1000 1000 1000 5000 32 10 2>&1 | grep -E "run |phase" 
timeout 300 python tools/bench_configs.py c1 c4 --iters 10 > gpurun_out/r2h_bench_configs.jsonl 2>&1; cut -c1-400 gpurun_out/r2h_bench_configs.jsonl
