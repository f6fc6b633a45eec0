#!/bin/bash
# round 2, call 3I (1 GPU): single-GPU record after the second half of the round - whole suite, smoke, bench line, launch
# list of the bench command, C4 / C5 rates, EM probe, slab probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r3i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3i_pytest.log; tail -12 gpurun_out/r3i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3i_smoke.log 2>&1; tail -2 gpurun_out/r3i_smoke.log
timeout 1200 python bench.py > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r3i_bench.err; cut -c1-250 gpurun_out/r3i_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r3i_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3i_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 600 python tools/bench_configs.py c4 c5 --iters 10 > gpurun_out/r3i_bench_c4_c5.jsonl 2>&1; cut -c1-200 gpurun_out/r3i_bench_c4_c5.jsonl
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r3i_em_probe.log 2>&1; cat gpurun_out/r3i_em_probe.log
timeout 600 python tools/perf_probe.py 4096 4096 128 8192 64 10 > gpurun_out/r3i_probe_c3slab.log 2>&1; tail -3 gpurun_out/r3i_probe_c3slab.log
