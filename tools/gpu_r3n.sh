#!/bin/bash
# round 2, call 3N (1 GPU): final whole-suite run and bench line of the round
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/r3n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3n_pytest.log; tail -8 gpurun_out/r3n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3n_smoke.log 2>&1; tail -1 gpurun_out/r3n_smoke.log
timeout 1200 python bench.py > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r3n_bench.err; cut -c1-250 gpurun_out/r3n_bench.json
