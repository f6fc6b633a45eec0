#!/bin/bash
# round 2, call 3R (1 GPU): last whole-suite run of the round
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/r3r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3r_pytest.log; tail -8 gpurun_out/r3r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3r_smoke.log 2>&1; tail -1 gpurun_out/r3r_smoke.log
