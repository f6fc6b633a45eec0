#!/bin/bash
# round 2, call 3S (1 GPU): quick sanity of the final build (CP + matrix family, smoke)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "script6 or cp_and_matrix or mttkrp_all_modes" > gpurun_out/r3s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3s_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
