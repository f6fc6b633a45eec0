#!/bin/bash
# round 2, call 3M (1 GPU): matrices carry a transposed copy (mode-2 product as a LEAD launch): parity of everything with
# matrices, slab probe
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "not parafac2 and not par2 and not prox_ and not config5" > gpurun_out/r3m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3m_pytest.log
timeout 600 python tools/perf_probe.py 4096 4096 128 8192 64 10 > gpurun_out/r3m_probe_c3slab.log 2>&1; tail -6 gpurun_out/r3m_probe_c3slab.log
