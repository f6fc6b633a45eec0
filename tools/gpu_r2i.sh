#!/bin/bash
# round 2, call I (1 GPU): EM kernel with the epilogue in the data layout; C4 launch list (why 3.1 ms vs 2.75 ms in round 1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "em_ or missing or znorm or gram" > gpurun_out/r2i_pytest_em.log 2>&1
echo "em tests rc=$?"; tail -4 gpurun_out/r2i_pytest_em.log
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r2i_em_probe.log 2>&1; cat gpurun_out/r2i_em_probe.log
timeout 300 python tools/em_probe.py 1000 32 > gpurun_out/r2i_em_probe_1000.log 2>&1; cat gpurun_out/r2i_em_probe_1000.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_c4_launches.csv python tools/bench_configs.py c4 --iters 4 > gpurun_out/r2i_c4_ncu.log 2>&1
echo "ncu c4 rc=$?"
timeout 300 python tools/bench_configs.py c4 --iters 20 2>&1 | cut -c1-300
