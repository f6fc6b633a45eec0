#!/bin/bash
# round 2, call E (1 GPU): reduced-precision tests (all), ncu launch list of a masked 512^3 iteration (EM pass), full bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reduced_precision" > gpurun_out/r2e_pytest_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/r2e_pytest_tc.log; tail -15 gpurun_out/r2e_pytest_tc.log
timeout 300 python tools/em_probe.py 512 32 > gpurun_out/r2e_em_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_em_launches.csv python tools/em_probe.py 512 32 > gpurun_out/r2e_em_ncu.log 2>&1
echo "ncu rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err; cut -c1-200 gpurun_out/r2e_bench.json
