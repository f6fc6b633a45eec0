#!/bin/bash
# round 2, call C (2 GPUs): sharded PARAFAC2 slices + DMMA EM pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=8 -x > gpurun_out/r2c_pytest_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2c_pytest_multi.log; tail -40 gpurun_out/r2c_pytest_multi.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "parafac2 or par2 or script1a or script2 or script14 or script11 or em_ or missing or config4 or znorm or release" > gpurun_out/r2c_pytest_par2_em.log 2>&1
echo "par2/em rc=$?" >> gpurun_out/r2c_pytest_par2_em.log; tail -6 gpurun_out/r2c_pytest_par2_em.log
timeout 300 python tools/em_probe.py > gpurun_out/r2c_em_probe.log 2>&1; tail -12 gpurun_out/r2c_em_probe.log
