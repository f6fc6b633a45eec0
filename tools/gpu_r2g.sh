#!/bin/bash
# round 2, call G (1 GPU): TF32 with the A operand in tensor memory vs in shared memory; ncu --set full of em_kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reduced_precision" > gpurun_out/r2g_pytest_tc.log 2>&1
echo "tc rc=$?"; tail -4 gpurun_out/r2g_pytest_tc.log
for cfg in "1000 1000 1000 5000 32" "4096 4096 256 8192 64"; do
  echo "== $cfg TF32 A in TMEM"; PROBE_PREC=1 timeout 200 python tools/perf_probe.py $cfg 5 2>&1 | grep -E "mttkrp mode|run "
  echo "== $cfg TF32 A in smem"; AOADMM_TC_A_SMEM=1 PROBE_PREC=1 timeout 200 python tools/perf_probe.py $cfg 5 2>&1 | grep -E "mttkrp mode|run "
  echo "== $cfg BF16";           PROBE_PREC=2 timeout 200 python tools/perf_probe.py $cfg 5 2>&1 | grep -E "mttkrp mode|run "
done > gpurun_out/r2g_tc_probe.log 2>&1
cat gpurun_out/r2g_tc_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:::em_kernel -c 2 -o gpurun_out/r2g_em_full python tools/em_probe.py 512 32 > gpurun_out/r2g_em_ncu.log 2>&1
echo "ncu em rc=$?"; ls -la gpurun_out/r2g_em_full.ncu-rep
