#!/bin/bash
# round 2, call K (1 GPU): makespan-based slab splits: MTTKRP parity, per-mode times at C2, a C3 slab (K=128) and c3k1024
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mttkrp or config2 or dimension_tree or golden or reduced_precision" > gpurun_out/r2k_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2k_pytest.log
for cfg in "1000 1000 1000 5000 32" "4096 4096 128 8192 64" "4096 4096 1024 8192 64" "2048 2048 2048 64 8"; do
  echo "== $cfg FP64"; timeout 300 python tools/perf_probe.py $cfg 5 2>&1 | grep -E "mttkrp mode|run "
done > gpurun_out/r2k_probe_fp64.log 2>&1
cat gpurun_out/r2k_probe_fp64.log
for cfg in "1000 1000 1000 5000 32" "4096 4096 128 8192 64"; do
  echo "== $cfg TF32"; PROBE_PREC=1 timeout 300 python tools/perf_probe.py $cfg 5 2>&1 | grep -E "mttkrp mode|run "
done > gpurun_out/r2k_probe_tf32.log 2>&1
cat gpurun_out/r2k_probe_tf32.log
