#!/bin/bash
# round 2, call B (2 GPUs): multi-GPU parity tests (one caller / N GPUs, torchrun workers), PARAFAC2 parity after the
# segmented prox, bench at N=2 in both launch forms (incl. the C3 full-size leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > gpurun_out/r2b_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=8 > gpurun_out/r2b_pytest_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2b_pytest_multi.log; tail -25 gpurun_out/r2b_pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "parafac2 or par2 or script1a or script2 or script14 or script11" > gpurun_out/r2b_pytest_par2.log 2>&1
echo "par2 rc=$?" >> gpurun_out/r2b_pytest_par2.log; tail -4 gpurun_out/r2b_pytest_par2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2_torchrun.json 2> gpurun_out/r2b_bench_n2_torchrun.err
echo "bench torchrun rc=$?"; tail -3 gpurun_out/r2b_bench_n2_torchrun.err; cut -c1-300 gpurun_out/r2b_bench_n2_torchrun.json
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 --no-c3-full > gpurun_out/r2b_bench_n2_single.json 2> gpurun_out/r2b_bench_n2_single.err
echo "bench single-process rc=$?"; tail -3 gpurun_out/r2b_bench_n2_single.err; cut -c1-300 gpurun_out/r2b_bench_n2_single.json
