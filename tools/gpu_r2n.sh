#!/bin/bash
# round 2, call N (N GPUs, N = $1): multi-GPU parity tests + bench in both launch forms
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > gpurun_out/r2n_gpu_n$N.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=8 > gpurun_out/r2n_pytest_multi_n$N.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2n_pytest_multi_n$N.log; tail -14 gpurun_out/r2n_pytest_multi_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n_bench_n${N}_torchrun.json 2> gpurun_out/r2n_bench_n${N}_torchrun.err
echo "bench torchrun rc=$?"; tail -2 gpurun_out/r2n_bench_n${N}_torchrun.err; cut -c1-220 gpurun_out/r2n_bench_n${N}_torchrun.json
if [ "$N" -le 2 ]; then
timeout 1200 python bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n_bench_n${N}_single.json 2> gpurun_out/r2n_bench_n${N}_single.err
echo "bench single-process rc=$?"; tail -2 gpurun_out/r2n_bench_n${N}_single.err; cut -c1-220 gpurun_out/r2n_bench_n${N}_single.json
fi
