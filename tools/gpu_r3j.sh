#!/bin/bash
# round 2, call 3J (N GPUs, N = $1): multi-GPU parity tests (N = 2 only) + bench under torchrun after the second half
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" -le 2 ]; then
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=4 > gpurun_out/r3j_pytest_multi_n$N.log 2>&1
echo "multi rc=$?" >> gpurun_out/r3j_pytest_multi_n$N.log; tail -8 gpurun_out/r3j_pytest_multi_n$N.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r3j_bench_n${N}_torchrun.json 2> gpurun_out/r3j_bench_n${N}_torchrun.err
echo "bench torchrun rc=$?"; tail -2 gpurun_out/r3j_bench_n${N}_torchrun.err; cut -c1-220 gpurun_out/r3j_bench_n${N}_torchrun.json
