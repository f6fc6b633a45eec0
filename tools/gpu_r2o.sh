#!/bin/bash
# round 2, call O (4 GPUs): bench under torchrun (the N=4 point of the scaling curve)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2o_bench_n4_torchrun.json 2> gpurun_out/r2o_bench_n4_torchrun.err
echo "bench torchrun rc=$?"; tail -2 gpurun_out/r2o_bench_n4_torchrun.err; cut -c1-220 gpurun_out/r2o_bench_n4_torchrun.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2o_bench_ref_n4.json 2> gpurun_out/r2o_bench_ref_n4.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/r2o_bench_ref_n4.json
