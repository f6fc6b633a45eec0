set -x
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > gpurun_out/dist_check_n$N.log 2>&1
grep -v "^\*\|OMP" gpurun_out/dist_check_n$N.log | tail -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_bench_c3k1024_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; grep "^{" gpurun_out/r01_bench_c3k1024_n$N.json | cut -c1-300
