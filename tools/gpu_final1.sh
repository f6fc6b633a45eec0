set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_c3k1024_v5.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
BENCH_WORKLOAD=c2 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_c2_v5.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref_v5.json 2> gpurun_out/bench_ref.err
timeout 900 python tools/bench_configs.py c1 c4 c5 --iters 10 > gpurun_out/r01_bench_configs.jsonl 2> gpurun_out/bench_configs.err
