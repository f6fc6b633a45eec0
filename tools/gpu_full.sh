set -x
timeout 900 python -m pytest tests -m gpu -q -k "orthonormal or quadratic or tparafac2" 2>&1 | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_c3k1024_v3.json 2> gpurun_out/bench_c3.err; tail -5 gpurun_out/bench_c3.err
