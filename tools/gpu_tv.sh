set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python tools/tv_probe.py 8
timeout 300 python tools/tv_probe.py 8 unimodality
BENCH_WORKLOAD=c2 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_c2_v4.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cut -c1-200 gpurun_out/r01_bench_c2_v4.json
