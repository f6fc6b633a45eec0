timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2_v7.json 2> gpurun_out/bench_c2_v7.err; echo "bench rc=$?"
