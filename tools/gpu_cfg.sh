set -x
timeout 300 python -m pytest tests -m gpu -q -k "tparafac2" 2>&1 | tail -5
timeout 1500 python tools/bench_configs.py c1 c4 c5 --iters 10 > gpurun_out/r01_bench_configs.jsonl 2> gpurun_out/bench_configs.err; tail -5 gpurun_out/bench_configs.err; cat gpurun_out/r01_bench_configs.jsonl | cut -c1-400
