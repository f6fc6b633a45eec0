timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3k1024_v8.json 2> gpurun_out/bench_c3k1024_v8.err; echo "bench rc=$?"
