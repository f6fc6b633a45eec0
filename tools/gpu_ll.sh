timeout 1200 python bench.py > gpurun_out/bench_c3k1024_v6.json 2> gpurun_out/bench_c3k1024_v6.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c3k1024_v6.err
