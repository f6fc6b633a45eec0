set -e
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3k1024_s3.json 2> gpurun_out/bench_c3k1024_s3.err
echo "plain rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_c3k1024.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_bench_c3k1024.csv
