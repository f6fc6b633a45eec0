set -e
timeout 300 python tools/nvecs_probe.py 1000 1000 1000 32 > gpurun_out/nvecs_probe_c2.log 2>&1
timeout 300 python tools/nvecs_probe.py 4096 4096 64 64 >> gpurun_out/nvecs_probe_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unfold_gram|gram_reduce|dgemm_small|jacobi" -c 400 --csv --log-file gpurun_out/nvecs_launches.csv python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/nvecs_ncu.log 2>&1
tail -3 gpurun_out/nvecs_ncu.log
