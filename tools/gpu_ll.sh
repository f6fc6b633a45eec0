timeout 900 python -m pytest tests -m gpu -q -k "nvecs or init_front" 2>&1 | tail -4
timeout 300 python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/nvecs_probe3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unfold_gram" -c 4 --csv --log-file gpurun_out/nvecs_launches3.csv python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/nvecs_ncu.log 2>&1
grep -E "unfold_gram" gpurun_out/nvecs_launches3.csv | awk -F'","' '{print $5, $(NF-1), $NF}'
cat gpurun_out/nvecs_probe3.log | cut -c1-120
