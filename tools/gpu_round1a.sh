set -x
nproc; free -g | head -2
BENCH_WORKLOAD=c2 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_c2_v2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_c3k1024_v2.json 2> gpurun_out/bench_c3.err
timeout 900 python tools/rank_sweep.py 2048 8 16 32 64 128 256 > gpurun_out/rank_sweep_2048.txt 2>&1
timeout 300 python tools/profile_target.py 1000 1000 1000 5000 32 > gpurun_out/plain_pt_c2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_c2_v2.csv python tools/profile_target.py 1000 1000 1000 5000 32 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mttkrp_(lead|inner|from_T)" -c 7 -o /tmp/mttkrp_c2_full python tools/profile_target.py 1000 1000 1000 5000 32 > gpurun_out/ncu_f1.log 2>&1
timeout 300 python tools/profile_target.py 4096 4096 64 8192 64 > gpurun_out/plain_pt_r64.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mttkrp_(lead|inner|from_T)" -c 7 -o /tmp/mttkrp_r64_full python tools/profile_target.py 4096 4096 64 8192 64 > gpurun_out/ncu_f2.log 2>&1
for n in mttkrp_c2_full mttkrp_r64_full; do
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/r01_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page details --csv > gpurun_out/r01_${n}_details.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv > gpurun_out/r01_${n}_source.csv 2>/dev/null
done
ls -la /tmp/*.ncu-rep
sz=$(du -sm /tmp/mttkrp_r64_full.ncu-rep | cut -f1); if [ "$sz" -lt 30 ]; then cp /tmp/mttkrp_r64_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out; ls -la gpurun_out
