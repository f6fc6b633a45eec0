#!/bin/bash
# round 2, call 3L (1 GPU): final whole-suite run (after the PARAFAC2 rank > 64 paths), ncu --set full of the TMA Gram
# kernel and of the register-resident PARAFAC2 polar-factor kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/r3l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3l_pytest.log; tail -8 gpurun_out/r3l_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:unfold_gram_tma_kernel' -c 2 -o gpurun_out/r3l_gram_tma_full -f python tools/nvecs_probe.py 4096 4096 64 64 > gpurun_out/r3l_ncu_gram.log 2>&1; echo "ncu gram rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:par2_B_step1_reg_kernel' -s 5 -c 2 -o gpurun_out/r3l_par2_reg_full -f python tools/c4_probe.py > gpurun_out/r3l_ncu_par2.log 2>&1; echo "ncu par2 rc=$?"
