#!/bin/bash
# Round-2 profile, second part (run under gpurun): launch list of one bench pass with the training step and the CRNN forward as
# single launches, and full captures of the kernels that changed late in the round.
set -x
mkdir -p gpurun_out
export SALSA_B200_CRNN_GRAPH=0
K='regex:salsa|stft|tracker|lite|eig_|iv_kernel|pcm16|conv_tc|conv_first|conv_wgrad|gru_|pack_input|avgpool2|freq_mean|head_finish|scaler|bn_|adam|seld_loss|augment|cutout'
B="python bench.py --clips 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 4 --train-batch 4 --no-train-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1500 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -o gpurun_out/r2b_prof_wgrad python scripts/run_wgrad_once.py > gpurun_out/ncu_full1.log 2>&1
B2="python bench.py --clips 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-other-configs --no-train --crnn-batch 2"
ncu --set full --clock-control none --import-source on -k regex:conv_first -s 0 -c 1 -o gpurun_out/r2b_prof_convfirst $B2 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_layer_mma -s 0 -c 1 -o gpurun_out/r2b_prof_gru $B2 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_train_bwd -s 0 -c 1 -o gpurun_out/r2b_prof_grubwd python scripts/run_train_once.py 32 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce -s 0 -c 1 -o gpurun_out/r2b_prof_bnreduce python scripts/run_train_once.py 8 > gpurun_out/ncu_full5.log 2>&1
for f in gpurun_out/ncu_full?.log; do tail -n 1 $f; done
ls -la gpurun_out/*.ncu-rep
