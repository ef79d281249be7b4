#!/bin/bash
# round 2, GPU call L: fused network ops (ConvGRU, conv+bias+act, head up-sampling): parity tests, training step, kernel breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_netops_gpu.py tests/test_cm_loss_gpu.py -m gpu -q -k "netops or conv_gru or conv_bias or upsample or fused_network or graphed or train" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -30 gpurun_out/r2l_pytest.log
for mode in graph eager; do
  timeout 300 python bench.py --workload train_128x128_b8 --steps 5 --warmup 3 --train-mode $mode >> gpurun_out/r2l_train.json 2>> gpurun_out/r2l.err
done
cut -c1-260 gpurun_out/r2l_train.json; tail -5 gpurun_out/r2l.err
timeout 400 python scripts/train_kernels.py --dtype f32 --top 45 > gpurun_out/r2l_train_kernels_f32.txt 2>&1
head -52 gpurun_out/r2l_train_kernels_f32.txt
