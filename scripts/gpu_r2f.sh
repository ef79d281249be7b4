#!/bin/bash
# round 2, GPU call F: interleaved forward/backward chains in iter_fwd_kernel (A/B), kernel breakdown of the graphed training step
mkdir -p gpurun_out
for wl in iterative_480x640_1Mev iterative_480x640_1Mev_edges iterative_128x128_b8_f4; do
  timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2f_variants.txt 2>&1
  TEF_B200_LIB=build_variants/libtef_chain2.so timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2f_variants.txt 2>&1
done
cat gpurun_out/r2f_variants.txt
timeout 400 python scripts/train_kernels.py --dtype f32 --top 60 > gpurun_out/r2f_train_kernels_f32.txt 2>&1
head -70 gpurun_out/r2f_train_kernels_f32.txt
