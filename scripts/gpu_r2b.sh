#!/bin/bash
# round 2, GPU call B: GPU tests with the quad-cell copies + exact deterministic sums, quad on/off, ncu of the event kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -E "Linf|passed|failed|FAILED|rc=" gpurun_out/r2b_pytest.log | tail -25
for wl in iterative_480x640_1Mev iterative_480x640_1Mev_edges iterative_128x128_b8_f4; do
  timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2b_variants.txt 2>&1
  TEF_QUAD=0 timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2b_variants.txt 2>&1
  TEF_B200_LIB=build_variants/libtef_bwd3.so timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2b_variants.txt 2>&1
done
cat gpurun_out/r2b_variants.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iter_ -s 6 -c 2 -o gpurun_out/r2b_quad python scripts/profile_step.py --workload iterative_480x640_1Mev --steps 5 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json
