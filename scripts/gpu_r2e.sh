#!/bin/bash
# round 2, GPU call E: polarity in the sort key (A/B), fused validation images, ncu sector counts of the final kernels, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2e_pytest.log | tail -15
for wl in iterative_480x640_1Mev iterative_480x640_1Mev_edges iterative_128x128_b8_f4 linear_480x640_1Mev; do
  timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2e_variants.txt 2>&1
  TEF_B200_LIB=build_variants/libtef_nopol.so timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2e_variants.txt 2>&1
done
cat gpurun_out/r2e_variants.txt
for wl in validation_480x640_100kev validation_480x640_500kev; do timeout 200 python bench.py --workload $wl --steps 10 >> gpurun_out/r2e_validation.json 2>> gpurun_out/r2e.err; done
cut -c1-200 gpurun_out/r2e_validation.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iter_ -s 6 -c 2 -o gpurun_out/r2e_final python scripts/profile_step.py --workload iterative_480x640_1Mev --steps 5 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2>> gpurun_out/r2e.err; tail -c 1500 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e.err
