#!/bin/bash
# round 2, GPU call C: GPU tests (Linear fast paths, quad test), Linear / config-1 kernel times, new bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2c_pytest.log | tail -15
for wl in linear_480x640_1Mev linear_128x128_b8_f4 iterative_480x640_1Mev; do
  timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2c_variants.txt 2>&1
done
cat gpurun_out/r2c_variants.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 3000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 300 python bench.py --workload linear_480x640_1Mev --steps 10 > gpurun_out/r2c_bench_linear.json 2>> gpurun_out/r2c_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_reference.json 2>> gpurun_out/r2c_bench.err; tail -c 600 gpurun_out/r2c_bench_reference.json
