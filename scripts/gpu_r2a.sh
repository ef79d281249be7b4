#!/bin/bash
# round 2, GPU call A: full GPU test-suite, microbenchmarks (smem vs red, quad fetch), merge variants, graphs, train variants, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
lscpu | head -25 >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt; nvidia-smi topo -m >> gpurun_out/r2a_smi.txt 2>&1; numactl -H >> gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python scripts/smem_vs_red.py > gpurun_out/r2a_smem_vs_red.txt 2>&1
timeout 200 python scripts/red_patterns.py > gpurun_out/r2a_red_patterns.txt 2>&1
for wl in iterative_480x640_1Mev iterative_480x640_1Mev_edges; do
  timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2a_variants.txt 2>&1
  TEF_B200_LIB=build_variants/libtef_merge_any.so timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2a_variants.txt 2>&1
done
cat gpurun_out/r2a_variants.txt
timeout 300 python scripts/graph_loss_bench.py > gpurun_out/r2a_graph_loss.txt 2>&1; cat gpurun_out/r2a_graph_loss.txt
timeout 600 python scripts/train_variants.py > gpurun_out/r2a_train_variants.txt 2>&1; cat gpurun_out/r2a_train_variants.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
