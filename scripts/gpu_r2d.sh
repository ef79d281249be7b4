#!/bin/bash
# round 2, GPU call D (N GPUs): validation rewrite tests + bench, then the multi-GPU arms at N = $1
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" = "single" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
  grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2d_pytest.log | tail -15
  for wl in validation_480x640_100kev validation_480x640_500kev; do timeout 200 python bench.py --workload $wl --steps 10 >> gpurun_out/r2d_validation.json 2>> gpurun_out/r2d.err; done
  cat gpurun_out/r2d_validation.json | cut -c1-400
  timeout 300 python bench.py --workload train_128x128_gb64 --steps 6 > gpurun_out/r2d_train_gb64_n1.json 2>> gpurun_out/r2d.err; cut -c1-300 gpurun_out/r2d_train_gb64_n1.json
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2d_bench_n$N.json 2>> gpurun_out/r2d.err; tail -c 1200 gpurun_out/r2d_bench_n$N.json
timeout 600 $TR bench.py --gpus $N --workload train_128x128_gb64 --steps 6 > gpurun_out/r2d_train_gb64_n$N.json 2>> gpurun_out/r2d.err; cut -c1-300 gpurun_out/r2d_train_gb64_n$N.json
timeout 600 $TR bench.py --gpus $N --workload train_128x128_gb64 --steps 6 --train-mode eager > gpurun_out/r2d_train_gb64_eager_n$N.json 2>> gpurun_out/r2d.err; cut -c1-300 gpurun_out/r2d_train_gb64_eager_n$N.json
tail -5 gpurun_out/r2d.err
