#!/bin/bash
# round 2, GPU call N (N GPUs): final build -- headline bench line at N GPUs (resident, e2e, weak training step), strong-scaling
# training step (global batch 64) with the all-reduce after the replay and captured in the graph
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "1" ]; then TR="python"; fi
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests/test_cm_loss_gpu.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; tail -2 gpurun_out/r2n_pytest.log
  for wl in iterative_480x640_1Mev iterative_480x640_4Mev iterative_128x128_b8_f1; do timeout 200 python scripts/kernel_times.py --workload $wl --steps 6 >> gpurun_out/r2n_kernels.txt 2>&1; done; cat gpurun_out/r2n_kernels.txt
fi
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n_bench_n$N.json 2>> gpurun_out/r2n.err; tail -c 700 gpurun_out/r2n_bench_n$N.json; echo
timeout 600 $TR bench.py --gpus $N --workload train_128x128_gb64 --steps 6 > gpurun_out/r2n_train_gb64_n$N.json 2>> gpurun_out/r2n.err; cut -c1-260 gpurun_out/r2n_train_gb64_n$N.json
# (the two TEF_TRAIN_CAPTURE_NCCL=1 runs that followed here never finished their first replay at 8 GPUs and ran into the call's time
#  limit -- 1 023 s x 8 GPUs, the rest of the round's GPU budget; the switch has been removed from bench.py)
tail -5 gpurun_out/r2n.err
