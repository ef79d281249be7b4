python -m pytest tests/test_cm_loss_gpu.py -x -q 2>&1 | tail -2
for wl in iterative_480x640_1Mev iterative_128x128_b8_f4; do
  python scripts/kernel_times.py --workload $wl 2>&1 | tail -1 | cut -c1-330
done
