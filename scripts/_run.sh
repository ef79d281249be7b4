python -m pytest tests/test_cm_loss_gpu.py tests/test_properties_fullsize_gpu.py -x -q 2>&1 | tail -3
for wl in iterative_480x640_1Mev iterative_128x128_b8_f4 iterative_480x640_1Mev_edges; do
  python scripts/kernel_times.py --workload $wl 2>&1 | tail -1 | cut -c1-330
done
