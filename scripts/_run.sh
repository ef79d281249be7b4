python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for wl in iterative_480x640_1Mev iterative_128x128_b8_f4 iterative_128x128_b8_f1 iterative_480x640_100kev linear_480x640_1Mev; do
  python scripts/kernel_times.py --workload $wl 2>&1 | tail -1 | cut -c1-340
done
