python -m pytest tests/test_flow_val_gpu.py tests/test_primitives_gpu.py -x -q 2>&1 | tail -15
for wl in validation_480x640_100kev validation_480x640_500kev; do python bench.py --workload $wl --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-700; done
