TEF_FUZZ_SEEDS=400 python -m pytest tests/test_fuzz_gpu.py -q -x 2>&1 | tail -30
