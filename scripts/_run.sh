python -m pytest tests/test_cm_loss_gpu.py -x -q 2>&1 | tail -40
