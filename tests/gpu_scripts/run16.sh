python -m pytest tests/test_gpu_refmodel.py -m gpu -x -q -s 2>&1 | grep -E "HybridMLP vs|passed|failed|^E " | head
