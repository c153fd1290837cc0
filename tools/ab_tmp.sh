python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_strip" 2>&1 | tail -5
