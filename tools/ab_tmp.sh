python -m pytest tests/test_gpu_parity.py -x -q -k "nms or full_size" 2>&1 | tail -4
