python -m pytest tests/test_gpu_parity.py -x -q -k "threshold_pack" 2>&1 | tail -2
