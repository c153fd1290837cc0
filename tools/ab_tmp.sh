python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shared_segment or golden" 2>&1 | tail -3
python bench.py --value-only --n-masks 4096 --tune gemm_shared_segments=1 2>&1 | tail -5 | cut -c1-600
