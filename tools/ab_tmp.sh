set -x
python -m pytest tests/test_gpu_parity.py tests/test_rle.py tests/test_filter_and_graph.py -x -q 2>&1 | tail -8
python bench.py --value-only --steps 8 --n-images 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2 us/img', round(d['us_per_image'],2))"
