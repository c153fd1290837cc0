python -m pytest tests/test_rle.py tests/test_runner_and_fill.py tests/test_sam2_seam.py tests/test_filter_and_graph.py -m gpu -x -q 2>&1 | tail -3
python tools/forward_timing.py 2>&1 | tail -8
