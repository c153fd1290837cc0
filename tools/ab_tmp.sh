python -m pytest tests/test_filter_and_graph.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_stage.py --latency --images 4 2>&1 | grep "graph replay"
