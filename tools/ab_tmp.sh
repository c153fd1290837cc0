python -m pytest tests/test_filter_and_graph.py -m gpu -x -q 2>&1 | tail -4
