python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --value-only --steps 8 --n-images 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('t_only us/img', round(d['us_per_image'],2))"
