python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --value-only "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'us/img', round(1e6/d['value'],2), d['clocks']['sm_mhz'])"; }
run
run --tune exp3=148
run --tune exp3=74
run --tune exp3=296
run --streams 24
run --streams 32
run --streams 8
run
