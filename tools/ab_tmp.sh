python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/profile_stage.py --latency --images 4 2>&1 | tail -3
python bench.py --value-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value us/img', round(1e6/d['value'],2), d['clocks']['sm_mhz'])"
python tools/forward_timing.py 2>&1 | grep "2nd\|segment\|C call" | cut -c1-70
