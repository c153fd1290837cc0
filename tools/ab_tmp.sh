python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "upsample or pipeline" 2>&1 | tail -2
for i in 1 2; do python bench.py --value-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value us/img', round(1e6/d['value'],2), d['clocks']['sm_mhz'])"; done
