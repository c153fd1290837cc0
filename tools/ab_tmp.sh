python tools/profile_stage.py --latency --images 4 2>&1 | tail -4
