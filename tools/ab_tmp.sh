python no-time-to-train_b200/build.py --force > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or similarity or full_size" 2>&1 | tail -2
python bench.py --value-only --steps 8 --n-images 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sim bn128 us/img', round(d['us_per_image'],2))"
NTTT_EXTRA_NVCC_FLAGS="-DNTTT_SIM_NO_BN128" python no-time-to-train_b200/build.py > /dev/null 2>&1
python bench.py --value-only --steps 8 --n-images 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sim bn64  us/img', round(d['us_per_image'],2))"
