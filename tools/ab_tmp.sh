for t in "exp7=0" "exp7=1" "exp7=1 --tune lowres_persistent=24" "exp7=1 --tune lowres_persistent=33" "exp0=1184" "exp0=296"; do
echo "== $t"; timeout 300 python tools/profile_stage.py --latency --images 4 --tune $t 2>&1 | grep "low_latency=True"
done
