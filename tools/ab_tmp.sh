set -e
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CMD="python bench.py --steps 1 --warmup 1 --n-images 16 --images-per-step 16 --value-only"
$CMD > gpurun_out/launchlist_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lowres|project_masks|gemm_|normalize|top1|nms_|upsample|ios_|decay|unpack|split_|fill_|proto_prepare|aa_" -c 900 --csv --log-file gpurun_out/launches_r2f_bench.csv $CMD > gpurun_out/launchlist_ncu.log 2>&1
wc -l gpurun_out/launches_r2f_bench.csv
