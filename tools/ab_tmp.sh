ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fill_|lowres|project|split|gemm|normalize|top1|nms|upsample|ios|decay|unpack|rle|proto|aa_" -c 900 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 1 --warmup 1 --n-images 16 --images-per-step 16 --value-only > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_r2_bench.csv
