set -e
python tools/profile_stage.py --images 3 > gpurun_out/prof_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upsample_pack2 -s 3 -c 1 -o gpurun_out/upsample_r2f -f python tools/profile_stage.py --images 3 > gpurun_out/prof_ncu.log 2>&1
ls -la gpurun_out/upsample_r2f.ncu-rep
