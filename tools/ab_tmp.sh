python tools/profile_stage.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:upsample_pack2_kernel -s 3 -c 1 -o gpurun_out/prof_upsample_v2b python tools/profile_stage.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
