set -x
python tools/profile_stage.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r2a.csv python tools/profile_stage.py > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upsample_pack_kernel -s 3 -c 2 -o gpurun_out/prof_upsample_staged python tools/profile_stage.py > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upsample_pack_kernel -s 3 -c 2 -o gpurun_out/prof_upsample_direct python tools/profile_stage.py --tune upsample_stage_bytes=0 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu2.log
