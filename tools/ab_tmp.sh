mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2f_n8b.json 2> gpurun_out/bench_r2f_n8b.err
tail -c 400 gpurun_out/bench_r2f_n8b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2f_n8b.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "per gpu us/img", round(8e6/d["value"],2), "e2e", round(d["e2e"]["value"],1), d["fill"]["fill_ms"], d["fill"]["allreduce_us"], d["fill"]["bit_identical_across_world"], d["run"]["ms_per_rank"], d["clocks"], d["run"]["cpu_affinity"])
PY
