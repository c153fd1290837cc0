python bench.py --steps 6 --n-masks 4096 --n-images 64 --images-per-step 64 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; tail -c 300 gpurun_out/bench_config5.err
python bench.py --steps 10 --n-classes 1203 --n-images 64 --images-per-step 128 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; tail -c 300 gpurun_out/bench_config4.err
python -c "
import json
for f in ('gpurun_out/bench_config5.json','gpurun_out/bench_config4.json'):
    d=json.load(open(f)); print(f, d['config']['workload'], round(d['value'],1), 'img/s', round(d['us_per_image'],1), 'us', 'e2e', round(d['e2e']['value'],1), 'fill ms', round(d['fill']['fill_ms'],2), 'roofline', round(d['roofline']['frac'],3))
"
