mkdir -p gpurun_out
for c in robot lili ouster64 dense at128; do timeout 600 python bench.py --config $c --steps 100 --warmup 10 --no-batched --no-shim > gpurun_out/head_$c.json 2> gpurun_out/head_$c.err; done
python -c "
import json
for f in ('robot','lili','ouster64','at128','dense'):
    d=json.load(open('gpurun_out/head_%s.json'%f)); print(f, d['config']['points_per_frame'], d['workload_sizes']['Q'], min(d['workload_sizes']['M_window']), max(d['workload_sizes']['M_window']), d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['exec']['eval_roundtrip_us'])
"
