mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2r_ref.json 2> gpurun_out/r2r_ref.err
FFLINS_QUIET=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2r_q0.json 2> gpurun_out/r2r_q0.err
timeout 600 python bench.py --config dense --steps 100 --warmup 10 --no-batched --no-shim > gpurun_out/r2r_dense.json 2> gpurun_out/r2r_dense.err
for c in robot lili ouster64; do timeout 600 python bench.py --config $c --steps 100 --warmup 10 --no-batched --no-shim > gpurun_out/r2r_$c.json 2> gpurun_out/r2r_$c.err; done
tail -c 300 gpurun_out/r2r_bench.err
python -c "
import json
for f in ('bench','q0','dense','robot','lili','ouster64'):
    try:
        d=json.load(open('gpurun_out/r2r_%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['workload_sizes'], (d.get('cpu_baseline') or {}).get('value'), (d.get('shim') or {}).get('queued',{}).get('shim_frames_per_s'))
    except Exception as e: print(f, 'ERR', e)
print(json.load(open('gpurun_out/r2r_ref.json'))['value'])
"
