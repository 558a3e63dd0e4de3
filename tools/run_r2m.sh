mkdir -p gpurun_out
timeout 120 tools/h2d_probe > gpurun_out/r2m_h2d.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2m_native.json 2> gpurun_out/r2m_native.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim --driver python > gpurun_out/r2m_python.json 2> gpurun_out/r2m_python.err
cat gpurun_out/r2m_h2d.txt; tail -c 400 gpurun_out/r2m_native.err
python -c "
import json
for f in ('native','python'):
    d=json.load(open('gpurun_out/r2m_%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['workload_sizes'], d['roofline']['residuals_per_launch'])
"
