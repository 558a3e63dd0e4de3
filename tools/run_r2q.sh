mkdir -p gpurun_out
FFLINS_COPY_TIMELINE=1 timeout 600 python bench.py --steps 50 --warmup 10 --repeats 1 --no-cpu --no-batched --no-shim > gpurun_out/r2q.json 2> gpurun_out/r2q.err
grep "copy timeline" gpurun_out/r2q.err | tail -4
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2q_tests.log
DRV_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -5 gpurun_out/r2q_tests.log
grep -A12 "host-buffer steps" gpurun_out/r2q_bench.err | cut -c1-130 | grep "assoc\|window_eval\|merge\|process"
python -c "
import json
d=json.load(open('gpurun_out/r2q_bench.json')); print(d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['exec']['e2e_ms_per_step_repetitions'])
"
