mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2n_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -8 gpurun_out/r2n_tests.log; tail -c 400 gpurun_out/r2n_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2n_bench.json')); print(d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['exec']['eval_path'], d['exec']['e2e_ms_per_step_repetitions'])
"
