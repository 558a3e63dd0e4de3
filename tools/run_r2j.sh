mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2j_tests.log
timeout 900 python bench.py --no-cpu --no-shim > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -8 gpurun_out/r2j_tests.log; tail -c 300 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_bench.json"))
print(d["value"], d["e2e"]["value"], json.dumps(d["roofline_batched"]))
PY
