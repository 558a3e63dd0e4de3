mkdir -p gpurun_out
(timeout 300 python tools/vox_bench.py at128 20; echo "rc=$?"; FFLINS_VOXEL_STAMPS=1 timeout 300 python tools/vox_bench.py at128 3 2>&1 | grep -v "n=5000\|n=819\|n=70000\|n=20000"; timeout 300 python tools/vox_bench.py dense 20; FFLINS_VOXEL_STAMPS=1 timeout 300 python tools/vox_bench.py dense 2 2>&1 | grep "n=768000") > gpurun_out/r2i_vox.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "voxel or merge or golden or async or end_to_end" 2>&1 | tail -8 > gpurun_out/r2i_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
cat gpurun_out/r2i_vox.txt; cat gpurun_out/r2i_tests.log; tail -c 300 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench.json"))
print(d["value"], d["e2e"]["value"], {k:(v["ms_per_call"],v["share"]) for k,v in d["stages"].items()})
PY
