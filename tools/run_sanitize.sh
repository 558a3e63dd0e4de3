mkdir -p gpurun_out
export FFLINS_NO_RESIDENT=1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/vox_bench.py robot 1 > gpurun_out/san_mem_vox.txt 2>&1; echo "memcheck vox rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/vox_bench.py robot 1 > gpurun_out/san_race_vox.txt 2>&1; echo "racecheck vox rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --config robot --steps 3 --warmup 3 --repeats 1 --settle 0 --prefill 4 --no-cpu --no-batched --no-shim > gpurun_out/san_mem_bench.txt 2>&1; echo "memcheck bench rc=$?"
for f in san_mem_vox san_race_vox san_mem_bench; do echo "== $f"; grep -i "ERROR SUMMARY\|RACECHECK SUMMARY\|Invalid\|hazard" gpurun_out/$f.txt | head -8; done
tail -3 gpurun_out/san_mem_bench.txt | cut -c1-300
