mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window10.py tests/test_gpu_parity.py tests/test_gpu_closed_loop.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2e_tests.log
FFLINS_RES_TIMELINE=1 FFLINS_TRACE=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --repeats 1 > gpurun_out/r2e_tl.json 2> gpurun_out/r2e_tl.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_tests.log; grep -E "fflins|^trace" gpurun_out/r2e_tl.err | head -12
