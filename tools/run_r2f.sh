mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2f_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
cat gpurun_out/r2f_tests.log; tail -c 400 gpurun_out/r2f_bench.err
