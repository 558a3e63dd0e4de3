mkdir -p gpurun_out
g++ -std=c++17 -O1 -o /tmp/shim_test tests/shim_test.cc -L ff-lins_b200 -lfflins_b200 -L oracle -loracle -Wl,-rpath,$PWD/ff-lins_b200 -Wl,-rpath,$PWD/oracle
/tmp/shim_test > gpurun_out/r2g_shim.txt 2>&1; echo "rc=$?" >> gpurun_out/r2g_shim.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_boundary.py::test_shim_runs_on_gpu 2>&1 | tail -15 > gpurun_out/r2g_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
cat gpurun_out/r2g_shim.txt; tail -8 gpurun_out/r2g_tests.log; tail -c 300 gpurun_out/r2g_bench.err
