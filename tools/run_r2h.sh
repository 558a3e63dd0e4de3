mkdir -p gpurun_out
g++ -std=c++17 -O1 -o /tmp/shim_test tests/shim_test.cc -L ff-lins_b200 -lfflins_b200 -L oracle -loracle -Wl,-rpath,$PWD/ff-lins_b200 -Wl,-rpath,$PWD/oracle
/tmp/shim_test > gpurun_out/r2h_shim.txt 2>&1; echo "rc=$?" >> gpurun_out/r2h_shim.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/r2h_tests.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err
FFLINS_NO_RESIDENT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 640 -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 3 --warmup 3 --settle 0 --no-cpu --no-batched --no-shim > gpurun_out/r2h_ncu.log 2>&1
cat gpurun_out/r2h_shim.txt; tail -8 gpurun_out/r2h_tests.log; tail -c 300 gpurun_out/r2h_bench.err; tail -c 600 gpurun_out/r2h_ref.json
