mkdir -p gpurun_out
timeout 120 ./tools/pcie_probe > gpurun_out/r2c_probe.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_window10.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c_tests.log
for cfg in "0 0" "1000 0" "3000 200" "0 200"; do
  set -- $cfg
  FFLINS_RES_POLL_SLEEP_NS=$1 FFLINS_RES_WORKER_SLEEP_NS=$2 FFLINS_TRACE=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --repeats 1 > gpurun_out/r2c_bench_$1_$2.json 2> gpurun_out/r2c_bench_$1_$2.err
done
cat gpurun_out/r2c_probe.txt gpurun_out/r2c_tests.log
