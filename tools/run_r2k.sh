mkdir -p gpurun_out
timeout 600 python tools/batched_only.py > gpurun_out/r2k_batched.json 2> gpurun_out/r2k_batched.err
FFLINS_VOXEL_STAMPS=1 timeout 600 python tools/batched_only.py > /dev/null 2> gpurun_out/r2k_stamps.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "undistort or golden or async or end_to_end or static or empty" 2>&1 | tail -8 > gpurun_out/r2k_tests.log
cat gpurun_out/r2k_batched.json; tail -5 gpurun_out/r2k_batched.err; grep "n=19968000\|n=768000" gpurun_out/r2k_stamps.txt | tail -4; cat gpurun_out/r2k_tests.log
