mkdir -p gpurun_out
DRV_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2p.json 2> gpurun_out/r2p.err
grep -A14 "host-buffer steps" gpurun_out/r2p.err | cut -c1-150
grep -A14 "device-resident steps" gpurun_out/r2p.err | cut -c1-150
