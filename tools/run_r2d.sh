FFLINS_RES_TIMELINE=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-batched --repeats 1 > gpurun_out/r2d_tl.json 2> gpurun_out/r2d_tl.err
grep fflins gpurun_out/r2d_tl.err
