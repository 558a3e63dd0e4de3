mkdir -p gpurun_out
K='regex:vox_coop|undistort|associate_knn|associate_plane|factor_kernel|chi2_kernel|hash_insert|vox_frame|ins_prep|project_batch'
FFLINS_NO_RESIDENT=1 timeout 900 ncu --set full --clock-control none -k "$K" --launch-skip 446 -c 27 -f -o /tmp/r02_pipeline python bench.py --steps 2 --warmup 3 --settle 0 --no-cpu --no-batched --no-shim > gpurun_out/r2l_ncu1.log 2>&1
echo "rc1=$?"
python tools/ncu_summary.py /tmp/r02_pipeline.ncu-rep > gpurun_out/r02_pipeline_summary.txt
ncu -i /tmp/r02_pipeline.ncu-rep --page raw --csv > gpurun_out/r02_pipeline_raw.csv
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:vox_coop|undistort|project_batch' --launch-skip 12 -c 12 -f -o /tmp/r02_batched python tools/batched_only.py > gpurun_out/r2l_ncu2.log 2>&1
echo "rc2=$?"
python tools/ncu_summary.py /tmp/r02_batched.ncu-rep > gpurun_out/r02_batched_summary.txt
ncu -i /tmp/r02_batched.ncu-rep --page raw --csv > gpurun_out/r02_batched_raw.csv
ls -la /tmp/*.ncu-rep
sz=$(stat -c %s /tmp/r02_batched.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/r02_batched.ncu-rep gpurun_out/; fi
FFLINS_NO_RESIDENT=1 FFLINS_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2l_nores.json 2> gpurun_out/r2l_nores.err
FFLINS_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2l_res.json 2> gpurun_out/r2l_res.err
timeout 600 python tools/batched_only.py > gpurun_out/r2l_batched.json 2> gpurun_out/r2l_batched.err
du -sh gpurun_out; tail -3 gpurun_out/r2l_batched.err; python -c "
import json
print(json.load(open('gpurun_out/r2l_batched.json')).get('associate'))
for f in ('nores','res'):
    d=json.load(open('gpurun_out/r2l_%s.json'%f)); print(f, d['value'], d['e2e']['value'])
"
grep "trace\[e2e\]" gpurun_out/r2l_nores.err | head -4; grep "trace\[e2e\]" gpurun_out/r2l_res.err | head -4
