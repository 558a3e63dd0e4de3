mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k 'regex:vox_coop|undistort|project_batch|associate_knn|associate_plane' --launch-skip 12 -c 14 -f -o /tmp/r02_head_batched python tools/batched_only.py > gpurun_out/head_ncu_batched.log 2>&1
python tools/ncu_summary.py /tmp/r02_head_batched.ncu-rep > gpurun_out/r02_head_ncu_batched_summary.txt
ncu -i /tmp/r02_head_batched.ncu-rep --page raw --csv > gpurun_out/r02_head_batched_raw.csv
grep -c . gpurun_out/r02_head_batched_raw.csv; tail -2 gpurun_out/head_ncu_batched.log | cut -c1-300
