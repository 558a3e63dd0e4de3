# Round-end evidence: GPU tests, full bench line, CPU arm, ncu launch list and --set full capture (summaries only travel back).
mkdir -p gpurun_out
T=${1:-r02_v16}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/${T}_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_ref.err
FFLINS_NO_RESIDENT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --settle 0 --no-cpu --no-batched --no-shim > gpurun_out/${T}_ncu_list.log 2>&1
K='regex:vox_coop|undistort|associate_knn|associate_plane|factor_kernel|chi2_kernel|hash_insert|vox_frame|ins_prep|project_batch'
FFLINS_NO_RESIDENT=1 timeout 900 ncu --set full --clock-control none -k "$K" --launch-skip 446 -c 27 -f -o /tmp/${T}_pipeline python bench.py --steps 2 --warmup 3 --settle 0 --no-cpu --no-batched --no-shim > gpurun_out/${T}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/${T}_pipeline.ncu-rep > gpurun_out/${T}_ncu_pipeline_summary.txt
ncu -i /tmp/${T}_pipeline.ncu-rep --page raw --csv > gpurun_out/${T}_pipeline_raw.csv
tail -4 gpurun_out/${T}_gpu_tests.txt; tail -c 300 gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['workload_sizes'], d['cpu_baseline']['value'], d['shim']['queued']['shim_frames_per_s'], d['roofline']['frac'])
print(json.load(open('gpurun_out/${T}_reference_arm.json'))['value'])
"
