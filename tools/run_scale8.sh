mkdir -p gpurun_out
lscpu | grep -i "thread\|core\|socket\|numa" | head -8
cat /sys/devices/system/cpu/cpu0/topology/thread_siblings_list /sys/devices/system/cpu/cpu1/topology/thread_siblings_list
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/s8_n8b.json 2> gpurun_out/s8_n8b.err
python -c "
import json
for f in ('n8b',):
    try:
        d=json.loads(open('gpurun_out/s8_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d['exec']['cpus_bound_per_replica'], d['exec']['ms_per_step_repetitions'], d['exec']['e2e_ms_per_step_repetitions'])
    except Exception as e: print(f,'ERR',e)
"
