mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/drv_n1.json 2> gpurun_out/drv_n1.err; echo "rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/drv_ref.json 2> gpurun_out/drv_ref.err; echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/drv_n2.json 2> gpurun_out/drv_n2.err; echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/drv_ref2.json 2> gpurun_out/drv_ref2.err; echo "rc=$?"
python -c "
import json
for f in ('n1','ref','n2','ref2'):
    try:
        lines=[l for l in open('gpurun_out/drv_%s.json'%f).read().strip().splitlines() if l.startswith('{')]
        d=json.loads(lines[-1]); print(f, len(lines), d.get('impl'), d['n_gpus'], d['steps'], d['value'], d['e2e']['value'], d.get('gpu_launches'), (d.get('exec') or {}).get('cpus_bound_per_replica'))
    except Exception as e: print(f,'ERR',e)
"
