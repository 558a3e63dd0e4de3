mkdir -p gpurun_out
for q in auto 1 0; do
  if [ $q = auto ]; then unset FFLINS_QUIET; else export FFLINS_QUIET=$q; fi
  timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-batched --no-shim > gpurun_out/r2o_$q.json 2> gpurun_out/r2o_$q.err
done
python -c "
import json
for q in ('auto','1','0'):
    d=json.load(open('gpurun_out/r2o_%s.json'%q)); print(q, d['value'], d['e2e']['value'], d['exec']['eval_roundtrip_us'], d['exec']['e2e_ms_per_step_repetitions'])
"
