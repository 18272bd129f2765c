set -u
mkdir -p gpurun_out
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/round_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/round_bench.log > gpurun_out/round_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/round_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])
for k,v in (d.get('other_configs') or {}).items():
    if isinstance(v, dict): print(k, v)
    else:
        for l in v: print(k, l.get('bench'), l.get('particles'), l.get('ms') or l.get('ms_per_step') or l.get('ms_total'))
PY
tail -4 gpurun_out/round_bench.log
