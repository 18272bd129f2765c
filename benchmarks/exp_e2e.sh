set -u
mkdir -p gpurun_out
for cfg in "1 5" "64 5" "64 20" "16 20"; do
  set -- $cfg
  ASP_E2E_CHUNKS=$1 ASP_E2E_STEPS=$2 timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); e=j['e2e']; print('chunks $1 steps $2', 'ms_per_step', j['ms_per_step'], 'e2e ms', e['ms_per_step'], 'e2e value', e['value'], 'n', e['particles'])
"
done
