# Developer tool: bench.py under a list of environment settings ("VAR=val VAR2=val2" per argument)
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f' % d['value'], 'ldlt ms', d['roofline']['phases']['ldlt_apply']['ms'])
"
done
