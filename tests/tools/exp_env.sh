# Developer tool: bench.py under a list of environment settings (one per argument, "A=1 B=2" form); prints it/s and the apply time
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg AAADMM_LDLT_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu 2> /tmp/err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f' % d['value'], 'ldlt ms', d['roofline']['phases']['ldlt_apply']['ms'])
"
  grep "^ldlt:" /tmp/err.log | tail -1 | cut -c1-400
done
