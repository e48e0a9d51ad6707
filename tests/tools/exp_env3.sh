# Developer tool: bench.py --config cfg3 under a list of environment settings
for cfg in "$@"; do
  echo "== cfg3 $cfg"
  env $cfg timeout 300 python bench.py --config cfg3 --steps 2 --warmup 1 --no-cpu 2> /tmp/err3.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f' % d['value'], 'ms/turn', d['roofline']['ms_per_turn'], 'e2e', d['e2e']['value'])
"
done
