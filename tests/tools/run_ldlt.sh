python -m pytest tests/test_gpu_parity.py -x -q -k "ldlt" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ldlt ms', d['roofline']['phases']['ldlt_apply']['ms'], 'frac', d['roofline']['frac'])
"
