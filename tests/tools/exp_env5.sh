for cfg in "X=1" "AAADMM_KSMALL=128 AAADMM_KSUBTREE=128" "AAADMM_KSMALL=96 AAADMM_KSUBTREE=96" "AAADMM_TASK_SLOTS=100" "AAADMM_TASK_SLOTS=400"; do
  echo "== cfg5 $cfg"
  env $cfg AAADMM_LDLT_VERBOSE=1 python bench.py --config cfg5 --steps 1 --warmup 1 --no-cpu 2>/tmp/e5.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('value', round(d['value'],1), 'scenes/s', round(d['scenes_per_s'],2), d['roofline']['phases']['ldlt_apply'])
"
  grep "^ldlt:" /tmp/e5.log | tail -1 | cut -c1-150
done
