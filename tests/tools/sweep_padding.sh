for cfg in "AAADMM_KSMALL=64" "AAADMM_KSMALL=32" "AAADMM_KSMALL=16" "AAADMM_KSMALL=8" "AAADMM_KSUBTREE=32" "AAADMM_KSUBTREE=16" "AAADMM_KSMALL=32 AAADMM_KSUBTREE=32" "AAADMM_KSMALL=16 AAADMM_KSUBTREE=16"; do
  echo "== $cfg"
  env $cfg AAADMM_LDLT_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu 2> /tmp/err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f' % d['value'], 'ldlt ms', d['roofline']['phases']['ldlt_apply']['ms'])
"
  grep "^ldlt:" /tmp/err.log | tail -1
done
