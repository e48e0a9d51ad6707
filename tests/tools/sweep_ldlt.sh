run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['config']['factor'], round(d['roofline']['phases']['ldlt_apply']['ms'],4))"; }
run AAADMM_KTINY=128
run AAADMM_KTINY=32
run AAADMM_KSMALL=96
run AAADMM_KSMALL=96 AAADMM_KTINY=128
run AAADMM_ND_LEAF=192 AAADMM_KSMALL=192 AAADMM_KTINY=128
run AAADMM_ND_LEAF=48 AAADMM_KSMALL=48
run AAADMM_KLONG=512
run AAADMM_KLONG=4096
