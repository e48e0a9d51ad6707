"""Developer tool: profiles/ldlt_dram_traffic.json from an ncu launch list of `bench.py` (tests/tools/profile_round2.sh):
mean dram__bytes_read.sum + dram__bytes_write.sum of k_fwd_front + k_bwd_front = the DRAM traffic of one LDL^T apply, which
bench.py reports as roofline.traffic (it is NOT re-measured in a bench run: a number taken under ncu is not a bench value,
and a bench run has no profiler)."""
import json
import sys
from summarize_launches import load

path, out = sys.argv[1], sys.argv[2]
L = load(path)
tot = {}
for e in L:
    for key in ("k_fwd_front", "k_bwd_front"):
        if key in e["k"]:
            t = tot.setdefault(key, [0, 0.0])
            t[0] += 1
            t[1] += e.get("rd", 0.0) + e.get("wr", 0.0)
per = {k: v[1] / v[0] for k, v in tot.items()}
json.dump({"bytes_per_apply": sum(per.values()), "per_kernel": per, "launches": {k: v[0] for k, v in tot.items()},
           "source": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu launch list %s" % path}, open(out, "w"), indent=1)
print(open(out).read())
