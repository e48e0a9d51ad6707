"""Developer tool: per-kernel summary (launches, mean time, share of the listed launches, mean DRAM bytes) of an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list, as markdown."""
import csv
import sys
from collections import OrderedDict

UNIT_T = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    L = OrderedDict()
    for r in rows[hdr + 1:]:
        d = dict(zip(H, r))
        e = L.setdefault(d["ID"], {"k": d["Kernel Name"], "grid": d.get("Grid Size", ""), "block": d.get("Block Size", "")})
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v * UNIT_T[d["Metric Unit"]]
        elif d["Metric Name"] == "dram__bytes_read.sum":
            e["rd"] = v * UNIT_B[d["Metric Unit"]]
        elif d["Metric Name"] == "dram__bytes_write.sum":
            e["wr"] = v * UNIT_B[d["Metric Unit"]]
    return list(L.values())


def short(name):
    name = name.split("(")[0]
    return name.replace("aaadmm::", "").replace("(anonymous namespace)::", "")


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    L = load(path)
    tot = sum(e.get("us", 0.0) for e in L)
    K = OrderedDict()
    for e in L:
        k = K.setdefault(short(e["k"]), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "grid": e["grid"], "block": e["block"]})
        k["n"] += 1
        k["us"] += e.get("us", 0.0)
        k["rd"] += e.get("rd", 0.0)
        k["wr"] += e.get("wr", 0.0)
    print("### %s" % title)
    print()
    print("%d launches, %.1f us in all (device time under ncu: cold-cache and serialised - compare the shares)" % (len(L), tot))
    print()
    print("| kernel | launches | mean us | share % | mean DRAM read MB | mean DRAM write MB | DRAM GB/s | grid x block |")
    print("|---|---:|---:|---:|---:|---:|---:|---|")
    for name, k in sorted(K.items(), key=lambda kv: -kv[1]["us"]):
        mu = k["us"] / k["n"]
        gbs = (k["rd"] + k["wr"]) / k["n"] / (mu * 1e-6) / 1e9 if mu > 0 else 0.0
        print("| %s | %d | %.1f | %.1f | %.2f | %.2f | %.0f | %s x %s |" % (name, k["n"], mu, 100.0 * k["us"] / tot, k["rd"] / k["n"] / 1e6,
                                                                         k["wr"] / k["n"] / 1e6, gbs, k["grid"], k["block"]))
    print()


if __name__ == "__main__":
    main()
