"""Developer tool: critical path of one LDL^T apply from gpurun_out/ldlt_trace.csv.gz (AAADMM_LDLT_TRACE=1,
tests/tools/exp_trace.sh). A task that waits on a counter is released by the LAST task that signals that counter; walking
these releases back from the last task of a sweep gives the chain that bounds the sweep. Per link: `gap` = release ->
the waiting task sees it, `run` = the task's own time after that."""
import sys
import pandas as pd

d = pd.read_csv(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ldlt_trace.csv.gz")
for sw in ("fwd", "bwd"):
    s = d[d.sweep == sw]
    t0 = s.t_start.min()
    last_signal = s[s.signal_idx >= 0].sort_values("t_done").groupby("signal_idx").tail(1).set_index("signal_idx")
    cur = s.loc[s.t_done.idxmax()]
    chain = []
    while True:
        pred = last_signal.loc[cur.wait_idx] if (cur.need > 0 and cur.wait_idx in last_signal.index) else None
        chain.append((cur, pred))
        if pred is None:
            break
        cur = pred
    print(sw, "span %.1f us, %d links" % ((s.t_done.max() - t0) / 1e3, len(chain)))
    print(" level   first    ns     k  shape cw  start |  start   deps   done |  gap(pred done->deps)  run  wait_before_deps")
    tot_gap = tot_run = 0.0
    for cur, pred in reversed(chain):
        gap = (cur.t_deps - pred.t_done) / 1e3 if pred is not None else float("nan")
        run = (cur.t_done - cur.t_deps) / 1e3
        tot_run += run
        tot_gap += 0.0 if pred is None else gap
        print(" %5d %7d %5d %5d %5d %3d %6d | %6.1f %6.1f %6.1f | %6.1f %6.1f %6.1f" % (
            cur.level, cur["first"], cur.ns, cur.k, cur["shape"], cur.cw, cur.start, (cur.t_start - t0) / 1e3, (cur.t_deps - t0) / 1e3,
            (cur.t_done - t0) / 1e3, gap, run, (cur.t_deps - cur.t_start) / 1e3))
    print(" sum of runs %.1f us, sum of gaps %.1f us" % (tot_run, tot_gap))
