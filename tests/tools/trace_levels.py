"""Developer tool: per-level summary of gpurun_out/ldlt_trace.csv.gz (AAADMM_LDLT_TRACE=1 solve_only.py)."""
import sys
import pandas as pd
d = pd.read_csv(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/ldlt_trace.csv.gz')
t0 = d.t_start.min()
for c in ['t_start', 't_primed', 't_deps', 't_done']:
    d[c] = (d[c] - t0) / 1000.0
d['kind'] = 'task'
d.loc[(d.sweep == 'fwd') & (d['shape'] == 0), 'kind'] = 'asm'
d.loc[(d.sweep == 'bwd') & (d.cw == 0), 'kind'] = 'asm'
print('total span us', d.t_done.max())
for sw in ['fwd', 'bwd']:
    x = d[d.sweep == sw]
    print(sw, 'span', x.t_start.min(), x.t_done.max(), 'tasks', len(x))
    x = x.assign(wait=x.t_deps - x.t_primed, work=x.t_done - x.t_deps)
    g = x.groupby(['level', 'kind']).agg(n=('task', 'size'), first_start=('t_start', 'min'), deps_first=('t_deps', 'min'), deps_last=('t_deps', 'max'),
                                         done_first=('t_done', 'min'), done_last=('t_done', 'max'), wait_avg=('wait', 'mean'), work_avg=('work', 'mean'), work_max=('work', 'max'))
    print(g.round(1).to_string())
