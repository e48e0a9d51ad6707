"""Developer tool: print the per-level table of gpurun_out/front_launches.csv (see level_profile.sh)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/front_launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; L = OrderedDict()
for r in rows[hdr + 1:]:
    d = dict(zip(H, r)); L.setdefault(d['ID'], {'k': d['Kernel Name']})[d['Metric Name']] = (d['Metric Value'], d['Metric Unit'])
tot = {'fwd': 0, 'bwd': 0}
for i, (k, v) in enumerate(L.items()):
    t = float(v['gpu__time_duration.sum'][0].replace(',', '')); u = v['gpu__time_duration.sum'][1]; t = t / 1000 if u == 'ns' else t
    b = float(v['dram__bytes_read.sum'][0].replace(',', '')); bu = v['dram__bytes_read.sum'][1]; b *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[bu]
    ins = float(v['smsp__inst_executed.sum'][0].replace(',', ''))
    kind = 'fwd' if 'Gather' in v['k'] or 'fwd' in v['k'] else 'bwd'
    tot[kind] += t
    print(i, kind, 'grid', v['launch__grid_size'][0], 't %.1f' % t, 'MB %.1f' % b, 'GB/s %.0f' % (b / t * 1e3), 'inst/16B %.1f' % (ins * 32 / (b * 1e6 / 16)),
          'issue%', v['smsp__issue_active.avg.pct_of_peak_sustained_active'][0], 'warps%', v['sm__warps_active.avg.pct_of_peak_sustained_active'][0], 'regs', v['launch__registers_per_thread'][0])
print(tot)
