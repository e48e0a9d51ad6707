"""Developer tool: summarise an `ncu --csv` launch list (time, DRAM bytes, grid, occupancy per launch)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
L = OrderedDict()
for r in data:
    d = dict(zip(H, r)); L.setdefault(d['ID'], {'k': d['Kernel Name']})[d['Metric Name']] = (d['Metric Value'], d['Metric Unit'])
tot = 0; totb = 0
for i, (k, v) in enumerate(L.items()):
    t = float(v['gpu__time_duration.sum'][0].replace(',', '')); u = v['gpu__time_duration.sum'][1]
    t = t / 1000 if u == 'ns' else (t * 1000 if u == 'ms' else t)
    b = float(v['dram__bytes_read.sum'][0].replace(',', '')); bu = v['dram__bytes_read.sum'][1]
    b *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[bu]
    tot += t; totb += b
    name = v['k'].split('::')[-1][:24]
    print(i, name, 'grid', v['launch__grid_size'][0], 't_us %.1f' % t, 'MB %.1f' % b, 'GB/s %.0f' % (b / t * 1e3 if t else 0),
          'warps%', v.get('sm__warps_active.avg.pct_of_peak_sustained_active', ('?',))[0], 'regs', v.get('launch__registers_per_thread', ('?',))[0])
print('total us %.1f  MB %.1f' % (tot, totb))
