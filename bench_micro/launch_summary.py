"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, count and median us."""
import collections, csv, statistics, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    d = dict(zip(H, r))
    v = float(d['Metric Value'].replace(',', ''))
    if d['Metric Unit'] in ('nsecond', 'ns'): v /= 1e3
    elif d['Metric Unit'] in ('msecond', 'ms'): v *= 1e3
    agg.setdefault(d['Kernel Name'][:70] + ' grid=' + d['Grid Size'], []).append(v)
for k, v in agg.items():
    print(f"{k:100s} n={len(v):4d} median={statistics.median(v):9.2f} us  min={min(v):9.2f}")
