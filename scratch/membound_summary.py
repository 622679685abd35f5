"""Summarise an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,...` capture:
per kernel name -> launches, mean duration, DRAM bytes per launch, achieved GB/s and its fraction of the measured
HBM peak (MEASURED_PEAKS.json, 6554 GB/s).  `python scratch/membound_summary.py capture.csv`"""
import collections
import csv
import re
import sys

PEAK = 6554.2
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.OrderedDict()
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    key = (d['ID'], re.sub(r'\(.*', '', d['Kernel Name']).replace('void missm::', '').replace('missm::', '')[:60])
    u = d['Metric Unit']
    m = d['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)       # -> us
    elif m.startswith('dram__bytes'):
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    per.setdefault(key, {})[m] = v
agg = collections.OrderedDict()
for (_, name), m in per.items():
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, occ=0.0, pct=0.0))
    a['n'] += 1
    a['us'] += m.get('gpu__time_duration.sum', 0.0)
    a['rd'] += m.get('dram__bytes_read.sum', 0.0)
    a['wr'] += m.get('dram__bytes_write.sum', 0.0)
    a['occ'] += m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0.0)
    a['pct'] += m.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0.0)
print(f"{'kernel':60s} {'n':>4s} {'us/launch':>10s} {'MB read':>9s} {'MB written':>10s} {'GB/s':>8s} {'of 6554':>8s} {'dram pct ctr':>12s} {'warps act %':>11s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    n = a['n']
    us = a['us'] / n
    gbs = (a['rd'] + a['wr']) / n / (us * 1e-6) / 1e9 if us > 0 else 0.0
    print(f"{name:60s} {n:4d} {us:10.1f} {a['rd'] / n / 1e6:9.2f} {a['wr'] / n / 1e6:10.2f} {gbs:8.0f} {gbs / PEAK:8.2f} {a['pct'] / n:12.1f} {a['occ'] / n:11.1f}")
