"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, legacy mma.sync -> HMMA).
    cuobjdump -sass missm-benchmark_b200/lib/libmissm_b200.so | python scratch/sass_evidence.py"""
import collections
import re
import subprocess
import sys

cur = None
counts = collections.defaultdict(collections.Counter)
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|LDTM|STTM|SYNCS|HMMA|LDGSTS|ELECT|MUFU|UCGABAR_ARV|UGETNEXTWORKID|REDG|RED)\b')
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        m = pat.search(line)
        if m:
            counts[cur][m.group(1)] += 1
names = list(counts)
dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.strip().splitlines() if names else []
rows = []
for k, d in zip(names, dem):
    d = re.sub(r'\(.*', '', d).replace('missm::', '')
    rows.append((d, counts[k]))
keys = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'UGETNEXTWORKID', 'HMMA', 'LDGSTS', 'MUFU', 'RED']
print(f"{'kernel':84s} " + ' '.join(f"{k:>8s}" for k in keys))
for d, c in sorted(rows):
    c['RED'] = c.get('RED', 0) + c.get('REDG', 0)
    if any(c.get(k) for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'HMMA')):
        print(f"{d[:84]:84s} " + ' '.join(f"{c.get(k, 0):8d}" for k in keys))
