"""ncu_summary.py <raw.csv>  -- one block per distinct kernel of an `ncu -i X.ncu-rep --page raw --csv` dump:
launch count, duration, tensor-pipe activity, DRAM bytes and throughput, achieved occupancy, registers,
taken over the launch of each kernel name + grid that ran longest (the bench-shaped one)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}


def pick(*cands):
    for c in cands:
        if c in col:
            return col[c]
    return None


C = dict(
    name=pick('Kernel Name'), grid=pick('Grid Size'), block=pick('Block Size'),
    dur=pick('gpu__time_duration.sum'),
    tens=pick('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
    tens_e=pick('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
    rd=pick('dram__bytes_read.sum'), wr=pick('dram__bytes_write.sum'),
    dpct=pick('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
    occ=pick('sm__warps_active.avg.pct_of_peak_sustained_active'), regs=pick('launch__registers_per_thread'),
    smem=pick('launch__shared_mem_per_block_dynamic'), l2hit=pick('lts__t_sector_hit_rate.pct'),
    sm_thr=pick('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
)


def num(r, k):
    i = C[k]
    if i is None or r[i] == '':
        return None
    try:
        v = float(r[i].replace(',', ''))
    except ValueError:
        return None
    u = units[i]
    if k == 'dur':
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'second': 1e6, 's': 1e6}.get(u, 1)
    if k in ('rd', 'wr'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    return v


groups = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    nm = re.sub(r'\(.*', '', r[C['name']])
    groups.setdefault(nm, []).append(r)

print(f"# source: {sys.argv[1]}  (ncu --set full --clock-control none; durations are cold-cache, serialised)")
for nm, rs in groups.items():
    best = max(rs, key=lambda r: num(r, 'dur') or 0)
    d = num(best, 'dur')
    rd, wr = num(best, 'rd') or 0, num(best, 'wr') or 0
    tot = sum(num(r, 'dur') or 0 for r in rs)
    tens, tens_e = num(best, 'tens'), num(best, 'tens_e')
    print(f"{nm}")
    print(f"  launches {len(rs)}, total {tot:.1f} us; longest launch: {d:.1f} us, grid {best[C['grid']]} x block {best[C['block']]}, "
          f"regs {best[C['regs']] if C['regs'] is not None else '?'}, dyn smem {best[C['smem']] if C['smem'] is not None else '?'}")
    line = f"  DRAM read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / d / 1e3:.0f} GB/s" if d else ""
    if C['dpct'] is not None:
        line += f" ({best[C['dpct']]} % of DRAM peak)"
    if tens is not None:
        line += f"; tensor pipe active {tens:.1f} % of active cycles ({tens_e:.1f} % of elapsed)"
    if C['sm_thr'] is not None:
        line += f"; SM throughput {best[C['sm_thr']]} %"
    if C['occ'] is not None:
        line += f"; warps active {best[C['occ']]} %"
    if C['l2hit'] is not None:
        line += f"; L2 hit {best[C['l2hit']]} %"
    print(line)
