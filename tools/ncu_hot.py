"""Summarise an ncu report's SASS source page: instruction mix and the top stall sites (development tool).
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [warps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
H = rows[1]
ai, ei, si = H.index('Source'), H.index('Instructions Executed'), H.index('# Samples')
cols = {k: H.index(k) for k in ('stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_math', 'stall_no_inst',
                                'stall_barrier', 'stall_not_selected', 'stall_branch_resolving', 'stall_selected')}
seq = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) < len(H) or r[0] == 'Address':
        continue
    seq.append(r)
tot = sum(int(r[ei]) for r in seq)
samples = sum(int(r[si]) for r in seq)
print(f"instructions/warp {tot / W:.1f}   samples {samples}")
agg = collections.Counter()
for r in seq:
    for k, c in cols.items():
        agg[k] += int(r[c])
print({k: v for k, v in agg.most_common()})
for key in ('stall_long_sb', 'stall_short_sb', 'stall_no_inst', 'stall_barrier', 'stall_wait'):
    print('---- top sites for', key)
    top = sorted(range(len(seq)), key=lambda j: -int(seq[j][cols[key]]))[:12]
    for j in sorted(top):
        r = seq[j]
        print(f"  #{j:5d} {r[ai].strip()[:70]:70s} exec/warp {int(r[ei]) / W:6.2f} {key}={r[cols[key]]}")
