"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid) count / mean / total
for the LAST training step in the file (from the last launch of the marker kernel on; default: k_prepare_fwd, the one-kernel table build, else k_plu_mask)."""
import csv, collections, sys
path = sys.argv[1]
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
marker = sys.argv[2] if len(sys.argv) > 2 else None   # first kernel of a step
idx = []
for mk in ([marker] if marker else ['k_prepare_fwd', 'k_plu_mask']):
    idx = [i for i, r in enumerate(rows) if mk in r['Kernel Name']]
    if idx:
        break
last = rows[idx[-1]:] if idx else rows
agg = collections.OrderedDict()
for r in last:
    n = r['Kernel Name'].split('(')[0].replace('nf::<unnamed>::', '').replace('void ', '')[:64]
    k = (n, r['Grid Size'])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value']) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches in the last step, {tot:.1f} us summed (ncu: cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:3d} x {v[1]/v[0]:7.2f} us = {v[1]:8.1f} us  {100*v[1]/tot:5.1f}%  {k[0]} grid={k[1]}")
