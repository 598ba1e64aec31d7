"""Opcode mix + hottest SASS per kernel from an `ncu --page source --csv` export."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
sections = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}; sections.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for sec in sections:
    hdr, data = sec['hdr'], sec['data']
    if not data: continue
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iW = hdr.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iW]) for r in data)
    print("==", sec['name'][:100]); print("total warp inst", tot, "sass lines", len(data), "stall samples", totS)
    op = collections.Counter(); ops = collections.Counter()
    for r in data:
        t = r[iS].split()
        o = t[1] if t[0].startswith('@') else t[0]
        o = o.split('.')[0]
        op[o] += int(r[iE]); ops[o] += int(r[iW])
    for k, v in op.most_common(18):
        print(f"  {k:10s} {100 * v / tot:5.1f}% inst   {100 * ops[k] / max(totS, 1):5.1f}% samples")
    mx = max(int(r[iE]) for r in data)
    print("  max exec per sass inst", mx, " sass insts with >=0.5*max:", sum(1 for r in data if int(r[iE]) >= 0.5 * mx), " >=0.9*max:", sum(1 for r in data if int(r[iE]) >= 0.9 * mx))
    if len(sys.argv) > 2:
        for r in sorted(data, key=lambda r: -int(r[iW]))[:int(sys.argv[2])]:
            print(f"  {int(r[iW]):7d} samples  {int(r[iE]):12d} exec  {r[iS].strip()[:90]}")
