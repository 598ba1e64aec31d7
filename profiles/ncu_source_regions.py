"""Aggregate an `ncu --page source --csv` export (SASS view): stall samples and executed instructions per block of 50 instructions,
the top stall reasons of the hottest blocks, and the hottest single instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data); totex = sum(int(r[iex] or 0) for r in data)
print(rows[0][1][:120]); print("instructions", len(data), "samples", tot, "warp instructions executed %.3f G" % (totex / 1e9))
for i in range(0, len(data), step):
    ch = data[i:i + step]
    s = sum(int(r[isamp] or 0) for r in ch); e = sum(int(r[iex] or 0) for r in ch)
    if s > tot * 0.02:
        d = {}
        for r in ch:
            for k in st:
                v = int(r[k] or 0)
                if v: d[hdr[k]] = d.get(hdr[k], 0) + v
        t = sum(d.values()) or 1
        top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / t) for k, v in sorted(d.items(), key=lambda x: -x[1])[:4])
        print("  instr %5d-%5d  samples %5.1f %%  executed %5.1f %%   %s" % (i, i + step - 1, 100 * s / tot, 100 * e / totex, top))
print("hottest instructions:")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:16]):
    print("  %5d  %-70s %6s" % (i, data[i][isrc][:70], data[i][isamp]))
