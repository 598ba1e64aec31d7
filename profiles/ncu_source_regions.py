"""Aggregate an `ncu --page source --csv` export (SASS view, one section per profiled kernel): stall samples and executed warp
instructions per block of N instructions with the top stall reasons of the hot blocks, and the hottest single instructions."""
import csv, sys


def section(name, hdr, data, step):
    data = [r for r in data if len(r) == len(hdr)]
    if not data:
        return
    isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data) or 1; totex = sum(int(r[iex] or 0) for r in data) or 1
    print("== " + name[:140]); print("   SASS instructions %d, stall samples %d, warp instructions executed %.3f G" % (len(data), tot, totex / 1e9))
    for i in range(0, len(data), step):
        ch = data[i:i + step]
        s = sum(int(r[isamp] or 0) for r in ch); e = sum(int(r[iex] or 0) for r in ch)
        if s > tot * 0.03:
            d = {}
            for r in ch:
                for k in st:
                    v = int(r[k] or 0)
                    if v:
                        d[hdr[k]] = d.get(hdr[k], 0) + v
            t = sum(d.values()) or 1
            top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / t) for k, v in sorted(d.items(), key=lambda x: -x[1])[:4])
            print("   instr %5d-%5d  samples %5.1f %%  executed %5.1f %%   %s" % (i, i + step - 1, 100 * s / tot, 100 * e / totex, top))
    print("   hottest instructions:")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:8]):
        print("     %5d  %-64s %7s samples" % (i, data[i][isrc][:64], data[i][isamp]))
    print()


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    seen = set()
    for a, b in zip(starts, starts[1:] + [len(rows)]):
        if rows[a][1] in seen:                           # the export lists every kernel once per source view
            continue
        seen.add(rows[a][1])
        section(rows[a][1], rows[a + 1], rows[a + 2:b], step)


if __name__ == "__main__":
    main()
