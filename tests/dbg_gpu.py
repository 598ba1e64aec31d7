import sys, numpy as np
sys.path.insert(0, '.')
from metamaps_b200 import capi, synth
from tests.golden.make_golden import small_workload
import tempfile
ctx = capi.Context(0)
db, fa, fq, names, reads = small_workload(tempfile.mkdtemp())
contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
rd = [synth.codes_to_ascii(r) for r in reads]
mode = sys.argv[1]
def build():
    ix = capi.Index(ctx, 16, 13); ix.add(contigs); ix.finalize(); return ix
try:
    if mode == 'A':      # index, destroy, index, map
        ix = build(); ix.close(); ix = build(); r = capi.map_reads(ctx, ix, rd); print('A ok', r['summary'])
    elif mode == 'B':    # index, fetch, then map on same index
        ix = build(); ix.fetch(); r = capi.map_reads(ctx, ix, rd); print('B ok', r['summary'])
    elif mode == 'C':    # index, lookup, then map
        ix = build(); ix.lookup(np.arange(10, dtype=np.uint32)); r = capi.map_reads(ctx, ix, rd); print('C ok', r['summary'])
    elif mode == 'D':    # two indexes alive
        ix = build(); ix2 = build(); r = capi.map_reads(ctx, ix2, rd); print('D ok', r['summary'])
    elif mode == 'E':    # stats only
        ix = build(); print(ix.stats()); r = capi.map_reads(ctx, ix, rd); print('E ok', r['summary'])
except Exception as e:
    print(mode, 'FAIL', e)
