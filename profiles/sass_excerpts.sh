#!/bin/bash
# SASS evidence for the two kernels VERDICT r1 asked about (run where the library is built; no GPU needed).
SO=metamaps_b200/libmetamaps_b200.so
cuobjdump -sass $SO > /tmp/mm_all.sass
fun() { awk -v f="Function : $1" 'index($0, f) {on=1; next} /Function : / {on=0} on' /tmp/mm_all.sass; }
{
echo "# cuobjdump -sass excerpts of $SO ($(date -u +%F)), sm_100a"
echo
echo "## l1_probe_filter_kernel<512> (K4 in one kernel): TMA bulk copy (cp.async.bulk -> UBLKCP) + mbarrier (SYNCS) staging of the key chunks; slot reads LDG.128, contig ids LDG.U16, survivors LDG.64"
fun '_ZN2mm22l1_probe_filter_kernelILi512EEEvPKNS_4SlotEjiPKjPKlPKiS9_PKtPKmNS_12HitKeyLayoutEijPyPmyPijP5uint2' | grep -E "UBLKCP|SYNCS|LDG|ATOMS|ATOMG|REDG" | sed 's/\s*\/\* 0x[0-9a-f]* \*\///' | head -48
echo
echo "## l2_classify_smem_kernel<true, 11> (K5a with the prune pass's counts): mnemonic counts; REDUX.SUM = the two warp-wide adds of packed 6-bit fields, STG.128 = the group record"
fun '_ZN2mm23l2_classify_smem_kernelILb1ELi11EEEvNS_12L2ClassifyFnEi' | grep -oE "REDUX[.A-Z0-9]*|LDG[.A-Z0-9]*|STG[.A-Z0-9]*|LDS[.A-Z0-9]*|STS[.A-Z0-9]*|BAR[.A-Z0-9]*" | sort | uniq -c
echo
echo "## l2_prune_warp_kernel (one warp per candidate): shuffles for the scans, REDUX for the bounds, shared-memory prefix pairs"
fun '_ZN2mm20l2_prune_warp_kernelENS_11L2PruneArgsEli' | grep -oE "REDUX[.A-Z0-9]*|SHFL[.A-Z0-9]*|LDG[.A-Z0-9]*|STG[.A-Z0-9]*|LDS[.A-Z0-9]*|STS[.A-Z0-9]*|MUFU[.A-Z0-9]*" | sort | uniq -c
echo
echo "## l2_sweep_band_kernel<256,4,1>: main loop (between the two warp votes): LDS/STS on the band state, LDS.64 from the event ring, LDGSTS (cp.async) refill"
fun '_ZN2mm20l2_sweep_band_kernelILi256ELi4ELi1EEEvNS_11L2SweepArgsEPKjlPKiS5_S5_PNS_8BandPartEPj' | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed 's/\s*\/\* 0x[0-9a-f]* \*\///' > /tmp/k5b_all.sass
python3 - <<'PY'
import re
L=open('/tmp/k5b_all.sass').read().splitlines()
heads=[i for i,l in enumerate(L) if re.search(r'VOTE\.ANY P\d, P\d', l)]
for h in heads:
    t=next((i for i in range(h,len(L)) if re.search(r'VOTE\.ANY R\d+, PT', L[i])), None)
    if t and t-h<400:
        print("(%d SASS instructions in the loop; %d in the kernel)"%(t-h,len(L)))
        print("\n".join(L[h-1:t+2]))
        break
PY
echo
echo "## em_round_kernel<32>: shuffles (SHFL.BFLY) for the per-read sums, shared-memory atomics (ATOMS) for the taxon sums, RED/ATOMG flush"
fun '_ZN2mm15em_round_kernelILi32EEEvPKiPKdPKllS4_PdiiPNS_7EmStateE' | grep -oE "SHFL\.[A-Z]+|ATOMS[.A-Z0-9]*|ATOMG[.A-Z0-9]*|RED[.A-Z0-9]*|DADD|DMUL|DFMA|MUFU[.A-Z0-9]*|LDG[.A-Z0-9]*|LDS[.A-Z0-9]*|STS[.A-Z0-9]*" | sort | uniq -c
echo
echo "## whole library: tensor-core / TMA mnemonics present"
cat /tmp/mm_all.sass | grep -oE "UBLKCP[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|SYNCS[.A-Z0-9]*|LDGSTS[.A-Z0-9]*|HMMA[.A-Z0-9]*|UTC[A-Z]*MMA[.A-Z0-9]*" | sort | uniq -c
} > profiles/r2_sass_excerpts.txt
