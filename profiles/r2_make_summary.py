#!/usr/bin/env python
"""Builds profiles/r2_summary.md, r2_traffic.json and the committed copies of the launch list / raw ncu pages from the files the
round-2 GPU calls left in gpurun_out/ (scripts profiles/r2_run_*.sh).  Run in the build container after the last call."""
import collections
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = "r2z"          # config 4 leg (call Z)
RTAG = "r2ba"        # ncu --set full raw / source pages of the committed tree (call BA)
BTAG = "r2av"        # bench line + launch list of the committed tree (call AV; reference arm: call AQ)


def last_json(fn):
    try:
        return json.loads(open(os.path.join(G, fn)).read().strip().splitlines()[-1])
    except Exception:
        return None


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        short = r[4].replace("void ", "").replace("mm::", "").split("(")[0][:72]
        agg.setdefault(short, [0, 0.0]); agg[short][0] += 1; agg[short][1] += float(r[14]) / 1e6
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | ms (ncu: cold cache, serialised) | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")
    out.append(f"| total ({len(rows)} launches in one step) | | {tot:.2f} | |")
    return "\n".join(out), agg, tot


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    return rows[2:], idx, units


def metric_block(path, names):
    data, idx, units = raw_rows(path)
    keys = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
            ("smsp__issue_active.avg.per_cycle_active", "issue slots busy"), ("smsp__inst_executed.sum", "warp instructions"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"), ("launch__registers_per_thread", "registers"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts")]
    st = [h for h in idx if "issue_stalled" in h and "per_issue_active" in h]
    out = []; seen = set(); traffic = {}
    for r in data:
        n = r[idx["Kernel Name"]]
        for pat, label in names.items():
            if pat in n and label not in seen:
                seen.add(label)
                out.append(f"**{label}** (`{n[:70]}`)")
                out.append("")
                out.append("| metric | value |"); out.append("|---|---|")
                for k, lab in keys:
                    if k in idx:
                        out.append(f"| {lab} | {r[idx[k]]} {units[idx[k]]} |")
                s = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in st), reverse=True)[:6]
                out.append("| top stalls (warps per issue) | " + ", ".join(f"{nm} {v:.2f}" for v, nm in s) + " |")
                out.append("")

                def gb(k):
                    return float(r[idx[k]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx[k]]]
                rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
                traffic[label] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "gpu_time_ms": float(r[idx["gpu__time_duration.sum"]])}
    return "\n".join(out), traffic


def main():
    md = ["# Round 2 — profiles and measurements (1 x B200 unless stated; scripts `profiles/r2_run_*.sh`)", ""]
    b = last_json(f"{BTAG}_bench.json")
    if b:
        md += ["## The bench line of the final tree (`bench.py --gpus 1 --steps 20 --warmup 5`, call AV)", "",
               f"`value` {b['value']:.0f} Mbp/s ({b['ms_per_step']:.2f} ms/step), `e2e` {b['e2e']['value']:.0f} Mbp/s, {b['gpu_launches']} launches in the timed region, "
               f"SM clock {b['clocks']['sm_mhz']:.0f} / {b['clocks']['sm_max_mhz']} MHz, throttle reasons {b['clocks']['reasons']}.", "",
               "| kernel / stage | ms | algorithmic GB (SURVEY 8d) | GB/s | fraction of 6549.8 GB/s | other roof |", "|---|---|---|---|---|---|"]
        for k, v in b["roofline"]["kernels"].items():
            md.append(f"| {k} | {v['ms']:.2f} | {v['bytes'] / 1e9:.2f} | {v['GBps']:.0f} | {v['frac']:.3f} | " + (f"IMAD pipe: {v['imad_frac']:.2f} of its roof ({v['imad_roof_ms']:.2f} ms)" if "imad_frac" in v else "") + " |")
        md += ["", "Extra legs of the same run: " + json.dumps({k: ({kk: vv for kk, vv in v.items() if kk in ("value", "seconds", "em_ms_per_round", "em_GBps", "em_frac_of_hbm_peak", "ratio", "gpu_cli_s", "ref_s", "files_identical", "GBps_per_rank_all_ranks_uploading", "mappings", "rounds_to_reference_stopping_rule")} if isinstance(v, dict) else v) for k, v in b.get("extra", {}).items()}), ""]
    lp = os.path.join(G, f"{BTAG}_launches.csv")
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(P, "r2_launches_config2.csv"))
        t, agg, tot = launch_table(lp)
        md += ["## Launch list of one config-2 step (`ncu --metrics gpu__time_duration.sum --clock-control none`, `profiles/r2_launches_config2.csv`)", "", t, ""]
        if b:
            dom = b["roofline"]["kernel"]; pat = dom.split()[1].replace("_kernel", "")
            kd = next((v[1] for k, v in agg.items() if pat in k), 0)
            md += [f"Share check: the dominant kernel of the bench line ({dom}) is {100 * kd / tot:.1f} % of the launch list and {100 * b['roofline']['kernels'][dom]['ms'] / b['detail']['kernel_ms_per_step']:.1f} % of the "
                   f"CUDA-event kernel time of a bench step ({b['detail']['kernel_ms_per_step']:.1f} ms).", ""]
    rp = os.path.join(G, f"{RTAG}_full_raw.csv")
    traffic = {"workload": "config2", "source": f"profiles/r2_full_raw.csv: ncu --set full --clock-control none, one step of bench.py --profile-step (round 2, call BA)"}
    if os.path.exists(rp):
        shutil.copy(rp, os.path.join(P, "r2_full_raw.csv"))
        t, tr = metric_block(rp, {"sketch_blockmin": "K1 sketch_blockmin_kernel", "read_sketch_block_kernel<(int)4": "K3 read_sketch_block_kernel<4>",
                                  "l1_probe_filter": "K4 l1_probe_filter_kernel", "l2_classify_smem": "K5a l2_classify_smem_kernel",
                                  "l1_sort_segments_warp": "K4 l1_sort_segments_warp_kernel", "l1_candidates_warp": "K4 l1_candidates_warp_kernel",
                                  "l2_prune_warp": "K5p l2_prune_warp_kernel", "l2_sweep_band": "K5b l2_sweep_band_kernel"})
        traffic.update(tr)
        md += ["## `ncu --set full` of the dominant kernels (`profiles/r2_full_raw.csv` = the raw page)", "", t]
        json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
    ep = os.path.join(G, f"{TAG}_em_raw.csv")
    if os.path.exists(ep):
        shutil.copy(ep, os.path.join(P, "r2_em_raw.csv"))
        t, _ = metric_block(ep, {"em_round_kernel": "K7 em_round_kernel (config 4 shape, 40 k reads x 200 mappings, T = 200)"})
        md += ["## The EM round kernel on config-4-shaped data (`profiles/r2_em_raw.csv`)", "", t]
    # A/B table
    ab = [("K5b ring 8, 8 warps / SM", "r2f_ring8_warps8.json"), ("K5b ring 8, 12 warps (default)", "r2f_ring8_warps12.json"), ("K5b ring 4, 12 warps", "r2f_ring4_warps12.json"),
          ("K5b ring 4, 16 warps", "r2f_ring4_warps16.json"), ("K5b ring 2, 12 warps", "r2f_ring2_warps12.json"), ("K5b ring 2, 16 warps", "r2f_ring2_warps16.json"),
          ("band 128, ring 4, 16 warps (call A build)", "r2a_bench_b128_r4_c0.json"), ("band 128, ring 4, 2 CTAs x 12 warps, 80 registers (call A build)", "r2a_bench_b128_r4_c1.json"),
          ("window skipping ON (call G)", "r2g_bench.json"), ("window skipping off, same build (call G)", "r2g_bench_noskip.json"),
          ("hash table load 0.25 (call B)", "r2b_bench.json"), ("hash table load 0.5, same build (call B)", "r2b_bench_mult2.json"),
          ("L1 filter: one list per 8 lanes (call C)", "r2c_bench_n1.json"), ("L1 filter: flattened walk, same build (call C)", "r2c_bench_flat.json"),
          ("K4 as two kernels (call L)", "r2l_bench_twokernels.json"), ("K4 fused, same build (call L)", "r2l_bench.json"),
          ("K3 radix 4 / 5 / 6 bits per pass (call M)", "r2m_k3radix4.json"), ("", "r2m_k3radix5.json"), ("", "r2m_k3radix6.json"),
          ("window pruning, ladder of 16 ranks + separate kernel, 4096 starts / segment (call N)", "r2n_bench.json"), ("no pruning, same build (call N)", "r2n_bench_noprune.json"),
          ("pruned, 256 window starts / segment (call P)", "r2p_seg256.json"), ("pruned, 512 (call P)", "r2p_seg512.json"), ("pruned, 1024 (call P)", "r2p_seg1024.json"), ("pruned, 2048 (call P)", "r2p_seg2048.json"),
          ("prune decision inside K5a, ballots (call Q)", "r2q_bench.json"), ("the same, REDUX counts + packed prefix pairs (call R)", "r2r2_bench.json"),
          ("prune decision in its own light kernel (call T)", "r2t_bench.json"), ("the same with DMA staging (e2e) (call T)", "r2t_bench_dma.json"),
          ("cooperative first-window build in K5b (call U)", "r2u_bench.json"),
          ("K5b ring 4 x 12 warps (call V)", "r2v_ring4_w12.json"), ("K5b ring 4 x 16 warps (call V)", "r2v_ring4_w16.json"), ("K5b ring 2 x 16 warps (call V)", "r2v_ring2_w16.json"),
          ("K5a 4096 buckets (call W)", "r2w_bits12.json"), ("K5a 8192 buckets (call W)", "r2w_bits13.json"), ("K5a buckets on 1 - (1 - x)^32 (call X)", "r2x_bench.json"),
          ("32-byte slots with inline contig ids (call AH)", "r2ah_bench.json"), ("the same, inline ids ignored (call AI)", "r2ai_inline0.json"),
          ("prune pass up to 512 groups per candidate (call AK)", "r2ak_g512_1.json"), ("up to 256, same build (call AK)", "r2ak_g256_1.json")]
    md += ["## A/B measurements (config 2, CUDA events inside `bench.py`)", "", "| variant | ms/step | K5a ms | K5b ms | L1 stage ms | `value` Mbp/s | `e2e` Mbp/s |", "|---|---|---|---|---|---|---|"]
    for name, fn in ab:
        d = last_json(fn)
        if d:
            st = d["detail"]["stage_ms"]
            md.append(f"| {name or fn} | {d['ms_per_step']:.2f} | {st['l2_classify_ms']:.2f} | {d['roofline']['kernels']['K5b l2_sweep_band_kernel']['ms']:.2f} | {st['l1_probe_ms']:.2f} | {d['value']:.0f} | {d['e2e']['value']:.0f} |")
    md.append("")
    # multi-GPU
    md += ["## Multi-GPU (final tree: calls Z, AC, AF; calls D / E: the build of call H, before the window pruning)", "", "| N | `value` Mbp/s | ms/step | `e2e` | config 3 (1 M reads) | contig shards `value` / ms | contig shards config 3 | config 5 slice ms/step |", "|---|---|---|---|---|---|---|---|"]
    for n, fn, c5 in ((1, f"{BTAG}_bench.json", "r2ae_c5_1.json"), ("2 (final tree, call AR)", "r2ar_bench_n2.json", None), ("2 (call AC, before the per-read survivor sort)", "r2ac_bench_n2.json", "r2ac_config5_n2.json"),
                      ("8 (call AF, before the per-read survivor sort, no extra legs)", "r2af_bench_n8.json", None), ("8, zero-copy staging (call AF)", "r2af_bench_n8_zerocopy.json", None),
                      ("2 (call D)", "r2d2_bench_n2.json", "r2c_config5_n2.json"), ("8 (call E)", "r2e_bench_n8.json", "r2e_config5_n8.json")):
        d = last_json(fn); c = last_json(c5) if c5 else None
        if d:
            ex = d.get("extra", {})
            sc = ex.get("shard_contigs", {}); c3 = ex.get("config3", {}); c3s = ex.get("config3_shard_contigs", {})
            md.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['e2e']['value']:.0f} | " + (f"{c3['value']:.0f} ({c3['seconds']:.3f} s)" if "value" in c3 else "—") + " | "
                      + (f"{sc['value']:.0f} / {sc['ms_per_step']:.1f}" if "value" in sc else "—") + " | " + (f"{c3s['value']:.0f}" if "value" in c3s else "—") + " | "
                      + (f"{c['ms_per_step']:.0f}" + (f" (check: {c['detail']['check']})" if c['detail'].get('check') else "") if c else "—") + " |")
    md.append("")
    for c4fn, c4title in (("r2b_config4.json", "## Config 4 at full size (`bench.py --workload config4`, call B, before the window pruning)"),
                          (f"{TAG}_config4_100k.json", "## Config 4 on 100 k reads, final tree (`bench.py --workload config4 --reads 100000`, call Z)")):
        c4 = last_json(c4fn)
        if not c4:
            continue
        dd = c4["detail"]
        md += [c4title, "",
               f"{dd['reads']} reads, {dd['mappings']} mappings ({dd['mappings_per_read']:.1f} per read), T = {dd['strains']}; EM {dd['em_rounds']} rounds = {dd['em_ms_total']:.1f} ms = "
               f"{dd['em_ms_per_round']:.3f} ms/round = {dd['em_GBps']:.0f} GB/s = {dd['em_frac_of_hbm_peak']:.3f} of the HBM peak; the reference's stopping rule needs "
               f"{dd['rounds_to_reference_stopping_rule']} rounds; mapping {dd['map_s']:.1f} s; whole workload {c4['value']:.0f} Mbp/s.", ""]
    open(os.path.join(P, "r2_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md)[:3000])


if __name__ == "__main__":
    main()
