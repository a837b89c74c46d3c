#!/usr/bin/env python3
"""Turn the scratch ncu captures under gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_profiles.py r01     # reads gpurun_out/r01_launches.csv, gpurun_out/r01_pb_tpr4.ncu-rep
Writes profiles/<tag>_launches.csv (our kernels only, per launch), profiles/<tag>_launches_summary.md,
profiles/<tag>_pb_tpr4_ncu.md (key counters of the dominant kernel) and profiles/traffic.json
(dram bytes per read of the dominant kernel, used by bench.py's roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list ------------------------------------------------------------------------------------
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag))))
hdr, ours, agg, total_all = None, [], collections.OrderedDict(), 0.0
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    ns = float(d["Metric Value"].replace(",", ""))
    total_all += ns
    if "moira" in d["Kernel Name"]:
        ours.append((d["ID"], d["Kernel Name"], d["Grid Size"], d["Block Size"], ns))
        a = agg.setdefault(d["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += ns
with open(os.path.join(out, "%s_launches.csv" % tag), "w") as fh:
    fh.write("id,kernel,grid,block,gpu__time_duration_ns\n")
    for o in ours:
        fh.write('%s,"%s","%s","%s",%.0f\n' % o)
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, "%s_launches_summary.md" % tag), "w") as fh:
    fh.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu`\n\n" % tag)
    fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n")
    fh.write("Only this repo's kernels are listed; the torch kernels of the synthetic-data generator (%.1f ms) are not.\n\n" % ((total_all - tot) / 1e6))
    fh.write("| kernel | launches | total ms | share of our kernels | mean us |\n|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write("| `%s` | %d | %.3f | %.1f %% | %.1f |\n" % (k.replace("moira::<unnamed>::", ""), a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e3))
    fh.write("\nA decision-mode step is the `pb_cascade<2,4>` group: `tpr_kernel<2,0,1,1>` twice (pilot over the first 151 552 reads, then\n"
             "the rest), `policy_kernel`, `tpr_kernel<4,0,1,1>` (the candidate not chosen: returns at once) and `tpr_kernel<4,0,1,0>` over the\n"
             "escalated reads (empty on this workload); the `cascade = 2` steps of the bench (roofline.single_sweep) are one full\n"
             "`tpr_kernel<4,0,1,1>` launch each. Exact-ee steps add the classifier (`tpr_kernel<2,2,1,0>`) and the ladder rungs; `tpr_kernel<1,1,1,1>`\n"
             "is the Poisson / expected-error sweep (modes section). `fp64_peak_kernel` is the roofline probe, outside the timed region.\n")

# ---- full capture of the dominant kernel ---------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", "%s_pb_tpr4.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, vals = rr[0], rr[1], rr[2]
d = dict(zip(h, vals))
u = dict(zip(h, units))
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]
def num(x):
    return float(x.replace(",", ""))
def to_bytes(key):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]
    return num(d[key]) * scale
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
with open(os.path.join(out, "%s_pb_tpr4_ncu.md" % tag), "w") as fh:
    fh.write("# %s: `ncu --set full --clock-control none` of the dominant kernel `%s`\n\n" % (tag, d.get("Kernel Name", "")))
    fh.write("Workload: %d reads x 253 bp (bench.py C2), one launch.  Raw report kept in gpurun_out/ (scratch).\n\n" % reads)
    fh.write("| metric | value | unit |\n|---|---|---|\n")
    for k in want:
        if k in d:
            fh.write("| %s | %s | %s |\n" % (k, d[k], u[k]))
    fh.write("\nDRAM traffic per launch: %.4g bytes = %.1f bytes/read (algorithmic: 272 B/read -> %.4g bytes).\n" % (traffic, traffic / reads, 272.0 * reads))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(src.splitlines()))
    sh = sr[1]
    ix = {n: i for i, n in enumerate(sh)}
    cnt = collections.Counter()
    for r in sr[2:]:
        if len(r) < len(sh) or r[0] in ("Address", "Kernel Name"):
            continue
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        cnt[op] += int(r[ix["Instructions Executed"]])
    base_warps = reads * 253 / 32
    fh.write("\nExecuted warp instructions per base-warp (reads x 253 / 32 = %.4g base-warps):\n\n| opcode | per base |\n|---|---|\n" % base_warps)
    for op, c in cnt.most_common(16):
        fh.write("| %s | %.3f |\n" % (op, c / base_warps))
    fh.write("| **total** | %.3f |\n" % (sum(cnt.values()) / base_warps))
    fh.write("\nDFMA: %.4f per base (%.1f per read and thread) -- all of it the IEEE division of the interpolation step\n"
             "(`__ddiv_rn`, once per read, bernoullimodule.c:172); the PMF recurrence itself is DMUL/DADD only.\n"
             % (cnt.get("DFMA", 0) / base_warps, cnt.get("DFMA", 0) / base_warps * 253))
tj_path = os.path.join(out, "traffic.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
tj["pb_tpr<K=4>"] = {"dram_bytes_per_read": traffic / reads, "source": "profiles/%s_pb_tpr4_ncu.md" % tag}
json.dump(tj, open(tj_path, "w"), indent=1)
print(open(os.path.join(out, "%s_pb_tpr4_ncu.md" % tag)).read())
print(open(os.path.join(out, "%s_launches_summary.md" % tag)).read())

# ---- contig kernel (gpurun_out/<tag>_contig.ncu-rep, tools/bench_contigs.py --no-cpu --pairs 131072 --steps 1) ----
rep = os.path.join(ROOT, "gpurun_out", "%s_contig.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
    want = ["gpu__time_duration.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg.per_second",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    pairs, cells = 131072, 251 * 251
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "contig_kernel"
    md = ["# %s: `ncu --set full --clock-control none` of the contig kernel `%s`" % (tag, name), "",
          "Workload: %d synthetic MiSeq V4 pairs, 2 x 251 bp (`tools/bench_contigs.py --no-cpu --pairs %d --steps 1`), one launch." % (pairs, pairs),
          "Raw report kept in gpurun_out/ (scratch).", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr:
            md.append("| %s | %s | %s |" % (w, vals[hdr.index(w)], units[hdr.index(w)]))

    def in_bytes(k):
        u = units[hdr.index(k)]
        return g(k) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]

    dram = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
    alu = g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")
    t_ms = g("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
    inst = g("smsp__inst_executed.sum")
    md += ["", "Binding resource: the integer ALU pipe (16 lanes per SM sub-partition: one warp instruction every 2 cycles) -- "
           "`sm__pipe_alu_cycles_active` = %.1f %% of peak, i.e. roofline fraction %.2f against the ALU issue peak.  DRAM traffic "
           "(%.1f KB per pair, almost all of it the traceback-bit lines being written back from L2) runs at ~%.0f GB/s, far from the "
           "HBM bound." % (alu, alu / 100, dram / pairs / 1e3, dram / (t_ms * 1e-3) / 1e9),
           "", "Executed warp instructions per pair: %.0f for %d matrix cells, i.e. %.2f warp instructions per cell with everything "
           "included -- wavefront fill and drain, traceback, consensus.  The recurrence itself is ~9.7 instructions per cell: ISETP + SEL + "
           "IADD for the diagonal, VIMNMX + VIADDMNMX for max(up, left) + gap against it, two ISETP + two SEL/IADD3 for the traceback bits."
           % (inst / pairs, cells, inst / pairs / cells)]
    cj = os.path.join(out, "%s_contigs.json" % tag)
    if os.path.exists(cj):
        d = json.load(open(cj))
        md += ["", "Bench of the same build (profiles/%s_contigs.json): kernel %.3g pairs/s (%.0f G cells/s), end to end %.3g pairs/s; unmodified "
               "reference aligner %.0f pairs/s on one core, C restatement of the whole pair -> contig step %.3g pairs/s on %d cores." % (
                   tag, d["kernel"]["value"], d["kernel"]["cells_per_s"] / 1e9, d["e2e"]["value"],
                   d["cpu_baseline"]["reference_nw_align_1core"]["value"], d["cpu_baseline"]["oracle_pair_to_contig_allcores"]["value"],
                   d["cpu_baseline"]["oracle_pair_to_contig_allcores"]["cores"])]
    open(os.path.join(out, "%s_contig_ncu.md" % tag), "w").write("\n".join(md) + "\n")
    print("wrote profiles/%s_contig_ncu.md" % tag)
