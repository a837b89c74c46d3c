#!/usr/bin/env python3
"""Round-2 ncu evidence -> profiles/: python tools/summarize_r02.py
Reads gpurun_out/r02_{k2,k4,k6}_{raw,source}.csv (exports of `ncu --set full --import-source on` captures made by
tools/gpu_profile_r02.sh), gpurun_out/r02_launches.csv (launch list of the bench command) and the ladder capture if present."""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def raw(tag):
    rows = list(csv.reader(open(os.path.join(G, tag + "_raw.csv"))))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def source(tag):
    rows = list(csv.reader(open(os.path.join(G, tag + "_source.csv"))))
    hdr = rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
    return out


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def kernel_summary(tag, title, what, reads, bases_per_read, traffic_key=None):
    d, u = raw(tag)
    src = source(tag)
    with open(os.path.join(P, tag + "_ncu.md"), "w") as fh:
        fh.write("# %s: `ncu --set full --clock-control none --import-source on` of `%s`\n\n%s\n\n" % (tag, d["Kernel Name"], what))
        fh.write("| metric | value | unit |\n|---|---|---|\n")
        for k in WANT:
            if k in d:
                fh.write("| %s | %s | %s |\n" % (k, d[k], u[k]))
        rd, wr = num(d["dram__bytes_read.sum"]), num(d["dram__bytes_write.sum"])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
        fh.write("\nDRAM traffic of this launch: %.4g bytes = %.1f bytes/read.\n\n" % (tot, tot / reads))
        wb = reads * bases_per_read / 32.0
        ops = collections.Counter()
        total = 0
        for s, n, _ in src:
            t = s.split()
            o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[o] += n
            total += n
        fh.write("Executed warp instructions per warp and base (%d reads x %d swept positions / 32 = %.4g):\n\n| opcode | per base |\n|---|---|\n" % (reads, bases_per_read, wb))
        for o, n in ops.most_common(12):
            fh.write("| %s | %.3f |\n" % (o, n / wb))
        fh.write("| **total** | %.3f |\n\n" % (total / wb))
        f64 = (ops["DMUL"] + ops["DADD"] + ops["DFMA"]) / wb
        fh.write("FP64 (DMUL + DADD + DFMA): %.2f per base; everything else: %.2f.  Issue model `2 F + O` = %.1f cycles per warp and base.\n"
                 % (f64, total / wb - f64, 2 * f64 + (total / wb - f64)))
        fh.write("DFMA appears only in the IEEE divisions of the per-read epilogue (interpolation, Newton bound): %.4f per base; the recurrence is DMUL / DADD only.\n" % (ops["DFMA"] / wb))
    if traffic_key:
        tj_path = os.path.join(P, "traffic.json")
        tj = json.load(open(tj_path))
        tj[traffic_key] = {"dram_bytes_per_read": tot / reads, "source": "profiles/%s_ncu.md" % tag}
        json.dump(tj, open(tj_path, "w"), indent=1)
    return d


def launches():
    rows = list(csv.reader(open(os.path.join(G, "r02_launches.csv"))))
    hdr, agg, total_all = None, collections.OrderedDict(), 0.0
    per = []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        ns = float(d["Metric Value"].replace(",", ""))
        total_all += ns
        name = d["Kernel Name"]
        if "moira" in name or "cub" in name.lower():
            per.append((d["ID"], name, d["Grid Size"], d["Block Size"], ns))
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ns
    with open(os.path.join(P, "r02_launches.csv"), "w") as fh:
        fh.write("id,kernel,grid,block,gpu__time_duration_ns\n")
        for o in per:
            fh.write('%s,"%s","%s","%s",%.0f\n' % o)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "r02_launches_summary.md"), "w") as fh:
        fh.write("# r02: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli`\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<this repo's kernels and CUB's>` (cold-cache, serialised:\n"
                 "compare shares).  C2, the real-profile workload, C3, C4 (+ collapse), C5 and their parity samples; the torch kernels of the\n"
                 "synthetic-data generators were not captured.  %d launches, %.1f ms.\n\n" % (len(per), tot / 1e6))
        fh.write("| kernel | launches | total ms | share | mean us |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f %% | %.1f |\n" % (k.replace("moira::<unnamed>::", "").replace("void ", "")[:110], a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e3))


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    kernel_summary("r02_k2", "two-entry sweep", "Main launch of `pb_cascade<2,4>` (tools/one_step.py: C2, 10 000 000 reads x 253 bp, decision mode, row marks given; "
                   "this launch sweeps the 9 848 448 reads behind the pilot's 151 552).", 9848448, 256, "pb_cascade<2,4>")
    kernel_summary("r02_k4", "four-entry sweep", "The single K = 4 sweep (`cascade = 2`), C2, 10 000 000 reads x 253 bp, row marks given.", 10000000, 256, "pb_tpr<K=4>")
    kernel_summary("r02_k6", "six-entry sweep", "The single K = 6 sweep (`cascade = 2`) over C3-shaped reads (10 000 000 x 420..480 bp, pitch 480), row marks given.",
                   10000000, 456, None)
    if os.path.exists(os.path.join(G, "r02_launches.csv")):
        launches()
    print("profiles/ updated")
