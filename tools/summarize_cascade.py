#!/usr/bin/env python3
"""profiles/ summaries of the cascaded first pass from the scratch captures in gpurun_out/:
  gpurun_out/r01_cascade_k2.ncu-rep     ncu --set full of the two-entry main launch (tools/one_step.py, 2nd step)
  gpurun_out/r01_cascade_launches.csv   launch list of the same program (our first-pass kernels only)
Writes profiles/r01_pb_cascade_k2_ncu.md, profiles/r01_cascade_launches.md and adds the group to profiles/traffic.json."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reads, pilot = 10_000_000, 2 * 148 * 16 * 32
out = os.path.join(ROOT, "profiles")
rep = os.path.join(ROOT, "gpurun_out", "r01_cascade_k2.ncu-rep")
rr = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, units, vals = rr[0], rr[1], rr[2]
d, u = dict(zip(h, vals)), dict(zip(h, units))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tb = lambda k: float(d[k].replace(",", "")) * scale[u[k]]
traffic = tb("dram__bytes_read.sum") + tb("dram__bytes_write.sum")
n_main = reads - pilot
sr = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
sh = sr[1]
ix = {n: i for i, n in enumerate(sh)}
cnt, bucket = collections.Counter(), collections.Counter()
for r in sr[2:]:
    if len(r) < len(sh):
        continue
    t = [x for x in r[ix["Source"]].split() if not x.startswith("@")]
    c = int(r[ix["Instructions Executed"]])
    cnt[t[0].split(".")[0]] += c
    bucket["per 16-base vector" if c >= 4_000_000 else "per 128-byte chunk" if c >= 500_000 else "per 32-read tile (epilogue, tail vector)" if c >= 250_000 else "other"] += c
bw = n_main * 253 / 32
with open(os.path.join(out, "r01_pb_cascade_k2_ncu.md"), "w") as fh:
    fh.write("# r01: `ncu --set full --clock-control none` of the two-entry sweep `%s`\n\n" % d.get("Kernel Name", "tpr_kernel<2,0,1,1>"))
    fh.write("The main launch of the `pb_cascade<2,4>` group (tools/one_step.py: config C2, 10 000 000 reads x 253 bp, decision mode; this launch sweeps\n"
             "the %d reads behind the pilot's %d).  Raw report kept in gpurun_out/ (scratch).\n\n| metric | value | unit |\n|---|---|---|\n" % (n_main, pilot))
    for k in want:
        if k in d:
            fh.write("| %s | %s | %s |\n" % (k, d[k], u[k]))
    fh.write("\nDRAM traffic of this launch: %.4g bytes = %.1f bytes/read (algorithmic: 272 B/read).\n" % (traffic, traffic / n_main))
    fh.write("\nExecuted warp instructions per base-warp (%d reads x 253 / 32 = %.4g base-warps):\n\n| opcode | per base |\n|---|---|\n" % (n_main, bw))
    for op, c in cnt.most_common(14):
        fh.write("| %s | %.3f |\n" % (op, c / bw))
    fh.write("| **total** | %.3f |\n\n| where | per base |\n|---|---|\n" % (sum(cnt.values()) / bw))
    for k, c in bucket.most_common():
        fh.write("| %s | %.3f |\n" % (k, c / bw))
    fp = (cnt["DMUL"] + cnt["DADD"] + cnt["DFMA"]) / bw
    fh.write("\nFP64 instructions: %.2f per base (3 DMUL + DSUB + DADD of the two-entry recurrence; the rest is the interpolation division\n"
             "and the Newton bound, once per read) against 10.2 for the K = 4 sweep (profiles/r01_pb_tpr4_ncu.md).  The pipe is busy %s %% of the\n"
             "time and the schedulers issue on %s %% of their cycles: with 5.4 of %.1f issue slots per base going to FP64 the kernel sits between the\n"
             "two limits; the remaining 6-7 slots per base are the lookup (PRMT + LDS.64), the N/n accounting (taken for 94 %% of the vectors\n"
             "because a warp holds a few noisy reads almost always), loop control and the per-tile epilogue.\n"
             % (fp, d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"], sum(cnt.values()) / bw))
tj = json.load(open(os.path.join(out, "traffic.json")))
tj["pb_cascade<2,4>"] = {"dram_bytes_per_read": traffic / n_main, "source": "profiles/r01_pb_cascade_k2_ncu.md"}
json.dump(tj, open(os.path.join(out, "traffic.json"), "w"), indent=1)

rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "r01_cascade_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
step = rows[-5:]
names = ["pilot: two-entry sweep of the first %d reads" % pilot, "verdict (`policy_kernel`): escalated fraction of the pilot <= 0.35 -> cascade",
         "two-entry sweep of the other reads (runs: the verdict chose it)", "four-entry sweep of the other reads (returns at once: not chosen)",
         "four-entry sweep over the escalated reads (queue 0; empty on this workload)"]
with open(os.path.join(out, "r01_cascade_launches.md"), "w") as fh:
    fh.write("# r01: launches of one decision-mode step over config C2 (`pb_cascade<2,4>`)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none "
             "-k regex:\"tpr_kernel|policy\" python tools/one_step.py` (second step; cold-cache and serialised: compare shares).\n\n"
             "| launch | kernel | grid x block | us |\n|---|---|---|---|\n")
    tot = sum(float(r[-1]) for r in step)
    for r, nm in zip(step, names):
        fh.write("| %s | `%s` | %s x %s | %.1f |\n" % (nm, r[4].replace("void unnamed>::", "").split("(")[0], r[8], r[7], float(r[-1]) / 1e3))
    fh.write("| **step** | | | %.1f |\n\nThe two-entry main launch is %.1f %% of the step.\n" % (tot / 1e3, 100 * float(step[2][-1]) / tot))
print(open(os.path.join(out, "r01_pb_cascade_k2_ncu.md")).read())
print(open(os.path.join(out, "r01_cascade_launches.md")).read())
