#!/usr/bin/env python3
"""profiles/r01_fastq_launches.md from gpurun_out/r01_fq_launches.csv (ncu launch list of tools/fq_step.py)."""
import collections, csv, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "r01_fq_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].replace("void unnamed>::", "").replace("unnamed>::", "").split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1])
tot = sum(a[1] for a in agg.values())
text_bytes, calls, chunks = 500_000 * 521, 2, None
with open(os.path.join(ROOT, "profiles", "r01_fastq_launches.md"), "w") as fh:
    fh.write("# r01: kernels of `moira_filter_fastq` (FASTQ text parsed on the device)\n\n"
             "`ncu --metrics gpu__time_duration.sum --clock-control none python tools/fq_step.py`: two calls over 500 000 synthetic 253-bp\n"
             "records = %.0f MB of text in pinned memory, four chunks of <= 64 MB per call (cold-cache, serialised: compare shares).\n\n"
             "| kernel | launches | total us | share | mean us |\n|---|---|---|---|---|\n" % (text_bytes / 1e6))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write("| `%s` | %d | %.1f | %.1f %% | %.1f |\n" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0] / 1e3))
    per_call = tot / calls / 1e3
    fh.write("| **all kernels, per call** | | %.1f | | |\n\n" % per_call)
    fh.write("One call moves %.0f MB over PCIe: %.1f ms at the measured 54 GB/s.  All kernels of the call together take %.2f ms (%.0f %% of the\n"
             "copy time) and run under the next chunk's copy, so the path is bound by the link (bench.py `e2e_parse`: 53.5 GB/s of text).\n"
             % (text_bytes / 1e6, text_bytes / 54e9 * 1e3, per_call / 1e3, 100 * per_call / 1e3 / (text_bytes / 54e9 * 1e3)))
print(open(os.path.join(ROOT, "profiles", "r01_fastq_launches.md")).read())
