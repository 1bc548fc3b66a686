#!/usr/bin/env python
"""Summarise ncu captures into the small text files kept under profiles/.

  python profiles/summarize.py full     <report.ncu-rep> <out.md>   # one `--set full` capture
  python profiles/summarize.py fullcsv  <raw.csv>        <out.md>   # the same from `ncu -i rep --page raw --csv`
                                                                    # (what travels back from the GPU box: < 64 MB)
  python profiles/summarize.py list     <launches.csv>   <out.md>   # a gpu__time_duration launch list

The .ncu-rep files themselves stay in gpurun_out/ (scratch, not tracked).
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def full(rep, out, from_csv=False):
    if from_csv:
        txt = open(rep, errors="replace").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu --set full summary of %s" % rep.split("/")[-1], ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("## %s  (id %s)" % (d.get("Kernel Name", "?")[:100], d.get("ID", "?")))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", ""))
            wr = float(d["dram__bytes_write.sum"].replace(",", ""))
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = units[hdr.index("gpu__time_duration.sum")]
            bu = units[hdr.index("dram__bytes_read.sum")]
            lines.append("| dram traffic (read+write) | %.6g | %s |" % (rd + wr, bu))
            lines.append("| duration | %.6g | %s |" % (t, tu))
        except Exception:
            pass
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        try:
            t = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        name = r[ik].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
        n += 1
    lines = ["# kernel launch list summary of %s (%d launches, ncu serialised, cold cache)" %
             (path.split("/")[-1], n), "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f%% |" % (name, c, t / 1e3, 100 * t / total))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"full": full, "fullcsv": lambda a, b: full(a, b, True), "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
