"""Turns one `ncu --set full` report under gpurun_out/ into a markdown summary under profiles/ (run here, no GPU):
key metrics, stall reasons over all sampled instructions, hottest SASS lines."""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
rep = os.path.join(ROOT, "gpurun_out", rep)


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, unit, val = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
lines = ["# " + title, ""]
if cmd:
    lines += ["Command: `%s`" % cmd, ""]
lines += ["| metric | unit | value |", "|---|---|---:|"]
for h, u, v in zip(hdr, unit, val):
    if h in want:
        lines.append("| %s | %s | %s |" % (h, u, v))
src = list(csv.reader(io.StringIO(ncu("source"))))
h2, data = src[1], src[2:]
si = h2.index("# Samples")
names = [n for n in h2 if n.startswith("stall_") and "Not Issued" not in n]
T = sum(float(r[si]) for r in data if r[si].replace(".", "").isdigit())
tot = {n: sum(float(r[h2.index(n)] or 0) for r in data) for n in names}
lines += ["", "Warp-state samples over the whole kernel (%d samples): " % T +
          ", ".join("%s %.1f %%" % (n.replace("stall_", ""), 100 * v / T) for n, v in sorted(tot.items(), key=lambda x: -x[1]) if v / T > 0.01), "",
          "Hottest instructions:", "", "| samples | executed | SASS |", "|---:|---:|---|"]
ie = h2.index("Instructions Executed")
for r in sorted(data, key=lambda r: -float(r[si] or 0))[:12]:
    lines.append("| %.1f %% | %s | `%s` |" % (100 * float(r[si]) / T, r[ie], r[1].strip()[:90]))
open(os.path.join(ROOT, "profiles", out), "w").write("\n".join(lines) + "\n")
print("ok")
