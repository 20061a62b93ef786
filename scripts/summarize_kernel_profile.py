"""Turns one `ncu --set full` report under gpurun_out/ into a markdown summary under profiles/ (run here, no GPU):
key metrics, DRAM traffic, stall reasons over all sampled instructions, instruction mix, hottest SASS lines.

    python scripts/summarize_kernel_profile.py <rep in gpurun_out/> <out.md in profiles/> "<title>" "<command>" [expected.json]

`expected.json` (written by scripts/prof_target.py in a run WITHOUT the profiler) holds the CUDA-event time of the launch
the capture claims to describe; the summary is REFUSED when ncu's gpu__time_duration is not within a factor 2 of it
(round 1 summarised a 48-iteration tuning launch under the title of the 1000-iteration one)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
expected = json.load(open(os.path.join(ROOT, "gpurun_out", sys.argv[5]))) if len(sys.argv) > 5 else None
rep = os.path.join(ROOT, "gpurun_out", rep)


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def num(v):
    return float(v.replace(",", ""))


UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, unit, val = rows[0], rows[1], rows[2]
if len(rows) > 3:
    print("note: the report holds %d launches; the first is summarised" % (len(rows) - 2))
d = {h: (u, v) for h, u, v in zip(hdr, unit, val)}
dur_us = num(d["gpu__time_duration.sum"][1]) * UNIT[d["gpu__time_duration.sum"][0]]
if expected is not None:
    exp_us = expected["expected_ms_per_launch"] * 1e3
    if not (0.5 * exp_us <= dur_us <= 2.0 * exp_us):
        raise SystemExit("REFUSED: the captured launch ran %.1f us under ncu, the benchmarked launch takes %.1f us (CUDA events): "
                         "this is not the launch the summary would claim to describe" % (dur_us, exp_us))
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
lines = ["# " + title, ""]
if cmd:
    lines += ["Command: `%s`" % cmd, ""]
if expected is not None:
    lines += ["The launch captured is the benchmarked one: gpu__time_duration %.1f us under ncu (cold cache, serialised replays) against "
              "%.1f us by CUDA events without the profiler (`%s`); a capture outside a factor 2 is refused by this script." % (
                  dur_us, expected["expected_ms_per_launch"] * 1e3, json.dumps({k: v for k, v in expected.items() if k != "expected_ms_per_launch"})), ""]
lines += ["| metric | unit | value |", "|---|---|---:|"]
for h in want:
    if h in d:
        lines.append("| %s | %s | %s |" % (h, d[h][0], d[h][1]))
traffic = None
try:
    rd = num(d["dram__bytes_read.sum"][1]) * UNIT[d["dram__bytes_read.sum"][0]]
    wr = num(d["dram__bytes_write.sum"][1]) * UNIT[d["dram__bytes_write.sum"][0]]
    traffic = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr), "ncu_duration_us": dur_us}
    lines += ["", "DRAM traffic of the launch: %.3f MB read + %.3f MB written." % (rd / 1e6, wr / 1e6)]
    if expected is not None and "bytes_per_iter" in expected and "iters" in expected:
        alg = expected["bytes_per_iter"] * expected["iters"]
        lines[-1] += (" Algorithmic bytes of the launch (SURVEY 8d: %d B x %d iterations) = %.3f GB: the matrix and the vectors are served "
                      "from L2 / shared memory after the first pass, DRAM carries %.4f of the algorithmic bytes." % (
                          expected["bytes_per_iter"], expected["iters"], alg / 1e9, (rd + wr) / alg))
        traffic.update(algorithmic_bytes_per_launch=int(alg))
except Exception as e:   # noqa: BLE001
    lines.append("(DRAM traffic unavailable: %s)" % e)
src = list(csv.reader(io.StringIO(ncu("source"))))
h2, data = src[1], src[2:]
si, ie, sc = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
names = [n for n in h2 if n.startswith("stall_") and "Not Issued" not in n]
T = sum(float(r[si] or 0) for r in data)
tot = {n: sum(float(r[h2.index(n)] or 0) for r in data) for n in names}
lines += ["", "## Warp-state samples over the whole kernel (%d samples)" % T, "", "| reason | share |", "|---|---:|"]
for n, v in sorted(tot.items(), key=lambda x: -x[1])[:9]:
    lines.append("| %s | %.1f %% |" % (n, 100 * v / T))
mix = collections.Counter()
for r in data:
    op = r[sc].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    mix[o.split(".")[0]] += int(float(r[ie] or 0))
ti = sum(mix.values())
lines += ["", "## Instruction mix (warp instructions executed: %d)" % ti, "", "| opcode | share |", "|---|---:|"]
for k, v in mix.most_common(12):
    lines.append("| %s | %.1f %% |" % (k, 100.0 * v / ti))
lines += ["", "## Hottest SASS lines by stall samples", "", "| samples | executed | SASS |", "|---:|---:|---|"]
for r in sorted(data, key=lambda r: -float(r[si] or 0))[:14]:
    lines.append("| %.1f %% | %s | `%s` |" % (100 * float(r[si] or 0) / T, r[ie], r[sc].strip()[:90]))
open(os.path.join(ROOT, "profiles", out), "w").write("\n".join(lines) + "\n")
if traffic is not None and expected is not None:
    traffic.update(workload=expected.get("workload"), iters_per_step=expected.get("iters"), kernel=expected.get("kernel"),
                   source="profiles/%s (ncu --set full, the benchmarked launch)" % out)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", out.replace("_ncu_full.md", "_traffic.json")), "w"))
print("ok: %s (%.1f us under ncu)" % (out, dur_us))
