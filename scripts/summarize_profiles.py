"""Turns gpurun_out/ ncu artefacts into the committed summaries under profiles/ (run here, no GPU)."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
launch_csv = sys.argv[2] if len(sys.argv) > 2 else "launches_r1c.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "prof_r1c_persistent.ncu-rep"

def launches():
    rows = list(csv.reader(open(os.path.join(GO, launch_csv))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv: continue
        v = float(r[mv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[mu], 1)
        a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["# %s -- ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --iters-per-step 200`" % tag, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (cold-cache, serialised: compare SHARES).",
             "The persistent kernel is ONE launch per step (200 fused iterations here); gather/scatter/eval are the",
             "boundary kernels of `mllp_pdhg_run`; `k_spmv`/`k_sumsq`/`k_scale_by_invnorm` are the one-off power iteration", "of `sigma_max` (outside the timed region).", "",
             "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], c, t / 1e6, 100 * t / tot))
    open(os.path.join(OUT, tag + "_launches.md"), "w").write("\n".join(lines) + "\n")

def ncu(page):
    return subprocess.run(["ncu", "-i", os.path.join(GO, rep), "--page", page, "--csv"], capture_output=True, text=True).stdout

def full():
    rows = list(csv.reader(io.StringIO(ncu("raw"))))
    hdr, unit, val = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sectors_srcunit_tex.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    lines = ["# %s -- `ncu --set full` of `k_pdhg_persistent` (osa-60, 1000 fused iterations in the launch = the bench configuration)" % tag, "",
             "Command: `ncu --set full --clock-control none --import-source on -k regex:k_pdhg_persistent -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --iters-per-step 1000`", "",
             "| metric | unit | value |", "|---|---|---:|"]
    d = {}
    for h, u, v in zip(hdr, unit, val):
        if h in want:
            lines.append("| %s | %s | %s |" % (h, u, v)); d[h] = v
    try:
        dur = float(d["gpu__time_duration.sum"].replace(",", ""))
        rd = float(d["dram__bytes_read.sum"].replace(",", "")); wr = float(d["dram__bytes_write.sum"].replace(",", ""))
        lines += ["", "Per launch (1000 iterations): DRAM traffic = %s + %s (units above); algorithmic bytes = 1000 x 44 866 664 B = 44.9 GB." % (d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]),
                  "DRAM traffic is ~1000x BELOW the algorithmic bytes: after the first iteration the matrix is served from L2 and shared memory.",
                  "The per-launch time under ncu is cold-cache/serialised; the bench number (CUDA events) is the one to quote."]
    except Exception as e:
        lines.append("(derived numbers unavailable: %s)" % e)
    rows = list(csv.reader(io.StringIO(ncu("source"))))
    hdr, data = rows[1], rows[2:]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in data)
    agg = {c: sum(int(r[hdr.index(c)]) for r in data) for c in cols}
    lines += ["", "## Warp stall sampling (all samples = %d)" % tot, "", "| reason | share |", "|---|---:|"]
    for c, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        lines.append("| %s | %.1f %% |" % (c, 100.0 * v / tot))
    lines += ["", "`stall_barrier` = warps parked at the two `bar.sync` of the grid barrier (waiting for the slowest warp of the CTA, then for the",
              "slowest CTA + barrier latency); `stall_long_sb` = waiting on gathers / vector loads.", "",
              "## Instruction mix (warp instructions executed: %d)" % sum(int(r[ie]) for r in data), "", "| opcode | share |", "|---|---:|"]
    h = collections.Counter()
    for r in data:
        op = r[src].strip().split()
        if not op: continue
        o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        h[o.split(".")[0]] += int(r[ie])
    ti = sum(h.values())
    for k, v in h.most_common(12):
        lines.append("| %s | %.1f %% |" % (k, 100.0 * v / ti))
    top = sorted(data, key=lambda r: -int(r[si]))[:12]
    lines += ["", "## Hottest SASS lines by stall samples", "", "| SASS | samples | share |", "|---|---:|---:|"]
    for r in top:
        lines.append("| `%s` | %s | %.1f %% |" % (r[src].strip()[:70], r[si], 100.0 * int(r[si]) / tot))
    open(os.path.join(OUT, tag + "_persistent_ncu_full.md"), "w").write("\n".join(lines) + "\n")

launches(); full()
for f in ("bench_r1e.json", "bench_r1e_ken18.json", "bench_r1e_ref.json", "bench_r1e_n2.json"):
    p = os.path.join(GO, f)
    if os.path.exists(p):
        txt = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if txt: open(os.path.join(OUT, tag + "_" + f.replace("_r1e", "")), "w").write(txt[-1] + "\n")
print("ok")
