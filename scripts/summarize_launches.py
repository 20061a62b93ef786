"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> profiles/<tag>_launches.md (run here, no GPU).
usage: python scripts/summarize_launches.py r02 gpurun_out/r02_launches.csv "<the profiled command>" """
import collections, csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, path, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[mu], 1)
    a = agg.setdefault(r[kn], [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
lines = ["# %s -- ncu launch list of `%s`" % (tag, cmd), "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolute times).",
         "The persistent kernel is ONE launch per 1000-iteration step; its short launches are the 48-iteration geometry / tuning",
         "measurements of `mllp_lp_create` (outside the timed region, reported as `create_s`); gather / scatter / eval are the boundary",
         "kernels of `mllp_pdhg_run`; `k_spmv` / `k_sumsq*` / `k_scale_by_invnorm` are the one-off power iteration of `sigma_max`.", "",
         "| kernel | launches | total ms | longest launch ms | share |", "|---|---:|---:|---:|---:|"]
for k, (c, t, mx) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append("| `%s` | %d | %.3f | %.3f | %.1f %% |" % (k[:90], c, t / 1e6, mx / 1e6, 100 * t / tot))
open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:16]))
