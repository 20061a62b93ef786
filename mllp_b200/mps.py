"""Netlib MPS reader -> general-form LP for the B200 path.

The reference never parses MPS (``netlib_mps/`` is only listed for instance names,
linear_program_data.py:23-24, :59-60); its arrays drop row senses and variable bounds, which
makes 18 instances unbounded and 2 infeasible.  This reader restores them:

    min c'x + offset   s.t.   A x - b in K_row (ylo <= y <= yhi),   lb <= x <= ub

Rows: ``E`` -> dual free, ``G`` (a'x >= b) -> y in [0, inf), ``L`` (a'x <= b) -> y in (-inf, 0].
RANGES rows become equality rows with one extra slack column boxed by the range: by default
a'x - s = 0, lo <= s <= hi; with ``range_form="dataset"`` exactly the representation of the
reference's raw arrays ``dataset/netlib_mps/*`` (Gurobi's): a'x + s = b,
0 <= s <= |R| on an ``L`` row, a'x - s = b on a ``G`` row, the extra columns appended in row order --
``A``, ``c`` and ``b`` then equal the reference's arrays bit for bit on all 97 files.
Dialect handled: fixed/free format with whitespace-separated fields, ``*``
comments, OBJSENSE (MAX is turned into MIN of -c), objective-row RHS = -offset, bound types
UP / LO / FX / FR / MI / PL / BV, negative UP with untouched lower bound -> lower = -inf.
"""
import gzip

import numpy as np
import scipy.sparse as sp

INF = float("inf")


def read_mps(path, range_form="boxed"):
    if range_form not in ("boxed", "dataset"):
        raise ValueError("range_form must be 'boxed' or 'dataset'")
    rows, row_sense, obj_row = {}, [], None
    cols, col_names = {}, []
    entries = []                 # (row, col, value)
    cobj = {}
    rhs, ranges = {}, {}
    lb, ub = {}, {}
    touched_lb = set()
    offset, maximize = 0.0, False
    section = None
    opener = (lambda p: gzip.open(p, "rt")) if str(path).endswith(".gz") else (lambda p: open(p, "r"))
    with opener(path) as fh:
        for raw in fh:
            if not raw.strip() or raw[0] == "*":
                continue
            if raw[0] not in " \t":
                tok = raw.split()
                section = tok[0].upper()
                if section == "OBJSENSE" and len(tok) > 1:
                    maximize = tok[1].upper().startswith("MAX")
                    section = None
                if section == "ENDATA":
                    break
                continue
            tok = raw.split()
            if section == "OBJSENSE":
                maximize = tok[0].upper().startswith("MAX")
            elif section == "ROWS":
                kind, name = tok[0].upper(), tok[1]
                if kind == "N":
                    if obj_row is None:
                        obj_row = name
                    continue            # further free rows are dropped
                rows[name] = len(row_sense)
                row_sense.append(kind)
            elif section == "COLUMNS":
                if len(tok) >= 3 and tok[1] == "'MARKER'":
                    continue
                name = tok[0]
                if name not in cols:
                    cols[name] = len(col_names)
                    col_names.append(name)
                j = cols[name]
                for k in range(1, len(tok) - 1, 2):
                    r, v = tok[k], float(tok[k + 1])
                    if r == obj_row:
                        cobj[j] = cobj.get(j, 0.0) + v
                    elif r in rows:
                        entries.append((rows[r], j, v))
            elif section in ("RHS", "RANGES"):
                pairs = tok[1:] if len(tok) % 2 == 1 else tok      # optional set name
                target = rhs if section == "RHS" else ranges
                for k in range(0, len(pairs) - 1, 2):
                    r, v = pairs[k], float(pairs[k + 1])
                    if section == "RHS" and r == obj_row:
                        offset = -v
                    elif r in rows:
                        target[rows[r]] = v
            elif section == "BOUNDS":
                kind = tok[0].upper()
                if kind in ("FR", "MI", "PL", "BV"):
                    name = tok[2] if len(tok) >= 3 else tok[1]
                    val = None
                else:
                    name, val = (tok[2], float(tok[3])) if len(tok) >= 4 else (tok[1], float(tok[2]))
                if name not in cols:
                    continue
                j = cols[name]
                if kind == "UP":
                    ub[j] = val
                    if val < 0 and j not in touched_lb:
                        lb[j] = -INF
                elif kind == "LO":
                    lb[j] = val; touched_lb.add(j)
                elif kind == "FX":
                    lb[j] = val; ub[j] = val; touched_lb.add(j)
                elif kind == "FR":
                    lb[j] = -INF; ub[j] = INF; touched_lb.add(j)
                elif kind == "MI":
                    lb[j] = -INF; touched_lb.add(j)
                elif kind == "PL":
                    ub[j] = INF
                elif kind == "BV":
                    lb[j] = 0.0; ub[j] = 1.0; touched_lb.add(j)
    m, n = len(row_sense), len(col_names)
    b = np.zeros(m)
    for i, v in rhs.items():
        b[i] = v
    ylo, yhi = np.full(m, -INF), np.full(m, INF)
    for i, s in enumerate(row_sense):
        if s == "G":
            ylo[i] = 0.0
        elif s == "L":
            yhi[i] = 0.0
    xl = np.zeros(n); xu = np.full(n, INF)
    for j, v in lb.items():
        xl[j] = v
    for j, v in ub.items():
        xu[j] = v
    c = np.zeros(n)
    for j, v in cobj.items():
        c[j] = v
    # RANGES: row i with range R: E -> [b, b+|R|] if R>0 else [b-|R|, b]; G -> [b, b+|R|]; L -> [b-|R|, b]
    extra_cols = []
    for i, R in sorted(ranges.items()):
        s = row_sense[i]
        if s == "E":
            lo, hi = (b[i], b[i] + abs(R)) if R >= 0 else (b[i] - abs(R), b[i])
        elif s == "G":
            lo, hi = b[i], b[i] + abs(R)
        else:
            lo, hi = b[i] - abs(R), b[i]
        extra_cols.append((i, lo, hi))
    ri = [e[0] for e in entries]; ci = [e[1] for e in entries]; vv = [e[2] for e in entries]
    for k, (i, lo, hi) in enumerate(extra_cols):
        ri.append(i); ci.append(n + k)
        if range_form == "dataset":
            # the row keeps its right-hand side; the slack (0 <= s <= hi - lo) closes the gap to it from the row's side
            upper = (row_sense[i] == "L") or (row_sense[i] == "E" and ranges[i] < 0)
            vv.append(1.0 if upper else -1.0)
            extra_cols[k] = (i, 0.0, hi - lo)
        else:
            vv.append(-1.0)        # a'x - s = 0
            b[i] = 0.0
        ylo[i] = -INF; yhi[i] = INF
    n2 = n + len(extra_cols)
    A = sp.csr_matrix((vv, (ri, ci)), shape=(m, n2))
    A.eliminate_zeros()   # explicit zeros in the file (standgub has one) are not entries; HiGHS drops them too
    A.sum_duplicates(); A.sort_indices()
    if extra_cols:
        c = np.concatenate([c, np.zeros(len(extra_cols))])
        xl = np.concatenate([xl, [e[1] for e in extra_cols]])
        xu = np.concatenate([xu, [e[2] for e in extra_cols]])
        col_names = col_names + ["__range_%s" % k for k in range(len(extra_cols))]
    if maximize:
        c, offset = -c, -offset
    # row_sense of a range row is reported as "E" (it is an equality row now); "row_sense_mps" keeps the file's letters
    sense_out = ["E" if i in ranges else s_ for i, s_ in enumerate(row_sense)]
    return {"A": A, "b": b, "c": c, "lb": xl, "ub": xu, "ylo": ylo, "yhi": yhi, "offset": offset,
            "maximize": maximize, "row_sense": sense_out, "row_sense_mps": row_sense, "col_names": col_names,
            "num_range_cols": len(extra_cols)}


def to_loader_tuple(lp):
    """(constrs, constr_weights, rhs, coefs) in the reference loader's representation
    (linear_program_data.py:75-77) plus the bound arrays the general-form kernels take."""
    A = lp["A"].tocsr()
    constrs = np.split(A.indices.astype(np.int32), A.indptr)[1:-1]
    return constrs, A.data, lp["b"], lp["c"], dict(lb=lp["lb"], ub=lp["ub"], ylo=lp["ylo"], yhi=lp["yhi"])
