"""CPU restatement (numpy/scipy) of the rule that produced the reference's `_norm` arrays -- TEST INFRASTRUCTURE, not
product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

The reference ships the arrays (dataset/netlib_mps_norm/*, consumed at linear_program_data.py:66-77) but not the
script that made them; the rule below was reverse-engineered from the data and is PINNED by the
data itself (tests/test_norm_rule.py, all 97 instances whose MPS file exists, from the MPS text):

    standard form : one slack column per inequality row, appended in row order (L: +1, G: -1); RANGES rows are
                    already equalities in the raw arrays
    row scaling   : r_i = ||(a_i, slack_i)||_2 with the squares added ONE BY ONE in ascending column order (plain
                    left-to-right accumulation, not numpy's pairwise reduction: this order reproduces the reference's
                    numbers to the bit);  |b_i| / r_i <= 5 : row and b_i are DIVIDED by r_i;
                    otherwise: multiplied by d_i = 5 / b_i (signed: a large negative b_i flips the row, b becomes +5)
    output        : A_norm = diag(d) [A | S],  rhs_norm = d * b,  coefs_norm = [c / ||c||_2, 0 ... 0]  (c as is if 0)

What the pin shows: the sparsity structure of all 97 `_constrs.npz` is reproduced exactly; every entry of a row
scaled by 1 / r_i is BIT-IDENTICAL; rows scaled by 5 / b_i, the right-hand sides and c / ||c|| agree to a few units
in the last place (<= 2e-15 relative; the generator's operation order for `5 / b` cannot be recovered -- each such
row matches a different one of the algebraically equal forms a*5/b, a/b*5, (a/r)/((b/r)/5) ...).
"""
import numpy as np
import scipy.sparse as sp


def standard_form(A, sense):
    """[A | S]: S has one column per row with sense 'L' (+1) or 'G' (-1), in row order."""
    A = sp.csr_matrix(A, dtype=np.float64)
    m, n = A.shape
    sense = np.asarray(sense)
    rows = np.flatnonzero((sense == "L") | (sense == "G"))
    sign = np.where(sense[rows] == "L", 1.0, -1.0)
    S = sp.csr_matrix((sign, (rows, np.arange(rows.size))), shape=(m, rows.size))
    out = sp.hstack([A, S], format="csr")
    out.sort_indices()
    return out


def row_norms_sequential(As):
    """sqrt of the left-to-right sum of squares of every row (ascending column order)."""
    sq = (As.data * As.data).tolist()
    ip = As.indptr
    r2 = np.empty(As.shape[0])
    for i in range(As.shape[0]):
        acc = 0.0
        for k in range(ip[i], ip[i + 1]):
            acc += sq[k]
        r2[i] = acc
    return np.sqrt(r2)


def norm_rule(A, b, c, sense):
    """Returns (A_norm csr, rhs_norm, coefs_norm, info) from the raw arrays and the row senses ('E' / 'L' / 'G');
    info: 'divided' = rows scaled by 1 / r_i (bool), 'r', 'c_norm2'."""
    As = standard_form(A, sense)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    m = As.shape[0]
    r = row_norms_sequential(As)
    with np.errstate(divide="ignore", invalid="ignore"):
        divided = (np.abs(b) / r <= 5.0) & (r > 0.0)
        d5 = 5.0 / b
    rows = np.repeat(np.arange(m), np.diff(As.indptr))
    with np.errstate(invalid="ignore"):       # 5 / 0 on rows that are divided anyway
        data = np.where(divided[rows], As.data / np.where(r > 0, r, 1.0)[rows], As.data * d5[rows])
        rhs = np.where(divided, b / np.where(r > 0, r, 1.0), b * d5)
    rhs = np.where(r > 0.0, rhs, b)            # an empty row keeps its right-hand side
    An = sp.csr_matrix((data, As.indices.copy(), As.indptr.copy()), shape=As.shape)
    cn = float(np.linalg.norm(c))
    coefs = np.concatenate([c / cn if cn > 0 else c, np.zeros(As.shape[1] - A.shape[1])])
    return An, rhs, coefs, {"divided": divided, "r": r, "c_norm2": cn}
