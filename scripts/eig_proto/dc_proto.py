"""numpy model of the divide-and-conquer tridiagonal eigensolver in csrc/eigh.cuh (same steps, same formulas).
Test infrastructure / design aid only: nothing in the product imports it."""
import numpy as np

EPS = np.finfo(np.float64).eps / 2  # LAPACK dlamch('E')


def jacobi_leaf(T):
    n = T.shape[0]
    A = T.copy(); V = np.eye(n)
    for sweep in range(30):
        off = np.sqrt(np.sum(np.tril(A, -1) ** 2))
        if off <= 1e-300 or off <= EPS * 1e-3 * np.sqrt(np.sum(np.diag(A) ** 2)):
            break
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = A[p, q]
                if apq == 0.0:
                    continue
                theta = (A[q, q] - A[p, p]) / (2 * apq)
                t = np.sign(theta) / (abs(theta) + np.sqrt(theta * theta + 1)) if theta != 0 else 1.0
                c = 1 / np.sqrt(t * t + 1); s = t * c
                J = np.eye(n); J[p, p] = c; J[q, q] = c; J[p, q] = s; J[q, p] = -s
                A = J.T @ A @ J; V = V @ J
    return np.diag(A).copy(), V


def deflate(d, z, rho):
    """d (any order), z (unit norm), rho > 0. Returns perm (new order: non-deflated ascending first, then deflated), rotations
    [(col_a, col_b, c, s)] on ORIGINAL column indices applied in order, K, dl (K), w (K), d_out (k, in new order)."""
    k = len(d)
    order = np.argsort(d, kind="stable")
    d = d.copy(); z = z.copy()
    tol = 8 * EPS * max(np.max(np.abs(d)), np.max(np.abs(z)))
    rots = []
    nd = []      # indices (original) of non-deflated, in ascending order
    defl = []
    pj = -1
    for idx in order:
        if rho * abs(z[idx]) <= tol:
            defl.append(idx)
            continue
        if pj < 0:
            pj = idx
            continue
        s = z[pj]; c = z[idx]
        tau = np.hypot(c, s)
        t = d[idx] - d[pj]
        c /= tau; s = -s / tau
        if abs(t * c * s) <= tol:
            z[idx] = tau; z[pj] = 0.0
            rots.append((pj, idx, c, s))
            tt = d[pj] * c * c + d[idx] * s * s
            d[idx] = d[pj] * s * s + d[idx] * c * c
            d[pj] = tt
            defl.append(pj)
            pj = idx
        else:
            nd.append(pj)
            pj = idx
    if pj >= 0:
        nd.append(pj)
    perm = np.array(nd + defl, dtype=np.int64)
    K = len(nd)
    return perm, rots, K, d[perm], z[perm]


def f64_bits(x):
    return np.asarray(x, dtype=np.float64).view(np.int64)


def secular(dl, w, rho):
    """roots of 1 + rho sum w_i^2/(dl_i - lam); returns origin index o_j, mu_j, and delta[j,i] = dl_i - lam_j"""
    K = len(dl)
    w2 = rho * w * w
    j = np.arange(K)
    nxt = np.minimum(j + 1, K - 1)
    gap = dl[nxt] - dl[j]
    gap[K - 1] = 0.0
    half = 0.5 * gap
    # f at the midpoint, in coordinates relative to dl_j
    D0 = dl[None, :] - dl[j][:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        fmid = 1.0 + np.sum(w2[None, :] / (D0 - half[:, None]), axis=1)
    use_left = fmid > 0
    use_left[K - 1] = True
    o = np.where(use_left, j, nxt)
    Dl = dl[None, :] - dl[o][:, None]      # Delta_i relative to the origin
    sgn = np.where(use_left, 1.0, -1.0)    # mu = sgn * nu, nu > 0
    hi = half.copy()
    hi[K - 1] = np.sum(w2)
    lo_b = np.zeros(K, dtype=np.int64)
    hi_b = f64_bits(hi).copy()
    # invariant: g(sgn*lo) has the sign "before the root", g(sgn*hi) "after"
    for _ in range(64):
        mid_b = lo_b + ((hi_b - lo_b) >> 1)
        nu = mid_b.view(np.float64)
        mu = sgn * nu
        with np.errstate(divide="ignore", invalid="ignore"):
            g = 1.0 + np.sum(w2[None, :] / (Dl - mu[:, None]), axis=1)
        # left origin: g increasing in nu (negative near 0); right origin: mu=-nu, g decreasing in nu (positive near 0)
        before = np.where(use_left, g < 0, g > 0)
        before |= (mid_b == 0)
        lo_b = np.where(before, mid_b, lo_b)
        hi_b = np.where(before, hi_b, mid_b)
        if np.all(hi_b - lo_b <= 1):
            break
    nu = hi_b.view(np.float64)
    mu = sgn * nu
    delta = Dl - mu[:, None]
    return o, mu, delta


def merge(d, Q, rho_z):
    """d: k child eigenvalues, Q: k x k block-diagonal child eigenvectors, rho_z = (rho, z)"""
    rho, z = rho_z
    k = len(d)
    perm, rots, K, dperm, zperm = deflate(d, z, rho)
    Q = Q.copy()
    for (a, b, c, s) in rots:
        qa = Q[:, a].copy(); qb = Q[:, b].copy()
        Q[:, a] = c * qa + s * qb     # dlaed2: drot(q_pj, q_nj, c, s): x' = c x + s y ; y' = c y - s x
        Q[:, b] = c * qb - s * qa
    Qp = Q[:, perm]
    Vt = np.eye(k)
    lam = dperm.copy()
    if K > 0:
        dl = dperm[:K]; w = zperm[:K]
        o, mu, delta = secular(dl, w, rho)
        lam[:K] = dl[o] + mu
        # Loewner: what_i^2 = prod_j (lam_j - dl_i) / prod_{j != i} (dl_j - dl_i)
        what = np.empty(K)
        for i in range(K):
            p = delta[i, i]   # dl_i - lam_i  (< 0)
            for jj in range(K):
                if jj != i:
                    p *= delta[jj, i] / (dl[i] - dl[jj])
            what[i] = np.copysign(np.sqrt(-p), w[i])
        V = what[None, :] / delta          # row j = eigenvector j
        V /= np.linalg.norm(V, axis=1)[:, None]
        Vt[:K, :K] = V
    return lam, Qp @ Vt.T, K


def dc_eigh(d, e, leaf=32):
    n = len(d)
    d = d.astype(np.float64).copy(); e = e.astype(np.float64).copy()
    scale = max(np.max(np.abs(d)), np.max(np.abs(e)) if n > 1 else 0.0)
    if scale == 0.0:
        return np.zeros(n), np.eye(n), []
    d /= scale; e /= scale
    # partition
    sizes = [n]
    while max(sizes) > leaf:
        sizes = [x for s_ in sizes for x in ((s_ // 2, s_ - s_ // 2) if s_ > leaf else (s_,))]
    # simpler: uniform binary tree
    def split(lo, hi):
        if hi - lo <= leaf:
            return ("leaf", lo, hi)
        m = (lo + hi) // 2
        m += (m - lo) & 1      # even left size -> aligned blocks
        return ("node", lo, hi, m, split(lo, m), split(m, hi))
    tree = split(0, n)
    # tear
    def tear(t):
        if t[0] == "node":
            m = t[3]
            b = abs(e[m - 1])
            d[m - 1] -= b; d[m] -= b
            tear(t[4]); tear(t[5])
    tear(tree)
    Q = np.zeros((n, n)); lam = np.zeros(n)
    stats = []
    def solve(t):
        if t[0] == "leaf":
            lo, hi = t[1], t[2]
            T = np.diag(d[lo:hi]) + np.diag(e[lo:hi - 1], 1) + np.diag(e[lo:hi - 1], -1)
            lam[lo:hi], Q[lo:hi, lo:hi] = jacobi_leaf(T)
            return
        _, lo, hi, m, l, r = t
        solve(l); solve(r)
        beta = e[m - 1]
        z = np.concatenate([Q[m - 1, lo:m], np.sign(beta) * Q[m, m:hi]]) / np.sqrt(2.0)
        if beta == 0:
            z = np.concatenate([Q[m - 1, lo:m], Q[m, m:hi]]) / np.sqrt(2.0)
        rho = 2 * abs(beta)
        if rho == 0.0:
            return
        lam[lo:hi], Q[lo:hi, lo:hi], K = merge(lam[lo:hi], Q[lo:hi, lo:hi], (rho, z))
        stats.append((hi - lo, K))
    solve(tree)
    order = np.argsort(lam, kind="stable")
    return lam[order] * scale, Q[:, order], stats


def check(d, e, name):
    n = len(d)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, Q, stats = dc_eigh(d, e)
    ref = np.linalg.eigvalsh(T)
    nrm = max(np.max(np.abs(ref)), 1e-300)
    print(f"{name:28s} n={n:5d} eig {np.max(np.abs(lam - ref)) / nrm:.2e} res {np.max(np.abs(T @ Q - Q * lam)) / nrm:.2e} "
          f"orth {np.max(np.abs(Q.T @ Q - np.eye(n))):.2e} top-merge K={stats[-1] if stats else None}")


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (5, 33, 64, 100, 257, 600):
        check(rng.standard_normal(n), rng.standard_normal(n - 1), "random")
    n = 300
    check(np.ones(n), np.zeros(n - 1), "identity")
    check(np.full(n, 2.0), np.full(n - 1, -1.0), "laplace")
    check(np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1), "wilkinson")
    # tridiagonal of a rank-deficient Gram matrix (clustered at 0)
    import scipy.linalg as sl
    X = rng.standard_normal((400, 12)) + 2
    G = (X @ X.T); nn = np.sqrt(np.diag(G)); G = (G / nn[:, None] / nn[None, :]) ** 2
    H = sl.hessenberg(G)
    check(np.diag(H).copy(), np.diag(H, 1).copy(), "gram rank-deficient")
    Xr = rng.standard_normal((400, 30)) / 3
    D2 = ((Xr[:, None, :] - Xr[None, :, :]) ** 2).sum(-1)
    H = sl.hessenberg(np.exp(-0.5 * D2))
    check(np.diag(H).copy(), np.diag(H, 1).copy(), "rbf")
    d = rng.standard_normal(200); e = rng.standard_normal(199) * 1e-9
    check(d, e, "tiny off-diagonals")
    e = rng.standard_normal(199); e[::7] = 0
    check(d, e, "exact zero couplings")
