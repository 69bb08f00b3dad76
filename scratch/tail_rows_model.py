"""Lane-level numpy model of bcr_tail_rows (dgpmp2_b200/csrc/bcr.cuh): the chain of nc kept nodes left after the
cyclic-reduction levels is a scalar SPD band system of n = nc*D rows (half bandwidth 2D-1).  One lane owns one
scalar row and keeps that row's band in a window c[0..2D-1] that slides with the pivot, so that every register
index in the CUDA code is static; the pivot's reciprocal square root travels by one shuffle, the column of L by
a shared-memory scratch line.  The model executes the same steps in the same order (including the software
pipelining of the trailing update) with one numpy array entry per lane.
"""
import numpy as np


def tail_rows(Db, Ub, rb, LP=None):
    """Db (nc,D,D) symmetric diagonal blocks, Ub (nc-1,D,D) couplings Lambda[e][e+1], rb (nc,D) -> x (nc,D)."""
    nc, D, _ = Db.shape
    n = nc * D
    BW = 2 * D - 1
    if LP is None:
        LP = 8 if n <= 8 else 16 if n <= 16 else 32
    assert n <= LP
    rows = np.arange(LP)
    e, a = rows // D, rows % D
    live = rows < n
    cb = D * np.maximum(e - 1, 0)
    c = np.zeros((LP, BW + 1))
    for i in range(LP):
        if not live[i]:
            continue
        for m in range(BW + 1):
            j = cb[i] + m
            if j <= i:
                ej, cj = j // D, j % D
                if ej == e[i]:
                    c[i, m] = Db[e[i], a[i], cj]
                elif ej == e[i] - 1:
                    c[i, m] = Ub[e[i] - 1, cj, a[i]]
    b = np.where(live, np.concatenate([rb.reshape(-1), np.zeros(LP - n)]), 0.0)
    q = np.zeros((LP, BW))
    rinv = np.zeros(LP)
    sl = np.zeros(LP + BW + 1)                 # scratch line: l of every lane, BW zeros, g_k
    dpiv = c[:, 0].copy()
    lprev = np.zeros(LP)
    ljprev = np.zeros((LP, BW + 1))
    gkprev = 0.0
    rkprev = 0.0
    ok = True
    for k in range(n + 1):
        if k < n:
            with np.errstate(all='ignore'):
                rk_own = 1.0 / np.sqrt(dpiv)
            rk = rk_own[k]                                         # shuffle from lane k
            ok = ok and dpiv[k] > 0
        if k > 0:                                                  # finish pivot k-1
            pa = (k - 1) >= cb
            for m in range(1, BW + 1):
                c[:, m] = np.where(pa, c[:, m] - lprev * ljprev[:, m], c[:, m])
            c[pa, :-1] = c[pa, 1:]
            c[pa, -1] = 0.0
            b = np.where(rows > k - 1, b - lprev * gkprev, b)
            own = rows == k - 1
            for m in range(1, BW + 1):
                q[own, BW - m] = ljprev[own, m]
            b[own] = gkprev
            rinv[own] = rkprev
        if k == n:
            break
        l = np.where((k >= cb) & (rows > k), c[:, 0] * rk, 0.0)
        gk_own = b * rk
        sl[:LP] = l
        sl[LP + BW] = gk_own[k]
        dpiv = c[:, 1] - l * l
        for m in range(1, BW + 1):
            ljprev[:, m] = sl[k + m]
        gkprev = sl[LP + BW]
        lprev = l
        rkprev = rk
    x = np.zeros(LP)
    for k in range(n - 1 + BW, -1, -1):
        if k < n:
            x = np.where(rows == k, b * rinv, x)
            xk = x[k]
        else:
            xk = 0.0
        act = (rows < k) & (k <= rows + BW)
        b = np.where(act, b - q[:, 0] * xk, b)
        q[act, :-1] = q[act, 1:]
    return x[:n].reshape(nc, D), ok
