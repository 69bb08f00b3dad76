"""Executable numpy model of the kernel's block cyclic reduction (dgpmp2_b200/csrc/bcr.cuh).

Mirrors the CUDA code item for item (level-ordered slots, E/F/g storage, kept-node
updates, back substitution) so the index algebra can be validated on the CPU.
Test infrastructure only.
"""
import numpy as np


def n_elim(T, s):
    return (T + s - 1) // (2 * s)


def n_kept(T, s):
    return (T + 2 * s - 1) // (2 * s)


def make_levels(T):
    off = [0, 0]
    s = 1
    while s < T:
        off.append(off[-1] + n_elim(T, s))
        s <<= 1
    return len(off) - 2, off


def slot(off, T, t):
    if t == 0:
        return T - 1
    l = (t & -t).bit_length()      # ctz(t) + 1
    return off[l] + (t >> l)


def state_of_slot(off, nlev, T, m):
    if m == T - 1:
        return 0
    l = 1
    while l < nlev and m >= off[l + 1]:
        l += 1
    return (2 * (m - off[l]) + 1) << (l - 1)


def bcr_solve(D, U, r):
    """D (T,d,d) SPD diagonal blocks, U (T-1,d,d) = Lambda[t,t+1], r (T,d) -> x (T,d)."""
    T, d, _ = D.shape
    nlev, off = make_levels(T)
    Dm = np.zeros((T, d, d)); Um = np.zeros((T, d, d)); Rm = np.zeros((T, d)); Em = np.zeros((T, d, d))
    Lm = [None] * T
    for m in range(T):
        t = state_of_slot(off, nlev, T, m)
        assert slot(off, T, t) == m
        Dm[m] = D[t]
        if t < T - 1:
            Um[m] = U[t]
        Rm[m] = r[t]
    for l in range(1, nlev + 1):
        s = 1 << (l - 1)
        ne = n_elim(T, s)
        for q in range(ne):
            j = s * (2 * q + 1)
            pj = off[l] + q
            assert pj == slot(off, T, j)
            pi = slot(off, T, j - s)
            has_right = (j + s) < T
            L = np.linalg.cholesky(Dm[pj])
            Lm[pj] = L
            Em[pj] = np.linalg.solve(L, Um[pi].T)
            Um[pj] = np.linalg.solve(L, Um[pj]) if has_right else 0.0
            Rm[pj] = np.linalg.solve(L, Rm[pj])
        nk = n_kept(T, s)
        for q in range(nk):
            i = 2 * s * q
            pi = slot(off, T, i)
            has_l, has_r = q > 0, (i + s) < T
            pl, pr = off[l] + q - 1, off[l] + q
            if has_l:
                Dm[pi] -= Um[pl].T @ Um[pl]
                Rm[pi] -= Um[pl].T @ Rm[pl]
            if has_r:
                Dm[pi] -= Em[pr].T @ Em[pr]
                Rm[pi] -= Em[pr].T @ Rm[pr]
            if (i + 2 * s) < T:
                Um[pi] = -Em[pr].T @ Um[pr]
    p0 = T - 1
    Rm[p0] = np.linalg.solve(Dm[p0], Rm[p0])
    for l in range(nlev, 0, -1):
        s = 1 << (l - 1)
        for q in range(n_elim(T, s)):
            j = s * (2 * q + 1)
            pj = off[l] + q
            pi = slot(off, T, j - s)
            v = Rm[pj] - Em[pj] @ Rm[pi]
            if (j + s) < T:
                v = v - Um[pj] @ Rm[slot(off, T, j + s)]
            Rm[pj] = np.linalg.solve(Lm[pj].T, v)
    x = np.zeros((T, d))
    for t in range(T):
        x[t] = Rm[slot(off, T, t)]
    return x
