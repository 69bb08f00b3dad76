"""Executable numpy model of the kernel's block cyclic reduction (dgpmp2_b200/csrc/bcr.cuh).

Mirrors the CUDA code item for item (level-ordered slots and their closed forms, E/F/g
storage, kept-node updates, the block-Thomas tail, back substitution) so the index algebra
can be validated on the CPU.
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


def slot_closed_form(T, t):
    """bcr_slot of bcr.cuh: off[l] = (T-1) - ((T-1) >> (l-1)) with l = ctz(t) + 1."""
    if t == 0:
        return T - 1
    z = (t & -t).bit_length() - 1
    return (T - 1) - ((T - 1) >> z) + (t >> (z + 1))


def state_of_slot_closed_form(T, m):
    """bcr_state_of_slot of bcr.cuh."""
    if m == T - 1:
        return 0
    l = 1
    while (T - 1) - ((T - 1) >> l) <= m:
        l += 1
    off = (T - 1) - ((T - 1) >> (l - 1))
    return (2 * (m - off) + 1) << (l - 1)


def make_plan(T, tail_max=4):
    """bcr_make_plan of bcr_plan.cuh: levels run, stride and length of the chain left for the tail."""
    nl = 0
    while ((T + (1 << nl) - 1) >> nl) > tail_max:
        nl += 1
    return nl, 1 << nl, (T + (1 << nl) - 1) >> nl


def bcr_solve(D, U, r, tail_max=4):
    """D (T,d,d) SPD diagonal blocks, U (T-1,d,d) = Lambda[t,t+1], r (T,d) -> x (T,d).
    nl elimination levels, then the block-Thomas tail on the chain t = S e (bcr_tail), then back substitution."""
    T, d, _ = D.shape
    nlev, off = make_levels(T)
    nl, S_t, nc = make_plan(T, tail_max)
    assert nl <= nlev and nc <= max(tail_max, 1) and (nc - 1) * S_t < T <= nc * S_t
    Dm = np.zeros((T, d, d)); Um = np.zeros((T, d, d)); Rm = np.zeros((T, d)); Em = np.zeros((T, d, d))
    Lm = [None] * T
    for m in range(T):
        t = state_of_slot(off, nlev, T, m)
        assert slot(off, T, t) == m == slot_closed_form(T, t) and state_of_slot_closed_form(T, m) == t
        Dm[m] = D[t]
        if t < T - 1:
            Um[m] = U[t]
        Rm[m] = r[t]
    for l in range(1, nl + 1):
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
    # tail: sequential block Cholesky over the chain t = S_t e, e = 0..nc-1 (root solve when nc == 1)
    prev = None
    for e in range(nc):
        pe = slot(off, T, S_t * e)
        if prev is not None:
            Dm[pe] -= Um[prev].T @ Um[prev]
            Rm[pe] -= Um[prev].T @ Rm[prev]
        L = np.linalg.cholesky(Dm[pe])
        Lm[pe] = L
        if e + 1 < nc:
            Um[pe] = np.linalg.solve(L, Um[pe])
        Rm[pe] = np.linalg.solve(L, Rm[pe])
        prev = pe
    xn = None
    for e in range(nc - 1, -1, -1):
        pe = slot(off, T, S_t * e)
        v = Rm[pe] if xn is None else Rm[pe] - Um[pe] @ xn
        Rm[pe] = np.linalg.solve(Lm[pe].T, v)
        xn = Rm[pe]
    for l in range(nl, 0, -1):
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
