// EXPERIMENT (not compiled into the library): row-per-lane scalar Cholesky for the chain left after the
// cyclic-reduction levels, measured on B200 against the shipped block-Thomas tail (bcr.cuh: bcr_tail).
//
//   B=1024, T=64 (7 problems per CTA):  block tail 17.38 us / step,  this tail 17.89 us  (first version, with
//   per-lane shift registers for the column of L and branches in the pivot loop: 20.74 us)
//   B=1, T=64:                           11.09 us vs 11.64 us;   B=1024, T=128: 37.07 vs 38.18 us
//
// All 70 schedule / parity tests pass with it (DGPMP2_TAILMODE=2 in the experiment build), i.e. it is correct; it is
// slower because a problem's lanes sit in ONE warp, a lone warp issues in order, and a pivot still costs ~72
// instructions (predicated window slides, selects) at ~4 cycles each: ~300 cycles per scalar pivot against the
// ~190 per pivot (756 per 4x4 block step) of the block tail -- the dependent chain (rsqrt, shuffle, multiply, fma
// ~ 100 cycles) is not what bounds it.  To try again: drop it next to bcr_tail in bcr.cuh and dispatch on a plan flag.
// Lane-level numpy model of the same algorithm: scratch/tail_rows_model.py.

// (d') the same chain as ONE scalar SPD band system of n = nc * D rows (half bandwidth 2D - 1), one lane per
// scalar row (plan.tail_mode == 2, n <= 32).  On the chain of a pivot there is only
// rsqrt -> one shuffle -> one multiply -> one fma:
//   * lane i keeps the band of row i in a window c[0..2D-1] that starts at column D*max(e-1, 0) (e = its node) and
//     slides by one column per pivot once the pivot has reached it, so every register index is static;
//   * the owner lane's 1/sqrt(a_kk) is broadcast by one shuffle, every lane forms its l_ik = c[0] * rk itself;
//   * column k of L (and g_k) is published as one 16-byte aligned LINE of 2D doubles in shared memory -- the lines
//     of node e's pivots overwrite the first 2 D^2 doubles of node e's record, whose contents are in registers by
//     then -- and fetched back with 128-bit loads for the trailing update, which is software-pipelined behind the
//     NEXT pivot's rsqrt except for the one entry that pivot depends on (the owner updates its diagonal from its
//     own l);
//   * back substitution: x_k is broadcast by one shuffle per step and lane i applies it with l_ki read from its own
//     line (row-oriented sweep, no register indexing by lane).
// The pivot loop is unrolled over the D pivots of a node and free of uniform branches, so the scheduler can fill the
// latency of one pivot's chain with the previous pivot's update: a problem's lanes are 8 / 16 / 32 consecutive
// lanes of ONE warp, and a lone warp only issues in order.
template <int D>
__device__ __forceinline__ void bcr_tail_rows(double* __restrict__ nodes, int T, int np,
                                              int S_t, int nc, int* fail) {
  using N = Node<D>;
  constexpr int S = N::kStride, BW = 2 * D - 1, LW = 2 * D;        // LW: doubles per line
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(D * LW <= N::oR, "the lines of a node's pivots must end before its x");
  const int n = nc * D;                                            // scalar rows (<= 32)
  const int lg = (n <= 8) ? 3 : (n <= 16) ? 4 : 5;                 // lanes per problem = 2^lg >= n
  const int LP = 1 << lg;
  const int lane32 = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int row = lane32 & (LP - 1), grp = lane32 >> lg, gbase = lane32 - row, ppw = 32 >> lg;
  const int e = row / D, a = row - e * D;
  const int cb = D * max(e - 1, 0);                                // first column of this lane's window
  const unsigned span = (unsigned)(row - cb);                      // this lane takes part in the pivots cb <= k < row
  for (int pb = warp * ppw; pb < np; pb += nwarps * ppw) {         // warp-uniform trip count (shuffles inside)
    const int p = pb + grp;
    const bool pon = p < np, live = pon && row < n;
    double* pn = nodes + (size_t)(pon ? p : pb) * N::problem_stride(T);   // idle groups never load or store
    const int ec = min(e, nc - 1);
    double* nd = pn + (size_t)bcr_slot(T, S_t * ec) * S;
    const double* ndp = pn + (size_t)bcr_slot(T, S_t * max(ec - 1, 0)) * S;
    double c[BW + 1], ljp[LW];
#pragma unroll
    for (int m = 0; m <= BW; ++m) {
      const int j = cb + m, ej = j / D, cj = j - ej * D;
      double v = 0.0;
      if (live && j <= row) v = (ej == e) ? nd[N::oD + a * D + cj] : ndp[N::oU + cj * D + a];   // D_e[a][cj] | U_{e-1}[cj][a]
      c[m] = v;
      ljp[m] = 0.0;
    }
    double b = live ? nd[N::oR + a] : 0.0;
    double rinv = 0.0, dpiv = c[0], lprev = 0.0;
    int bad = 0;
    __syncwarp();                                                  // every row is in registers: the records are free
#pragma unroll 1
    for (int eb = 0; eb < nc; ++eb) {
      double* rec = pn + (size_t)bcr_slot(T, S_t * eb) * S;
#pragma unroll
      for (int ab = 0; ab < D; ++ab) {
        const int k = eb * D + ab;
        double* line = rec + ab * LW;                              // line[0] = g_k, line[m] = l_{k+m,k}
        const double rk_own = fast_rsqrt(dpiv);                    // meaningful in lane k only
        bad |= (row == k && !(dpiv > 0.0)) ? 1 : 0;
        const double rk = __shfl_sync(FULL, rk_own, gbase + k);
        // trailing update of pivot k - 1 (its line is back by now); nothing happens at k = 0 (lprev = 0, k - 1 < cb)
        if (k - 1 >= cb) {
#pragma unroll
          for (int m = 1; m <= BW; ++m) c[m - 1] = fnma(lprev, ljp[m], c[m]);   // update and slide
          c[BW] = 0.0;
        }
        b = fnma(lprev, ljp[0], b);                                // lprev = 0 in the rows pivot k - 1 does not touch
        const double c0m = ((unsigned)(k - cb) < span) ? c[0] : 0.0;
        const double l = __dmul_rn(c0m, rk);                       // l_{row,k}
        dpiv = fnma(l, l, c[1]);                                   // lane k + 1: its diagonal after this pivot
        const double gk = __dmul_rn(b, rk);
        if (row == k) { b = gk; rinv = rk; }                       // lane k keeps g_k and 1 / l_kk
        const int off = row - k;
        if (pon && (unsigned)off <= (unsigned)BW) line[off] = (off == 0) ? gk : l;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < LW; m += 2) {
          const double2 t = lds2(line + m);
          ljp[m] = t.x;
          ljp[m + 1] = t.y;
        }
        lprev = l;
      }
    }
    if (live && bad) atomicMax(&fail[p], S_t * e + 1);
    const double* myline = nd + a * LW;                            // myline[m] = l_{row+m,row}
    double x = 0.0;
#pragma unroll 2
    for (int k = n - 1; k >= 0; --k) {
      const int off = k - row;
      const double lk = (live && (unsigned)(off - 1) < (unsigned)BW) ? myline[off] : 0.0;   // l_{k,row}
      if (row == k) x = __dmul_rn(b, rinv);
      const double xk = __shfl_sync(FULL, x, gbase + k);
      b = fnma(lk, xk, b);
    }
    if (live) nd[N::oR + a] = x;
  }
}

