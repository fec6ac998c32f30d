// dfm_kernels_em.cuh -- GENERAL (any r, p with k = r*p <= 48, any N, T, missing data) kernels of the
// state-space EM, row a' of SURVEY.md section 8.  No reference code exists for this path
// (dfm_functions.ipynb:23 is an empty placeholder); the spec is oracle/kalman_em.py.
// One EM iteration = k_em_prep -> k_em_contract -> k_em_filter_smooth -> k_em_mstep_series.
// The fused small-k fast path lives in dfm_kernels_fused.cuh; this file is also the path the host-emulation
// tests exercise.
#pragma once
#include "dfm_common.cuh"

namespace dfm {

struct EmState {           // one per panel
  double ll, ll_prev;
  int iters, done, status, has_missing, conv_pending, pad;
};

#define DFM_LOG2PI 1.8378770664093454835606594728112

// any NaN among the series that are in the model?  grid (N, B)
__global__ void k_em_scan(const double* __restrict__ X, const double* __restrict__ Lam, int T, int N, int r, EmState* st) {
  int i = DFM_BX, b = DFM_BY;
  if (is_nan(Lam[(size_t)b * N * r + i])) return;     // excluded series
  const double* x = X + ((size_t)b * N + i) * T;
  int miss = 0;
  for (int t = DFM_TID; t < T; t += DFM_NT) if (is_nan(x[t])) miss = 1;
  if (miss) st[b].has_missing = 1;     // benign race: all writers store 1
}

__global__ void k_em_state_init(EmState* st) {
  if (DFM_TID != 0) return;
  int b = DFM_BX;
  st[b].ll = 0.0; st[b].ll_prev = 0.0; st[b].iters = 0; st[b].done = 0; st[b].status = 0;
  st[b].has_missing = 0; st[b].conv_pending = 0;
}

// Default prior P0 = sum_j M^j Qt M'^j by doubling (oracle lyapunov_doubling, 12 steps).
// grid (B), one block; shared 3*k*k doubles.
__global__ void k_lyapunov(const double* __restrict__ Aall, const double* __restrict__ Qall, int r, int p,
                           double* __restrict__ P0all, int steps) {
  DFM_SMEM(sm);
  int b = DFM_BX, k = r * p, kk = k * k;
  const double* A = Aall + (size_t)b * r * k; const double* Q = Qall + (size_t)b * r * r;
  double* P = sm; double* Mj = P + kk; double* T1 = Mj + kk;
  for (int e = DFM_TID; e < kk; e += DFM_NT) {
    int i = e % k, j = e / k;
    P[e] = (i < r && j < r) ? Q[i + r * j] : 0.0;
    Mj[e] = (i < r) ? A[i + r * j] : ((j == i - r) ? 1.0 : 0.0);
  }
  DFM_SYNC();
  for (int s = 0; s < steps; ++s) {
    bm_gemm(T1, k, Mj, k, false, P, k, false, k, k, k, 1.0, 0.0);      // T1 = Mj P
    bm_gemm(P, k, T1, k, false, Mj, k, true, k, k, k, 1.0, 1.0);       // P += T1 Mj'
    bm_gemm(T1, k, Mj, k, false, Mj, k, false, k, k, k, 1.0, 0.0);     // T1 = Mj Mj
    bm_copy(Mj, k, T1, k, k, k);
  }
  bm_symmetrize(P, k, k);
  for (int e = DFM_TID; e < kk; e += DFM_NT) P0all[(size_t)b * kk + e] = P[e];
}

// Per-iteration preparation: closes the previous iteration (convergence / iteration count) and builds
// W = Lam / R, logR, C = Lam' R^-1 Lam for the next E-step.  grid (B), one block.
__global__ void k_em_prep(const double* __restrict__ LamAll, const double* __restrict__ Rall, int N, int r, int p,
                          double* __restrict__ Wall, double* __restrict__ logRall, double* __restrict__ Call,
                          double* __restrict__ A, const double* __restrict__ Anew, double* __restrict__ Q,
                          const double* __restrict__ Qnew, EmState* st, int max_iter, int closing) {
  int b = DFM_BX;
  int was_done = st[b].done;
  DFM_SYNC();
  if (closing && !was_done) {
    int rk = r * r * p, rr = r * r;                    // commit the transition M-step of this iteration
    for (int e = DFM_TID; e < rk; e += DFM_NT) A[(size_t)b * rk + e] = Anew[(size_t)b * rk + e];
    for (int e = DFM_TID; e < rr; e += DFM_NT) Q[(size_t)b * rr + e] = Qnew[(size_t)b * rr + e];
    if (DFM_TID == 0) {
      st[b].iters += 1;
      if (st[b].conv_pending || st[b].iters >= max_iter || st[b].status == 3) st[b].done = 1;
    }
    DFM_SYNC();
  }
  if (st[b].done) return;
  const double* Lam = LamAll + (size_t)b * N * r; const double* R = Rall + (size_t)b * N;
  double* W = Wall + (size_t)b * N * r; double* logR = logRall + (size_t)b * N; double* C = Call + (size_t)b * r * r;
  for (int i = DFM_TID; i < N; i += DFM_NT) {
    bool use = !is_nan(Lam[i]) && !is_nan(R[i]);
    double rinv = use ? 1.0 / R[i] : 0.0;
    for (int a = 0; a < r; ++a) W[i + (size_t)N * a] = use ? Lam[i + (size_t)N * a] * rinv : DFM_NAN;
    logR[i] = use ? log(R[i]) : 0.0;
    if (use && !(R[i] > 0.0)) st[b].status = 3;
  }
  DFM_SYNC();
  for (int e = DFM_TID; e < r * r; e += DFM_NT) {
    int a = e % r, c = e / r;
    if (a < c) continue;
    double s = 0.0;
    for (int i = 0; i < N; ++i) { double w = W[i + (size_t)N * a]; if (!is_nan(w)) s += w * Lam[i + (size_t)N * c]; }
    C[a + r * c] = s; C[c + r * a] = s;
  }
}

// E-step contraction: b_t = Lam' R^-1 x_t, q_t = x_t' R^-1 x_t, n_t, sum_obs log R_i and (when data are
// missing) the packed information matrix C_t = C - sum_{i missing} lam_i lam_i'/R_i.
// One THREAD per period t (coalesced column-major reads); per-thread workspace in shared memory.
// grid (ceil(T/NT), B); shared (r + np) * NT doubles.
__global__ void k_em_contract(const double* __restrict__ Xall, const double* __restrict__ LamAll,
                              const double* __restrict__ Wall, const double* __restrict__ Rall,
                              const double* __restrict__ logRall, const double* __restrict__ Call, int T, int N, int r,
                              double* __restrict__ Bt, double* __restrict__ qt, double* __restrict__ slr,
                              int* __restrict__ nt_, double* __restrict__ Ct, const EmState* st) {
  DFM_SMEM(sm);
  int b = DFM_BY;
  if (st[b].done || !st[b].has_missing) return;   // balanced panels: k_em_contract_bal
  int np = r * (r + 1) / 2, nt = DFM_NT;
  double* c = sm + DFM_TID;                       // c[a*nt]
  double* A = sm + (size_t)r * nt + DFM_TID;      // A[e*nt]
  const double* X = Xall + (size_t)b * T * N; const double* Lam = LamAll + (size_t)b * N * r;
  const double* W = Wall + (size_t)b * N * r; const double* R = Rall + (size_t)b * N;
  const double* logR = logRall + (size_t)b * N; const double* C = Call + (size_t)b * r * r;
  int hm = st[b].has_missing;
  for (int t = DFM_BX * nt + DFM_TID; t < T; t += DFM_GX * nt) {
    for (int a = 0; a < r; ++a) c[a * nt] = 0.0;
    if (hm) for (int a = 0; a < r; ++a) for (int cc = 0; cc <= a; ++cc) A[pidx(a, cc) * nt] = C[a + r * cc];
    double q = 0.0, sl = 0.0; int n = 0;
    for (int i = 0; i < N; ++i) {
      double w0 = W[i];
      if (is_nan(w0)) continue;                   // series excluded from the model
      double x = X[t + (size_t)T * i];
      if (!is_nan(x)) {
        ++n; q += x * x / R[i]; sl += logR[i];
        for (int a = 0; a < r; ++a) c[a * nt] += x * W[i + (size_t)N * a];
      } else if (hm) {
        for (int a = 0; a < r; ++a) { double wa = W[i + (size_t)N * a]; for (int cc = 0; cc <= a; ++cc) A[pidx(a, cc) * nt] -= wa * Lam[i + (size_t)N * cc]; }
      }
    }
    for (int a = 0; a < r; ++a) Bt[(size_t)b * T * r + t + (size_t)T * a] = c[a * nt];
    qt[(size_t)b * T + t] = q; slr[(size_t)b * T + t] = sl; nt_[(size_t)b * T + t] = n;
    if (hm) for (int e = 0; e < np; ++e) Ct[(size_t)b * T * np + t + (size_t)T * e] = A[e * nt];
  }
}

// max over the block of two values at once.  red: >= 80 doubles of shared scratch.  All threads get the results.
__device__ __forceinline__ void block_max2(double& a, double& b, double* red) {
#ifndef DFM_EMU
  for (int o = 16; o > 0; o >>= 1) { a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { red[w] = a; red[40 + w] = b; }
  __syncthreads();
  double x = red[0], y = red[40];
  for (int i = 1; i < nw; ++i) { x = fmax(x, red[i]); y = fmax(y, red[40 + i]); }
  __syncthreads();
  a = x; b = y;
#else
  (void)red;
#endif
}

// E-step contraction of a BALANCED panel (no NaN among the series in the model): b_t = W' x_t, q_t, sum log R, n_t; the
// information matrix is the constant C.  32 periods x 8 component groups per block: lane = period (coalesced reads of
// the column-major panel), warp = component group g (components g, g + 8, ...: the W loads are warp-uniform broadcasts).
// r <= 64.  grid (ceil(T / 32), B), 256 threads.
__global__ void k_em_contract_bal(const double* __restrict__ Xall, const double* __restrict__ Wall, const double* __restrict__ Rall,
                                  const double* __restrict__ logRall, int T, int N, int r, double* __restrict__ Bt,
                                  double* __restrict__ qt, double* __restrict__ slr, int* __restrict__ nt_, const EmState* st) {
  DFM_SMEM(part);                                 // [8][32][3] partial (q, sum log R, n) of the component groups
  int b = DFM_BY;
  if (st[b].done || st[b].has_missing) return;
  const double* X = Xall + (size_t)b * T * N; const double* W = Wall + (size_t)b * N * r;
  const double* R = Rall + (size_t)b * N; const double* logR = logRall + (size_t)b * N;
  for (int idx = DFM_TID; idx < 256; idx += DFM_NT) {
    const int tl = idx & 31, g = idx >> 5, t = DFM_BX * 32 + tl;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    double q = 0.0, sl = 0.0, n = 0.0;
    if (t < T)
      for (int i = 0; i < N; ++i) {
        if (is_nan(W[i])) continue;               // series excluded from the model (uniform over the block)
        const double x = X[t + (size_t)T * i];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int a = g + 8 * j; if (a < r) acc[j] += x * W[i + (size_t)N * a]; }
        if ((i & 7) == g) { n += 1.0; q += x * x / R[i]; sl += logR[i]; }       // the scalar sums are split over the groups
      }
    if (t < T) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int a = g + 8 * j; if (a < r) Bt[(size_t)b * T * r + t + (size_t)T * a] = acc[j]; }
    }
    part[(g * 32 + tl) * 3] = q; part[(g * 32 + tl) * 3 + 1] = sl; part[(g * 32 + tl) * 3 + 2] = n;
  }
  DFM_SYNC();
  for (int tl = DFM_TID; tl < 32; tl += DFM_NT) {
    const int t = DFM_BX * 32 + tl;
    if (t >= T) continue;
    double q = 0.0, sl = 0.0, n = 0.0;
    for (int g = 0; g < 8; ++g) { q += part[(g * 32 + tl) * 3]; sl += part[(g * 32 + tl) * 3 + 1]; n += part[(g * 32 + tl) * 3 + 2]; }
    qt[(size_t)b * T + t] = q; slr[(size_t)b * T + t] = sl; nt_[(size_t)b * T + t] = (int)n;
  }
}

// shared-memory footprint (doubles) of k_em_filter_smooth
__host__ __device__ inline size_t em_fs_smem_doubles(int r, int p) {
  size_t k = (size_t)r * p, kk = k * k, rr = (size_t)r * r, rk = (size_t)r * k;
  return 9 * kk + 7 * rr + 3 * rk + 6 * k + 2 * r + 64 + 64;
}

// Kalman filter + RTS smoother + transition M-step for one panel.  grid (B), one block.
// Scratch (global, per panel): zp, zf [T x k]; Pp, Pf [T x k x k].
// Outputs: Fs [T x r], PsF packed [T x np] (smoothed Var f_t), SffAll [r x r] = sum_t E f f',
// Anew [r x k], Qnew [r x r], loglik path.
__global__ void k_em_filter_smooth(const double* __restrict__ Aall, const double* __restrict__ Qall,
                                   const double* __restrict__ P0all, const double* __restrict__ Call,
                                   const double* __restrict__ Bt_, const double* __restrict__ qt_,
                                   const double* __restrict__ slr_, const int* __restrict__ nt__,
                                   const double* __restrict__ Ct_, int T, int r, int p,
                                   double* __restrict__ zp_, double* __restrict__ zf_, double* __restrict__ Pp_,
                                   double* __restrict__ Pf_, double* __restrict__ Fs_, double* __restrict__ PsF_,
                                   double* __restrict__ SffAll_, double* __restrict__ Anew_, double* __restrict__ Qnew_,
                                   double* __restrict__ loglik_, int max_iter, double tol, EmState* st, int* __restrict__ src_) {
  // FROZEN STEPS.  Where the information matrix C_t does not change (a balanced panel: everywhere; missing data: between
  // two changes of the observation pattern) the covariance recursion is data independent and converges to its steady
  // state; once P_{t|t-1} repeats (relative 1e-14) the factorisations, the gain and P_{t|t} of the following periods
  // are the ones already in shared memory, and only the r- and k-sized mean updates remain.  src[t] = the period whose
  // stored covariances period t uses.  The smoother gain J_t depends on (P_{src[t]|src[t]}, P_{src[t+1]|src[t+1]-1})
  // only and is reused likewise; the smoothed covariance recursion freezes the same way (C3: T = 2000 periods, ~25
  // explicit steps in each direction).
  DFM_SMEM(sm);
  int b = DFM_BX;
  if (st[b].done) return;
  int k = r * p, kk = k * k, rr = r * r, rk = r * k, np = r * (r + 1) / 2;
  double* M = sm;            double* Pp = M + kk;      double* Pf = Pp + kk;    double* T1 = Pf + kk;
  double* T2 = T1 + kk;      double* T3 = T2 + kk;     double* Psn = T3 + kk;   double* S00 = Psn + kk;
  double* Ps = S00 + kk;
  double* Q = Ps + kk;       double* C = Q + rr;       double* L = C + rr;      double* S = L + rr;
  double* Sff2 = S + rr;     double* SffA = Sff2 + rr; double* T4 = SffA + rr;
  double* Tm = T4 + rr;      double* Wm = Tm + rk;     double* S11 = Wm + rk;
  double* zp = S11 + rk;     double* zf = zp + k;      double* zsn = zf + k;    double* zs = zsn + k;
  double* dv = zs + k;       double* tv = dv + k;
  double* g = tv + k;        double* bt = g + r;
  double* red = bt + r;      // 40 (+ 80 for block_max2 behind `info`)
  int* info = (int*)(red + 44);                          // [0] Cholesky flag, [2], [3]: "C changed" flags of even / odd periods
  double* red2 = red + 48;   // 80
  int* src = src_ + (size_t)b * T;
  const double* A = Aall + (size_t)b * rk; const double* Qg = Qall + (size_t)b * rr;
  const double* P0 = P0all + (size_t)b * kk; const double* Cg = Call + (size_t)b * rr;
  const double* Bt = Bt_ + (size_t)b * T * r; const double* qt = qt_ + (size_t)b * T;
  const double* slr = slr_ + (size_t)b * T; const int* ntv = nt__ + (size_t)b * T;
  const double* Ct = Ct_ + (size_t)b * T * np;
  double* zpg = zp_ + (size_t)b * T * k; double* zfg = zf_ + (size_t)b * T * k;
  double* Ppg = Pp_ + (size_t)b * T * kk; double* Pfg = Pf_ + (size_t)b * T * kk;
  double* Fs = Fs_ + (size_t)b * T * r; double* PsF = PsF_ + (size_t)b * T * np;
  int hm = st[b].has_missing;
  if (DFM_TID == 0) { info[0] = 0; info[2] = 0; info[3] = 0; }
  for (int e = DFM_TID; e < kk; e += DFM_NT) {
    int i = e % k, j = e / k;
    M[e] = (i < r) ? A[i + r * j] : ((j == i - r) ? 1.0 : 0.0);
    S00[e] = 0.0;
  }
  for (int e = DFM_TID; e < rr; e += DFM_NT) { Q[e] = Qg[e]; C[e] = Cg[e]; Sff2[e] = 0.0; SffA[e] = 0.0; }
  for (int e = DFM_TID; e < rk; e += DFM_NT) S11[e] = 0.0;
  DFM_SYNC();
  double ll = 0.0;
  // ------------------------------------------------------------------ forward: Kalman filter
  int frozen = 0, last_src = 0;          // (uniform over the block)
  double ldS = 0.0;                      // thread 0: 2 sum log diag chol(S) of the last explicit step
  for (int t = 0; t < T; ++t) {
    // information matrix of this period; did it change?  (flag of this period's parity; thread 0 clears the other one)
    if (DFM_TID == 0) info[2 + ((t + 1) & 1)] = 0;
    if (hm) {
      for (int e = DFM_TID; e < rr; e += DFM_NT) {
        int a = e % r, c = e / r;
        double cn = (a >= c) ? Ct[t + (size_t)T * pidx(a, c)] : Ct[t + (size_t)T * pidx(c, a)];
        if (cn != C[e]) info[2 + (t & 1)] = 1;
        C[e] = cn;
      }
    }
    for (int e = DFM_TID; e < r; e += DFM_NT) bt[e] = Bt[t + (size_t)T * e];
    DFM_SYNC();
    const int cchg = (t == 0) ? 1 : info[2 + (t & 1)];
    if (frozen && !cchg) {
      // ---- frozen step: same Pp, L, S, Tm, Wm, Pf as period last_src; only the means move
      for (int i = DFM_TID; i < k; i += DFM_NT) { double s = 0.0; for (int l = 0; l < k; ++l) s += M[i + k * l] * zf[l]; tv[i] = s; }
      DFM_SYNC();
      for (int e = DFM_TID; e < k; e += DFM_NT) zp[e] = tv[e];
      DFM_SYNC();
      for (int a = DFM_TID; a < r; a += DFM_NT) { double s = bt[a]; for (int c = 0; c < r; ++c) s -= C[a + r * c] * zp[c]; g[a] = s; }
      DFM_SYNC();
      for (int i = DFM_TID; i < k; i += DFM_NT) { double s = zp[i]; for (int a = 0; a < r; ++a) s += Pf[i + k * a] * g[a]; zf[i] = s; }
      DFM_SYNC();
      if (DFM_TID == 0) {
        // quad = q - 2 zp'b + zp'C zp - g'Pff g  with  C zp = b - g  and  Pff g = (zf - zp)[0:r]:  O(r) instead of O(r^2)
        double quad = qt[t];
        for (int a = 0; a < r; ++a) quad += -2.0 * zp[a] * bt[a] + zp[a] * (bt[a] - g[a]) - g[a] * (zf[a] - zp[a]);
        ll += -0.5 * ((double)ntv[t] * DFM_LOG2PI + slr[t] + ldS + quad);
        src[t] = last_src;
      }
      for (int e = DFM_TID; e < k; e += DFM_NT) { zpg[(size_t)t * k + e] = zp[e]; zfg[(size_t)t * k + e] = zf[e]; }
      DFM_SYNC();
      continue;
    }
    frozen = 0;
    if (t == 0) {
      for (int e = DFM_TID; e < kk; e += DFM_NT) Pp[e] = P0[e];
      for (int e = DFM_TID; e < k; e += DFM_NT) zp[e] = 0.0;
      DFM_SYNC();
    } else {
      for (int e = DFM_TID; e < kk; e += DFM_NT) T3[e] = Pp[e];                 // previous P_{t|t-1}: freeze test below
      bm_gemm(T1, k, M, k, false, Pf, k, false, k, k, k, 1.0, 0.0);            // M Pf
      bm_gemm(Pp, k, T1, k, false, M, k, true, k, k, k, 1.0, 0.0);             // (M Pf) M'
      for (int e = DFM_TID; e < rr; e += DFM_NT) { int i = e % r, j = e / r; Pp[i + k * j] += Q[e]; }
      for (int i = DFM_TID; i < k; i += DFM_NT) { double s = 0.0; for (int l = 0; l < k; ++l) s += M[i + k * l] * zf[l]; tv[i] = s; }
      DFM_SYNC();
      bm_symmetrize(Pp, k, k);
      for (int e = DFM_TID; e < k; e += DFM_NT) zp[e] = tv[e];
      DFM_SYNC();
    }
    for (int e = DFM_TID; e < rr; e += DFM_NT) { int i = e % r, j = e / r; L[e] = Pp[i + k * j]; }
    DFM_SYNC();
    bm_chol(L, r, r, info);                                                     // Pff = L L'
    bm_gemm(T4, r, C, r, false, L, r, false, r, r, r, 1.0, 0.0);               // C L
    bm_gemm(S, r, L, r, true, T4, r, false, r, r, r, 1.0, 0.0);                // L' C L
    for (int e = DFM_TID; e < r; e += DFM_NT) S[e + r * e] += 1.0;
    DFM_SYNC();
    bm_symmetrize(S, r, r);
    bm_chol(S, r, r, info);                                                     // S = Ls Ls'
    for (int e = DFM_TID; e < rk; e += DFM_NT) { int i = e % r, j = e / r; Tm[e] = Pp[i + k * j]; }
    DFM_SYNC();
    bm_trsm_lower(L, r, r, Tm, r, k);                                           // Tm = L^-1 Pp[0:r,:]
    bm_copy(Wm, r, Tm, r, r, k);
    bm_trsm_lower(S, r, r, Wm, r, k);                                           // Wm = Ls^-1 Tm
    for (int e = DFM_TID; e < kk; e += DFM_NT) {                                // Pf = Pp - Tm'Tm + Wm'Wm
      int i = e % k, j = e / k;
      if (i < j) continue;
      double s = 0.0;
      for (int a = 0; a < r; ++a) s += Wm[a + r * i] * Wm[a + r * j] - Tm[a + r * i] * Tm[a + r * j];
      double v = Pp[i + k * j] + s;
      Pf[i + k * j] = v; Pf[j + k * i] = v;
    }
    for (int a = DFM_TID; a < r; a += DFM_NT) { double s = bt[a]; for (int c = 0; c < r; ++c) s -= C[a + r * c] * zp[c]; g[a] = s; }
    DFM_SYNC();
    for (int i = DFM_TID; i < k; i += DFM_NT) { double s = zp[i]; for (int a = 0; a < r; ++a) s += Pf[i + k * a] * g[a]; zf[i] = s; }
    if (DFM_TID == 0) {
      ldS = 0.0;
      for (int a = 0; a < r; ++a) ldS += 2.0 * log(S[a + r * a]);
      double quad = qt[t];
      for (int a = 0; a < r; ++a) {
        quad -= 2.0 * zp[a] * bt[a];
        double cz = 0.0, pg = 0.0;
        for (int c = 0; c < r; ++c) { cz += C[a + r * c] * zp[c]; pg += Pf[a + k * c] * g[c]; }
        quad += zp[a] * cz - g[a] * pg;
      }
      ll += -0.5 * ((double)ntv[t] * DFM_LOG2PI + slr[t] + ldS + quad);
      src[t] = t;
    }
    DFM_SYNC();
    for (int e = DFM_TID; e < kk; e += DFM_NT) { Ppg[(size_t)t * kk + e] = Pp[e]; Pfg[(size_t)t * kk + e] = Pf[e]; }
    for (int e = DFM_TID; e < k; e += DFM_NT) { zpg[(size_t)t * k + e] = zp[e]; zfg[(size_t)t * k + e] = zf[e]; }
    last_src = t;
    if (t >= 1 && !cchg) {                                                      // steady state reached?  (T3 = previous Pp)
      double dm = 0.0, pm = 0.0;
      for (int e = DFM_TID; e < kk; e += DFM_NT) { dm = fmax(dm, fabs(Pp[e] - T3[e])); pm = fmax(pm, fabs(Pp[e])); }
      block_max2(dm, pm, red2);
      frozen = (dm <= 1e-14 * pm) ? 1 : 0;
    }
    DFM_SYNC();
  }
#ifdef DFM_EMU
  if (getenv("DFM_DEBUG_FREEZE")) { int nf = 0; for (int t = 0; t < T; ++t) nf += (src[t] != t); printf("[freeze] b=%d T=%d k=%d frozen forward steps %d\n", b, T, k, nf); }
#endif
  // ------------------------------------------------------------------ backward: RTS smoother
  for (int e = DFM_TID; e < kk; e += DFM_NT) Psn[e] = Pf[e];
  for (int e = DFM_TID; e < k; e += DFM_NT) zsn[e] = zf[e];
  DFM_SYNC();
  for (int e = DFM_TID; e < r; e += DFM_NT) Fs[(T - 1) + (size_t)T * e] = zsn[e];
  for (int e = DFM_TID; e < rr; e += DFM_NT) {
    int a = e % r, c = e / r;
    if (a >= c) PsF[(T - 1) + (size_t)T * pidx(a, c)] = Psn[a + k * c];
    SffA[e] = zsn[a] * zsn[c] + Psn[a + k * c];
  }
  DFM_SYNC();
  int jpp = -1, jpf = -1, ps_frozen = 0;           // periods whose covariances built the gain in T3; smoothed covariance frozen?
  for (int t = T - 2; t >= 0; --t) {
    const int sp = src[t + 1], sf = src[t];
    const bool newJ = (sp != jpp) || (sf != jpf);
    if (newJ) ps_frozen = 0;
    for (int e = DFM_TID; e < k; e += DFM_NT) { zp[e] = zpg[(size_t)(t + 1) * k + e]; zf[e] = zfg[(size_t)t * k + e]; }
    if (!ps_frozen) for (int e = DFM_TID; e < kk; e += DFM_NT) { T1[e] = Ppg[(size_t)sp * kk + e]; Pf[e] = Pfg[(size_t)sf * kk + e]; }
    DFM_SYNC();
    if (newJ) {
      bm_copy(T2, k, T1, k, k, k);
      bm_chol(T2, k, k, info);                                                  // Pp(t+1) = Lp Lp'
      bm_gemm(T3, k, M, k, false, Pf, k, false, k, k, k, 1.0, 0.0);            // M Pf(t)
      bm_trsm_lower(T2, k, k, T3, k, k);
      bm_trsm_lowerT(T2, k, k, T3, k, k);                                       // T3 = J' = Pp^-1 M Pf
      jpp = sp; jpf = sf;
    }
    for (int e = DFM_TID; e < k; e += DFM_NT) dv[e] = zsn[e] - zp[e];
    if (!ps_frozen) for (int e = DFM_TID; e < kk; e += DFM_NT) T1[e] = Psn[e] - T1[e];     // D = Ps(t+1) - Pp(t+1)
    DFM_SYNC();
    for (int i = DFM_TID; i < k; i += DFM_NT) { double s = zf[i]; for (int l = 0; l < k; ++l) s += T3[l + k * i] * dv[l]; zs[i] = s; }
    if (!ps_frozen) {
      bm_gemm(T2, k, T1, k, false, T3, k, false, k, k, k, 1.0, 0.0);           // D J'
      bm_copy(Ps, k, Pf, k, k, k);
      bm_gemm(Ps, k, T3, k, true, T2, k, false, k, k, k, 1.0, 1.0);            // Pf + J D J'
      bm_symmetrize(Ps, k, k);
      bm_gemm(Tm, r, Psn, k, false, T3, k, false, r, k, k, 1.0, 0.0);          // Pc[0:r,:] = Ps(t+1)[0:r,:] J'
      if (!newJ) {                                                              // smoothed covariance at its steady state?
        double dm = 0.0, pm = 0.0;
        for (int e = DFM_TID; e < kk; e += DFM_NT) { dm = fmax(dm, fabs(Ps[e] - Psn[e])); pm = fmax(pm, fabs(Ps[e])); }
        block_max2(dm, pm, red2);
        ps_frozen = (dm <= 1e-14 * pm) ? 1 : 0;      // from the next period on: Ps = Psn, Tm as they are
      }
    } else DFM_SYNC();
    for (int e = DFM_TID; e < rk; e += DFM_NT) { int i = e % r, j = e / r; S11[e] += zsn[i] * zs[j] + Tm[e]; }
    for (int e = DFM_TID; e < kk; e += DFM_NT) { int i = e % k, j = e / k; S00[e] += zs[i] * zs[j] + Ps[e]; }
    for (int e = DFM_TID; e < rr; e += DFM_NT) {
      int a = e % r, c = e / r;
      Sff2[e] += zsn[a] * zsn[c] + Psn[a + k * c];
      SffA[e] += zs[a] * zs[c] + Ps[a + k * c];
      if (a >= c) PsF[t + (size_t)T * pidx(a, c)] = Ps[a + k * c];
    }
    for (int e = DFM_TID; e < r; e += DFM_NT) Fs[t + (size_t)T * e] = zs[e];
    DFM_SYNC();
    for (int e = DFM_TID; e < kk; e += DFM_NT) Psn[e] = Ps[e];
    for (int e = DFM_TID; e < k; e += DFM_NT) zsn[e] = zs[e];
    DFM_SYNC();
  }
  // ------------------------------------------------------------------ transition M-step
  // A = S11 S00^-1 ;  Q = (Sff2 - A S11') / (T-1)
  bm_copy(T2, k, S00, k, k, k);
  bm_chol(T2, k, k, info);
  for (int e = DFM_TID; e < rk; e += DFM_NT) { int i = e % r, j = e / r; T3[j + k * i] = S11[e]; }   // S11' (k x r)
  DFM_SYNC();
  bm_trsm_lower(T2, k, k, T3, k, r);
  bm_trsm_lowerT(T2, k, k, T3, k, r);                                           // T3 = A' (k x r)
  for (int e = DFM_TID; e < rr; e += DFM_NT) {
    int a = e % r, c = e / r;
    double s = Sff2[e];
    for (int l = 0; l < k; ++l) s -= T3[l + k * a] * S11[c + r * l];           // (A S11')[a,c]
    T4[e] = s / (double)(T - 1);
  }
  DFM_SYNC();
  bm_symmetrize(T4, r, r);
  for (int e = DFM_TID; e < rk; e += DFM_NT) { int i = e % r, j = e / r; Anew_[(size_t)b * rk + e] = T3[j + k * i]; }
  for (int e = DFM_TID; e < rr; e += DFM_NT) { Qnew_[(size_t)b * rr + e] = T4[e]; SffAll_[(size_t)b * rr + e] = SffA[e]; }
  if (DFM_TID == 0) {
    int it = st[b].iters;
    loglik_[(size_t)b * max_iter + it] = ll;
    st[b].ll_prev = st[b].ll; st[b].ll = ll;
    if (it >= 1 && fabs(ll - st[b].ll_prev) <= tol * 0.5 * (fabs(ll) + fabs(st[b].ll_prev))) st[b].conv_pending = 1;
    if (*info || !(ll == ll)) st[b].status = 3;
  }
}

// Measurement M-step: one block per series.  S_ff^(i) = SffAll - sum_{t missing} E_t.
// Lam_i = S_ff^(i)^-1 S_xf^(i);  R_i = (S_xx - 2 lam'S_xf + lam' S_ff lam) / T_i.   grid (N, B).
__global__ void k_em_mstep_series(const double* __restrict__ Xall, const double* __restrict__ Fs_,
                                  const double* __restrict__ PsF_, const double* __restrict__ SffAll_, int T, int N,
                                  int r, double* __restrict__ LamAll, double* __restrict__ Rall, EmState* st) {
  DFM_SMEM(sm);
  int i = DFM_BX, b = DFM_BY;
  if (st[b].done) return;
  int np = r * (r + 1) / 2;
  double* Lam = LamAll + (size_t)b * N * r; double* R = Rall + (size_t)b * N;
  if (is_nan(Lam[i]) || is_nan(R[i])) return;         // excluded series stay excluded
  const double* x = Xall + ((size_t)b * N + i) * T;
  const double* Fs = Fs_ + (size_t)b * T * r; const double* PsF = PsF_ + (size_t)b * T * np;
  const double* SffA = SffAll_ + (size_t)b * r * r;
  double* A = sm; double* c = A + np; double* sc = c + r; double* A0 = sc + 4;
  int hm = st[b].has_missing;
  int nwork = np + r + 2;
  for (int e = DFM_TID; e < nwork; e += DFM_NT) {
    double s = 0.0;
    if (e < np) {
      int a = 0; while ((a + 1) * (a + 2) / 2 <= e) ++a;
      int cc = e - a * (a + 1) / 2;
      if (hm) for (int t = 0; t < T; ++t) if (is_nan(x[t])) s += Fs[t + (size_t)T * a] * Fs[t + (size_t)T * cc] + PsF[t + (size_t)T * e];
      A[e] = SffA[a + r * cc] - s; A0[e] = A[e];
    } else if (e < np + r) {
      const double* fa = Fs + (size_t)T * (e - np);
      for (int t = 0; t < T; ++t) { double v = x[t]; if (!is_nan(v)) s += v * fa[t]; }
      c[e - np] = s;
    } else if (e == np + r) { for (int t = 0; t < T; ++t) if (!is_nan(x[t])) s += 1.0; sc[0] = s; }
    else { for (int t = 0; t < T; ++t) { double v = x[t]; if (!is_nan(v)) s += v * v; } sc[1] = s; }
  }
  DFM_SYNC();
  if (DFM_TID != 0) return;
  if (sc[0] < 1.0) return;
  double sxf[64];
  for (int a = 0; a < r; ++a) sxf[a] = c[a];
  if (chol_solve_packed(A, c, r, 1)) { st[b].status = 3; return; }
  double q1 = 0.0, q2 = 0.0;
  for (int a = 0; a < r; ++a) {
    q1 += c[a] * sxf[a];
    for (int cc = 0; cc < r; ++cc) q2 += c[a] * c[cc] * ((a >= cc) ? A0[pidx(a, cc)] : A0[pidx(cc, a)]);
  }
  for (int a = 0; a < r; ++a) Lam[i + (size_t)N * a] = c[a];
  R[i] = (sc[1] - 2.0 * q1 + q2) / sc[0];
}

// unpack PsF (packed, [T x np]) to r x r x T column-major for the API output.
__global__ void k_unpack_psf(const double* __restrict__ PsF_, int T, int r, double* __restrict__ out) {
  int b = DFM_BY, np = r * (r + 1) / 2;
  const double* PsF = PsF_ + (size_t)b * T * np;
  double* o = out + (size_t)b * T * r * r;
  for (long long e = (long long)DFM_BX * DFM_NT + DFM_TID; e < (long long)T * r * r; e += (long long)DFM_GX * DFM_NT) {
    int a = (int)(e % r), c = (int)((e / r) % r), t = (int)(e / ((long long)r * r));
    o[e] = (a >= c) ? PsF[t + (size_t)T * pidx(a, c)] : PsF[t + (size_t)T * pidx(c, a)];
  }
}

__global__ void k_em_count_active(const EmState* st, int B, int* out) {
  DFM_SMEM(sm);
  double n = 0.0;
  for (int b = DFM_TID; b < B; b += DFM_NT) n += st[b].done ? 0.0 : 1.0;
  n = block_sum(n, sm);
  if (DFM_TID == 0) *out = (int)n;
}

__global__ void k_em_collect(const EmState* st, int* iters, int* status) {
  if (DFM_TID != 0) return;
  int b = DFM_BX;
  iters[b] = st[b].iters; status[b] = st[b].status;
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long e = (long long)DFM_BX * DFM_NT + DFM_TID; e < n; e += (long long)DFM_GX * DFM_NT) p[e] = v;
}

}  // namespace dfm
