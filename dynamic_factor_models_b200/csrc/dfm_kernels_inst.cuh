// dfm_kernels_inst.cuh -- instability tests of the factor loadings: HAC covariance, Chow (Wald)
// statistic at a given break date and QLR = sup of the Chow statistics over the central break dates, batched over the
// series of a panel.  Reference: dfm_functions.ipynb form_kernel / form_hscrc / hac / regress_hac / compute_chow /
// compute_qlr and the per-series loop of Stock_Watson.ipynb Table 4(a); CPU restatement oracle/dfm_ref.py
// (instability_tests), pinned on the notebook's stored Table 4(a).
// One CTA per series: the rows that survive drop_missing_row([y X]) are compacted into shared memory once; every break
// date then needs  Z'Z = [[S, S2], [S2, S2]] (S2 = sum over the post-break rows), its Cholesky factor, the residuals, the
// Bartlett-weighted autocovariances of g_t = z_t e_t (DMMA tile products over the time axis) and two K x K solves.
#pragma once
#include "dfm_common.cuh"

namespace dfm {

__host__ __device__ inline size_t inst_smem_doubles(int T, int r) {
  const size_t K = 2 * (size_t)r, ldg = em_lds((int)K);
  return (size_t)T * (r + 2) + (size_t)(T + 8) * ldg + 6 * K * K + 2 * (size_t)r * r + 6 * K + 64 + (size_t)T / 2 + 8;
}

// Chow statistic with HAC(q) covariance for the break after `tb` of the Td compacted rows.  All threads call it; the value
// is returned to every thread.  Workspace pointers: see k_instability.
__device__ inline double inst_chow(int Td, int r, int q, int tb, const double* yd, const double* Xd, const double* S, const double* Xy,
                                   double* S2, double* xy2, double* ZtZ, double* Lz, double* V, double* Fm, double* G0, double* W,
                                   double* beta, double* dinv, double* ev, double* G, int ldg, double* red, int* info) {
  const int K = 2 * r;
  // post-break moment sums
  for (int e = DFM_TID; e < r * r + r; e += DFM_NT) {
    double s = 0.0;
    if (e < r * r) { const int a = e % r, c = e / r; for (int t = tb; t < Td; ++t) s += Xd[t + (size_t)Td * a] * Xd[t + (size_t)Td * c]; S2[e] = s; }
    else { const int a = e - r * r; for (int t = tb; t < Td; ++t) s += Xd[t + (size_t)Td * a] * yd[t]; xy2[a] = s; }
  }
  DFM_SYNC();
  for (int e = DFM_TID; e < K * K; e += DFM_NT) {
    const int i = e % K, j = e / K;
    const double v = (i < r && j < r) ? S[i + r * j] : S2[(i % r) + r * (j % r)];
    ZtZ[e] = v; Lz[e] = v;
  }
  for (int i = DFM_TID; i < K; i += DFM_NT) beta[i] = (i < r) ? Xy[i] : xy2[i - r];
  DFM_SYNC();
  bc_chol(Lz, K, K, dinv, info);
  bt_trsm_lower(Lz, K, K, dinv, beta, 1, 1);
  bt_trsm_lowerT(Lz, K, K, dinv, beta, 1, 1);                                 // beta = (Z'Z)^-1 Z'y
  // residuals and g_t = z_t e_t  (rows beyond Td zero: the lagged products read up to q rows past the end)
  for (int t = DFM_TID; t < Td + q + 1; t += DFM_NT) {
    double e_ = 0.0;
    if (t < Td) {
      e_ = yd[t];
      for (int a = 0; a < r; ++a) { const double x = Xd[t + (size_t)Td * a]; e_ -= x * beta[a]; if (t >= tb) e_ -= x * beta[r + a]; }
    }
    for (int a = 0; a < r; ++a) {
      const double x = (t < Td) ? Xd[t + (size_t)Td * a] : 0.0;
      G[(size_t)t * ldg + a] = x * e_; G[(size_t)t * ldg + r + a] = (t >= tb && t < Td) ? x * e_ : 0.0;
    }
  }
  for (int e = DFM_TID; e < K * K; e += DFM_NT) Fm[e] = 0.0;
  DFM_SYNC();
  // Fm = sum_j w_j Gamma_j,  Gamma_j[a][c] = sum_t g_t[a] g_{t+j}[c]  (form_hscrc's first loop);  V = Fm + Fm' - w_0 Gamma_0
  for (int j = 0; j <= q; ++j) {
    const double wj = 1.0 - (double)j / (double)(q + 1);
    wt_gemm(G, 1, ldg, G + (size_t)j * ldg, 1, ldg, K, K, Td - j, [&](int a, int c, double v) {
      Fm[a + K * c] += wj * v;
      if (j == 0) G0[a + K * c] = v;
    });
  }
  DFM_SYNC();
  for (int e = DFM_TID; e < K * K; e += DFM_NT) { const int a = e % K, c = e / K; V[e] = Fm[a + K * c] + Fm[c + K * a] - G0[e]; }
  DFM_SYNC();
  // vbeta = (Z'Z)^-1 V (Z'Z)^-1: row-wise solves, transpose, row-wise solves
  bt_trsm_lower(Lz, K, K, dinv, V, K, K);
  bt_trsm_lowerT(Lz, K, K, dinv, V, K, K);                                    // V <- V (Z'Z)^-1   (row i: V[i,:] (Z'Z)^-1)
  for (int e = DFM_TID; e < K * K; e += DFM_NT) { const int a = e % K, c = e / K; W[a + K * c] = V[c + K * a]; }
  DFM_SYNC();
  bt_trsm_lower(Lz, K, K, dinv, W, K, K);
  bt_trsm_lowerT(Lz, K, K, dinv, W, K, K);                                    // W = vbeta
  // chow = gamma' v1^-1 gamma,  gamma = beta[r:], v1 = vbeta[r:, r:]
  for (int e = DFM_TID; e < r * r; e += DFM_NT) { const int a = e % r, c = e / r; S2[e] = 0.5 * (W[(r + a) + K * (r + c)] + W[(r + c) + K * (r + a)]); }
  for (int a = DFM_TID; a < r; a += DFM_NT) xy2[a] = beta[r + a];
  DFM_SYNC();
  bc_chol(S2, r, r, dinv, info);
  bt_trsm_lower(S2, r, r, dinv, xy2, 1, 1);                                   // y = L^-1 gamma ; chow = y'y
  if (DFM_TID == 0) { double s = 0.0; for (int a = 0; a < r; ++a) s += xy2[a] * xy2[a]; red[40] = s; }
  DFM_SYNC();
  const double out = red[40];
  DFM_SYNC();
  return out;
}

// grid (ns), 256 threads.  data: T x ns column-major (NaN = missing); F: T x r column-major (NaN rows outside the
// estimation window).  chow[i], qlr[i] (HAC(q)) and qlr0[i] (q = 0; may be NULL) are NaN for series with fewer than
// min_obs observations on either side of the break row T_break (counted on y alone, as the notebook does).
__global__ void k_instability(const double* __restrict__ data, const double* __restrict__ Fall, int T, int ns, int r, int q,
                              int T_break, double ccut, int min_obs, double* __restrict__ chow, double* __restrict__ qlr,
                              double* __restrict__ qlr0, int* __restrict__ status) {
  DFM_SMEM(sm);
  const int i = DFM_BX, K = 2 * r, ldg = em_lds(K);
  const double* y = data + (size_t)i * T;
  double* yd = sm;                        // [T]
  double* Xd = yd + T;                    // [T][r] column-major with leading dimension Td (<= T)
  double* ev = Xd + (size_t)T * r;        // [T]
  double* G = ev + T;                     // [T + 8][ldg]
  double* ZtZ = G + (size_t)(T + 8) * ldg; double* Lz = ZtZ + K * K; double* V = Lz + K * K; double* Fm = V + K * K;
  double* G0 = Fm + K * K; double* W = G0 + K * K;
  double* S = W + K * K; double* S2 = S + r * r;
  double* Xy = S2 + r * r; double* xy2 = Xy + K; double* beta = xy2 + K; double* dinv = beta + K;   // K each (dinv: K)
  double* red = dinv + 2 * K;             // 48
  int* info = (int*)(red + 44);
  int* idx = (int*)(red + 48);            // [T] kept rows
  int* cnt = info + 1;                    // [1] Td, [2] n_pre, [3] n_post
  // row flags in parallel (bit 0: y observed, bit 1: y and every factor observed), then one thread compacts them
  for (int t = DFM_TID; t < T; t += DFM_NT) {
    const bool yok = !is_nan(y[t]);
    bool ok = yok;
    for (int a = 0; a < r && ok; ++a) ok = !is_nan(Fall[t + (size_t)T * a]);
    idx[t] = (yok ? 1 : 0) | (ok ? 2 : 0);
  }
  DFM_SYNC();
  if (DFM_TID == 0) {
    int Td = 0, npre = 0, npost = 0;
    for (int t = 0; t < T; ++t) {
      const int f = idx[t];
      if (f & 1) { if (t < T_break) ++npre; else ++npost; }
      if (f & 2) idx[Td++] = t;                       // (Td <= t: in-place compaction)
    }
    info[0] = 0; cnt[0] = Td; cnt[1] = npre; cnt[2] = npost;
  }
  DFM_SYNC();
  const int Td = cnt[0];
  const bool eligible = cnt[1] >= min_obs && cnt[2] >= min_obs && Td > 2 * K + q && T_break > r && T_break < Td - r;
  if (!eligible) {
    if (DFM_TID == 0) { chow[i] = DFM_NAN; qlr[i] = DFM_NAN; if (qlr0) qlr0[i] = DFM_NAN; }
    return;
  }
  for (int e = DFM_TID; e < Td * (r + 1); e += DFM_NT) {
    const int t = e % Td, a = e / Td, ts = idx[t];
    if (a < r) Xd[t + (size_t)Td * a] = Fall[ts + (size_t)T * a]; else yd[t] = y[ts];
  }
  DFM_SYNC();
  for (int e = DFM_TID; e < r * r + r; e += DFM_NT) {
    double s = 0.0;
    if (e < r * r) { const int a = e % r, c = e / r; for (int t = 0; t < Td; ++t) s += Xd[t + (size_t)Td * a] * Xd[t + (size_t)Td * c]; S[e] = s; }
    else { const int a = e - r * r; for (int t = 0; t < Td; ++t) s += Xd[t + (size_t)Td * a] * yd[t]; Xy[a] = s; }
  }
  DFM_SYNC();
  const double c_ = inst_chow(Td, r, q, T_break, yd, Xd, S, Xy, S2, xy2, ZtZ, Lz, V, Fm, G0, W, beta, dinv, ev, G, ldg, red, info);
  const int n1t = (int)floor(ccut * (double)Td), n2t = Td - n1t;
  double lmr = -1e300, lm = -1e300;
  for (int tb = n1t; tb <= n2t; ++tb) {
    const double v = inst_chow(Td, r, q, tb, yd, Xd, S, Xy, S2, xy2, ZtZ, Lz, V, Fm, G0, W, beta, dinv, ev, G, ldg, red, info);
    lmr = (v > lmr) ? v : lmr;
    if (qlr0) { const double v0 = inst_chow(Td, r, 0, tb, yd, Xd, S, Xy, S2, xy2, ZtZ, Lz, V, Fm, G0, W, beta, dinv, ev, G, ldg, red, info); lm = (v0 > lm) ? v0 : lm; }
  }
  if (DFM_TID == 0) {
    chow[i] = c_; qlr[i] = lmr; if (qlr0) qlr0[i] = lm;
    if (info[0] && status) status[i] = 3;
  }
}

// Fitted-value correlations of Table 4(a): for every series, yhat = F bhat and yhat_alt = F_alt bhat_alt with bhat from
// ols_skipmissing(y, F, Balanced()) (rows with a missing y or factor dropped; no intercept), correlation over the rows where
// both fitted values exist.  NaN for series that fail the min_obs rule of the instability loop.  grid (ns), 128 threads;
// shared 2 (r*r + 2 r) + 64 + 48 doubles.
__global__ void k_fit_corr(const double* __restrict__ data, const double* __restrict__ Fall, const double* __restrict__ Falt, int T, int ns,
                           int r, int T_break, int min_obs, double* __restrict__ cor, int* __restrict__ status) {
  DFM_SMEM(sm);
  const int i = DFM_BX;
  const double* y = data + (size_t)i * T;
  double* S = sm; double* Sa = S + r * r; double* c = Sa + r * r; double* ca = c + r; double* dv = ca + r;   // dv: r (shared by the two factorisations)
  double* red = dv + 64;
  int* info = (int*)(red + 44);
  double npre = 0.0, npost = 0.0;
  for (int t = DFM_TID; t < T; t += DFM_NT) if (!is_nan(y[t])) { if (t < T_break) npre += 1.0; else npost += 1.0; }
  npre = block_sum(npre, red); npost = block_sum(npost, red);
  if (DFM_TID == 0) info[0] = 0;
  if (npre < (double)min_obs || npost < (double)min_obs) { if (DFM_TID == 0) cor[i] = DFM_NAN; return; }
  // normal equations of the two regressions (rows with y and all factors of that set present)
  for (int e = DFM_TID; e < 2 * (r * r + r); e += DFM_NT) {
    const int which = e / (r * r + r), ee = e % (r * r + r);
    const double* F = which ? Falt : Fall;
    double s = 0.0;
    for (int t = 0; t < T; ++t) {
      if (is_nan(y[t])) continue;
      bool ok = true;
      for (int a = 0; a < r && ok; ++a) ok = !is_nan(F[t + (size_t)T * a]);
      if (!ok) continue;
      if (ee < r * r) s += F[t + (size_t)T * (ee % r)] * F[t + (size_t)T * (ee / r)];
      else s += F[t + (size_t)T * (ee - r * r)] * y[t];
    }
    if (ee < r * r) (which ? Sa : S)[ee] = s; else (which ? ca : c)[ee - r * r] = s;
  }
  DFM_SYNC();
  bc_chol(S, r, r, dv, info);
  bt_trsm_lower(S, r, r, dv, c, 1, 1); bt_trsm_lowerT(S, r, r, dv, c, 1, 1);
  bc_chol(Sa, r, r, dv, info);
  bt_trsm_lower(Sa, r, r, dv, ca, 1, 1); bt_trsm_lowerT(Sa, r, r, dv, ca, 1, 1);
  // correlation of the fitted values over the rows where both factor sets exist (two passes: means, then moments)
  double n = 0.0, s1 = 0.0, s2 = 0.0;
  for (int t = DFM_TID; t < T; t += DFM_NT) {
    double a1 = 0.0, a2 = 0.0; bool ok = true;
    for (int a = 0; a < r; ++a) { const double f = Fall[t + (size_t)T * a], g = Falt[t + (size_t)T * a]; if (is_nan(f) || is_nan(g)) ok = false; a1 += f * c[a]; a2 += g * ca[a]; }
    if (ok) { n += 1.0; s1 += a1; s2 += a2; }
  }
  n = block_sum(n, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red);
  const double m1 = s1 / n, m2 = s2 / n;
  double v11 = 0.0, v22 = 0.0, v12 = 0.0;
  for (int t = DFM_TID; t < T; t += DFM_NT) {
    double a1 = 0.0, a2 = 0.0; bool ok = true;
    for (int a = 0; a < r; ++a) { const double f = Fall[t + (size_t)T * a], g = Falt[t + (size_t)T * a]; if (is_nan(f) || is_nan(g)) ok = false; a1 += f * c[a]; a2 += g * ca[a]; }
    if (ok) { v11 += (a1 - m1) * (a1 - m1); v22 += (a2 - m2) * (a2 - m2); v12 += (a1 - m1) * (a2 - m2); }
  }
  v11 = block_sum(v11, red); v22 = block_sum(v22, red); v12 = block_sum(v12, red);
  if (DFM_TID == 0) { cor[i] = (n >= 2.0) ? v12 / sqrt(v11 * v22) : DFM_NAN; if (info[0] && status) status[i] = 3; }
}

}  // namespace dfm
