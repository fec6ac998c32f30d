/* kalman_em.c -- plain-C port of oracle/kalman_em.py (Kalman filter + RTS smoother + EM for the
 * state-space DFM) and of the ALS sweep of oracle/dfm_ref.py:estimate_factor.
 * ORACLE / CPU-BASELINE INFRASTRUCTURE ONLY: used by tests (validated against the numpy spec) and
 * timed by bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.
 * PARITY UNPINNED for the Kalman part (the reference has no such code, dfm_functions.ipynb:23).
 * Arrays are ROW-major (numpy C order): X[t*N+i], Lam[i*r+a], A[a*k+j].  OpenMP over panels.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG2PI 1.8378770664093454835606594728112

static int chol(double* A, int n) { /* in place lower, row-major n x n; upper zeroed */
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int c = 0; c < j; ++c) d -= A[j * n + c] * A[j * n + c];
    if (!(d > 0)) return 1;
    d = sqrt(d); A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int c = 0; c < j; ++c) s -= A[i * n + c] * A[j * n + c];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) A[i * n + j] = 0;
  return 0;
}
/* B (n x m, row-major) <- L^-1 B */
static void trsm_l(const double* L, int n, double* B, int m) {
  for (int i = 0; i < n; ++i) {
    for (int l = 0; l < i; ++l) { double f = L[i * n + l]; for (int c = 0; c < m; ++c) B[i * m + c] -= f * B[l * m + c]; }
    double inv = 1.0 / L[i * n + i];
    for (int c = 0; c < m; ++c) B[i * m + c] *= inv;
  }
}
/* B <- L^-T B */
static void trsm_lt(const double* L, int n, double* B, int m) {
  for (int i = n - 1; i >= 0; --i) {
    for (int l = i + 1; l < n; ++l) { double f = L[l * n + i]; for (int c = 0; c < m; ++c) B[i * m + c] -= f * B[l * m + c]; }
    double inv = 1.0 / L[i * n + i];
    for (int c = 0; c < m; ++c) B[i * m + c] *= inv;
  }
}
/* C (m x n) = A (m x kk) * B (kk x n) */
static void mm(double* C, const double* A, const double* B, int m, int kk, int n) {
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) C[i * n + j] = 0;
    for (int l = 0; l < kk; ++l) { double a = A[i * kk + l]; for (int j = 0; j < n; ++j) C[i * n + j] += a * B[l * n + j]; }
  }
}
/* C (m x n) = A (m x kk) * B' (B is n x kk) */
static void mmt(double* C, const double* A, const double* B, int m, int kk, int n) {
  for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int l = 0; l < kk; ++l) s += A[i * kk + l] * B[j * kk + l]; C[i * n + j] = s; }
}
static void symm(double* A, int n) { for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) { double v = 0.5 * (A[i * n + j] + A[j * n + i]); A[i * n + j] = v; A[j * n + i] = v; } }

void kem_lyapunov(const double* A, const double* Q, int r, int p, double* P0, int steps) {
  int k = r * p, kk = k * k;
  double* M = calloc(3 * kk, 8); double* T1 = M + kk; double* T2 = T1 + kk;
  for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) M[i * k + j] = A[i * k + j];
  for (int i = r; i < k; ++i) M[i * k + (i - r)] = 1.0;
  memset(P0, 0, kk * 8);
  for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) P0[i * k + j] = Q[i * r + j];
  for (int s = 0; s < steps; ++s) {
    mm(T1, M, P0, k, k, k); mmt(T2, T1, M, k, k, k);
    for (int e = 0; e < kk; ++e) P0[e] += T2[e];
    mm(T1, M, M, k, k, k); memcpy(M, T1, kk * 8);
  }
  symm(P0, k);
  free(M);
}

/* one panel: EM.  Returns status 0 ok / 3 not PD. */
int kem_panel(const double* X, int T, int N, int r, int p, double* Lam, double* R, double* A, double* Q,
              const double* P0, int max_iter, double tol, double* Fout, double* PFout, double* loglik, int* iters) {
  int k = r * p, kk = k * k, rr = r * r, rk = r * k;
  size_t nd = (size_t)N * r + (size_t)T * r + 3 * (size_t)T + 2 * (size_t)T * k + 2 * (size_t)T * kk + (size_t)T * k + (size_t)T * kk +
              12 * (size_t)kk + 8 * (size_t)rr + 4 * (size_t)rk + 8 * (size_t)k + (size_t)N * 4 + (size_t)T * rr + 64;
  double* buf = malloc(nd * 8);
  if (!buf) return 5;
  double* W = buf; double* Bt = W + (size_t)N * r; double* qv = Bt + (size_t)T * r; double* slr = qv + T; double* nobs = slr + T;
  double* zp = nobs + T; double* zf = zp + (size_t)T * k; double* Pp = zf + (size_t)T * k; double* Pf = Pp + (size_t)T * kk;
  double* zs = Pf + (size_t)T * kk; double* Ps = zs + (size_t)T * k;
  double* M = Ps + (size_t)T * kk; double* T1 = M + kk; double* T2 = T1 + kk; double* T3 = T2 + kk; double* S00 = T3 + kk;
  double* Lp = S00 + kk; double* J = Lp + kk; double* D = J + kk; double* Pst = D + kk; double* Pc = Pst + kk; double* Mx = Pc + kk; double* Mx2 = Mx + kk;
  double* C = Mx2 + kk; double* Cf = C + rr; double* L = Cf + rr; double* S = L + rr; double* T4 = S + rr; double* Sff2 = T4 + rr; double* SffA = Sff2 + rr; double* Qn = SffA + rr;
  double* Tm = Qn + rr; double* Wm = Tm + rk; double* S11 = Wm + rk; double* An = S11 + rk;
  double* g = An + rk; double* tv = g + k; double* dv = tv + k;
  double* logR = dv + 6 * k; double* Sxx = logR + N; double* Ti = Sxx + N; double* use = Ti + N;
  double* E = use + N; /* T x rr */
  int has_missing = 0, status = 0, it = 0;
  for (int i = 0; i < N; ++i) use[i] = (Lam[i * r] == Lam[i * r] && R[i] == R[i]) ? 1.0 : 0.0;
  for (int t = 0; t < T && !has_missing; ++t) for (int i = 0; i < N; ++i) if (use[i] && X[(size_t)t * N + i] != X[(size_t)t * N + i]) { has_missing = 1; break; }
  for (it = 0; it < max_iter; ++it) {
    /* ---- E-step preparation */
    memset(M, 0, kk * 8);
    for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) M[i * k + j] = A[i * k + j];
    for (int i = r; i < k; ++i) M[i * k + (i - r)] = 1.0;
    memset(Cf, 0, rr * 8);
    for (int i = 0; i < N; ++i) {
      if (!use[i]) { logR[i] = 0; continue; }
      if (!(R[i] > 0)) status = 3;
      double ri = 1.0 / R[i]; logR[i] = log(R[i]);
      for (int a = 0; a < r; ++a) W[i * r + a] = Lam[i * r + a] * ri;
      for (int a = 0; a < r; ++a) for (int c = 0; c < r; ++c) Cf[a * r + c] += W[i * r + a] * Lam[i * r + c];
    }
    if (status) break;
    for (int t = 0; t < T; ++t) {
      const double* x = X + (size_t)t * N; double* b = Bt + (size_t)t * r;
      for (int a = 0; a < r; ++a) b[a] = 0;
      double q = 0, sl = 0; int n = 0;
      for (int i = 0; i < N; ++i) {
        double v = x[i];
        if (!use[i] || v != v) continue;
        ++n; q += v * v / R[i]; sl += logR[i];
        for (int a = 0; a < r; ++a) b[a] += v * W[i * r + a];
      }
      qv[t] = q; slr[t] = sl; nobs[t] = n;
    }
    /* ---- filter */
    double ll = 0;
    for (int t = 0; t < T; ++t) {
      double* zpt = zp + (size_t)t * k; double* zft = zf + (size_t)t * k; double* Ppt = Pp + (size_t)t * kk; double* Pft = Pf + (size_t)t * kk;
      if (t == 0) { memset(zpt, 0, k * 8); memcpy(Ppt, P0, kk * 8); }
      else {
        const double* zf1 = zf + (size_t)(t - 1) * k; const double* Pf1 = Pf + (size_t)(t - 1) * kk;
        for (int i = 0; i < k; ++i) { double s = 0; for (int l = 0; l < k; ++l) s += M[i * k + l] * zf1[l]; zpt[i] = s; }
        mm(T1, M, Pf1, k, k, k); mmt(Ppt, T1, M, k, k, k);
        for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) Ppt[i * k + j] += Q[i * r + j];
        symm(Ppt, k);
      }
      if (has_missing) {
        memcpy(C, Cf, rr * 8);
        const double* x = X + (size_t)t * N;
        for (int i = 0; i < N; ++i) if (use[i] && x[i] != x[i]) for (int a = 0; a < r; ++a) for (int c = 0; c < r; ++c) C[a * r + c] -= W[i * r + a] * Lam[i * r + c];
      } else memcpy(C, Cf, rr * 8);
      for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) L[i * r + j] = Ppt[i * k + j];
      if (chol(L, r)) { status = 3; break; }
      mm(T4, C, L, r, r, r);
      for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) { double s = (i == j); for (int l = 0; l < r; ++l) s += L[l * r + i] * T4[l * r + j]; S[i * r + j] = s; }
      symm(S, r);
      if (chol(S, r)) { status = 3; break; }
      for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) Tm[i * k + j] = Ppt[i * k + j];
      trsm_l(L, r, Tm, k); memcpy(Wm, Tm, rk * 8); trsm_l(S, r, Wm, k);
      for (int i = 0; i < k; ++i) for (int j = 0; j <= i; ++j) {
        double s = 0;
        for (int a = 0; a < r; ++a) s += Wm[a * k + i] * Wm[a * k + j] - Tm[a * k + i] * Tm[a * k + j];
        double v = Ppt[i * k + j] + s; Pft[i * k + j] = v; Pft[j * k + i] = v;
      }
      const double* b = Bt + (size_t)t * r;
      for (int a = 0; a < r; ++a) { double s = b[a]; for (int c = 0; c < r; ++c) s -= C[a * r + c] * zpt[c]; g[a] = s; }
      for (int i = 0; i < k; ++i) { double s = zpt[i]; for (int a = 0; a < r; ++a) s += Pft[i * k + a] * g[a]; zft[i] = s; }
      double ld = slr[t], quad = qv[t];
      for (int a = 0; a < r; ++a) {
        ld += 2.0 * log(S[a * r + a]);
        quad -= 2.0 * zpt[a] * b[a];
        double cz = 0, pg = 0;
        for (int c = 0; c < r; ++c) { cz += C[a * r + c] * zpt[c]; pg += Pft[a * k + c] * g[c]; }
        quad += zpt[a] * cz - g[a] * pg;
      }
      ll += -0.5 * (nobs[t] * LOG2PI + ld + quad);
    }
    if (status) break;
    loglik[it] = ll;
    /* ---- smoother */
    memcpy(zs + (size_t)(T - 1) * k, zf + (size_t)(T - 1) * k, k * 8);
    memcpy(Ps + (size_t)(T - 1) * kk, Pf + (size_t)(T - 1) * kk, kk * 8);
    memset(S00, 0, kk * 8); memset(S11, 0, rk * 8); memset(Sff2, 0, rr * 8);
    for (int t = T - 2; t >= 0; --t) {
      const double* Pp1 = Pp + (size_t)(t + 1) * kk; const double* Pft = Pf + (size_t)t * kk;
      const double* zs1 = zs + (size_t)(t + 1) * k; const double* Ps1 = Ps + (size_t)(t + 1) * kk;
      double* zst = zs + (size_t)t * k; double* Pstt = Ps + (size_t)t * kk;
      memcpy(Lp, Pp1, kk * 8);
      if (chol(Lp, k)) { status = 3; break; }
      mm(J, M, Pft, k, k, k); trsm_l(Lp, k, J, k); trsm_lt(Lp, k, J, k);   /* J holds J' = Pp^-1 M Pf */
      for (int i = 0; i < k; ++i) dv[i] = zs1[i] - zp[(size_t)(t + 1) * k + i];
      for (int i = 0; i < k; ++i) { double s = zf[(size_t)t * k + i]; for (int l = 0; l < k; ++l) s += J[l * k + i] * dv[l]; zst[i] = s; }
      for (int e = 0; e < kk; ++e) D[e] = Ps1[e] - Pp1[e];
      mm(T1, D, J, k, k, k);                                             /* D J' */
      for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) { double s = Pft[i * k + j]; for (int l = 0; l < k; ++l) s += J[l * k + i] * T1[l * k + j]; Pstt[i * k + j] = s; }
      symm(Pstt, k);
      mm(Pc, Ps1, J, k, k, k);                                           /* Ps(t+1) J' */
      for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) S11[i * k + j] += zs1[i] * zst[j] + Pc[i * k + j];
      for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) S00[i * k + j] += zst[i] * zst[j] + Pstt[i * k + j];
      for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) Sff2[i * r + j] += zs1[i] * zs1[j] + Ps1[i * k + j];
    }
    if (status) break;
    /* ---- M-step */
    memset(SffA, 0, rr * 8);
    for (int t = 0; t < T; ++t) for (int a = 0; a < r; ++a) for (int c = 0; c < r; ++c) {
      double v = zs[(size_t)t * k + a] * zs[(size_t)t * k + c] + Ps[(size_t)t * kk + a * k + c];
      E[(size_t)t * rr + a * r + c] = v; SffA[a * r + c] += v;
    }
    for (int i = 0; i < N; ++i) {
      if (!use[i]) continue;
      double sxf[64], lam[64], Sf[64 * 64];
      double sxx = 0; int ti = 0;
      for (int a = 0; a < r; ++a) sxf[a] = 0;
      memcpy(Sf, SffA, rr * 8);
      for (int t = 0; t < T; ++t) {
        double v = X[(size_t)t * N + i];
        if (v != v) { for (int e = 0; e < rr; ++e) Sf[e] -= E[(size_t)t * rr + e]; continue; }
        ++ti; sxx += v * v;
        for (int a = 0; a < r; ++a) sxf[a] += v * zs[(size_t)t * k + a];
      }
      if (ti == 0) continue;
      double Lc[64 * 64]; memcpy(Lc, Sf, rr * 8);
      if (chol(Lc, r)) { status = 3; break; }
      memcpy(lam, sxf, r * 8); trsm_l(Lc, r, lam, 1); trsm_lt(Lc, r, lam, 1);
      double q1 = 0, q2 = 0;
      for (int a = 0; a < r; ++a) { q1 += lam[a] * sxf[a]; for (int c = 0; c < r; ++c) q2 += lam[a] * lam[c] * Sf[a * r + c]; }
      for (int a = 0; a < r; ++a) Lam[i * r + a] = lam[a];
      R[i] = (sxx - 2 * q1 + q2) / ti;
    }
    if (status) break;
    memcpy(Lp, S00, kk * 8);
    if (chol(Lp, k)) { status = 3; break; }
    for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) T1[j * r + i] = S11[i * k + j];   /* S11' (k x r) */
    trsm_l(Lp, k, T1, r); trsm_lt(Lp, k, T1, r);                                             /* A' */
    for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) An[i * k + j] = T1[j * r + i];
    for (int a = 0; a < r; ++a) for (int c = 0; c < r; ++c) { double s = Sff2[a * r + c]; for (int l = 0; l < k; ++l) s -= An[a * k + l] * S11[c * k + l]; Qn[a * r + c] = s / (T - 1); }
    symm(Qn, r);
    memcpy(A, An, rk * 8); memcpy(Q, Qn, rr * 8);
    if (it >= 1 && fabs(loglik[it] - loglik[it - 1]) <= tol * 0.5 * (fabs(loglik[it]) + fabs(loglik[it - 1]))) { ++it; break; }
  }
  if (it > max_iter) it = max_iter;
  *iters = it;
  if (Fout) for (int t = 0; t < T; ++t) for (int a = 0; a < r; ++a) Fout[(size_t)t * r + a] = zs[(size_t)t * k + a];
  if (PFout) for (int t = 0; t < T; ++t) for (int a = 0; a < r; ++a) for (int c = 0; c < r; ++c) PFout[(size_t)t * rr + a * r + c] = Ps[(size_t)t * kk + a * k + c];
  free(buf);
  return status;
}

/* batch of B panels, OpenMP over panels.  All arrays batched row-major.  P0 may be NULL. */
int kem_batch(const double* X, int B, int T, int N, int r, int p, double* Lam, double* R, double* A, double* Q,
              const double* P0, int max_iter, double tol, double* F, double* loglik, int* iters, int* status, int nthreads) {
  int k = r * p;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    double* P0b = malloc((size_t)k * k * 8);
    if (P0) memcpy(P0b, P0 + (size_t)b * k * k, (size_t)k * k * 8);
    else kem_lyapunov(A + (size_t)b * r * k, Q + (size_t)b * r * r, r, p, P0b, 12);
    for (int j = 0; j < max_iter; ++j) loglik[(size_t)b * max_iter + j] = NAN;
    status[b] = kem_panel(X + (size_t)b * T * N, T, N, r, p, Lam + (size_t)b * N * r, R + (size_t)b * N, A + (size_t)b * r * k,
                          Q + (size_t)b * r * r, P0b, max_iter, tol, F ? F + (size_t)b * T * r : NULL, NULL,
                          loglik + (size_t)b * max_iter, iters + b);
    free(P0b);
  }
  return 0;
}

int kem_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
