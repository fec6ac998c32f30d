// dfm_kernels_rep.cuh -- replication-level device kernels: the reference has no
// Monte-Carlo / bootstrap / RNG code at all, so these are new components of the named path.
//   k_simulate_panels   frozen synthetic DGP of SURVEY.md 8d, one panel per CTA
//   k_bootstrap_panels  residual bootstrap of a fitted non-parametric model (config C4), one draw per CTA
//   k_percentiles       percentile bands over the replication axis (bitonic sort per statistic)
// Random numbers: counter-based Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (element, element_hi,
// replication id, stream tag) -- a draw is a pure function of (seed, replication id, stream, element), so panel b is
// bit-identical whatever the batch split, launch geometry or GPU count.  oracle/dgp.py restates the stream in numpy.
#pragma once
#include "dfm_common.cuh"

namespace dfm {

enum { RNG_LAM = 0, RNG_AR = 1, RNG_S2 = 2, RNG_ETA = 3, RNG_E = 4, RNG_BIDX = 5, RNG_BETA = 6 };

struct philox4 { uint32_t x, y, z, w; };
__host__ __device__ inline philox4 philox4x32_10(philox4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    philox4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0; n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1; n.w = (uint32_t)p0;
    c = n;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
// two uniforms in (0, 1) with 53 random bits each from one Philox block
__host__ __device__ inline void rng_u2(unsigned long long seed, unsigned long long rep, int stream, unsigned long long ctr, double& u0, double& u1) {
  philox4 c; c.x = (uint32_t)ctr; c.y = (uint32_t)(ctr >> 32); c.z = (uint32_t)rep; c.w = ((uint32_t)(rep >> 32) << 8) | (uint32_t)stream;
  const philox4 o = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  u0 = ((double)(((uint64_t)(o.x >> 5) << 26) | (o.y >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
  u1 = ((double)(((uint64_t)(o.z >> 5) << 26) | (o.w >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
}
__host__ __device__ inline double rng_uniform(unsigned long long seed, unsigned long long rep, int stream, unsigned long long e) {
  double u0, u1; rng_u2(seed, rep, stream, e, u0, u1); return u0;
}
// standard normal number `e` of a stream (Box-Muller; elements 2m and 2m+1 share Philox block m)
__host__ __device__ inline double rng_normal(unsigned long long seed, unsigned long long rep, int stream, unsigned long long e) {
  double u0, u1; rng_u2(seed, rep, stream, e >> 1, u0, u1);
  const double rad = sqrt(-2.0 * log(u0)), ang = 6.283185307179586476925286766559 * u1;
  return (e & 1) ? rad * sin(ang) : rad * cos(ang);
}

// ---------------------------------------------------------------------------------------------------------------
// Frozen DGP of SURVEY.md 8d:  Lam_ij ~ N(0,1);  f_t = diag(a) f_{t-1} + eta_t, a_j ~ U(.2,.8), eta ~ N(0, I), burn-in
// 100;  e_it ~ N(0, s2_i), s2_i ~ U(.5,1.5);  x = Lam f + e, column-standardised (population std, as standardize_data
// dfm_functions.ipynb:501-509).  grid = panels; a warp owns a series at a time (lanes over periods).
// X: [B][N][T] column-major panels;  Ftrue (optional): [B][T*r] column-major true factors (also the kernel's scratch).
__global__ void k_simulate_panels(unsigned long long seed, long long rep0, int T, int N, int r, double* __restrict__ X,
                                  double* __restrict__ Fscr) {
  const int b = DFM_BX;
  const unsigned long long rep = (unsigned long long)(rep0 + b);
  double* F = Fscr + (size_t)b * T * r;
  double* Xb = X + (size_t)b * T * N;
  // factors: r independent AR(1) chains, 100 burn-in periods
  for (int j = DFM_TID; j < r; j += DFM_NT) {
    const double a = 0.2 + 0.6 * rng_uniform(seed, rep, RNG_AR, (unsigned long long)j);
    double f = 0.0;
    for (int t = 0; t < T + 100; ++t) {
      f = a * f + rng_normal(seed, rep, RNG_ETA, (unsigned long long)t * r + j);
      if (t >= 100) F[(t - 100) + (size_t)T * j] = f;
    }
  }
  DFM_SYNC();
#ifdef DFM_EMU
  const int wid = 0, nw = 1, lane = 0, wsz = 1;
#else
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31, wsz = 32;
#endif
  for (int i = wid; i < N; i += nw) {
    double lam[64];                                            // r <= 64 (checked by the host)
    for (int j = 0; j < r; ++j) lam[j] = rng_normal(seed, rep, RNG_LAM, (unsigned long long)i * r + j);
    const double sd = sqrt(0.5 + rng_uniform(seed, rep, RNG_S2, (unsigned long long)i));
    double* x = Xb + (size_t)i * T;
    double s = 0.0;
    for (int t = lane; t < T; t += wsz) {
      double v = sd * rng_normal(seed, rep, RNG_E, (unsigned long long)i * T + t);
      for (int j = 0; j < r; ++j) v += F[t + (size_t)T * j] * lam[j];
      x[t] = v; s += v;
    }
#ifndef DFM_EMU
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncwarp();
#endif
    const double mean = s / T;
    double q = 0.0;
    for (int t = lane; t < T; t += wsz) { const double dv = x[t] - mean; q += dv * dv; }
#ifndef DFM_EMU
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
#endif
    const double inv = 1.0 / sqrt(q / T);
    for (int t = lane; t < T; t += wsz) x[t] = (x[t] - mean) * inv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Residual bootstrap of a fitted non-parametric model (config C4; the reference has no bootstrap: SURVEY.md 8d defines
// it): resample the factor-VAR residuals with replacement and rebuild f* through the VAR (`betahat`, dfm_functions.ipynb
// :463), draw the idiosyncratic AR(L) processes from (uar_coef, uar_ser) (:405-412), x* = Lam f* + u*, and re-impose the
// missing pattern of the original data.  One draw per CTA; thread per series for the AR recursions.
struct BootArgs {
  const double* F0;        // [Tw][r] col-major: fitted factors (first p rows start the recursion)
  const double* resid;     // [nres][r] col-major: VAR residuals to resample
  const double* beta;      // [K][r] col-major, K = 1 + r p: [const; lag 1; ...; lag p]
  const double* lam;       // [ns][r] col-major (NaN row = series not fitted)
  const double* uar_coef;  // [ns][L] col-major
  const double* uar_ser;   // [ns]
  const double* data;      // [Tw][ns] col-major original data (only its NaN pattern is used)
  double* X;               // [B][ns][Tw] out
  int Tw, ns, r, p, L, nres, burn;
  unsigned long long seed; long long rep0;
};
__global__ void k_bootstrap_panels(BootArgs a) {
  DFM_SMEM(sm);
  const int Tw = a.Tw, ns = a.ns, r = a.r, p = a.p, L = a.L, K = 1 + r * p;
  const unsigned long long rep = (unsigned long long)(a.rep0 + DFM_BX);
  double* fs = sm;                       // [Tw][r] row-major bootstrap factors
  double* Xb = a.X + (size_t)DFM_BX * ns * Tw;
  for (int e = DFM_TID; e < p * r; e += DFM_NT) { const int t = e / r, j = e % r; fs[t * r + j] = a.F0[t + (size_t)Tw * j]; }
  DFM_SYNC();
  for (int t = p; t < Tw; ++t) {         // f*_t = [1, f*_{t-1}, ..., f*_{t-p}] beta + resid[idx_t]
    for (int j = DFM_TID; j < r; j += DFM_NT) {
      int idx = (int)(rng_uniform(a.seed, rep, RNG_BIDX, (unsigned long long)(t - p)) * a.nres);
      if (idx >= a.nres) idx = a.nres - 1;
      double v = a.beta[(size_t)K * j] + a.resid[idx + (size_t)a.nres * j];
      for (int l = 1; l <= p; ++l)
        for (int c = 0; c < r; ++c) v += fs[(t - l) * r + c] * a.beta[1 + (l - 1) * r + c + (size_t)K * j];
      fs[t * r + j] = v;
    }
    DFM_SYNC();
  }
  for (int i = DFM_TID; i < ns; i += DFM_NT) {
    double* x = Xb + (size_t)i * Tw;
    bool ok = !is_nan(a.uar_ser[i]);
    for (int c = 0; c < r && ok; ++c) if (is_nan(a.lam[i + (size_t)ns * c])) ok = false;
    if (!ok) { for (int t = 0; t < Tw; ++t) x[t] = DFM_NAN; continue; }
    double u[16];                          // last L values of the idiosyncratic process (L <= 16), u[0] = newest
    for (int l = 0; l < L; ++l) u[l] = 0.0;
    const double ser = a.uar_ser[i];
    for (int t = 0; t < Tw + a.burn; ++t) {
      double acc = ser * rng_normal(a.seed, rep, RNG_BETA, (unsigned long long)i * (Tw + a.burn) + t);
      for (int l = 0; l < L; ++l) acc += a.uar_coef[i + (size_t)ns * l] * u[l];
      for (int l = L - 1; l > 0; --l) u[l] = u[l - 1];
      u[0] = acc;
      if (t >= a.burn) {
        const int tt = t - a.burn;
        double v = acc;
        for (int c = 0; c < r; ++c) v += fs[tt * r + c] * a.lam[i + (size_t)ns * c];
        x[tt] = is_nan(a.data[tt + (size_t)Tw * i]) ? DFM_NAN : v;
      }
    }
  }
}

// Sign convention of re-estimated factors: column j of every panel's F is flipped when it correlates negatively with
// the reference factors F0 (PCA / ALS factors are identified up to sign).  grid (B).  NaN factors stay NaN.
__global__ void k_sign_align(double* __restrict__ Fall, const double* __restrict__ F0, int T, int r) {
  DFM_SMEM(red);
  double* F = Fall + (size_t)DFM_BX * T * r;
  for (int j = 0; j < r; ++j) {
    double s = 0.0;
    for (int t = DFM_TID; t < T; t += DFM_NT) s += F[t + (size_t)T * j] * F0[t + (size_t)T * j];
    s = block_sum(s, red);
    if (s < 0.0) for (int t = DFM_TID; t < T; t += DFM_NT) F[t + (size_t)T * j] = -F[t + (size_t)T * j];
    DFM_SYNC();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Percentile bands over the replication axis (numpy.percentile's default linear interpolation: position q/100 (n-1)).
// recs: [n][d] row-major (replication-major records);  grid = d statistics, one bitonic sort of <= npad values in shared
// memory per statistic; NaN records (failed replications) sort last and are not counted.  out: [nq][d].
__global__ void k_percentiles(const double* __restrict__ recs, int n, int d, const double* __restrict__ q, int nq, int npad,
                              double* __restrict__ out) {
  DFM_SMEM(v);
  const int e = DFM_BX;
  int* cnt = (int*)(v + npad);
  if (DFM_TID == 0) *cnt = 0;
  DFM_SYNC();
  int c = 0;
  for (int i = DFM_TID; i < npad; i += DFM_NT) {
    double x = (i < n) ? recs[(size_t)i * d + e] : DFM_NAN;
    if (is_nan(x)) x = HUGE_VAL; else ++c;
    v[i] = x;
  }
  if (c) atomicAdd(cnt, c);
  DFM_SYNC();
#ifdef DFM_EMU
  for (int i = 1; i < npad; ++i) { double x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
#else
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = v[i], b = v[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { v[i] = b; v[l] = a; }
        }
      }
      __syncthreads();
    }
#endif
  const int m = *cnt;
  for (int k = DFM_TID; k < nq; k += DFM_NT) {
    double r_ = DFM_NAN;
    if (m > 0) {
      const double pos = q[k] / 100.0 * (double)(m - 1);
      int lo = (int)floor(pos); if (lo < 0) lo = 0; if (lo > m - 1) lo = m - 1;
      const int hi = (lo + 1 < m) ? lo + 1 : lo;
      const double fr = pos - (double)lo;
      r_ = v[lo] + fr * (v[hi] - v[lo]);
    }
    out[(size_t)k * d + e] = r_;
  }
}

}  // namespace dfm
