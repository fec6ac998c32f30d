// dev micro-benchmark + numerical check of the warp-level 8x8 helpers (register/DMMA versions vs a plain reference)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../dynamic_factor_models_b200/csrc/dfm_kernels_fused.cuh"
using namespace dfm;
__global__ void k(double* out, long long* cyc, double* chk) {
  __shared__ double m[16 * 64]; __shared__ double tmp[16]; __shared__ int bad;
  constexpr int R = 8;
  double* A = m; double* B = m + 64; double* C = m + 128; double* D = m + 192; double* E = m + 256; double* F = m + 320;
  for (int e = threadIdx.x; e < 64; e += 32) { int i = e / 8, j = e % 8; A[e] = (i == j) ? 2.0 + 0.1 * i : 0.01 * (i + j) + 0.003 * i * j; B[e] = (i == j) ? 1.0 : 0.02 * (i - j) + 0.001 * i; }
  if (threadIdx.x == 0) bad = 0;
  __syncwarp();
  // ---- numerical checks (thread 0 computes plain references)
  double ld = w_inv<R>(D, A, tmp, &bad);
  w_gemm<R>(C, A, false, B, false); w_gemm<R>(E, A, true, B, false); w_gemm<R>(F, A, false, B, true);
  w_gemm<R>(m + 384, A, false, D, false);      // A * inv(A) ~ I
  __syncwarp();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
      double s1 = 0, s2 = 0, s3 = 0;
      for (int l = 0; l < 8; ++l) { s1 += A[i * 8 + l] * B[l * 8 + j]; s2 += A[l * 8 + i] * B[l * 8 + j]; s3 += A[i * 8 + l] * B[j * 8 + l]; }
      e1 = fmax(e1, fabs(s1 - C[i * 8 + j])); e2 = fmax(e2, fabs(s2 - E[i * 8 + j])); e3 = fmax(e3, fabs(s3 - F[i * 8 + j]));
      e4 = fmax(e4, fabs((m + 384)[i * 8 + j] - (i == j ? 1.0 : 0.0)));
    }
    chk[0] = e1; chk[1] = e2; chk[2] = e3; chk[3] = e4; chk[4] = ld; chk[5] = bad;
  }
  long long t0 = clock64();
  for (int rep = 0; rep < 20; ++rep) w_gemm<R>(C, A, false, B, false);
  long long t1 = clock64();
  for (int rep = 0; rep < 20; ++rep) ld += w_inv<R>(D, A, tmp, &bad);
  long long t2 = clock64();
  for (int rep = 0; rep < 20; ++rep) w_sym<R>(C);
  long long t3 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / 20; cyc[1] = (t2 - t1) / 20; cyc[2] = (t3 - t2) / 20; }
  out[blockIdx.x * 32 + threadIdx.x] = ld + C[threadIdx.x] + D[threadIdx.x];
}
int main() {
  double* out; long long* cyc; double* chk; cudaMalloc(&out, 8 * 32 * 2048); cudaMalloc(&cyc, 64); cudaMalloc(&chk, 64);
  for (int grid : {1, 444}) {
    k<<<grid, 32>>>(out, cyc, chk); cudaDeviceSynchronize(); k<<<grid, 32>>>(out, cyc, chk); cudaDeviceSynchronize();
    long long h[8]; double c[8]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost); cudaMemcpy(c, chk, 48, cudaMemcpyDeviceToHost);
    printf("grid %4d: w_gemm %lld  w_inv %lld  w_sym %lld cycles | err AB %.1e A'B %.1e AB' %.1e A*inv(A)-I %.1e logdet %.12f bad %.0f\n", grid, h[0], h[1], h[2], c[0], c[1], c[2], c[3], c[4], c[5]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
