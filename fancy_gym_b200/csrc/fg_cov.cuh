// Trajectory covariance of the probabilistic MPs (K5 of SURVEY.md §2.1 / §8 row a20):
//   Sigma_y = Psi Sigma_w Psi^T + reg * max(diag) * I,   Sigma_w = L L^T,   Psi = blockdiag_dof(Bm),  Bm [T, Kc]
// (mp_pytorch ProMP / ProDMP get_traj_pos_cov, App. B.5; no call site inside fancy_gym).  Row / column order is dof-major
// (d*T + t).  Per output element the contraction depth is only Kc (5..6), i.e. 2*Kc flop per 4 stored bytes: the full
// matrix is HBM-STORE bound (4 MB per env at dof 5, T 200), so the work is organised around the store stream:
//
//   out[d1*T + t1, c] = sum_j Bm[t1, j] * G_d1[j, c],      G_d1 = Sigma_w[d1*Kc .. d1*Kc+Kc, :] Psi^T   ([Kc, dof*T])
//
// A block owns (env, d1, a chunk of rows t1): it builds its Sigma_w slab from L and G_d1 in shared memory (~10 % of its
// flops), then every thread keeps the Kc x 4 G values of its four output columns in registers and streams rows:
// Kc*4 FMAs and one coalesced 16-byte store per row.  The diagonal (needed for the regulariser BEFORE anything can be
// written) comes from a cheap pre-pass: diag[d, t] = || Bm[t] L_d ||^2.
#pragma once
#include "fg_device.cuh"
#include "fg_dispatch.h"

namespace fg {

constexpr int kCovThreads = 256;
constexpr int kCovMaxD = 96;      // dof * Kc

// ---- pre-pass: diag and its maxima --------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCovThreads) k_cov_diag(const __grid_constant__ CovArgs a) {
  extern __shared__ float sm[];
  const int D = a.N * a.Kc, T = a.T;
  float* Ls = sm;                 // [D, D]
  float* Bs = Ls + D * D;         // [T, Kc]
  __shared__ float red[kCovThreads / 32];
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < D * D; i += kCovThreads) {
    const int r = i / D, c = i - r * D;
    Ls[i] = (c <= r) ? a.L[b * D * D + i] : 0.f;      // lower triangular factor: the strict upper part is ignored
  }
  for (int i = threadIdx.x; i < T * a.Kc; i += kCovThreads) {
    const int t = i / a.Kc, k = i - t * a.Kc;
    Bs[i] = a.basis[t * a.ld + a.c0 + k];
  }
  __syncthreads();
  float mx = 0.f;
  for (int i = threadIdx.x; i < a.N * T; i += kCovThreads) {
    const int d = i / T, t = i - d * T;
    float s = 0.f;
    for (int m = 0; m < D; ++m) {
      float u = 0.f;
      for (int j = 0; j < a.Kc; ++j) u = fmaf(Bs[t * a.Kc + j], Ls[(d * a.Kc + j) * D + m], u);
      s = fmaf(u, u, s);
    }
    a.diag[b * a.N * T + i] = s;
    mx = fmaxf(mx, s);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kCovThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    a.envmax[b] = mx;
    atomicMax(reinterpret_cast<int*>(a.gmax), __float_as_int(mx));     // mx >= 0: integer order == float order
  }
}

__global__ void __launch_bounds__(kCovThreads) k_cov_std(const __grid_constant__ CovArgs a, const long long B) {
  const long long i = (long long)blockIdx.x * kCovThreads + threadIdx.x;     // over [B, T, N]
  const long long per = (long long)a.T * a.N;
  if (i >= B * per) return;
  const long long b = i / per;
  const int r = (int)(i - b * per), t = r / a.N, d = r - t * a.N;
  const float regterm = a.reg * (a.batch_scope ? *a.gmax : a.envmax[b]);
  a.stdv[i] = sqrtf(a.diag[b * per + (long long)d * a.T + t] + regterm);
}

// ---- full covariance on the CUDA cores ------------------------------------------------------------------------------
template <int KC>    // KC > 0: compile-time basis count; 0: run-time (<= 16)
__global__ void __launch_bounds__(kCovThreads) k_cov_simt(const __grid_constant__ CovArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int Kc = (KC > 0) ? KC : a.Kc;
  const int N = a.N, T = a.T, D = N * Kc, NT = N * T;
  float* Ls = sm;                          // [D, D]
  float* Bs = Ls + ((D * D + 3) & ~3);     // [T, Kc]   (every region starts 16-byte aligned)
  float* Ss = Bs + ((T * Kc + 3) & ~3);    // [Kc, D]   rows d1*Kc.. of Sigma_w
  float* Gs = Ss + ((Kc * D + 3) & ~3);    // [Kc, NT]
  const int KP = (Kc + 1) & ~1;            // basis pairs per row, padded to whole float4
  float2* Bs2 = reinterpret_cast<float2*>(Gs + ((Kc * NT + 3) & ~3));   // [T, KP] (b, b): operand pairs for FFMA2
  const long long b = blockIdx.z;
  const int d1 = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_block, r1 = min(T, r0 + a.rows_per_block);
  const int tid = threadIdx.x;
  for (int i = tid; i < D * D; i += kCovThreads) {
    const int r = i / D, c = i - r * D;
    Ls[i] = (c <= r) ? a.L[b * D * D + i] : 0.f;
  }
  for (int i = tid; i < T * Kc; i += kCovThreads) {
    const int t = i / Kc, k = i - t * Kc;
    Bs[i] = a.basis[t * a.ld + a.c0 + k];
  }
  __syncthreads();
  for (int i = tid; i < Kc * D; i += kCovThreads) {        // Sigma_w slab = L_d1 L^T
    const int j = i / D, c = i - j * D;
    const float* lr = Ls + (d1 * Kc + j) * D;
    const float* lc = Ls + c * D;
    float s = 0.f;
    for (int m = 0; m < D; ++m) s = fmaf(lr[m], lc[m], s);
    Ss[i] = s;
  }
  __syncthreads();
  for (int c = tid; c < NT; c += kCovThreads) {            // G = Sigma_w slab * Psi^T, one output column per thread
    const int d2 = c / T, t2 = c - d2 * T;
    constexpr int KR0 = (KC > 0) ? KC : 16;
    float bcol[KR0];
#pragma unroll
    for (int k = 0; k < KR0; ++k) bcol[k] = (k < Kc) ? Bs[t2 * Kc + k] : 0.f;
    for (int j = 0; j < Kc; ++j) {
      const float* sr = Ss + j * D + d2 * Kc;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KR0; ++k)
        if (k < Kc) s = fmaf(sr[k], bcol[k], s);
      Gs[j * NT + c] = s;
    }
  }
  for (int i = tid; i < (r1 - r0) * KP; i += kCovThreads) {
    const int t = r0 + i / KP, k = i % KP;
    const float v = (k < Kc) ? Bs[t * Kc + k] : 0.f;
    Bs2[t * KP + k] = make_float2(v, v);
  }
  __syncthreads();
  const float regterm = a.reg * (a.batch_scope ? *a.gmax : a.envmax[b]);
  float* out = a.cov + (b * NT + (long long)d1 * T) * NT;
  constexpr int KR = (KC > 0) ? KC : 16;
  if ((NT & 3) == 0) {
    // narrow matrices (dof*T/4 < 256 column quads): the spare threads take interleaved rows
    const int nc4 = NT / 4;
    const int row_lanes = nc4 < kCovThreads ? kCovThreads / nc4 : 1;
    const int rl = (row_lanes > 1) ? tid / nc4 : 0;
    const int c_first = (row_lanes > 1) ? tid - rl * nc4 : tid;
    const int c_step = (row_lanes > 1) ? nc4 : kCovThreads;
    for (int c4 = c_first; c4 < nc4 && rl < row_lanes; c4 += c_step) {
      // the thread's Kc x 4 slab of G as float2 pairs: every row costs Kc*2 packed FMAs (FFMA2, sm_100) + one 16-byte store
      float2 glo[KR], ghi[KR];
#pragma unroll
      for (int k = 0; k < KR; ++k)
        if (k < Kc) {
          const float4 g = *reinterpret_cast<const float4*>(Gs + k * NT + 4 * c4);
          glo[k] = make_float2(g.x, g.y);
          ghi[k] = make_float2(g.z, g.w);
        }
      auto row = [&](int t1) -> float4 {
        float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
        const float4* bp = reinterpret_cast<const float4*>(Bs2 + t1 * KP);
#pragma unroll
        for (int k2 = 0; k2 < (KR + 1) / 2; ++k2)
          if (2 * k2 < Kc) {
            const float4 bb = bp[k2];           // (b_k, b_k, b_k+1, b_k+1)
            lo = __ffma2_rn(make_float2(bb.x, bb.y), glo[2 * k2], lo);
            hi = __ffma2_rn(make_float2(bb.x, bb.y), ghi[2 * k2], hi);
            if (2 * k2 + 1 < Kc) {
              lo = __ffma2_rn(make_float2(bb.z, bb.w), glo[2 * k2 + 1], lo);
              hi = __ffma2_rn(make_float2(bb.z, bb.w), ghi[2 * k2 + 1], hi);
            }
          }
        return make_float4(lo.x, lo.y, hi.x, hi.y);
      };
      for (int t1 = r0 + rl; t1 < r1; t1 += row_lanes)
        __stcs(reinterpret_cast<float4*>(out + (long long)t1 * NT + 4 * c4), row(t1));      // written once, never re-read here
      // regulariser: at most four of this thread's elements lie on the diagonal; those rows are rewritten with it
      const int dcol = 4 * c4 - d1 * T;       // row t1 == dcol + q meets column 4*c4 + q
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t1 = dcol + q;
        if (t1 >= r0 && t1 < r1 && (t1 - r0 - rl) % row_lanes == 0) {
          float4 o = row(t1);
          if (q == 0) o.x += regterm; else if (q == 1) o.y += regterm; else if (q == 2) o.z += regterm; else o.w += regterm;
          __stcs(reinterpret_cast<float4*>(out + (long long)t1 * NT + 4 * c4), o);
        }
      }
    }
  } else {
    for (int c = tid; c < NT; c += kCovThreads)
      for (int t1 = r0; t1 < r1; ++t1) {
        float o = 0.f;
        for (int k = 0; k < Kc; ++k) o = fmaf(Bs[t1 * Kc + k], Gs[k * NT + c], o);
        if (c == d1 * T + t1) o += regterm;
        out[(long long)t1 * NT + c] = o;
      }
  }
}

}  // namespace fg
