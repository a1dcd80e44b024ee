// Trajectory covariance on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), path 2 of fg_traj_cov.
//
// Same factorisation as the CUDA-core kernel (fg_cov.cuh):   out[t1, c] = sum_j Bm[t1, j] * G_d1[j, c]
// as a GEMM with M = 128 time points, N = 256 output columns per accumulator, and a contraction depth of only Kc (5..6).
// Float32 accuracy on TF32 tensor cores comes from the 3xTF32 split, stacked along K:
//     A' = [ A_hi | A_hi | A_lo ]   B' = [ B_hi ; B_lo ; B_hi ]      (x_hi = x with the low 13 mantissa bits cleared — what the
//     tensor core reads anyway — and x_lo = x - x_hi, exact in float32; the dropped A_lo*B_lo term is ~2^-22 relative)
// so K' = 3*Kc padded to a multiple of 8 = 2..3 tcgen05.mma.kind::tf32 instructions (K = 8 each) per 128 x 256 tile.
//
// One CTA per (env, d1, 128-row tile), 9 warps:
//   all      build the Sigma_w slab, G_d1 and the operand tiles in shared memory (canonical K-major, no-swizzle core matrices)
//   warp 8   allocates 512 TMEM columns (two 128 x 256 fp32 accumulators), one lane issues the MMAs for n-tile i+1 while
//   warps 0-7  drain n-tile i (warp w: TMEM lane quarter w % 4, column half w / 4): tcgen05.ld (32 lanes x 32 columns, the
//            next batch is already in flight) -> + regulariser on the diagonal -> 128B-swizzled staging tile in shared
//            memory -> ONE 3-D TMA tensor store per 32 x 32 block (cp.async.bulk.tensor), which also clips rows >= T and
//            columns >= dof*T.
// Synchronisation: mbarriers full[2] (tcgen05.commit -> epilogue) and empty[2] (epilogue -> MMA issuer).
// The matrix is store bound (4 bytes out per 2*Kc flop): the tensor pipe is idle most of the time by construction; the
// point of this path is that the SM's CUDA cores only move data (profiles/README.md has the measured pipe utilisations).
#pragma once
#include <cuda.h>

#include "fg_device.cuh"
#include "fg_dispatch.h"

namespace fg {

constexpr int kUmmaThreads = 288;     // warps 0-7: epilogue (TMEM lane quarter w % 4, column half w / 4), warp 8: MMA issuer
constexpr int kUmmaEpiWarps = 8;
constexpr int kUmmaM = 128, kUmmaN = 256;

namespace umma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE ("interleave"): 8-row x 16-byte core matrices;
// LBO = byte distance of the two 16-byte K chunks of one instruction, SBO = byte distance of consecutive 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  return d;                    // base offset 0, LBO mode 0, layout type 0 = no swizzle
}
// instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {      // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
}  // namespace umma

// shared-memory plan (bytes), computed identically on host and device
struct UmmaSmem {
  int KK, n_pad, off_a, off_b, off_stage, off_misc, off_f32, total;
};
__host__ __device__ inline UmmaSmem umma_smem_plan(int Kc, int N, int T) {
  UmmaSmem s;
  const int D = N * Kc, NT = N * T;
  s.KK = ((3 * Kc + 7) / 8) * 8;
  s.n_pad = ((NT + kUmmaN - 1) / kUmmaN) * kUmmaN;
  s.off_stage = 0;                                        // 8 warps x 2 buffers x 4 KB, 1024-byte aligned (128B swizzle)
  s.off_a = s.off_stage + kUmmaEpiWarps * 2 * 4096;       // A' [KK/4 chunks][16 row groups][8 rows][16 B]
  s.off_b = s.off_a + kUmmaM * s.KK * 4;                  // B' [KK/4 chunks][n_pad/8 row groups][8 rows][16 B]
  s.off_misc = s.off_b + s.n_pad * s.KK * 4;              // mbarriers (4 x 8 B) + TMEM base (4 B)
  s.off_f32 = s.off_misc + 64;                            // Ls [D*D], Bs [T*Kc], Ss [Kc*D], Gs [Kc*NT]
  s.total = s.off_f32 + 4 * (D * D + ((T * Kc + 3) & ~3) + ((Kc * D + 3) & ~3) + Kc * NT);
  return s;
}

__global__ void __launch_bounds__(kUmmaThreads, 1)
k_cov_umma(const __grid_constant__ CovArgs a, const __grid_constant__ CUtensorMap out_map) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using namespace umma;
  const int Kc = a.Kc, N = a.N, T = a.T, D = N * Kc, NT = N * T;
  const UmmaSmem sp = umma_smem_plan(Kc, N, T);
  const int KK = sp.KK, n_pad = sp.n_pad;
  float* stage = reinterpret_cast<float*>(smem + sp.off_stage);
  float* As = reinterpret_cast<float*>(smem + sp.off_a);
  float* Bt = reinterpret_cast<float*>(smem + sp.off_b);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.off_misc);       // full[0], full[1], empty[0], empty[1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sp.off_misc + 32);
  float* Ls = reinterpret_cast<float*>(smem + sp.off_f32);
  float* Bs = Ls + D * D;
  float* Ss = Bs + ((T * Kc + 3) & ~3);
  float* Gs = Ss + ((Kc * D + 3) & ~3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long b = blockIdx.z;
  const int d1 = blockIdx.y, m0 = blockIdx.x * kUmmaM;

  // ---- TMEM allocation + barriers (warp 4) while the other warps start on the operands ----
  if (warp == kUmmaEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);       // full: one tcgen05.commit
      mbar_init(&bars[2], kUmmaEpiWarps); mbar_init(&bars[3], kUmmaEpiWarps);     // empty: the epilogue warps
      fence_barrier_init();
    }
  }

  // ---- operands: Sigma_w slab, G_d1 (as fg_cov.cuh), then the split, K-stacked tiles in the canonical UMMA layout ----
  for (int i = tid; i < D * D; i += kUmmaThreads) {
    const int r = i / D, c = i - r * D;
    Ls[i] = (c <= r) ? a.L[b * D * D + i] : 0.f;
  }
  for (int i = tid; i < T * Kc; i += kUmmaThreads) {
    const int t = i / Kc, k = i - t * Kc;
    Bs[i] = a.basis[t * a.ld + a.c0 + k];
  }
  __syncthreads();
  for (int i = tid; i < Kc * D; i += kUmmaThreads) {
    const int j = i / D, c = i - j * D;
    const float* lr = Ls + (d1 * Kc + j) * D;
    const float* lc = Ls + c * D;
    float s = 0.f;
    for (int m = 0; m < D; ++m) s = fmaf(lr[m], lc[m], s);
    Ss[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < Kc * NT; i += kUmmaThreads) {
    const int j = i / NT, c = i - j * NT, d2 = c / T, t2 = c - d2 * T;
    float s = 0.f;
    for (int k = 0; k < Kc; ++k) s = fmaf(Ss[j * D + d2 * Kc + k], Bs[t2 * Kc + k], s);
    Gs[i] = s;
  }
  __syncthreads();
  // element (row, kk) of a K-major no-swizzle tile with `rows` rows lives at
  //   (kk/4) * (rows*16 B) + (row/8) * 128 B + (row%8) * 16 B + (kk%4) * 4 B
  auto tile_index = [](int rows, int row, int kk) { return (kk >> 2) * (rows * 4) + (row >> 3) * 32 + (row & 7) * 4 + (kk & 3); };
  for (int i = tid; i < kUmmaM * KK; i += kUmmaThreads) {        // A' = [hi | hi | lo] of Bm rows m0 .. m0+127
    const int kk = i / kUmmaM, m = i - kk * kUmmaM, t1 = m0 + m;
    float v = 0.f;
    if (t1 < T && kk < 3 * Kc) {
      const float x = Bs[t1 * Kc + (kk % Kc)], hi = tf32_hi(x);
      v = (kk < 2 * Kc) ? hi : x - hi;
    }
    As[tile_index(kUmmaM, m, kk)] = v;
  }
  for (int i = tid; i < n_pad * KK; i += kUmmaThreads) {         // B' = [hi ; lo ; hi] of G columns
    const int kk = i / n_pad, n = i - kk * n_pad;
    float v = 0.f;
    if (n < NT && kk < 3 * Kc) {
      const float x = Gs[(kk % Kc) * NT + n], hi = tf32_hi(x);
      v = (kk >= Kc && kk < 2 * Kc) ? x - hi : hi;
    }
    Bt[tile_index(n_pad, n, kk)] = v;
  }
  fence_proxy_async();            // generic-proxy writes of the operand tiles -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = n_pad / kUmmaN;
  const float regterm = a.reg * (a.batch_scope ? *a.gmax : a.envmax[b]);

  if (warp == kUmmaEpiWarps) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32(kUmmaM, kUmmaN);
      const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bt);
      const uint32_t a_lbo = kUmmaM * 16, b_lbo = (uint32_t)n_pad * 16;
      for (int i = 0; i < n_tiles; ++i) {
        const int buf = i & 1;
        if (i >= 2) mbar_wait(&bars[2 + buf], ((i >> 1) - 1) & 1);      // the epilogue has drained this accumulator
        tc_fence_after();
        for (int ks = 0; ks < KK / 8; ++ks) {
          const uint64_t ad = smem_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128);
          const uint64_t bd = smem_desc(b_addr + ks * 2 * b_lbo + (uint32_t)i * (kUmmaN / 8) * 128, b_lbo, 128);
          mma_tf32(tmem_base + buf * kUmmaN, ad, bd, idesc, ks > 0);
        }
        mma_commit(&bars[buf]);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> swizzled staging tile -> TMA tensor store =====
    const int quarter = warp & 3, half = warp >> 2;            // TMEM lanes 32*quarter.., columns 128*half.. of each tile
    const int row = quarter * 32 + lane, t1 = m0 + row;
    const bool warp_has_rows = m0 + quarter * 32 < T;
    const int diag_col = d1 * T + t1;
    float* st = stage + warp * 2 * 1024;
    constexpr int kBatches = kUmmaN / 2 / 32;                   // 32-column batches per warp and tile
    int sbuf = 0;
    for (int i = 0; i < n_tiles; ++i) {
      const int buf = i & 1;
      mbar_wait(&bars[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kUmmaN + half * (kUmmaN / 2));
      const int tile_col0 = i * kUmmaN + half * (kUmmaN / 2);
      uint32_t r[2][32];
      tmem_ld32_issue(t_addr, r[0]);
#pragma unroll
      for (int j = 0; j < kBatches; ++j) {
        const int col0 = tile_col0 + j * 32;
        tmem_ld_wait();                                          // batch j has landed in r[j & 1]
        if (j + 1 < kBatches) tmem_ld32_issue(t_addr + (j + 1) * 32, r[(j + 1) & 1]);      // batch j+1 in flight meanwhile
        if (warp_has_rows && col0 < NT) {
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[j & 1][q]);
          const int e = diag_col - col0;
          if (e >= 0 && e < 32) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (q == e) v[q] += regterm;
          }
          bulk_wait_read<1>();                         // the store issued two batches ago has read this staging buffer
          __syncwarp();
          float* sb = st + sbuf * 1024;
#pragma unroll
          for (int q = 0; q < 8; ++q)                  // 128B swizzle: 16-byte chunk q of row r sits at chunk q ^ (r & 7)
            *reinterpret_cast<float4*>(sb + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&out_map, sb, col0, m0 + quarter * 32, (int)(b * N + d1));
            bulk_commit();
          }
          sbuf ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[2 + buf]);
    }
    bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kUmmaEpiWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace fg
