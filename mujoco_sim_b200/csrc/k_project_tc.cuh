// k_project_tc.cuh — tensor-core variant of the projection B = J M^-1 (SURVEY.md 8a' row s6, the "true small GEMM";
// north_star: "tensor cores used only for the dense M and J M^-1 J^T blocks").  For one-tree models whose compact
// constraint row spans every dof (a PR2: nv = 49), M^-1 is densified once per environment (k_dense_minv, sparse LtDL
// solves of the unit vectors) and every environment's [rows x nv] . [nv x nv] product runs on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32, M = 128 rows, N = 64, K = 8 per instruction, accumulator in TMEM, read back
// with tcgen05.ld.  fp32 accuracy is recovered by the 3xTF32 split  A B ~= Ah Bh + Ah Bl + Al Bh  (hi = value rounded
// to tf32, lo = the tf32-rounded remainder): 21 MMAs per environment tile.
//
// Operands are K-major, no swizzle ("interleave") in shared memory: a core matrix is 8 rows x 16 bytes stored
// contiguously; core matrices adjacent along M/N are SBO = 128 bytes apart, adjacent along K are LBO = (rows / 8) * 128
// bytes apart (descriptor layout: cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS tree).
#pragma once
#include "k_args.h"
#include "k_common.cuh"

namespace b2 {

namespace tc {
constexpr int TM = 128, TN = 64, TK = 56;   // tile rows, columns (>= nv), padded K (>= nv, multiple of 8)
constexpr int PITCH = 64;                   // floats per row of the environment-major J / M^-1 / B arrays
constexpr uint32_t A_BYTES = TM * TK * 4, B_BYTES = TN * TK * 4;
constexpr uint32_t LBO_A = (TM / 8) * 128, LBO_B = (TN / 8) * 128, SBO = 128;

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  // start address [0,14) | LBO [16,30) | SBO [32,46) (all >> 4) | version = 1 at [46,48) | no swizzle
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32 (bits 4-5 = 1), A = B = tf32 (bits 7-9, 10-12 = 2), K-major both, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// byte offset of element (row, k) of a K-major interleaved tile with `rows` rows
__device__ __forceinline__ uint32_t tile_off(int row, int k, int rows) {
  return (uint32_t)(((k >> 2) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
// 32 consecutive TMEM columns of this thread's lane -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
}  // namespace tc

// Dense M^-1 per environment for the GEMM, from the sparse factor M = L^T D L (L unit, row i non-zero on the ancestors
// of dof i): M^-1 = L^-1 D^-1 L^-T.  L^-1 has the sparsity of L, so it is built in place of the staged factor
// (phase 1, one column per warp at a time: Linv[i][a] = -sum_j L[i][j] Linv[j][a] over the ancestors j of i inside the
// subtree of a), and M^-1[r][b] is the sum over the common ancestors k of r and b of Linv[r][k] / D[k] * Linv[b][k]
// (phase 2, one row per warp at a time).  Cost ~ sum_i depth_i^2 (about 5 k multiply-adds for a PR2) instead of nv full solves (48 k).
// A CTA owns 32 environments (lane = environment); a thread writes row a of its environment's [64][64] matrix.
template <typename T, int ROWS>
__global__ void __launch_bounds__(32 * ROWS) k_dense_minv(const KArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the loops below are driven by the model's integer tables: staged once per CTA (one TMA bulk copy)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int nwords = a.model_words;
  stage_model(blob, a.model, nwords, bar);
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  const int nM = h.nM, nv = h.nv;
  T* LDs = reinterpret_cast<T*>(smem_raw + 16 + (((size_t)nwords * 4 + 15) & ~(size_t)15));          // [nM][32] factor
  T* LIs = LDs + (size_t)nM * 32;                   // [nM][32] L^-1, layout of the factor (entry q of row i: ancestor dof_anc[adr_i + q])
  T* dis = LIs + (size_t)nM * 32;                   // [nv][32]
  T* accs = dis + (size_t)nv * 32;                  // [ROWS][nv][32]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int ntiles = a.nenvp / 32;
  auto adr_of = [&](int i) { return m.i(h.o_dof_Madr, i); };
  auto cnt_of = [&](int i) { return m.i(h.o_dof_Mcnt, i); };
  // position of ancestor `anc` in row i's entry list, or -1 when anc is not an ancestor of (or equal to) i
  auto pos_of = [&](int i, int anc) { const int p = cnt_of(i) - cnt_of(anc); return (p >= 0 && m.i(h.o_dof_anc, adr_of(i) + p) == anc) ? p : -1; };
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * 32 + lane;
    __syncthreads();
    for (int i = wrp; i < nM; i += ROWS) LDs[i * 32 + lane] = a.qLD[(long long)i * S + env];
    for (int i = wrp; i < nv; i += ROWS) dis[i * 32 + lane] = a.qLDiagInv[(long long)i * S + env];
    __syncthreads();
    SArr<T> LD{LDs + lane, 32}, LI{LIs + lane, 32}, dinv{dis + lane, 32};
    // phase 1: columns of L^-1
    for (int c = wrp; c < nv; c += ROWS) {
      LI[adr_of(c)] = 1;
      for (int i = c + 1; i < nv; i++) {
        const int pc = pos_of(i, c);
        if (pc < 0) continue;
        const int ai = adr_of(i);
        T s0 = 0;
        for (int q = 1; q <= pc; q++) {   // ancestors j of i down to c: L[i][j] * Linv[j][c]
          const int j = m.i(h.o_dof_anc, ai + q);
          s0 += LD[ai + q] * LI[adr_of(j) + (cnt_of(j) - cnt_of(c))];
        }
        LI[ai + pc] = -s0;
      }
    }
    __syncthreads();
    // phase 2: rows of M^-1
    SArr<T> acc{accs + (size_t)wrp * nv * 32 + lane, 32};
    T* out = a.minv_em + (long long)env * 64 * 64;
    for (int r = wrp; r < nv; r += ROWS) {
      for (int b = 0; b < nv; b++) acc[b] = 0;
      // M^-1[r][b] = sum over the common ancestors k of r and b (themselves included) of Linv[r][k] / D[k] * Linv[b][k]:
      // for every ancestor k of r, every dof b of k's subtree receives its term
      const int ar = adr_of(r), cr = cnt_of(r);
      for (int q = 0; q < cr; q++) {
        const int k = m.i(h.o_dof_anc, ar + q);
        const T c0 = LI[ar + q] * dinv[k];
        for (int b = k; b < nv; b++) {
          const int pk = pos_of(b, k);
          if (pk >= 0) acc[b] += c0 * LI[adr_of(b) + pk];
        }
      }
      for (int b = 0; b < nv; b++) out[r * 64 + b] = acc[b];
    }
  }
}

// D[env][r][0..63] = sum_k A[env][r][k] * Bm[env][n][k]   (A = J rows, Bm = M^-1 (symmetric), both [.][PITCH] fp32,
// environment-major).  One CTA of 128 threads per environment at a time (persistent over environments, two CTAs per SM
// overlap fill / MMA / epilogue); mtiles row tiles of 128 per environment.  passes = 3: 3xTF32, passes = 1: plain TF32.
__global__ void __launch_bounds__(128) k_project_tc(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D,
                                                    const int* __restrict__ nrows, int nenv, int mtiles, int arows, int passes) {
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  unsigned char* sAh = smem_tc;
  unsigned char* sAl = sAh + A_BYTES;
  unsigned char* sBh = sAl + A_BYTES;
  unsigned char* sBl = sBh + B_BYTES;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_a = smem_u32(&bar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_sh;
  uint32_t phase = 0;
  for (int env = blockIdx.x; env < nenv; env += gridDim.x) {
    const int ne = nrows ? min(nrows[env], arows) : arows;   // rows of this environment that carry a Jacobian
    if (ne <= 0) continue;
    // M^-1 of the environment: hi / lo images (n = output column, k = summation index; symmetric, so row n of M^-1 is
    // column n)
    const float* bsrc = Bm + (size_t)env * TN * PITCH;
    for (int idx = tid; idx < TN * (TK / 4); idx += 128) {
      const int kc = idx / TN, n = idx % TN;
      const float4 v = *reinterpret_cast<const float4*>(bsrc + (size_t)n * PITCH + 4 * kc);
      float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      float4 lo = make_float4(tf32_hi(v.x - hi.x), tf32_hi(v.y - hi.y), tf32_hi(v.z - hi.z), tf32_hi(v.w - hi.w));
      const uint32_t o = tile_off(n, 4 * kc, TN);
      *reinterpret_cast<float4*>(sBh + o) = hi;
      *reinterpret_cast<float4*>(sBl + o) = lo;
    }
    for (int mt = 0; mt < mtiles && mt * TM < ne; mt++) {
      const float* asrc = A + ((size_t)env * arows + (size_t)mt * TM) * PITCH;
      for (int idx = tid; idx < TM * (TK / 4); idx += 128) {
        const int kc = idx / TM, r = idx % TM;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mt * TM + r < ne) v = *reinterpret_cast<const float4*>(asrc + (size_t)r * PITCH + 4 * kc);
        float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        float4 lo = make_float4(tf32_hi(v.x - hi.x), tf32_hi(v.y - hi.y), tf32_hi(v.z - hi.z), tf32_hi(v.w - hi.w));
        const uint32_t o = tile_off(r, 4 * kc, TM);
        *reinterpret_cast<float4*>(sAh + o) = hi;
        *reinterpret_cast<float4*>(sAl + o) = lo;
      }
      // generic-proxy writes -> visible to the tensor core's async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);
        uint32_t acc = 0;
        for (int p = 0; p < passes; p++) {
          const uint32_t aa = p == 2 ? aL : aH, bb = p == 1 ? bL : bH;   // Ah Bh, Ah Bl, Al Bh
          for (int ks = 0; ks < TK / 8; ks++) {
            mma_tf32(tmem, smem_desc(aa + ks * 2 * LBO_A, LBO_A, SBO), smem_desc(bb + ks * 2 * LBO_B, LBO_B, SBO), IDESC, acc);
            acc = 1;
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
      }
      tc::mbar_wait(bar_a, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // epilogue: thread = row (TMEM lane), 64 columns
      uint32_t v[64];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      tc::tmem_ld32(taddr, v);
      tc::tmem_ld32(taddr + 32, v + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = mt * TM + tid;
      if (row < ne) {
        float4* dst = reinterpret_cast<float4*>(D + ((size_t)env * arows + row) * PITCH);
#pragma unroll
        for (int q = 0; q < 16; q++)
          dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();   // TMEM and the A images are free again
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace b2
