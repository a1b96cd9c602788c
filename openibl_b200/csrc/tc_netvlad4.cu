// Fused NetVLAD on clusters of four SMs (reference ibl/models/netvlad.py:44-61).
//
// The one-SM kernel (tc_netvlad.cu) cannot keep a 128-pixel x 512-channel tile (256 KiB as bf16 hi/lo planes)
// in shared memory, so it streams every tile twice (once per contraction) plus the 128 KiB of W per tile:
// 640 KiB of L2->SM traffic per 128 pixels, which is what bounds it.  Here the CHANNELS are split over the
// four CTAs of a cluster: CTA r owns channels [128 r, 128 r + 128).
//
//   per CTA, resident:   W[:, own channels] as [W_hi ; W_lo] rows (32 KiB)
//   per tile, once:      x[128 px, own channels] hi/lo (64 KiB, TMA, double buffered)
//   GEMM 1 (K-major)     Z_r[px, k]   = x_r . W_r^T            partial logits over the CTA's channels
//   exchange (DSMEM)     pixel group q (32 px) of every CTA's Z_r goes to CTA q, which adds the four partials,
//                        does the softmax for its 32 pixels and broadcasts a' = a/|x| (bf16 hi/lo, MN-major
//                        SW128 operand rows) into the a' buffer of all four CTAs
//   GEMM 2 (MN-major)    V_r[c, k]   += x_r^T . a'             the same shared-memory tile, read as MN-major
//
// Both contractions are bf16x3-grade: the B operand is the concatenation [hi | lo] (N = 128), multiplied by the
// hi and by the lo plane of A (the extra lo.lo term is harmless), so every MMA is M = 128, N = 128.
// V (128 channels x [k | k] fp32 = 128 TMEM columns) stays resident across the unit's tiles; logits are double
// buffered in TMEM so GEMM 1 of tile i+1 overlaps the exchange/softmax of tile i.
//
// All cross-CTA traffic is st.async (remote shared-memory store that completes transaction bytes on the
// DESTINATION CTA's mbarrier, like a TMA write): no fences, and the tensor core may read the a' rows as soon
// as the barrier flips.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 "senders" (logits
// TMEM -> owner CTA; V read-out at the end of a unit), warps 6-9 "softmax" (4 threads per pixel).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ibl {

using namespace tc;

struct Nv4Args {
  int B, S, G, T;                 // images, pixels per image, units per image, 128-pixel tiles per image
  int ssq_parts;
  const float* ssq;               // [ssq_parts][B*S]
  int normalize_input;
  float* part;                    // [B*G][64][512]
  float* asum_part;               // [B*G*4][64]   one partial per CTA of the cluster
  unsigned long long* dbg;        // optional [gridDim][64] globaltimer stamps (IBL_NV_DEBUG=1)
};

constexpr int NV4_W = 0;                       // 2 chunks x ([W_hi 64 rows][W_lo 64 rows]) x 128 B
constexpr int NV4_X = 32768;                   // 2 buffers x (hi c0 | hi c1 | lo c0 | lo c1), 16 KiB each
constexpr int NV4_A = NV4_X + 2 * 65536;       // a_hi | a_lo, [128 px][64 k] MN-major, 16 KiB each
constexpr int NV4_EX = NV4_A + 32768;          // 4 sender slots of partial logits
constexpr int NV4_EX_SLOT = 8192;          // [32 px][64 k] fp32
constexpr int NV4_BARS = NV4_EX + 4 * NV4_EX_SLOT;
constexpr int NV4_ASUM = NV4_BARS + 256;       // [4 warps][64] fp32
constexpr int NV4_SMEM = NV4_ASUM + 1024;

__device__ __forceinline__ uint64_t nv4_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// remote (or local) shared-memory store through the async proxy; completes 16 tx bytes on `mbar` (an mbarrier
// of the SAME destination CTA, shared::cluster address)
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar)
               : "memory");
}

__device__ __forceinline__ unsigned long long nv4_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define NV4_STAMP(slot)                                                                             \
  do {                                                                                              \
    if (a.dbg && q == 0 && lane == 0 && (slot) < 128) a.dbg[blockIdx.x * 128 + (slot)] = nv4_now();   \
  } while (0)

struct Nv4Iter {   // the cluster's tiles in processing order: units cid, cid + n_clusters, ...; tiles g, g + G, ...
  int u, t, useq, G, T, n_units, stride;
  __device__ void init(int cid, int n_clusters, int G_, int T_, int n_units_) {
    G = G_; T = T_; n_units = n_units_; stride = n_clusters;
    u = cid; useq = 0; t = u % G;
  }
  __device__ bool valid() const { return u < n_units; }
  __device__ bool first() const { return t < G; }
  __device__ bool last() const { return t + G >= T; }
  __device__ void next() {
    t += G;
    if (t >= T) { u += stride; ++useq; t = u % G; }
  }
};

__global__ void __launch_bounds__(320, 1)
netvlad_c4_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                  const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo,
                  const Nv4Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NV4_BARS);
  uint64_t* w_full = bars;            // 1
  uint64_t* x_full = bars + 1;        // [2] 1 + tx
  uint64_t* x_empty = bars + 3;       // [2] 1 (GEMM 2 commit)
  uint64_t* z_full = bars + 5;        // [2] 1 (GEMM 1 commit)
  uint64_t* z_empty = bars + 7;       // [2] 4 warps
  uint64_t* p_full = bars + 9;        // 1 + 32 KiB tx: partial logits of my 32 pixels from the four CTAs
  uint64_t* p_empty = bars + 10;      // [4] (per destination CTA q) 4: the softmax warps of CTA q
  uint64_t* a_full = bars + 14;       // 1 + 32 KiB tx: a' rows from the four CTAs
  uint64_t* a_empty = bars + 15;      // 4: GEMM 2 commits of the 4 CTAs (multicast)
  uint64_t* d_full = bars + 16;       // 1
  uint64_t* d_empty = bars + 17;      // 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* asum_sm = reinterpret_cast<float*>(smem + NV4_ASUM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;
  const int n_units = a.B * a.G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xhi); tma_prefetch_desc(&tm_xlo); tma_prefetch_desc(&tm_whi); tma_prefetch_desc(&tm_wlo);
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
      mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 4);
    }
    mbar_init(p_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&p_empty[i], 4);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 4);
    mbar_init(d_full, 1);
    mbar_init(d_empty, 4);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_v = tmem_base + 256;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 32768);
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(smem + NV4_W + c * 16384, &tm_whi, w_full, (int)rank * 128 + c * 64, 0);
        tma_load_2d(smem + NV4_W + c * 16384 + 8192, &tm_wlo, w_full, (int)rank * 128 + c * 64, 0);
      }
      Nv4Iter it;
      it.init(cid, n_clusters, a.G, a.T, n_units);
      for (int i = 0; it.valid(); it.next(), ++i) {
        const int buf = i & 1;
        const int b = it.u / a.G, p0 = it.t * 128;
        if (it.first()) {   // pull the unit's later tiles into L2 while the first ones are processed
          for (int t2 = it.t + a.G; t2 < a.T; t2 += a.G)
            for (int c = 0; c < 2; ++c) {
              tma_prefetch_3d(&tm_xhi, (int)rank * 128 + c * 64, t2 * 128, b);
              tma_prefetch_3d(&tm_xlo, (int)rank * 128 + c * 64, t2 * 128, b);
            }
        }
        mbar_wait(&x_empty[buf], ((i >> 1) & 1) ^ 1);
        uint8_t* xb = smem + NV4_X + buf * 65536;
        mbar_arrive_expect_tx(&x_full[buf], 65536);
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(xb + c * 16384, &tm_xhi, &x_full[buf], (int)rank * 128 + c * 64, p0, b);
          tma_load_3d(xb + 32768 + c * 16384, &tm_xlo, &x_full[buf], (int)rank * 128 + c * 64, p0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16_f32(128, 128);
      constexpr uint32_t idesc2 = umma_idesc_bf16_f32(128, 128) | (1u << 15) | (1u << 16);   // A and B MN-major
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t w_addr = smem_u32(smem + NV4_W);
      const uint32_t a_addr = smem_u32(smem + NV4_A);
      Nv4Iter g1, g2;
      g1.init(cid, n_clusters, a.G, a.T, n_units);
      g2.init(cid, n_clusters, a.G, a.T, n_units);
      int i1 = 0, i2 = 0;
      while (g2.valid()) {
        if (g1.valid()) {
          // ---- GEMM 1 of tile i1: Z[128 px, hi|lo x 64 k] over this CTA's 128 channels ----
          const int buf = i1 & 1;
          mbar_wait(&x_full[buf], (i1 >> 1) & 1);
          mbar_wait(&z_empty[buf], ((i1 >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t xb = smem_u32(smem + NV4_X + buf * 65536);
          const uint32_t t_z = tmem_base + buf * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint64_t xh = umma_desc_kmajor_sw128(xb + c * 16384);
            const uint64_t xl = umma_desc_kmajor_sw128(xb + 32768 + c * 16384);
            const uint64_t wd = umma_desc_kmajor_sw128(w_addr + c * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);
              umma_bf16(t_z, xl + ko, wd + ko, idesc1, (c > 0 || k > 0) ? 1u : 0u);
              umma_bf16(t_z, xh + ko, wd + ko, idesc1, 1u);
            }
          }
          umma_commit(&z_full[buf]);
          g1.next();
          ++i1;
        }
        if (i2 < i1 - 1 || !g1.valid()) {
          // ---- GEMM 2 of tile i2: V[128 c, hi|lo x 64 k] += X^T a' ----
          const int buf = i2 & 1;
          if (g2.first()) {
            mbar_wait(d_empty, (g2.useq & 1) ^ 1);     // the previous unit's V has been read out
            tc_fence_after();
          }
          mbar_arrive_expect_tx(a_full, 32768);        // 4 CTAs x 32 px x (128 B hi + 128 B lo)
          mbar_wait_cluster(a_full, i2 & 1);
          tc_fence_after();
          const uint32_t xb = smem_u32(smem + NV4_X + buf * 65536);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {             // 16 pixel rows (2048 B) per MMA
            const uint32_t off = ks * 2048;
            const uint64_t xh = nv4_desc_mn(xb + off, 16384);
            const uint64_t xl = nv4_desc_mn(xb + 32768 + off, 16384);
            const uint64_t ad = nv4_desc_mn(a_addr + off, 16384);
            umma_bf16(t_v, xl, ad, idesc2, (g2.first() && ks == 0) ? 0u : 1u);
            umma_bf16(t_v, xh, ad, idesc2, 1u);
          }
          umma_commit(&x_empty[buf]);
          umma_commit_mc(a_empty, 0xF);                // every CTA's a' buffer is written by all four
          if (g2.last()) umma_commit(d_full);
          g2.next();
          ++i2;
        }
      }
    }
  } else if (warp < 6) {
    // ================= senders (warps 2..5): partial logits -> owner CTA; V read-out =================
    const int q = warp & 3;                            // TMEM lane quarter == pixel group == destination CTA
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t ex_remote = mapa_u32(smem_u32(smem + NV4_EX) + rank * NV4_EX_SLOT + (uint32_t)lane * 256u, (uint32_t)q);
    const uint32_t p_full_remote = mapa_u32(smem_u32(p_full), (uint32_t)q);
    Nv4Iter it;
    it.init(cid, n_clusters, a.G, a.T, n_units);
    NV4_STAMP(0);
    for (int i = 0; it.valid(); it.next(), ++i) {
      const int ds = 1 + 4 * i;
      const int buf = i & 1;
      mbar_wait(&z_full[buf], (i >> 1) & 1);
      NV4_STAMP(ds);
      tc_fence_after();
      const uint32_t t_z = tmem_base + buf * 128 + lane_base;
      float p[64];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_z + h * 32, r0);               // x . W_hi
        tmem_ld_32x32(t_z + 64 + h * 32, r1);          // x . W_lo
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) p[h * 32 + j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_empty[buf]);
      NV4_STAMP(ds + 1);
      mbar_wait_cluster(&p_empty[q], (i & 1) ^ 1);     // CTA q has read the previous tile's partials
      NV4_STAMP(ds + 2);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        st_async_v4(ex_remote + j * 16, __float_as_uint(p[4 * j]), __float_as_uint(p[4 * j + 1]),
                    __float_as_uint(p[4 * j + 2]), __float_as_uint(p[4 * j + 3]), p_full_remote);
      NV4_STAMP(ds + 3);
      if (it.last()) {
        // ---- unit epilogue: this CTA's 128-channel slice of V^T ----
        const int unit = it.u;
        mbar_wait(d_full, it.useq & 1);
        tc_fence_after();
        float* po = a.part + (long long)unit * 64 * 512 + (int)rank * 128 + q * 32 + lane;   // lane == channel
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_v + lane_base + h * 32, r0);        // x . a_hi
          tmem_ld_32x32(t_v + lane_base + 64 + h * 32, r1);   // x . a_lo
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) po[(long long)(h * 32 + k) * 512] = __uint_as_float(r0[k]) + __uint_as_float(r1[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
      }
    }
  } else {
    // ================= softmax (warps 6..9): 4 threads per pixel of this CTA's group of 32 =================
    const int q = warp - 6;
    const int pxl = 8 * q + (lane >> 2);               // pixel inside this CTA's group
    const int qq = lane & 3;                           // clusters [16 qq, 16 qq + 16)
    const int row = 32 * (int)rank + pxl;              // pixel row inside the 128-pixel tile
    uint32_t pe_remote[4], af_remote[4], abuf_remote[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pe_remote[j] = mapa_u32(smem_u32(&p_empty[rank]), (uint32_t)j);
      af_remote[j] = mapa_u32(smem_u32(a_full), (uint32_t)j);
      abuf_remote[j] = mapa_u32(smem_u32(smem + NV4_A), (uint32_t)j);
    }
    float asum[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) asum[e] = 0.f;
    Nv4Iter it;
    it.init(cid, n_clusters, a.G, a.T, n_units);
    for (int i = 0; it.valid(); it.next(), ++i) {
      const int ds = 64 + 1 + 4 * i;
      const int b = it.u / a.G;
      const int s = it.t * 128 + row;
      const bool valid = s < a.S;
      float inv = 1.f;
      if (valid && a.normalize_input) {
        float ss = 0.f;
        for (int p = 0; p < a.ssq_parts; ++p) ss += __ldg(a.ssq + (long long)p * a.B * a.S + (long long)b * a.S + s);
        inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      }
      if (warp == 6 && lane == 0) mbar_arrive_expect_tx(p_full, 32768);   // 4 CTAs x 32 px x 256 B
      NV4_STAMP(ds);
      mbar_wait_cluster(p_full, i & 1);
      NV4_STAMP(ds + 1);
      float z[16];
      {
        const uint32_t ex = smem_u32(smem + NV4_EX) + (uint32_t)(pxl * 256 + qq * 64);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 v0 = lds_v4(ex + e * 16), v1 = lds_v4(ex + NV4_EX_SLOT + e * 16),
                       v2 = lds_v4(ex + 2 * NV4_EX_SLOT + e * 16), v3 = lds_v4(ex + 3 * NV4_EX_SLOT + e * 16);
          z[4 * e + 0] = ((v0.x + v1.x) + (v2.x + v3.x)) * inv;
          z[4 * e + 1] = ((v0.y + v1.y) + (v2.y + v3.y)) * inv;
          z[4 * e + 2] = ((v0.z + v1.z) + (v2.z + v3.z)) * inv;
          z[4 * e + 3] = ((v0.w + v1.w) + (v2.w + v3.w)) * inv;
        }
      }
      float m = z[0];
#pragma unroll
      for (int e = 1; e < 16; ++e) m = fmaxf(m, z[e]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      {
        // "My slots may be overwritten": the arrive must not be ISSUED before the loads above have RETURNED
        // (the tensor cores can stall shared-memory reads for longer than the remote round trip), so its
        // address is made to depend on the loaded data through an opaque zero.
        uint32_t dep;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(dep) : "r"(__float_as_uint(m)));
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mbar_arrive_remote_relaxed(pe_remote[j] + dep);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) { z[e] = expf(z[e] - m); sum += z[e]; }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float rs = valid ? 1.f / sum : 0.f;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a0 = z[2 * e] * rs, a1 = z[2 * e + 1] * rs;
        asum[2 * e] += a0;
        asum[2 * e + 1] += a1;
        const float x0 = a0 * inv, x1 = a1 * inv;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
        __nv_bfloat162 hh(h0, h1);
        __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
        hi[e] = *reinterpret_cast<uint32_t*>(&hh);
        lo[e] = *reinterpret_cast<uint32_t*>(&ll);
      }
      mbar_wait_cluster(a_empty, (i & 1) ^ 1);         // GEMM 2 of the previous tile is done in all four CTAs
      NV4_STAMP(ds + 2);
      // MN-major SW128 row `row`: 8 chunks of 8 clusters, chunk j stored at j ^ (row & 7); mine: 2qq, 2qq+1
      const uint32_t o0 = (uint32_t)row * 128u + (uint32_t)(((2 * qq) ^ (row & 7)) * 16);
      const uint32_t o1 = (uint32_t)row * 128u + (uint32_t)(((2 * qq + 1) ^ (row & 7)) * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        st_async_v4(abuf_remote[j] + o0, hi[0], hi[1], hi[2], hi[3], af_remote[j]);
        st_async_v4(abuf_remote[j] + o1, hi[4], hi[5], hi[6], hi[7], af_remote[j]);
        st_async_v4(abuf_remote[j] + 16384u + o0, lo[0], lo[1], lo[2], lo[3], af_remote[j]);
        st_async_v4(abuf_remote[j] + 16384u + o1, lo[4], lo[5], lo[6], lo[7], af_remote[j]);
      }
      NV4_STAMP(ds + 3);
      if (it.last()) {
        // ---- this CTA's share of sum_s a for the unit ----
#pragma unroll
        for (int e = 0; e < 16; ++e) {                 // add the warp's 8 pixels (lanes differing in bits 2..4)
          float v = asum[e];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (lane < 4) asum_sm[q * 64 + 16 * qq + e] = v;
          asum[e] = 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x - 192 < 64) {
          const int k = threadIdx.x - 192;
          a.asum_part[((long long)it.u * 4 + rank) * 64 + k] =
              (asum_sm[k] + asum_sm[64 + k]) + (asum_sm[128 + k] + asum_sm[192 + k]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // no CTA may exit while a peer can still write into it
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// number of resident 4-CTA clusters (<= 37 on 148 SMs; GPC boundaries usually allow fewer)
int netvlad_c4_clusters() {
  static int n = 0;
  if (n) return n;
  cudaFuncSetAttribute(netvlad_c4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NV4_SMEM + 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = NV4_SMEM + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, netvlad_c4_kernel, &cfg) != cudaSuccess || nc < 1) {
    cudaGetLastError();
    nc = 32;
  }
  n = nc;
  if (getenv("IBL_NV_DEBUG")) fprintf(stderr, "[nv4] resident 4-CTA clusters: %d\n", n);
  return n;
}

int netvlad_c4_units(int B, int S) {
  const int T = cdiv(S, 128);
  int G = netvlad_c4_clusters() / (B > 0 ? B : 1);
  if (G < 1) G = 1;
  if (G > T) G = T;
  return G;
}

int launch_netvlad_c4(const CUtensorMap& mx_hi, const CUtensorMap& mx_lo, const CUtensorMap& mw_hi,
                      const CUtensorMap& mw_lo, int B, int S, int G, const float* ssq, int ssq_parts,
                      bool normalize_input, float* part, float* asum_part, cudaStream_t s) {
  Nv4Args a{};
  a.B = B; a.S = S; a.T = cdiv(S, 128); a.G = G;
  a.ssq = ssq; a.ssq_parts = ssq_parts; a.normalize_input = normalize_input ? 1 : 0;
  a.part = part; a.asum_part = asum_part;
  static unsigned long long* dbg_dev = nullptr;
  static const bool dbg_on = [] { const char* v = getenv("IBL_NV_DEBUG"); return v && atoi(v) != 0; }();
  if (dbg_on && !dbg_dev) cudaMalloc(&dbg_dev, 148 * 128 * 8);
  if (dbg_on) cudaMemsetAsync(dbg_dev, 0, 148 * 128 * 8, s);
  a.dbg = dbg_on ? dbg_dev : nullptr;
  const int units = B * G;
  const int nc = netvlad_c4_clusters();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4 * (units < nc ? units : nc));
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = NV4_SMEM + 1024;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IBL_CUDA_OK(cudaLaunchKernelEx(&cfg, netvlad_c4_kernel, mx_hi, mx_lo, mw_hi, mw_lo, a));
  if (dbg_on) {   // phase stamps of the first cluster, ns relative to the earliest stamp; 8 per tile:
    // z_full seen | partials read from TMEM | p_empty seen | partials sent
    cudaStreamSynchronize(s);
    static unsigned long long h[148 * 128];
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 148 * 128; ++i) if (h[i] && h[i] < t0) t0 = h[i];
    for (int c = 0; c < 4; ++c) {
      fprintf(stderr, "[nv4] cta %d start %lld\n", c, (long long)(h[c * 128] - t0));
      for (int t = 0; t < 10; ++t) {
        fprintf(stderr, "[nv4]   tile %d:", t);
        for (int j = 0; j < 4; ++j) fprintf(stderr, " %6lld", h[c * 128 + 1 + 4 * t + j] ? (long long)(h[c * 128 + 1 + 4 * t + j] - t0) : -1ll);
        fprintf(stderr, "   |");   // softmax warps: loop top | p_full seen | a_empty seen | a' sent
        for (int j = 0; j < 4; ++j) fprintf(stderr, " %6lld", h[c * 128 + 65 + 4 * t + j] ? (long long)(h[c * 128 + 65 + 4 * t + j] - t0) : -1ll);
        fprintf(stderr, "\n");
      }
    }
  }
  return IBL_OK;
}

}  // namespace ibl
