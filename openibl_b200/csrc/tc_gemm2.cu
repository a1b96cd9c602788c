// Distance + running top-16 on a PAIR of SMs: tcgen05.mma.cta_group::2 (UMMA M = 256).
//
// The one-SM kernel (tc_gemm.cu) is bound by shared-memory bandwidth: an M=128 MMA re-reads its whole B tile
// from the issuing SM's shared memory while TMA is writing the next stages into the same banks.  With
// cta_group::2 the two CTAs of a cluster form one 256-row tile: each CTA stages its own 128 query rows (A) and
// only HALF of the database tile (B, 128 of 256 rows); the tensor cores of both SMs read A locally and the two
// B halves across the pair.  Per SM and 16-wide K step: 8 KiB of operand reads per 128 clk (64 B/clk) instead
// of 12 KiB (96 B/clk), and the TMA fill per stage drops from 96 to 64 KiB.
//
//   * both CTAs issue TMA (cta_group::2 form) into their own shared memory; all bytes are accounted on the
//     LEADER's (rank 0) full barrier, on which the peer's producer also arrives remotely;
//   * one thread of the leader issues the MMAs for the pair and multicasts its tcgen05.commit to the empty
//     and accumulator-full barriers of both CTAs;
//   * each CTA's epilogue warps drain their own 128 accumulator rows from their own TMEM and arrive on the
//     leader's accumulator-empty barrier (count 8).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ibl {

using namespace tc;

struct Gemm2Args {
  int M, N, K;
  int n_tiles, nt_per_item, items_per_mpair, total_items, n_valid;
  const float* an;
  const float* bn;
  float* cand_d;        // [items_per_mpair][M][16]
  long long* cand_i;
};

constexpr int G2_BN = 256, G2_BK = 64, G2_STAGES = 3;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;            // 16 KiB per plane: this CTA's 128 query rows
constexpr int G2_BH_BYTES = (G2_BN / 2) * G2_BK * 2;   // 16 KiB per plane: this CTA's half of the B tile
constexpr int G2_STAGE = 2 * G2_A_BYTES + 2 * G2_BH_BYTES;   // 64 KiB

__global__ void __launch_bounds__(192, 1)
gemm2_top16_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                   const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                   const Gemm2Args g) {
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int unit0 = blockIdx.x >> 1, unit_stride = gridDim.x >> 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE);
  uint64_t* full_bar = bars;                       // leader's are used: count 1 (leader's arrive.expect_tx for both CTAs; see tc_dist1.cu)
  uint64_t* empty_bar = bars + G2_STAGES;          // local, count 1 (multicast commit)
  uint64_t* tfull_bar = bars + 2 * G2_STAGES;      // local, count 1 (multicast commit)
  uint64_t* tempty_bar = bars + 2 * G2_STAGES + 2; // leader's are used: count 8 (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_ahi); tma_prefetch_desc(&tm_alo); tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo);
    for (int i = 0; i < G2_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 8); mbar_init(&tempty_bar[1], 8);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {   // both CTAs, same warp id: one allocation spanning the pair
    tmem_alloc_2sm(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& mp, int& nt0, int& ntn) {
    mp = item / g.items_per_mpair;
    const int sub = item - mp * g.items_per_mpair;
    nt0 = sub * g.nt_per_item;
    ntn = (nt0 + g.nt_per_item <= g.n_tiles) ? g.nt_per_item : (g.n_tiles - nt0);
  };
  const int kiters = g.K / G2_BK;

  // Producer and MMA issuer: convergent warps, one elected lane issues a stage's TMA / tcgen05 instructions, ring
  // position and bases warp-uniform (tc_conv.cu, MMA issuer, explains what `if (lane == 0)` costs per instruction).
  if (warp == 0) {
    const uint32_t smem_a = warp_uniform(smem_u32(smem));
    const uint32_t full_a = smem_a + G2_STAGES * G2_STAGE, empty_a = full_a + 8 * G2_STAGES;
    const uint32_t full_c = warp_uniform(mapa_u32(full_a, 0));   // the leader's barriers, shared::cluster addresses
    const int rank_u = (int)warp_uniform(rank);
    int stage = 0; uint32_t phase = 0;
    for (int item = unit0; item < g.total_items; item += unit_stride) {
      int mp, nt0, ntn;
      decode(item, mp, nt0, ntn);
      const int row0 = (int)warp_uniform((uint32_t)((mp * 2 + rank_u) * 128));
      for (int nt = nt0; nt < nt0 + ntn; ++nt) {
        const int col0 = (int)warp_uniform((uint32_t)(nt * G2_BN + rank_u * (G2_BN / 2)));
        for (int kit = 0; kit < kiters; ++kit) {
          const uint32_t sg = warp_uniform((uint32_t)stage);
          mbar_wait_warp_a(empty_a + 8 * sg, phase ^ 1);
          const uint32_t st = smem_a + sg * G2_STAGE, fb = full_c + 8 * sg;
          const int k0 = (int)warp_uniform((uint32_t)(kit * G2_BK));
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx_a(full_a + 8 * sg, 2 * G2_STAGE);   // bytes of BOTH CTAs; the peer only loads
            tma_load_2d_2sm_a(st, &tm_ahi, fb, k0, row0);
            tma_load_2d_2sm_a(st + G2_A_BYTES, &tm_alo, fb, k0, row0);
            tma_load_2d_2sm_a(st + 2 * G2_A_BYTES, &tm_bhi, fb, k0, col0);
            tma_load_2d_2sm_a(st + 2 * G2_A_BYTES + G2_BH_BYTES, &tm_blo, fb, k0, col0);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (warp_uniform(leader ? 1u : 0u)) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(256, G2_BN);
      const uint32_t tmem_u = warp_uniform(tmem_base);
      const uint32_t smem_a = warp_uniform(smem_u32(smem));
      const uint32_t full_a = smem_a + G2_STAGES * G2_STAGE, empty_a = full_a + 8 * G2_STAGES;
      const uint32_t tfull_a = full_a + 16 * G2_STAGES, tempty_a = tfull_a + 16;
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = unit0; item < g.total_items; item += unit_stride) {
        int mp, nt0, ntn;
        decode(item, mp, nt0, ntn);
        for (int nt = nt0; nt < nt0 + ntn; ++nt, ++it) {
          const uint32_t as = warp_uniform((uint32_t)(it & 1));
          const uint32_t aphase = (it >> 1) & 1;
          mbar_wait_warp_a(tempty_a + 8 * as, aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + as * G2_BN;
          for (int kit = 0; kit < kiters; ++kit) {
            const uint32_t sg = warp_uniform((uint32_t)stage);
            mbar_wait_warp_a(full_a + 8 * sg, phase);
            tc_fence_after();
            const uint32_t sa = smem_a + sg * G2_STAGE;
            if (elect_one()) {
              const uint64_t a_hi = umma_desc_kmajor_sw128(sa), a_lo = umma_desc_kmajor_sw128(sa + G2_A_BYTES);
              const uint64_t b_hi = umma_desc_kmajor_sw128(sa + 2 * G2_A_BYTES);
              const uint64_t b_lo = umma_desc_kmajor_sw128(sa + 2 * G2_A_BYTES + G2_BH_BYTES);
#pragma unroll
              for (int k = 0; k < G2_BK / 16; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                umma_bf16_2sm(d_tmem, a_lo + ko, b_hi + ko, idesc, (kit > 0 || k > 0) ? 1u : 0u);
                umma_bf16_2sm(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                umma_bf16_2sm(d_tmem, a_hi + ko, b_hi + ko, idesc, 1u);
              }
              umma_commit_2sm_mc_a(empty_a + 8 * sg, 0x3);
              if (kit == kiters - 1) umma_commit_2sm_mc_a(tfull_a + 8 * as, 0x3);   // same elected thread as the MMAs
            }
            __syncwarp();
            if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int rloc = q * 32 + lane;
    int it = 0;
    for (int item = unit0; item < g.total_items; item += unit_stride) {
      int mp, nt0, ntn;
      decode(item, mp, nt0, ntn);
      const int row = (mp * 2 + (int)rank) * 128 + rloc;
      const bool row_ok = row < g.M;
      const float an = row_ok ? __ldg(g.an + row) : 0.f;
      float td[16];
      int ti[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { td[j] = INFINITY; ti[j] = -1; }
      for (int nt = nt0; nt < nt0 + ntn; ++nt, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * G2_BN;
#pragma unroll 1
        for (int ch = 0; ch < G2_BN / 32; ++ch) {
          const int col0 = nt * G2_BN + ch * 32;
          // one coalesced load of the chunk's |d|^2 terms + shuffles instead of 32 dependent loads (see tc_dist1.cu)
          const float bmine = (col0 + lane < g.n_valid) ? __ldg(g.bn + col0 + lane) : INFINITY;
          uint32_t raw[32];
          tmem_ld_32x32(t_row + ch * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            const float d = fmaf(-2.f, __uint_as_float(raw[j]), an + __shfl_sync(0xffffffffu, bmine, j));
            if (d < td[15]) {       // chain-free sorted insert (see tc_dist1.cu)
              int pos = 0;
#pragma unroll
              for (int s = 0; s < 16; ++s) pos += (td[s] <= d) ? 1 : 0;
#pragma unroll
              for (int s = 15; s > 0; --s) {
                const bool shift = s > pos, here = s == pos;
                td[s] = shift ? td[s - 1] : (here ? d : td[s]);
                ti[s] = shift ? ti[s - 1] : (here ? col : ti[s]);
              }
              if (pos == 0) { td[0] = d; ti[0] = col; }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        }
      }
      if (row_ok) {
        const int sub = item % g.items_per_mpair;
        float* od = g.cand_d + ((long long)sub * g.M + row) * 16;
        long long* oi = g.cand_i + ((long long)sub * g.M + row) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) { od[j] = td[j]; oi[j] = ti[j]; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int sms2() { return device_sm_count(); }   // per device: one process may drive several GPUs

static int pick_runs2(int m_pairs, int n_tiles) {
  const int G = sms2() / 2;
  int best = 1;
  double best_eff = -1.0;
  for (int r = 1; r <= n_tiles && r <= 64; ++r) {
    const int per = cdiv(n_tiles, r), runs = cdiv(n_tiles, per);
    const long long total = (long long)m_pairs * runs, waves = (total + G - 1) / G;
    const double eff = (double)m_pairs * n_tiles / ((double)waves * G * per);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = runs; }
  }
  return best;
}

int dist_top16_2sm_max_runs(int m, int n_valid) {
  return pick_runs2(cdiv(cdiv(m, 128), 2), cdiv(n_valid > 0 ? n_valid : 1, G2_BN));
}

int launch_dist_top16_2sm(const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, const float* qn, int m,
                          const __nv_bfloat16* d_hi, const __nv_bfloat16* d_lo, const float* dn, int n,
                          int n_valid, int K, float* cand_d, long long* cand_i, int* runs_out, cudaStream_t s) {
  IBL_REQUIRE(K % 64 == 0, "tcgen05 distance needs dim % 64 == 0");
  CUtensorMap maps[4];
  {
    uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)m}, dims_b[2] = {(uint64_t)K, (uint64_t)n};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    IBL_RET(make_tmap(&maps[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_hi, dims_a, str, box));
    IBL_RET(make_tmap(&maps[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_lo, dims_a, str, box));
    IBL_RET(make_tmap(&maps[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_hi, dims_b, str, box));
    IBL_RET(make_tmap(&maps[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_lo, dims_b, str, box));
  }
  Gemm2Args g{};
  g.M = m; g.N = n; g.K = K;
  g.n_tiles = cdiv(n_valid > 0 ? n_valid : 1, G2_BN);
  const int m_pairs = cdiv(cdiv(m, 128), 2);
  const int runs = pick_runs2(m_pairs, g.n_tiles);
  g.nt_per_item = cdiv(g.n_tiles, runs);
  g.items_per_mpair = cdiv(g.n_tiles, g.nt_per_item);
  g.total_items = m_pairs * g.items_per_mpair;
  g.n_valid = n_valid;
  g.an = qn; g.bn = dn; g.cand_d = cand_d; g.cand_i = cand_i;
  *runs_out = g.items_per_mpair;
  const int smem = G2_STAGES * G2_STAGE + 1024 + 256;
  static DeviceOnce attr_done;   // the attribute is per device
  if (!attr_done.done()) {
    IBL_CUDA_OK(cudaFuncSetAttribute(gemm2_top16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.mark();
  }
  const int pairs = sms2() / 2;
  const int units = g.total_items < pairs ? g.total_items : pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * units);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IBL_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm2_top16_kernel, maps[0], maps[1], maps[2], maps[3], g));
  return IBL_OK;
}

}  // namespace ibl
