// Hardware probe: does a K-major SWIZZLE_128B UMMA operand tolerate (a) a start address that is 128-byte
// but not 1024-byte aligned and (b) a stride between 8-row groups (SBO) that is not a multiple of 1024?
// That is what a convolution needs to read all nine taps out of ONE halo tile staged by a single TMA box
// ([rows of (TW+2) pixels][64 channels]): tap (kh,kw) starts (kh*(TW+2)+kw) rows into the tile and 8-pixel
// row groups are (TW+2) rows apart.
//
//   A_halo : [rows][64] bf16 dense, loaded by TMA with the 128B swizzle (row r at byte r*128, 16-byte chunk
//            j stored at chunk position j ^ (r & 7))
//   view   : row m of the 128-row operand = halo row  s0 + (m / 8) * group_rows + (m % 8)
//   D[m,n] = sum_k view[m,k] * B[n,k]
// The caller compares D with the expected product for several (s0, group_rows, base_offset mode).
#include "common.cuh"
#include "tc_common.cuh"

namespace ibl {

using namespace tc;

__global__ void __launch_bounds__(128, 1)
umma_strided_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                          int rows, int s0, int group_rows, int base_mode, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_sm = smem;                       // rows * 128 B (<= 32 KiB)
  uint8_t* b_sm = smem + 32768;               // 64 rows * 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], rows * 128 + 8192);
    tma_load_2d(a_sm, &tm_a, &bars[0], 0, 0);
    tma_load_2d(b_sm, &tm_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(a_sm) + (uint32_t)s0 * 128u;
    uint64_t da = 0;
    da |= (uint64_t)((a_addr >> 4) & 0x3fffu);
    da |= (uint64_t)1 << 16;
    da |= (uint64_t)(((uint32_t)group_rows * 128u) >> 4) << 32;       // SBO = group_rows * 128 B
    da |= (uint64_t)1 << 46;
    if (base_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;   // base_offset = start phase
    da |= (uint64_t)2 << 61;
    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(b_sm));
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k > 0);
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    uint32_t r0[32], r1[32];
    const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16);
    tmem_ld_32x32(t, r0);
    tmem_ld_32x32(t + 32, r1);
    tmem_ld_wait();
    float* o = D + (size_t)(warp * 32 + lane) * 64;
    for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(r0[j]); o[32 + j] = __uint_as_float(r1[j]); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// A: [rows][64] bf16 bits, B: [64][64] bf16 bits (device), D: [128][64] fp32
int debug_umma_strided(const void* A, int rows, const void* B, int s0, int group_rows, int base_mode, float* D,
                       cudaStream_t s) {
  IBL_REQUIRE(rows >= 8 && rows <= 256 && s0 >= 0 && group_rows >= 8 && s0 + 15 * group_rows + 8 <= rows,
              "probe view does not fit the halo tile");
  CUtensorMap ma, mb;
  {
    uint64_t dims[2] = {64, (uint64_t)rows};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, (uint32_t)rows};
    IBL_RET(make_tmap(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, str, box));
  }
  {
    uint64_t dims[2] = {64, 64};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 64};
    IBL_RET(make_tmap(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, str, box));
  }
  const int smem = 32768 + 8192 + 1024 + 64;
  static DeviceOnce attr_done;   // the attribute is per device
  if (!attr_done.done()) {
    IBL_CUDA_OK(cudaFuncSetAttribute(umma_strided_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.mark();
  }
  umma_strided_probe_kernel<<<1, 128, smem, s>>>(ma, mb, rows, s0, group_rows, base_mode, D);
  IBL_CUDA_OK(cudaGetLastError());
  return IBL_OK;
}

}  // namespace ibl
