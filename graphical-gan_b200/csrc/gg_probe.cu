// gg_probe.cu — two small hardware probes exercised by tests/test_gpu_probes.py.  They pin down the
// TMA / tcgen05 behaviours the implicit-GEMM convolution kernels (gg_conv_tc.cu) are built on:
//   * a tiled tensor map with elementStrides = 2 gathers the stride-2 input pixels of a conv tap and
//     zero-fills out-of-range (padding) coordinates;
//   * kind::tf32 UMMA with K-major and MN-major 128B-swizzled shared-memory operands written by TMA.
#include "gg_tc_common.cuh"

using namespace gg;

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_tma_strided_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int w0, int h0,
                                                                int b, int nfloats, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) tile[i] = -12345.f;  // poison: detect unwritten elements
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)nfloats * 4u);
    tma_load_4d(tile, &tmap, &bar, c0, w0, h0, b);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = tile[i];
}

extern "C" int gg_probe_tma_strided(const float* x, int B, int H, int W, int C, int b, int h0, int w0, int c0, int hb, int wb,
                                    int swizzle128, float* out, void* stream) {
  GG_REQUIRE(hb > 0 && wb > 0 && hb * wb * 32 * 4 <= 64 * 1024, "gg_probe_tma_strided");
  CUtensorMap tmap;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
  uint32_t box[4] = {32, (uint32_t)(2 * wb), (uint32_t)(2 * hb), 1};
  uint32_t es[4] = {1, 2, 2, 1};
  int rc = encode_tmap(&tmap, x, 4, dims, strides, box, es, swizzle128);
  if (rc) return rc;
  int nfloats = hb * wb * 32;
  size_t smem = (size_t)nfloats * 4 + 1024;
  cudaFuncSetAttribute(probe_tma_strided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_tma_strided_kernel<<<1, 128, smem, as_stream(stream)>>>(tmap, c0, w0, h0, b, nfloats, out);
  return check_launch("gg_probe_tma_strided");
}

// -------------------------------------------------------------------------------------------------
// D[128,N] = A[128,K] * B[K,N]; one CTA, serial k-blocks of 32 (no pipelining: this is a semantics probe).
__global__ void __launch_bounds__(128) probe_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                         float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sA = reinterpret_cast<float*>(base);                   // 128 x 32 fp32 = 16 KB
  float* sB = reinterpret_cast<float*>(base + 128 * 32 * 4);    // N x 32 fp32
  __shared__ uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_sh;
  int warp = threadIdx.x >> 5;
  uint32_t ncols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_sh, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_d = tmem_base_sh;
  uint32_t idesc = make_idesc_tf32(128, N, a_mn, b_mn);
  int nkb = K / 32;
  for (int kb = 0; kb < nkb; ++kb) {
    uint32_t par = kb & 1;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full_bar, (uint32_t)(128 + N) * 32u * 4u);
      if (!a_mn) tma_load_2d(sA, &tmA, &full_bar, kb * 32, 0);
      else
        for (int mb = 0; mb < 4; ++mb) tma_load_2d(sA + mb * 1024, &tmA, &full_bar, mb * 32, kb * 32);
      if (!b_mn) tma_load_2d(sB, &tmB, &full_bar, kb * 32, 0);
      else
        for (int nb = 0; nb < N / 32; ++nb) tma_load_2d(sB + nb * 1024, &tmB, &full_bar, nb * 32, kb * 32);
      mbar_wait(&full_bar, par);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // K-major: 8-row x 128B swizzle atoms, k-step = 32 bytes inside the row.  MN-major (tf32): 4-row x 128B
        // atoms (32-byte swizzle chunks), one MMA (K=8) spans two atoms: SBO = 512, k-step = 1024 bytes.
        uint64_t ad = a_mn ? make_smem_desc(smem_u32(sA) + j * 1024, 4096, 512, kLayoutSW128_32B)
                           : make_smem_desc(smem_u32(sA) + j * 32, 16, 1024);
        uint64_t bd = b_mn ? make_smem_desc(smem_u32(sB) + j * 1024, 4096, 512, kLayoutSW128_32B)
                           : make_smem_desc(smem_u32(sB) + j * 32, 16, 1024);
        umma_tf32(tmem_d, ad, bd, idesc, (kb | j) != 0);
      }
      umma_commit(&mma_bar);
      mbar_wait(&mma_bar, par);  // smem is reused by the next k-block
    }
    __syncthreads();
  }
  tc_fence_after();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32)
  int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(long long)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, ncols);
}

extern "C" int gg_probe_umma_tf32(const float* A, const float* Bm, float* D, int N, int K, int a_mn_major, int b_mn_major,
                                  int tma_tf32_convert, void* stream) {
  GG_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 32 == 0 && K >= 32, "gg_probe_umma_tf32");
  CUtensorMap tmA, tmB;
  int rc;
  bool cv = tma_tf32_convert != 0;
  if (!a_mn_major) {
    uint64_t dims[2] = {(uint64_t)K, 128}, str[2] = {1, (uint64_t)K};
    uint32_t box[2] = {32, 128};
    rc = encode_tmap(&tmA, A, 2, dims, str, box, nullptr, 1, cv);
  } else {
    uint64_t dims[2] = {128, (uint64_t)K}, str[2] = {1, 128};
    uint32_t box[2] = {32, 32};
    rc = encode_tmap(&tmA, A, 2, dims, str, box, nullptr, 2, cv);
  }
  if (rc) return rc;
  if (!b_mn_major) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[2] = {1, (uint64_t)K};
    uint32_t box[2] = {32, (uint32_t)N};
    rc = encode_tmap(&tmB, Bm, 2, dims, str, box, nullptr, 1, cv);
  } else {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)K}, str[2] = {1, (uint64_t)N};
    uint32_t box[2] = {32, 32};
    rc = encode_tmap(&tmB, Bm, 2, dims, str, box, nullptr, 2, cv);
  }
  if (rc) return rc;
  size_t smem = (size_t)(128 + N) * 32 * 4 + 1024;
  cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_umma_kernel<<<1, 128, smem, as_stream(stream)>>>(tmA, tmB, D, N, K, a_mn_major, b_mn_major);
  return check_launch("gg_probe_umma_tf32");
}
