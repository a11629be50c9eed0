// One-tile GEMM through the exact tcgen05 building blocks of the patch Encoder
// (tc05.cuh): swizzled operand layouts, matrix / instruction descriptors, bulk
// async copy of a host-packed weight image, tcgen05.mma, commit, tcgen05.ld.
// D[128,N] = A[128,K] * B[N,K]^T with operands rounded to the selected type.
// Exposed through the C ABI so the GPU tests can pin every layout assumption
// against a torch matmul before the Encoder kernels rely on them.
#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int ST_THREADS = 128;
constexpr int ST_TMEM_COLS = 512;
constexpr int ST_A_COL = 256;  // A-in-TMEM mode: the A operand lives at columns [256, 256 + K / 2)

template <uint32_t FMT>
__global__ void __launch_bounds__(ST_THREADS, 1)
umma_selftest_kernel(const float* __restrict__ a, const unsigned char* __restrict__ a_packed,
                     const float* __restrict__ b, float* __restrict__ d, int N, int K, int b_mn, int split, int a_tmem) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kchunks = K / 64;
  const uint32_t a_bytes = (uint32_t)kchunks * 16384u;      // one copy of A
  const uint32_t b_bytes = (uint32_t)N * (uint32_t)K * 2u;  // one copy of B
  unsigned char* sA = smem;                                 // [split][kchunks][128 x 128 B]
  unsigned char* sB = sA + (size_t)split * a_bytes;         // [split][...]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)split * b_bytes);  // [0] load, [1] mma
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled atoms need 1024-byte alignment

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<ST_TMEM_COLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  // ---- operands into shared memory ----
  if (a_tmem) {
    // A in TENSOR MEMORY (tcgen05.mma with [a_tmem]): row m of A = lane m, two consecutive K elements per 32-bit
    // column (element k in the low half).  Thread m (warp w owns lanes 32 w .. 32 w + 31) converts its row and
    // stores it 8 columns (K = 16) at a time.
    const int m = tid;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t lo = to_operand<FMT>(a[(size_t)m * K + k0 + 2 * j]);
        const uint32_t hi = to_operand<FMT>(a[(size_t)m * K + k0 + 2 * j + 1]);
        r[j] = lo | (hi << 16);
      }
      tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ST_A_COL + k0 / 2), r);
    }
    tmem_wait_st();
  } else if (a_packed) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[0], (uint32_t)split * a_bytes);
      for (int s = 0; s < split; ++s)
        for (int kc = 0; kc < kchunks; ++kc)
          bulk_g2s(sA + (size_t)s * a_bytes + (size_t)kc * 16384, a_packed + ((size_t)kc * split + s) * 16384, 16384u,
                   &bars[0]);  // image order [kc][split], as in the Encoder weight blob
    }
  } else {
    for (int e = tid; e < 128 * K; e += ST_THREADS) {
      const int m = e / K, k = e - m * K;
      const float v = a[e];
      const uint16_t hi = to_operand<FMT>(v);
      const uint32_t off = (uint32_t)(k >> 6) * 16384u + sw128_kmajor_off(m, k & 63);
      *reinterpret_cast<uint16_t*>(sA + off) = hi;
      if (split == 2) *reinterpret_cast<uint16_t*>(sA + a_bytes + off) = to_operand<FMT>(v - from_operand<FMT>(hi));
    }
  }
  const uint32_t mn_block = (uint32_t)(K / 8) * 1024u;  // MN-major: bytes between 64-point blocks
  for (int e = tid; e < N * K; e += ST_THREADS) {
    const int n = e / K, k = e - n * K;
    const float v = b[e];
    const uint16_t hi = to_operand<FMT>(v);
    const uint32_t off = b_mn ? sw128_mnmajor_off(n, k, mn_block)
                              : (uint32_t)(k >> 6) * ((uint32_t)N * 128u) + sw128_kmajor_off(n, k & 63);
    *reinterpret_cast<uint16_t*>(sB + off) = hi;
    if (split == 2) *reinterpret_cast<uint16_t*>(sB + b_bytes + off) = to_operand<FMT>(v - from_operand<FMT>(hi));
  }
  fence_proxy_async_smem();
  fence_before_sync();  // (orders the tcgen05.st of the A-in-TMEM mode before the barrier)
  __syncthreads();

  // ---- MMA: one thread issues ----
  if (tid == 0) {
    if (a_packed) mbar_wait(&bars[0], 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc(FMT, 128, N, b_mn);
    const uint32_t aaddr = smem_u32(sA), baddr = smem_u32(sB);
    uint32_t acc = 0;
    for (int kc = 0; kc < kchunks; ++kc)
      for (int k16 = 0; k16 < 4; ++k16)
        for (int pass = 0; pass < (split == 2 ? 3 : 1); ++pass) {
          const int sa = pass == 2 ? 1 : 0, sb = pass == 1 ? 1 : 0;  // hi*hi, hi*lo, lo*hi
          const uint64_t adesc = make_sdesc(aaddr + sa * a_bytes + kc * 16384u + k16 * 32u, 16u, 1024u);
          const uint32_t a_taddr = tbase + (uint32_t)(ST_A_COL + (kc * 64 + k16 * 16) / 2);
          uint64_t bdesc;
          if (b_mn) {
            const uint32_t k0 = (uint32_t)kc * 64u + (uint32_t)k16 * 16u;
            bdesc = make_sdesc(baddr + sb * b_bytes + (k0 >> 3) * 1024u, mn_block, 1024u);
          } else {
            bdesc = make_sdesc(baddr + sb * b_bytes + kc * ((uint32_t)N * 128u) + k16 * 32u, 16u, 1024u);
          }
          if (a_tmem) umma_f16_ta(tbase, a_taddr, bdesc, idesc, acc);
          else umma_f16(tbase, adesc, bdesc, idesc, acc);
          acc = 1;
        }
    umma_commit(&bars[1]);
  }

  // ---- epilogue: TMEM -> global ----
  mbar_wait(&bars[1], 0);
  fence_after_sync();
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) d[(size_t)m * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<ST_TMEM_COLS>(tbase);
}

}  // namespace

// mode: bits [0,2) = PPT_ENC_FP16 / PPT_ENC_BF16 / PPT_ENC_FP16X3; bit 2 = B operand MN-major;
// bit 3 = `a` points at a packed operand image (ppt_b200/encoder_pack.py: pack_kmajor) instead of fp32;
// bit 4 = the A operand is placed in tensor memory (tcgen05.st) and the MMA reads it from there.
extern "C" PPT_EXPORT int ppt_selftest_umma(const float* a, const float* b, float* d, int N, int K, int mode,
                                            void* stream) {
  if (!a || !b || !d) return PPT_EINVAL;
  const int prec = mode & 3, b_mn = (mode >> 2) & 1, packed = (mode >> 3) & 1, a_tmem = (mode >> 4) & 1;
  if (a_tmem && (packed || prec == PPT_ENC_FP16X3 || K > 512)) return PPT_EINVAL;
  if (prec > PPT_ENC_FP16X3) return PPT_EINVAL;
  if (N < 32 || N > 256 || (N % 32) != 0 || K < 64 || (K % 64) != 0) return PPT_ERANGE;
  if (b_mn && (N % 64) != 0) return PPT_ERANGE;
  const int split = prec == PPT_ENC_FP16X3 ? 2 : 1;
  const size_t smem = (size_t)split * ((size_t)(K / 64) * 16384 + (size_t)N * K * 2) + 64;
  if (smem > 220 * 1024) return PPT_ERANGE;
  const float* af = packed ? nullptr : a;
  const unsigned char* ap = packed ? reinterpret_cast<const unsigned char*>(a) : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (prec != PPT_ENC_BF16) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<tc05::FMT_F16>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    umma_selftest_kernel<tc05::FMT_F16><<<1, ST_THREADS, smem, st>>>(af, ap, b, d, N, K, b_mn, split, a_tmem);
  } else {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<tc05::FMT_BF16>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    umma_selftest_kernel<tc05::FMT_BF16><<<1, ST_THREADS, smem, st>>>(af, ap, b, d, N, K, b_mn, split, a_tmem);
  }
  return ppt_launch_status();
}
