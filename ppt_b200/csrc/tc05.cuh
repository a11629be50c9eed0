// Thin PTX wrappers for the Blackwell (sm_100a) tensor path used by the patch
// Encoder: mbarrier, 1-D bulk async copy (TMA engine, UBLKCP), tcgen05
// alloc / mma / commit / ld, and the 128-byte-swizzled shared-memory operand
// layout with its matrix descriptors.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef PPT_PRODUCER_SLEEP_NS
#define PPT_PRODUCER_SLEEP_NS 64
#endif

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spins on try_wait (which itself sleeps in hardware for a bounded time).  A pipeline bug would
// otherwise hang the GPU: after ~2^26 failed probes (seconds) the kernel traps, so the host sees an
// error instead of a stuck stream.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// Same, for a single long-waiting thread (the copy producer): sleeps between probes so that its spin
// does not take issue slots from the warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#if PPT_PRODUCER_SLEEP_NS > 0
    __nanosleep(PPT_PRODUCER_SLEEP_NS);
#endif
    if (++spins > (1u << 26)) __trap();
  }
}

// Generic-proxy shared-memory writes -> visible to the async proxy (tensor core, bulk copies).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tensor memory ---------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp, same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base_lane + i), columns c..c+31.
// A warp may only touch the 32-lane quadrant (warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait, so that several loads are in flight before one tmem_wait_ld().
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 8 columns store: thread i of the warp writes lane (base_lane + i), columns c..c+7.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------
// Operand formats of tcgen05.mma.kind::f16 (instruction descriptor bits [7,10) / [10,13)).
constexpr uint32_t FMT_F16 = 0, FMT_BF16 = 1;

// Instruction descriptor: fp32 accumulate, both operands `fmt`, A K-major,
// B K-major (b_mn = 0) or MN-major (b_mn = 1), shape M x N (K = 16 per instruction).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int M, int N, int b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle (layout type 2), sm_100 version bit set.
//   K-major : rows (M or N index) are 128 B apart inside an 8-row, 1024-byte atom;
//             stride_bytes = distance between consecutive 8-row atoms; leading offset unused (1).
//   MN-major: an atom is 8 K-rows of 64 contiguous MN elements; stride_bytes = distance between
//             K-atoms, leading_bytes = distance between 64-element MN blocks.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t leading_bytes, uint32_t stride_bytes) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((leading_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((stride_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}

// The same descriptor as two 32-bit words, so that an issue loop can step the start address (and keep
// everything else) with one 32-bit add: lo = start>>4 | leading>>4 << 16, hi = stride>>4 | version | layout.
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr, uint32_t leading_bytes) {
  return ((smem_addr & 0x3ffffu) >> 4) | (((leading_bytes >> 4) & 0x3fffu) << 16);
}
__host__ __device__ constexpr uint32_t sdesc_hi(uint32_t stride_bytes) {
  return ((stride_bytes >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint64_t sdesc_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[TMEM] * B[smem]: the A operand (M = 128 rows = lanes, K = 16 = 8 columns of fp16 pairs) is read
// from tensor memory, so the instruction costs half the shared-memory bandwidth of the smem x smem form.
__device__ __forceinline__ void umma_f16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Warp-converged variants: every lane executes the call with warp-uniform operands (so ptxas keeps the
// descriptor arithmetic in uniform registers), one elected lane issues.  The elected lane is the same
// for every call of a converged warp, so "ops issued so far by this thread" stays meaningful for commit.
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster run one M = 256 MMA ----------------------
// CTA r of the pair holds rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 (r + 1)) of B at the SAME
// shared-memory offsets, and receives rows [128 r, ...) of D in its own tensor memory.  The leader
// (cluster rank 0) issues; commits are multicast to the barrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_pair_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// Wait with cluster-scope acquire: the arrivals may come from the other CTA of the pair.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}
// Generic-proxy writes (local or to the peer CTA's shared memory) -> visible to the async proxy everywhere.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Address of `local_smem_addr` in CTA `rank`'s shared memory (shared::cluster window).
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// Arrive on the barrier at the same offset in CTA `rank` of the cluster (release at cluster scope).
// Wait of a converged warp whose loop branch is warp-uniform (vote): ptxas then treats the code after the wait
// as convergent and keeps loop counters, barrier addresses and MMA descriptors in uniform registers.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if (++spins > (1u << 26)) __trap();
  }
}
// Arrival on the barrier at the same offset in CTA `rank` with the default (CTA-scope release) semantics:
// the cheap form -- no MEMBAR.GPU / ERRBAR -- for hand-offs whose payload is shared memory that the
// arriving thread has already made visible to the async proxy of its own SM (fence.proxy.async.shared::cta)
// or tensor memory it has finished reading (tcgen05.wait::ld + tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mbar_arrive_peer(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

// Arrives on `bar` once every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- operand layout in shared memory ---------------------------------------------------
// K-major, 128-byte swizzle.  A "chunk" is ROWS x 64 16-bit elements = ROWS x 128 B; row r lives at
// r * 128 and its 16-byte piece c is stored at position c ^ (r & 7).  Chunks must be 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_kmajor_off(int row, int k_in_chunk /*0..63*/) {
  const uint32_t piece = (uint32_t)(k_in_chunk >> 3) ^ (uint32_t)(row & 7);
  return (uint32_t)row * 128u + piece * 16u + (uint32_t)(k_in_chunk & 7) * 2u;
}
// MN-major, 128-byte swizzle: 1024-byte atoms of 8 K-rows x 64 MN elements; K-atoms are 1024 B apart
// (stride), 64-element MN blocks are `mn_block_bytes` apart (leading).
__device__ __forceinline__ uint32_t sw128_mnmajor_off(int mn, int k, uint32_t mn_block_bytes) {
  const uint32_t piece = (uint32_t)((mn & 63) >> 3) ^ (uint32_t)(k & 7);
  return (uint32_t)(mn >> 6) * mn_block_bytes + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u + piece * 16u +
         (uint32_t)(mn & 7) * 2u;
}

// ---- fp32 -> operand conversion -----------------------------------------------------------
// Two fp32 values -> one packed pair of operand elements (a in the low half = lower address), with an
// optional fused ReLU.  fp16 saturates to +-65504 instead of producing inf (SURVEY.md F15 range guard).
// One F2FP instruction either way.
template <uint32_t FMT, bool RELU>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  if (FMT == FMT_F16) {
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}
template <uint32_t FMT>
__device__ __forceinline__ float2 unpack2(uint32_t r) {
  if (FMT == FMT_F16) return __half22float2(*reinterpret_cast<const __half2*>(&r));
  return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
}
template <uint32_t FMT>
__device__ __forceinline__ uint16_t to_operand(float v) {
  return (uint16_t)(pack2<FMT, false>(v, 0.f) & 0xffffu);
}
template <uint32_t FMT>
__device__ __forceinline__ float from_operand(uint16_t b) {
  if (FMT == FMT_F16) return __half2float(__ushort_as_half(b));
  return __bfloat162float(__ushort_as_bfloat16(b));
}

}  // namespace tc05
