// mini-PointNet patch Encoder + reduce_dim on tcgen05 tensor cores (sm_100a).
//
// Replaces Encoder.forward in eval mode (models/pointbert/dvae.py:201-215) and
// reduce_dim (models/pointbert/point_encoder.py:133,239).  The host folds BatchNorm,
// composes first_conv.3 with the per-point half of second_conv.0 and packs every
// weight into swizzled operand images (ppt_b200/encoder_pack.py).  Four launches:
//
//   stage1        per 128-point tile: h1 = relu(W1'x + b1') as one K = 16 tcgen05.mma (hi/lo parts along K;
//                 encoder_stage1_tc_kernel) ->
//                 tcgen05: W2 h1 -> max over each 32-point group -> g   [groups, 256]
//   group_linear  c = W3a' g + bias_c                                   [groups, 512] fp32
//   stage2        per tile: relu(W32 h1 + c) -> h3 (shared memory only) ->
//                 tcgen05: W4 h3 -> max over each group -> t            [groups, 256]
//   group_linear  tokens = Wr t + bias_tok                              [groups, 384] fp32
//
// Orientation: the WEIGHTS are the A operand (128 output channels = 128 TMEM lanes)
// and the ACTIVATIONS the B operand (points = TMEM columns).  A thread of the
// epilogue therefore owns one output channel: its bias is one register and the
// max over a 32-point group is a max over 32 of its own registers -- no shuffles.
// h1 is written by point-owning threads as a K-major operand (16-byte stores along the
// channels); h3 is written by channel-owning threads as an MN-major operand (16-byte
// stores along the points), each pair of values converted, ReLU'd and saturated by one
// F2FP instruction.
//
// Also here: the cls / pos_embed token assembly (pos_hidden_kernel, group_linear_kernel<ASSEMBLE>) and the
// train-mode BatchNorm path (bn_moments / bn_fold1 / Gram statistics / bn_fold2, encoder_stage_kernel<BN_APPLY>);
// DESIGN.md section 9.  (A cta_group::2 variant of stage 2 was built and measured in round 1 -- same time, more
// code -- and removed in round 2; DESIGN.md section 8 keeps what it showed.)
//
// Pipeline per CTA (persistent over tiles): warp 0 streams 16 KB weight images from
// L2 with 1-D bulk async copies into a ring (full/empty mbarriers); warp 1 issues
// tcgen05.mma into two alternating TMEM accumulators; warps 2-9 build h1, drain the
// accumulators (tcgen05.ld), apply bias/ReLU/max and write the next operand.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

// stage kernels: producer warp, MMA warp, EPW epilogue warps (8, or 16 where the epilogue side is the bottleneck)
constexpr int LIN_EPW = 4;                           // group_linear: producer, MMA, 4 epilogue warps (8 measured the same:
constexpr int LIN_THREADS = (LIN_EPW + 2) * 32;      // 39 vs 39 us -- the output stores are not what limits these kernels)
constexpr int LIN_EPI = LIN_EPW * 32;
constexpr int LIN_JPW = 4 / (LIN_EPW / 4);           // 32-column chunks of a 128-column accumulator per epilogue warp
constexpr uint32_t IMG = 16384;       // one operand image: 128 rows x 64 K x 2 B
constexpr uint32_t MNBLK = 65536;     // MN-major h3: bytes between 64-point blocks (64 K-atoms x 1 KB)

// ---- packed weight blob (ppt_b200/encoder_pack.py) -----------------------------------
struct BlobLayout {
  uint32_t split;
  __host__ __device__ uint32_t w1() const { return 0; }          // [128][4] fp32
  __host__ __device__ uint32_t bias_c() const { return 2048; }   // [512]
  __host__ __device__ uint32_t b4() const { return 4096; }       // [256]
  __host__ __device__ uint32_t bias_tok() const { return 5120; } // [384]
  // [8] powers of two: 1/scale of {stage1, linear c, stage2 W32, stage2 W4, linear tokens} accumulators,
  // scale of the point activations (h1, h3), scale of the group operands (g, t), unused
  __host__ __device__ uint32_t scales() const { return 6656; }
  __host__ __device__ uint32_t W2() const { return 8192; }                       // 2 units x 2 chunks
  __host__ __device__ uint32_t W3A() const { return W2() + 4 * split * IMG; }    // 4 x 4
  __host__ __device__ uint32_t W32() const { return W3A() + 16 * split * IMG; }  // 4 x 2
  __host__ __device__ uint32_t W4() const { return W32() + 8 * split * IMG; }    // 2 x 8
  __host__ __device__ uint32_t WR() const { return W4() + 16 * split * IMG; }    // 3 x 4
  __host__ __device__ uint32_t W1T() const { return WR() + 12 * split * IMG; }   // layer-1 image (1, unsplit)
  __host__ __device__ uint32_t W32F() const { return W1T() + IMG; }             // W32 as fp32 [512][128] (train statistics)
  __host__ __device__ uint32_t total() const { return W32F() + 512u * 128u * 4u; }
};

struct Ring {  // position in a ring of mbarrier-guarded stages
  uint32_t it = 0;
  template <int NSTAGE> __device__ uint32_t stage() const { return it % NSTAGE; }
  template <int NSTAGE> __device__ uint32_t parity() const { return (it / NSTAGE) & 1u; }
};

// One 64-wide K chunk = four K=16 instructions (x3 passes in the hi/lo split mode: hi*hi, hi*lo, lo*hi).
// The issuing thread's own instruction stream is the critical path of the tensor pipe, so descriptors
// are stepped as 32-bit words (tc05.cuh: sdesc_lo / sdesc_hi): `a_lo` / `b_lo` describe the first K=16
// slice of the chunk, each further slice adds a constant (in 16-byte units) to the start-address field.
//   A (weights, K-major):      +2 per slice (32 B),  lo part +IMG/16
//   B K-major  (h1, g, t):     +2 per slice
//   B MN-major (h3):           +128 per slice (two 1 KB K-atoms)
template <int SPLIT, bool B_MN>
__device__ __forceinline__ void issue_k64(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t b_split_step,
                                          uint32_t idesc, bool first) {
  constexpr uint32_t HI = sdesc_hi(1024u);
  constexpr uint32_t B_STEP = B_MN ? 128u : 2u;
#pragma unroll
  for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
    for (int pass = 0; pass < (SPLIT == 2 ? 3 : 1); ++pass) {
      const uint32_t sa = pass == 2 ? 1 : 0, sb = pass == 1 ? 1 : 0;
      const uint64_t ad = sdesc_join(a_lo + sa * (IMG >> 4) + (uint32_t)k16 * 2u, HI);
      const uint64_t bd = sdesc_join(b_lo + sb * b_split_step + (uint32_t)k16 * B_STEP, HI);
      umma_f16_elect(d_tmem, ad, bd, idesc, (first && k16 == 0 && pass == 0) ? 0u : 1u);
    }
  }
}

// Writes one value as operand element(s) (hi, and lo in split mode) at byte offset `off`.
template <uint32_t FMT, int SPLIT>
__device__ __forceinline__ void store_operand(unsigned char* base, uint32_t off, uint32_t split_bytes, float v) {
  const uint16_t hi = to_operand<FMT>(v);
  *reinterpret_cast<uint16_t*>(base + off) = hi;
  if (SPLIT == 2) *reinterpret_cast<uint16_t*>(base + split_bytes + off) = to_operand<FMT>(v - from_operand<FMT>(hi));
}

// Eight consecutive operand elements (16 bytes) from eight fp32 values, ReLU fused; lo parts in split mode.
template <uint32_t FMT, int SPLIT>
__device__ __forceinline__ void store_relu8(unsigned char* base, uint32_t off, uint32_t split_bytes, const float* v) {
  uint32_t hi[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) hi[t] = pack2<FMT, true>(v[2 * t], v[2 * t + 1]);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (SPLIT == 2) {
    uint32_t lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 h = unpack2<FMT>(hi[t]);
      lo[t] = pack2<FMT, false>(fmaxf(v[2 * t], 0.f) - h.x, fmaxf(v[2 * t + 1], 0.f) - h.y);
    }
    *reinterpret_cast<uint4*>(base + split_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// store_relu8 that also returns the sum of the eight values AS STORED (operand rounding included): the train-mode
// statistics are those of the activations the tensor core actually multiplies.
template <uint32_t FMT, int SPLIT>
__device__ __forceinline__ float store_relu8_sum(unsigned char* base, uint32_t off, uint32_t split_bytes, const float* v) {
  uint32_t hi[4];
  float s = 0.f;
  float2 h[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    hi[t] = pack2<FMT, true>(v[2 * t], v[2 * t + 1]);
    h[t] = unpack2<FMT>(hi[t]);
    s += h[t].x + h[t].y;
  }
  *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (SPLIT == 2) {
    uint32_t lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      lo[t] = pack2<FMT, false>(fmaxf(v[2 * t], 0.f) - h[t].x, fmaxf(v[2 * t + 1], 0.f) - h[t].y);
      const float2 l = unpack2<FMT>(lo[t]);
      s += l.x + l.y;
    }
    *reinterpret_cast<uint4*>(base + split_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  return s;
}

// Gram matrix of one 64-point chunk of h1: G[i][j] += sum_p h1[p][i] h1[p][j].  The MN-major h1 buffer (points =
// MN, channels = K) read with the roles swapped is a K-major operand with channels as rows and points as K (same
// bytes, same swizzle), so the chunk serves as BOTH operands of an M128 x N128 x K64 product.  Split mode:
// hi*hi + hi*lo + lo*hi like every other product.
template <int SPLIT>
__device__ __forceinline__ void issue_gram_k64(uint32_t d_tmem, uint32_t h_lo, uint32_t split_step, uint32_t idesc,
                                               bool first) {
  constexpr uint32_t HI = sdesc_hi(1024u);
#pragma unroll
  for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
    for (int pass = 0; pass < (SPLIT == 2 ? 3 : 1); ++pass) {
      const uint32_t sa = pass == 2 ? 1 : 0, sb = pass == 1 ? 1 : 0;
      const uint64_t ad = sdesc_join(h_lo + sa * split_step + (uint32_t)k16 * 2u, HI);
      const uint64_t bd = sdesc_join(h_lo + sb * split_step + (uint32_t)k16 * 2u, HI);
      umma_f16_elect(d_tmem, ad, bd, idesc, (first && k16 == 0 && pass == 0) ? 0u : 1u);
    }
  }
}

// ======================================================================================
// stage kernels
// ======================================================================================
// BN (stage 2 only; train-mode BatchNorm of second_conv.1, DESIGN.md "train mode"):
//   BN_EVAL   the blob's W32 / c carry the folded running statistics (the inference path);
//   BN_APPLY  training step: h3 = relu(s[ch] * y + t[ch]) with s = bn_vec[ch], t = bn_vec[512 + ch] from the
//             batch statistics -- folded into the per-unit accumulator scale and the prefetched c values, so
//             the inner loop is the same single FFMA per element as BN_EVAL.
constexpr int BN_EVAL = 0, BN_APPLY = 2;

// CLK: measurement build of the same kernel (ppt_set_clock_trace).  A separate instantiation because even these few
// instructions at kernel entry / exit changed ptxas's schedule of the MMA-issue loop and cost 7 % (0.59 -> 0.63 ms).
template <uint32_t FMT, int SPLIT, int NT, int STAGE, int EPW, int BN = BN_EVAL, bool CLK = false>
__global__ void __launch_bounds__((EPW + 2) * 32, 1)
encoder_stage_kernel(const float* __restrict__ nbhd, const unsigned char* __restrict__ blob,
                     const float* __restrict__ cbuf,          // stage 2: [groups_pad, 512]
                     unsigned char* __restrict__ out_img,     // stage 1: g images, stage 2: t images
                     float* __restrict__ features_out,        // stage 2, nullable: [groups, 256]
                     long long num_groups, int num_tiles,
                     double* __restrict__ /*unused*/ = nullptr, const float* __restrict__ bn_vec = nullptr,
                     long long* __restrict__ clock_acc = nullptr) {
  static_assert(BN == BN_EVAL || STAGE == 2, "batch statistics belong to stage 2");

  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  constexpr int GPT = NT / 32;                       // groups per tile
  constexpr int NUNITS = STAGE == 1 ? 2 : 6;
  constexpr int NH1 = STAGE == 1 ? 2 : 1;            // h1 buffers (stage 1 builds one tile ahead)
  constexpr uint32_t H1_BYTES = 2u * NT * 128u;      // one split part of one buffer: 2 K-chunks, K-major
  constexpr uint32_t H1_BUF = SPLIT * H1_BYTES;
  constexpr uint32_t H3_BYTES = STAGE == 2 ? (NT / 64) * MNBLK : 0u;  // one split part, MN-major
  constexpr uint32_t STAGE_BYTES = SPLIT * IMG;
  // Accumulators in rotation (unit g of the CTA's unit sequence uses accumulator g % NACC).  Stage 2 keeps FOUR: the
  // round trip accumulator full -> epilogue wakes -> tcgen05.ld -> convert -> proxy fence -> arrive -> issuer wakes is
  // ~900 cycles of pure latency, longer than one W32 unit runs (~740), so with two accumulators every W32 unit waited
  // for the epilogue of the unit before last; with four, the four W32 units of a tile issue back to back.
  constexpr int NACC = STAGE == 2 ? 4 : 2;
  constexpr int TCOLS = NACC * NT;
  constexpr int EPI_THREADS = EPW * 32;
  constexpr int CPW = NT / (EPW / 4);                // accumulator columns per epilogue thread
  constexpr int GH = CPW / 32;                       // groups per epilogue thread
  static_assert(EPW % 4 == 0 && CPW >= 32, "each TMEM lane quadrant needs EPW/4 warps of >= 32 columns");

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* h1buf = smem;                                   // [NH1][SPLIT][2][NT x 128 B]
  unsigned char* h3buf = h1buf + NH1 * H1_BUF;                   // [SPLIT][NT/64][64][1 KB]
  unsigned char* ring = h3buf + SPLIT * H3_BYTES;                // [NSTAGE][SPLIT][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE_BYTES);
  uint64_t* full = bars;                   // [NSTAGE]
  uint64_t* empty = full + NSTAGE;         // [NSTAGE]
  uint64_t* acc_full = empty + NSTAGE;     // [NACC]
  uint64_t* acc_empty = acc_full + NACC;   // [NACC]
  uint64_t* h1_ready = acc_empty + NACC;   // [2]
  uint64_t* h3_ready = h1_ready + 2;       // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h3_ready + 4);
  static_assert((2 * 4 + 2 * NACC + 2 + 4) * 8 + 4 <= 192, "barrier block");
  float4* w1s = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(bars) + 256);  // [128]
  long long* clk0 = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(bars) + 192);  // [2], see below

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  const BlobLayout L{(uint32_t)SPLIT};
  const float* sc = reinterpret_cast<const float*>(blob + L.scales());
  // measurement aid (ppt_set_clock_trace): CTA 0 adds its lifetime in ns and in SM cycles to clock_acc[0..1];
  // the start values wait in shared memory so that no register stays live across the kernel
  if (CLK && clock_acc && blockIdx.x == 0 && tid == 0) {
    long long ns0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    clk0[0] = ns0;
    clk0[1] = clock64();
  }

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS); }
    for (int i = 0; i < 2; ++i) mbar_init(&h1_ready[i], EPI_THREADS);
    for (int i = 0; i < 4; ++i) mbar_init(&h3_ready[i], EPI_THREADS);
    mbar_fence_init();
  }
  if (tid < 128) {  // W1' rows pre-multiplied by the activation scale (a power of two: exact)
    float4 w = __ldg(reinterpret_cast<const float4*>(blob + L.w1()) + tid);
    const float s = __ldg(sc + 5);
    w.x *= s; w.y *= s; w.z *= s; w.w *= s;
    w1s[tid] = w;
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      Ring r;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int u = 0; u < NUNITS; ++u) {
          const bool g3 = STAGE == 2 && u >= 4;
          const int nkc = g3 ? 8 : 2;
          const uint32_t sec = STAGE == 1 ? L.W2() : (g3 ? L.W4() : L.W32());
          const int blk = g3 ? u - 4 : u;
          for (int kc = 0; kc < nkc; ++kc) {
            const uint32_t s = r.stage<NSTAGE>();
            mbar_wait_relaxed(&empty[s], r.parity<NSTAGE>() ^ 1u);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(ring + s * STAGE_BYTES, blob + sec + (size_t)(blk * nkc + kc) * STAGE_BYTES, STAGE_BYTES, &full[s]);
            ++r.it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, one elected lane issues =====================
    {
      const uint32_t idesc_k = make_idesc(FMT, 128, NT, 0), idesc_mn = make_idesc(FMT, 128, NT, 1);
      const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u);            // stage s: + s * STAGE_BYTES/16
      const uint32_t h1_lo0 = sdesc_lo(smem_u32(h1buf), 16u);          // K-major: chunk kc: + kc * NT*128/16
      const uint32_t h3_lo0 = sdesc_lo(smem_u32(h3buf), MNBLK);        // MN-major: chunk kc: + kc * 8 KB/16
      uint32_t it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait(&h1_ready[tile_it % NH1], (tile_it / NH1) & 1u);
        const uint32_t h1_lo = h1_lo0 + (tile_it % NH1) * (H1_BUF >> 4);
#pragma unroll
        for (int u = 0; u < NUNITS; ++u) {
          const bool g3 = STAGE == 2 && u >= 4;
          const int nkc = g3 ? 8 : 2;
          // unit g of this CTA's sequence: accumulator g % NACC, whose n-th use (n = g / NACC) has parity n & 1
          const uint32_t g = tile_it * NUNITS + (uint32_t)u;
          const uint32_t buf = g % NACC, use = g / NACC;
          mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);
          fence_after_sync();
          const uint32_t d_tmem = tbase + (uint32_t)(buf * NT);
#pragma unroll
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            if (STAGE == 2 && u == 4 && (kc & 1) == 0) {  // h3 channels of K-chunks 2i, 2i+1 come from unit i
              mbar_wait(&h3_ready[kc >> 1], tile_it & 1u);
            }
            const uint32_t s = it % NSTAGE;
            mbar_wait(&full[s], (it / NSTAGE) & 1u);
            fence_after_sync();
            const uint32_t a_lo = a_lo0 + s * (STAGE_BYTES >> 4);
            if (g3)
              issue_k64<SPLIT, true>(d_tmem, a_lo, h3_lo0 + (uint32_t)kc * 512u, H3_BYTES >> 4, idesc_mn, kc == 0);
            else
              issue_k64<SPLIT, false>(d_tmem, a_lo, h1_lo + (uint32_t)kc * ((NT * 128u) >> 4), H1_BYTES >> 4,
                                      idesc_k, kc == 0);
            umma_commit_elect(&empty[s]);
          }
          umma_commit_elect(&acc_full[buf]);
        }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. EPW+1) =====================
    const int e = tid - 64;                 // 0 .. EPI_THREADS-1
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;       // which slice (CPW columns) of the accumulator
    const int m = quad * 32 + lane;         // output-channel row inside a 128-row unit
    const int col0 = half * CPW;
    const float* bias_b4 = reinterpret_cast<const float*>(blob + L.b4());
    const float inv_p1 = __ldg(sc + 0), inv_g3 = __ldg(sc + 3);
    const float act_scale = __ldg(sc + 5), grp_scale = __ldg(sc + 6);
    const float inv_p2s = __ldg(sc + 2) * act_scale;  // accumulator -> scaled activation, one FFMA per element
    // per-unit accumulator scale and shift of this thread's channels (BN_APPLY: batch-statistics BatchNorm)
    float u_scale[4], u_shift[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      u_scale[u] = BN == BN_APPLY ? __ldg(bn_vec + u * 128 + m) : 1.f;
      u_shift[u] = BN == BN_APPLY ? __ldg(bn_vec + 512 + u * 128 + m) : 0.f;
    }

    // This thread's point of a tile, fetched one build ahead: the load comes from HBM (the neighbourhoods
    // are read here for the first time) and would otherwise stall the first FFMA of every build.
    constexpr int TPP = EPI_THREADS / NT;            // threads per point
    constexpr int CH = 128 / TPP;                    // channels per thread
    const int p = e % NT, ch0 = (e / NT) * CH;       // ch0 is warp-uniform: w1s reads broadcast
    float nx = 0.f, ny = 0.f, nz = 0.f;              // coordinates for the NEXT build_h1 call
    auto fetch_point = [&](int tile) {
      nx = ny = nz = 0.f;
      const long long gp = (long long)tile * NT + p;
      if (tile < num_tiles && gp < num_groups * 32) {
        const float* src = nbhd + gp * 3;
        nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
      }
    };
    // builds h1 of `tile` from the prefetched point, then prefetches the point of `tile_after`
    auto build_h1 = [&](int tile, uint32_t slot, int tile_after) {
      // h1[p][ch] = relu(w.x + b), K = 3 on CUDA cores; written as the K-major B operand.
      unsigned char* dst = h1buf + slot * H1_BUF;
      const float x = nx, y = ny, z = nz;
      fetch_point(tile_after);
      (void)tile;
#pragma unroll 4
      for (int c8 = 0; c8 < CH; c8 += 8) {
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 w = w1s[ch0 + c8 + t];
          v[t] = fmaf(w.z, z, fmaf(w.y, y, fmaf(w.x, x, w.w)));
        }
        const int ch = ch0 + c8;
        store_relu8<FMT, SPLIT>(dst, (uint32_t)(ch >> 6) * (NT * 128u) + sw128_kmajor_off(p, ch & 63), H1_BYTES, v);
      }
      fence_proxy_async_smem();
      mbar_arrive(&h1_ready[slot]);
    };

    float cnext[4][GH];
    auto load_c = [&](int tile) {
#pragma unroll
      for (int u = 0; u < (STAGE == 2 ? 4 : 0); ++u)
#pragma unroll
        for (int jj = 0; jj < GH; ++jj) {
          const long long g = (long long)tile * GPT + half * GH + jj;
          const float cv = g < num_groups ? __ldg(cbuf + g * 512 + u * 128 + m) : 0.f;
          cnext[u][jj] = (BN == BN_APPLY ? fmaf(cv, u_scale[u], u_shift[u]) : cv) * act_scale;
        }
    };

    if ((int)blockIdx.x < num_tiles) {
      fetch_point(blockIdx.x);
      load_c(blockIdx.x);
      build_h1(blockIdx.x, 0, blockIdx.x + gridDim.x);
      if (STAGE == 1 && (int)(blockIdx.x + gridDim.x) < num_tiles)
        build_h1(blockIdx.x + gridDim.x, 1, blockIdx.x + 2 * gridDim.x);
    }

    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const long long g0 = (long long)tile * GPT + half * GH;  // first group this thread touches
      float ccur[4][GH];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int jj = 0; jj < GH; ++jj) ccur[u][jj] = STAGE == 2 ? cnext[u][jj] : 0.f;

#pragma unroll
      for (int u = 0; u < NUNITS; ++u) {
        const uint32_t gu = tile_it * NUNITS + (uint32_t)u;
        const uint32_t buf = gu % NACC;
        const bool relu_unit = STAGE == 2 && u < 4;
        mbar_wait(&acc_full[buf], (gu / NACC) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * NT + col0);

        if (relu_unit) {
          // h3[p][ch] = relu(acc + c[group][ch]); ch = u*128 + m is this thread's K index (MN-major B).
          const int ch = u * 128 + m;
          const float a_unit = BN == BN_APPLY ? inv_p2s * u_scale[u < 4 ? u : 0] : inv_p2s;
          const uint32_t krow = (uint32_t)(ch >> 3) * 1024u + (uint32_t)(ch & 7) * 128u;
          const uint32_t sw = (uint32_t)(ch & 7);
          // both 32-column loads in flight before the first use: one tensor-memory round trip per unit, not two
          uint32_t raw[GH][32];
#pragma unroll
          for (int jj = 0; jj < GH; ++jj) tmem_ld32_async(t_addr + jj * 32, raw[jj]);
          tmem_wait_ld();
#pragma unroll
          for (int jj = 0; jj < GH; ++jj) {
            float v[32];
            const float cv = ccur[u < 4 ? u : 0][jj];
            // packed f32x2 FMA (FFMA2): half the instructions of this issue-paced epilogue's arithmetic
            const float2 a2 = make_float2(a_unit, a_unit), c2 = make_float2(cv, cv);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 y = __ffma2_rn(make_float2(__uint_as_float(raw[jj][i]), __uint_as_float(raw[jj][i + 1])), a2, c2);
              v[i] = y.x;
              v[i + 1] = y.y;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int n = col0 + jj * 32 + q * 8;
              const uint32_t off = (uint32_t)(n >> 6) * MNBLK + krow + ((((uint32_t)(n & 63) >> 3) ^ sw) << 4);
              store_relu8<FMT, SPLIT>(h3buf, off, H3_BYTES, v + q * 8);
            }
          }
          fence_proxy_async_smem();
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          mbar_arrive(&h3_ready[u & 3]);
        } else {
          // per-group max over the 32 points (columns) this thread holds for its channel
          const int blk = STAGE == 1 ? u : u - 4;
          const int ch = blk * 128 + m;  // 0..255
          const float inv = STAGE == 1 ? inv_p1 : inv_g3;
#pragma unroll
          for (int jj = 0; jj < GH; ++jj) {
            float v[32];
            tmem_ld32(t_addr + jj * 32, v);
            float mx = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
            mx *= inv;  // exact: power of two
            const long long g = g0 + jj;
            if (g < num_groups) {
              // operand image for group_linear: tile of 128 groups, K = 256 -> 4 chunks
              const size_t img = ((size_t)(g >> 7) * 4 + (size_t)(ch >> 6)) * (SPLIT * IMG);
              store_operand<FMT, SPLIT>(out_img + img, sw128_kmajor_off((int)(g & 127), ch & 63), IMG,
                                        mx * grp_scale);
              if (STAGE == 2 && features_out) features_out[g * 256 + ch] = mx + __ldg(bias_b4 + ch);
            }
          }
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
        }

        // Stage 2: every MMA that reads h1 has completed once unit 3's accumulator is full -> rebuild it
        // for the next tile while the tensor pipe runs the W4 units; fetch that tile's c values too.
        if (STAGE == 2 && u == 3) {
          const int next = tile + gridDim.x;
          if (next < num_tiles) {
            load_c(next);
            build_h1(next, 0, next + gridDim.x);
          }
        }
      }
      // Stage 1: both units of this tile are drained, so its h1 buffer is free for tile + 2.
      if (STAGE == 1) {
        const int next2 = tile + 2 * gridDim.x;
        if (next2 < num_tiles) build_h1(next2, tile_it & 1u, next2 + gridDim.x);
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
  if (CLK && clock_acc && blockIdx.x == 0 && threadIdx.x == 0) {
    long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    atomicAdd(reinterpret_cast<unsigned long long*>(clock_acc), (unsigned long long)(ns1 - clk0[0]));
    atomicAdd(reinterpret_cast<unsigned long long*>(clock_acc) + 1, (unsigned long long)(clock64() - clk0[1]));
  }
}

// ======================================================================================
// stage 1 with layer 1 on the tensor core as well
// ======================================================================================
// Same result as encoder_stage_kernel<STAGE 1>, but h1 = relu(W1'x + b1') is ONE K=16 tcgen05.mma per
// tile instead of 1500 FFMA + 500 LDS per tile on the CUDA cores (which bounded that kernel): the blob's
// W1T image holds [W_hi | W_hi | W_lo | b_hi | b_lo] per channel along K and the epilogue warps write
// [x_hi | x_lo | x_hi | 1 | 1] per point (encoder_pack.layer1_image), so the fp32 accumulator equals the
// fp32 product up to the dropped W_lo.x_lo term.  The accumulator (channels = lanes, points = columns) is
// drained by channel-owning threads, so h1 becomes an MN-major operand like h3.
//
// Software pipeline per CTA (tile index n): the MMA warp issues L1(n+1) BEFORE the two W2 units of tile n,
// and the epilogue warps convert L1(n+1) while those units run:
//   MMA      : L1(n+1) | unit0(n) unit1(n)
//   epilogue : convert(n+1) -> h1[(n+1)%2] | point rows X(n+2) | max-epilogue unit0(n), unit1(n)
//
// TRAIN (model.train(), batch-statistics BatchNorm of second_conv.1): the same pass also produces what the
// statistics of y = W32 h1 + c need, so that no extra pass over the points exists:
//   * the Gram matrix G = sum_p h1_p h1_p^T of this CTA's points, accumulated on the tensor core in a fourth
//     accumulator (128 x 128 fp32 in tensor memory for the CTA's whole lifetime; the h1 buffer doubles as both
//     operands) and written once at the end to gram_out[blockIdx.x];
//   * the per-group MEAN of h1 (sum of the operand values as stored / 32) as operand images s_img for the
//     small per-group GEMM W32 s (group_stats_kernel).
// Padding points (beyond the last group) get all-zero operand rows (no bias slot), so their h1 is exactly 0.
template <uint32_t FMT, int SPLIT, int NT, int EPW, bool TRAIN = false>
__global__ void __launch_bounds__((EPW + 2) * 32, 1)
encoder_stage1_tc_kernel(const float* __restrict__ nbhd, const unsigned char* __restrict__ blob,
                         unsigned char* __restrict__ out_img, long long num_groups, int num_tiles,
                         unsigned char* __restrict__ s_img, float* __restrict__ gram_out) {
  constexpr int GPT = NT / 32;
  constexpr int EPI_THREADS = EPW * 32;
  constexpr int CPW = NT / (EPW / 4);
  constexpr int GH = CPW / 32;
  static_assert(EPW % 4 == 0 && CPW >= 32, "each TMEM lane quadrant needs EPW/4 warps of >= 32 columns");
  constexpr uint32_t XBYTES = NT * 128u;                       // one point-row operand: NT rows x 128 B (32 B used)
  constexpr uint32_t H1BLK = 16384u;                           // MN-major h1: 16 K-atoms x 1 KB per 64 points
  constexpr uint32_t H1_BYTES = (NT / 64) * H1BLK;             // one split part of one buffer
  constexpr uint32_t H1_BUF = SPLIT * H1_BYTES;
  constexpr uint32_t STAGE_BYTES = SPLIT * IMG;
  constexpr int TCOLS = (TRAIN || NT >= 128) ? 512 : 256;      // two unit accumulators + layer 1 (+ the Gram matrix)

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* w1img = smem;                                   // [16 KB]
  unsigned char* xbuf = w1img + IMG;                             // [2][XBYTES]
  unsigned char* h1buf = xbuf + 2 * XBYTES;                      // [2][SPLIT][H1_BYTES]
  unsigned char* w2img = h1buf + 2 * H1_BUF;                     // [4][SPLIT][16 KB]: all of W2, resident
  uint64_t* bars = reinterpret_cast<uint64_t*>(w2img + 4 * STAGE_BYTES);
  uint64_t* acc_full = bars;               // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* h1_ready = acc_empty + 2;      // [2]
  uint64_t* x_ready = h1_ready + 2;        // [2]
  uint64_t* l1_full = x_ready + 2;         // [1]
  uint64_t* l1_empty = l1_full + 1;        // [1]
  uint64_t* w1_full = l1_empty + 1;        // [1]
  uint64_t* gram_full = w1_full + 1;       // [1] TRAIN: every Gram product of this CTA is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gram_full + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  const BlobLayout L{(uint32_t)SPLIT};
  const float* sc = reinterpret_cast<const float*>(blob + L.scales());

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&h1_ready[i], EPI_THREADS); mbar_init(&x_ready[i], EPI_THREADS); }
    mbar_init(l1_full, 1);
    mbar_init(l1_empty, EPI_THREADS);
    mbar_init(w1_full, 1);
    mbar_init(gram_full, 1);
    mbar_fence_init();
  }
  // the unused 96 bytes of every point row must be finite zeros (they sit in K slots the MMA never reads,
  // but keep the buffer defined); done once
  for (int i = tid; i < (int)(2 * XBYTES / 16); i += (EPW + 2) * 32)
    reinterpret_cast<uint4*>(xbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t l1_tmem = tbase + 2u * NT;
  const uint32_t gram_tmem = tbase + 3u * NT;  // TRAIN: 128 columns

  if (warp == 0) {
    // ===================== producer: every weight of this kernel, once =====================
    // W2 is 4 images (64 KB): it stays resident, so nothing is streamed per tile (a 4-stage ring
    // cannot cover the ~1600-cycle latency of a bulk copy at this kernel's consumption rate).
    if (lane == 0) {
      mbar_arrive_expect_tx(w1_full, IMG + 4 * STAGE_BYTES);
      bulk_g2s(w1img, blob + L.W1T(), IMG, w1_full);
      for (int c = 0; c < 4; ++c)  // unit u = c / 2, K chunk kc = c % 2: images are consecutive in the blob
        bulk_g2s(w2img + c * STAGE_BYTES, blob + L.W2() + (size_t)c * STAGE_BYTES, STAGE_BYTES, w1_full);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, one elected lane issues =====================
    const uint32_t idesc_l1 = make_idesc(FMT, 128, NT, 0), idesc_mn = make_idesc(FMT, 128, NT, 1);
    constexpr uint32_t HI = sdesc_hi(1024u);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(w2img), 16u);
    const uint32_t w1_lo = sdesc_lo(smem_u32(w1img), 16u), x_lo0 = sdesc_lo(smem_u32(xbuf), 16u);
    const uint32_t h1_lo0 = sdesc_lo(smem_u32(h1buf), H1BLK);
    auto issue_l1 = [&](uint32_t n) {  // layer 1 of tile index n: one K=16 slice
      mbar_wait(&x_ready[n & 1u], (n >> 1) & 1u);
      mbar_wait(l1_empty, (n & 1u) ^ 1u);  // the n-th use waits for the drain of use n-1
      fence_after_sync();
      umma_f16_elect(l1_tmem, sdesc_join(w1_lo, HI), sdesc_join(x_lo0 + (n & 1u) * (XBYTES >> 4), HI), idesc_l1, 0u);
      umma_commit_elect(l1_full);
    };
    mbar_wait(w1_full, 0);
    uint32_t n = 0;
    if ((int)blockIdx.x < num_tiles) issue_l1(0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      if (tile + (int)gridDim.x < num_tiles) issue_l1(n + 1);
      mbar_wait(&h1_ready[n & 1u], (n >> 1) & 1u);
      const uint32_t h1_lo = h1_lo0 + (n & 1u) * (H1_BUF >> 4);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        mbar_wait(&acc_empty[u], (n & 1u) ^ 1u);
        fence_after_sync();
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
          issue_k64<SPLIT, true>(tbase + (uint32_t)(u * NT), a_lo0 + (uint32_t)(u * 2 + kc) * (STAGE_BYTES >> 4),
                                 h1_lo + (uint32_t)kc * 512u, H1_BYTES >> 4, idesc_mn, kc == 0);
        umma_commit_elect(&acc_full[u]);
        if (TRAIN && u == 0) {
          // Gram products of this tile, between the two units: unit 1's commit (acc_full[1]) then also covers
          // them, and that is what the epilogue warps wait for before this h1 buffer is overwritten
          const uint32_t g_lo = sdesc_lo(smem_u32(h1buf) + (n & 1u) * H1_BUF, 16u);
#pragma unroll
          for (int kc = 0; kc < NT / 64; ++kc)
            issue_gram_k64<SPLIT>(gram_tmem, g_lo + (uint32_t)kc * (H1BLK >> 4), H1_BYTES >> 4,
                                  make_idesc(FMT, 128, 128, 0), n == 0 && kc == 0);
        }
      }
    }
    if (TRAIN) umma_commit_elect(gram_full);
  } else {
    // ===================== epilogue warps =====================
    const int e = tid - 64;
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const int m = quad * 32 + lane;      // channel
    const int col0 = part * CPW;
    const float inv_p1 = __ldg(sc + 0), grp_scale = __ldg(sc + 6);

    float nx = 0.f, ny = 0.f, nz = 0.f;  // this thread's point for the NEXT build_x (threads e < NT)
    bool nvalid = false;
    auto fetch_point = [&](int tile) {
      nx = ny = nz = 0.f;
      nvalid = false;
      const long long gp = (long long)tile * NT + e;
      if (e < NT && tile < num_tiles && gp < num_groups * 32) {
        const float* src = nbhd + gp * 3;
        nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
        nvalid = true;
      }
    };
    // point rows of tile index n: [x_hi y_hi z_hi | x_lo y_lo z_lo | x_hi y_hi z_hi | 1 | 1 | 0 ...]
    auto build_x = [&](uint32_t n, int tile_after) {
      if (e < NT) {
        const uint32_t hxy = pack2<FMT, false>(nx, ny), hz1 = pack2<FMT, false>(nz, 1.0f);
        const float2 fxy = unpack2<FMT>(hxy), fz = unpack2<FMT>(hz1);
        const uint32_t lxy = pack2<FMT, false>(nx - fxy.x, ny - fxy.y), lz = pack2<FMT, false>(nz - fz.x, 0.f);
        // TRAIN: a padding point gets no bias slot either, so its h1 is exactly zero (Gram matrix, group sums)
        const uint32_t hx = hxy & 0xffffu, hy = hxy >> 16, hz = hz1 & 0xffffu, one = (TRAIN && !nvalid) ? 0u : hz1 >> 16;
        const uint32_t lx = lxy & 0xffffu, ly = lxy >> 16, lzz = lz & 0xffffu;
        // k: 0 xh 1 yh | 2 zh 3 xl | 4 yl 5 zl | 6 xh 7 yh || 8 zh 9 one | 10 one 11 0 | 0 | 0
        const uint4 p0 = make_uint4(hx | (hy << 16), hz | (lx << 16), ly | (lzz << 16), hx | (hy << 16));
        const uint4 p1 = make_uint4(hz | (one << 16), one, 0u, 0u);
        unsigned char* row = xbuf + (n & 1u) * XBYTES + (uint32_t)e * 128u;
        *reinterpret_cast<uint4*>(row + ((0u ^ (uint32_t)(e & 7)) << 4)) = p0;
        *reinterpret_cast<uint4*>(row + ((1u ^ (uint32_t)(e & 7)) << 4)) = p1;
      }
      fetch_point(tile_after);
      fence_proxy_async_smem();
      mbar_arrive(&x_ready[n & 1u]);
    };
    // layer-1 accumulator of tile index n -> relu -> h1[n % 2] (MN-major: this thread's channel is the K row)
    auto convert = [&](uint32_t n) {
      mbar_wait(l1_full, n & 1u);
      fence_after_sync();
      unsigned char* dst = h1buf + (n & 1u) * H1_BUF;
      const uint32_t krow = (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u, sw = (uint32_t)(m & 7);
      const uint32_t t_addr = l1_tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
#pragma unroll
      for (int jj = 0; jj < GH; ++jj) {
        float v[32];
        tmem_ld32(t_addr + jj * 32, v);
        float gsum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int pt = col0 + jj * 32 + q * 8;
          const uint32_t off = (uint32_t)(pt >> 6) * H1BLK + krow + ((((uint32_t)(pt & 63) >> 3) ^ sw) << 4);
          if (TRAIN) gsum += store_relu8_sum<FMT, SPLIT>(dst, off, H1_BYTES, v + q * 8);
          else store_relu8<FMT, SPLIT>(dst, off, H1_BYTES, v + q * 8);
        }
        if (TRAIN) {
          // group mean of channel m (in act_scale units, like h1 itself): K-major operand image, K = 128 -> 2 chunks
          const long long g = ((long long)blockIdx.x + (long long)n * gridDim.x) * GPT + part * GH + jj;
          if (g < num_groups) {
            const size_t img = ((size_t)(g >> 7) * 2 + (size_t)(m >> 6)) * (SPLIT * IMG);
            store_operand<FMT, SPLIT>(s_img + img, sw128_kmajor_off((int)(g & 127), m & 63), IMG, gsum * 0.03125f);
          }
        }
      }
      fence_proxy_async_smem();
      fence_before_sync();
      mbar_arrive(&h1_ready[n & 1u]);
      mbar_arrive(l1_empty);
    };

    const int t0 = blockIdx.x, stride = gridDim.x;
    if (t0 < num_tiles) {
      fetch_point(t0);
      build_x(0, t0 + stride);
      if (t0 + stride < num_tiles) build_x(1, t0 + 2 * stride);
      convert(0);
    }
    uint32_t n = 0;
    for (int tile = t0; tile < num_tiles; tile += stride, ++n) {
      if (tile + stride < num_tiles) convert(n + 1);
      // X(n) was consumed by L1(n), complete before convert(n): its slot is free for tile index n+2.
      // Built BEFORE the unit epilogues: L1(n+2) is the first thing the MMA warp issues next iteration.
      if (tile + 2 * stride < num_tiles) build_x(n + 2, tile + 3 * stride);
      const long long g0 = (long long)tile * GPT + part * GH;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        mbar_wait(&acc_full[u], n & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(u * NT + col0);
        const int ch = u * 128 + m;
#pragma unroll
        for (int jj = 0; jj < GH; ++jj) {
          float v[32];
          tmem_ld32(t_addr + jj * 32, v);
          float mx = v[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
          mx *= inv_p1;
          const long long g = g0 + jj;
          if (g < num_groups) {
            const size_t img = ((size_t)(g >> 7) * 4 + (size_t)(ch >> 6)) * (SPLIT * IMG);
            store_operand<FMT, SPLIT>(out_img + img, sw128_kmajor_off((int)(g & 127), ch & 63), IMG, mx * grp_scale);
          }
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[u]);
      }
    }
    if (TRAIN && t0 < num_tiles) {
      // this CTA's Gram matrix: row m (this thread's channel), columns [64 part, 64 part + 64)
      mbar_wait(gram_full, 0);
      fence_after_sync();
      float* grow = gram_out + ((size_t)blockIdx.x * 128 + m) * 128 + part * (128 / (EPW / 4));
#pragma unroll
      for (int jj = 0; jj < 128 / (EPW / 4) / 32; ++jj) {
        float v[32];
        tmem_ld32(gram_tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * (128 / (EPW / 4)) + jj * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) grow[jj * 32 + i] = v[i];
      }
      fence_before_sync();
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

// ======================================================================================
// group_linear: out[row(g)][o] = sum_k W[o][k] * act[g][k] + bias[o], act given as operand images
// (K = 64 KCH).  row(g) = g, or -- rows_per_cloud = G > 0, the token-assembly layout of
// models/pointbert/point_encoder.py:245-246 -- g + g / G + 1: every cloud's G rows follow one row that
// belongs to the cls token.  G >= 32 so that a run of 32 consecutive groups crosses at most one cloud boundary.
// ======================================================================================
//
// STATS (train-mode BatchNorm of second_conv.1): act = per-group means of h1 (s_img of
// encoder_stage1_tc_kernel<TRAIN>), W = W32, `bias` = the per-group term c [groups, 512].  Nothing is stored per
// group; instead, with t = W32 . mean_g(h1) for channel o, the sums over the group's 32 points of y = W32 h1 + c
//   sum y   = 32 (t + c),      sum y^2 = (Gram term, bn_fold2_kernel) + 64 c t + 32 c^2
// are accumulated per channel (fp32 within a 128-group tile, fp64 across) and added to stats[o] / stats[512 + o].
template <uint32_t FMT, int SPLIT, int NUNITS, int KCH = 4, bool ASSEMBLE = false, bool STATS = false>
__global__ void __launch_bounds__(LIN_THREADS, 1)
group_linear_kernel(const unsigned char* __restrict__ act_img, const unsigned char* __restrict__ wsec,
                    const float* __restrict__ bias, const float* __restrict__ inv_scale_ptr, float* __restrict__ out,
                    long long num_groups, int num_tiles, int rows_per_cloud, int out_f16) {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  constexpr int NT = 128;                      // groups per tile (MMA N)
  constexpr int NOUT = NUNITS * 128;
  constexpr uint32_t STAGE_BYTES = SPLIT * IMG;
  constexpr uint32_t B_BYTES = (uint32_t)KCH * SPLIT * IMG;  // [kc][split][16 KB]
  constexpr int TCOLS = 2 * NT;

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* bbuf = smem;
  unsigned char* ring = bbuf + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint64_t* b_empty = b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], LIN_EPI); }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      Ring r;
      uint32_t tile_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait_relaxed(b_empty, (tile_it & 1u) ^ 1u);
        mbar_arrive_expect_tx(b_full, B_BYTES);
        for (int kc = 0; kc < KCH; ++kc)
          bulk_g2s(bbuf + kc * STAGE_BYTES, act_img + ((size_t)tile * KCH + kc) * STAGE_BYTES, STAGE_BYTES, b_full);
#pragma unroll 1
        for (int u = 0; u < NUNITS; ++u)
          for (int kc = 0; kc < KCH; ++kc) {
            const uint32_t s = r.stage<NSTAGE>();
            mbar_wait_relaxed(&empty[s], r.parity<NSTAGE>() ^ 1u);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(ring + s * STAGE_BYTES, wsec + (size_t)(u * KCH + kc) * STAGE_BYTES, STAGE_BYTES, &full[s]);
            ++r.it;
          }
      }
    }
  } else if (warp == 1) {
    {  // converged warp, one elected lane issues
      const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
      const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u), b_lo0 = sdesc_lo(smem_u32(bbuf), 16u);
      uint32_t it = 0, tile_it = 0, unit_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait(b_full, tile_it & 1u);
#pragma unroll 1
        for (int u = 0; u < NUNITS; ++u, ++unit_it) {
          const int buf = unit_it & 1;
          mbar_wait(&acc_empty[buf], ((unit_it >> 1) & 1u) ^ 1u);
          fence_after_sync();
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc, ++it) {
            const uint32_t s = it % NSTAGE;
            mbar_wait(&full[s], (it / NSTAGE) & 1u);
            fence_after_sync();
            // activation image order is [kc][split]: the lo copy sits IMG bytes after the hi copy
            issue_k64<SPLIT, false>(tbase + (uint32_t)(buf * NT), a_lo0 + s * (STAGE_BYTES >> 4),
                                    b_lo0 + (uint32_t)kc * (STAGE_BYTES >> 4), IMG >> 4, idesc, kc == 0);
            umma_commit_elect(&empty[s]);
          }
          umma_commit_elect(&acc_full[buf]);
          if (u == NUNITS - 1) umma_commit_elect(b_empty);  // all MMAs reading this tile's activations are done
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const float inv_scale = __ldg(inv_scale_ptr);
    uint32_t unit_it = 0;
    // STATS: this thread's fp64 accumulators, one pair per unit, in its own shared-memory slots (the unit loop is not
    // unrolled -- unrolling it tripled the run time of every instantiation -- so they cannot live in registers)
    double* sacc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(bars) + 128) + (tid - 64) * 2;
    if (STATS)
      for (int u = 0; u < NUNITS; ++u) sacc[u * 2 * LIN_EPI] = sacc[u * 2 * LIN_EPI + 1] = 0.0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int u = 0; u < NUNITS; ++u, ++unit_it) {
        const int buf = unit_it & 1;
        const int o = u * 128 + m;
        const float bo = STATS ? 0.f : __ldg(bias + o);
        mbar_wait(&acc_full[buf], (unit_it >> 1) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * NT);
        float a1 = 0.f, a2 = 0.f;
        const int j0 = ((warp - 2) >> 2) * LIN_JPW;
#pragma unroll
        for (int j = j0; j < j0 + LIN_JPW; ++j) {
          float v[32];
          tmem_ld32(t_addr + j * 32, v);
          const long long g0 = (long long)tile * NT + j * 32;
          if (STATS) {
            // c [groups, 512]: per group a warp reads 128 contiguous bytes.  All 32 loads are issued before the first
            // use (they miss to L2 / HBM: one at a time this epilogue took 186 us for 134 MB)
            float cv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) cv[i] = g0 + i < num_groups ? __ldg(bias + (g0 + i) * 512 + o) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (g0 + i < num_groups) {
                const float t = v[i] * inv_scale;
                a1 += t + cv[i];
                a2 = fmaf(cv[i], fmaf(2.f, t, cv[i]), a2);
              }
          } else if (!ASSEMBLE && out_f16) {
            // 16-bit rows (PPT_TOKENS_F16): same values rounded once more to fp16 (saturating), half the bytes
            uint16_t* out16 = reinterpret_cast<uint16_t*>(out);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (g0 + i < num_groups) out16[(g0 + i) * NOUT + o] = to_operand<FMT_F16>(fmaf(v[i], inv_scale, bo));
          } else if (!ASSEMBLE) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (g0 + i < num_groups) out[(g0 + i) * NOUT + o] = fmaf(v[i], inv_scale, bo);
          } else {
            const long long q = g0 / rows_per_cloud;
            const int wrap = (int)((q + 1) * rows_per_cloud - g0);  // first i that belongs to the next cloud
            const long long left = num_groups - g0;
            const int limit = left < 32 ? (int)left : 32;
            float* p0 = out + (g0 + q + 1) * NOUT + o;  // rows of this cloud
            float* p1 = p0 + NOUT;                      // rows of the next one: skip its cls row
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < limit) (i < wrap ? p0 : p1)[i * NOUT] = fmaf(v[i], inv_scale, bo);
          }
        }
        if (STATS) { sacc[u * 2 * LIN_EPI] += (double)a1; sacc[u * 2 * LIN_EPI + 1] += (double)a2; }
        fence_before_sync();
        mbar_arrive(&acc_empty[buf]);
      }
    }
    if (STATS) {
      double* stats = reinterpret_cast<double*>(out);
      for (int u = 0; u < NUNITS; ++u) {
        atomicAdd(stats + u * 128 + m, 32.0 * sacc[u * 2 * LIN_EPI]);
        atomicAdd(stats + 512 + u * 128 + m, 32.0 * sacc[u * 2 * LIN_EPI + 1]);
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

// ======================================================================================
// train mode: c = W3a g + bias_c AND the second_conv.1 channel sums, one kernel (single-part operands)
// ======================================================================================
// group_linear(c) and group_linear<STATS> fused: per 128-group tile and 128-channel unit, two accumulators -- W3a . g
// (K = 256) and W32 . mean_g(h1) (K = 128) -- are complete at the same time, so the epilogue writes c and forms
// c t and c^2 from registers.  Run separately, the STATS kernel re-read the 134 MB of c it needs with 128-byte pieces
// of 2 KB rows and took 187 us; fused, the statistics cost no memory traffic at all.  Four accumulators (two pairs,
// 512 tensor-memory columns) so that the units of a tile still pipeline against the epilogue.
template <uint32_t FMT>
__global__ void __launch_bounds__(LIN_THREADS, 1)
group_c_stats_kernel(const unsigned char* __restrict__ g_img, const unsigned char* __restrict__ s_img,
                     const unsigned char* __restrict__ w3a_sec, const unsigned char* __restrict__ w32_sec,
                     const float* __restrict__ bias_c, const float* __restrict__ scales, float* __restrict__ cbuf,
                     double* __restrict__ stats, long long num_groups, int num_tiles) {
  constexpr int NSTAGE = 4, NT = 128, KG = 4, KS = 2;
  constexpr uint32_t B_BYTES = (uint32_t)(KG + KS) * IMG;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* bbuf = smem;                    // [g: 4 chunks | s: 2 chunks] x 16 KB
  unsigned char* ring = bbuf + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * IMG);
  uint64_t* full = bars;
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;   // [2] pair p: both accumulators of a unit complete
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint64_t* b_full = acc_empty + 2;
  uint64_t* b_empty = b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + 1);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], LIN_EPI); }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      Ring r;
      uint32_t tile_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait_relaxed(b_empty, (tile_it & 1u) ^ 1u);
        mbar_arrive_expect_tx(b_full, B_BYTES);
        for (int kc = 0; kc < KG; ++kc) bulk_g2s(bbuf + kc * IMG, g_img + ((size_t)tile * KG + kc) * IMG, IMG, b_full);
        for (int kc = 0; kc < KS; ++kc)
          bulk_g2s(bbuf + (KG + kc) * IMG, s_img + ((size_t)tile * KS + kc) * IMG, IMG, b_full);
#pragma unroll 1
        for (int u = 0; u < 4; ++u)
          for (int kc = 0; kc < KG + KS; ++kc) {
            const uint32_t s = r.stage<NSTAGE>();
            mbar_wait_relaxed(&empty[s], r.parity<NSTAGE>() ^ 1u);
            mbar_arrive_expect_tx(&full[s], IMG);
            const unsigned char* src = kc < KG ? w3a_sec + (size_t)(u * KG + kc) * IMG
                                               : w32_sec + (size_t)(u * KS + kc - KG) * IMG;
            bulk_g2s(ring + s * IMG, src, IMG, &full[s]);
            ++r.it;
          }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u), b_lo0 = sdesc_lo(smem_u32(bbuf), 16u);
    uint32_t it = 0, tile_it = 0, unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(b_full, tile_it & 1u);
#pragma unroll 1
      for (int u = 0; u < 4; ++u, ++unit_it) {
        const int p = unit_it & 1;
        mbar_wait(&acc_empty[p], ((unit_it >> 1) & 1u) ^ 1u);
        fence_after_sync();
#pragma unroll
        for (int kc = 0; kc < KG + KS; ++kc, ++it) {
          const uint32_t s = it % NSTAGE;
          mbar_wait(&full[s], (it / NSTAGE) & 1u);
          fence_after_sync();
          const uint32_t d = tbase + (uint32_t)((2 * p + (kc < KG ? 0 : 1)) * NT);
          issue_k64<1, false>(d, a_lo0 + s * (IMG >> 4), b_lo0 + (uint32_t)kc * (IMG >> 4), IMG >> 4, idesc,
                              kc == 0 || kc == KG);
          umma_commit_elect(&empty[s]);
        }
        umma_commit_elect(&acc_full[p]);
        if (u == 3) umma_commit_elect(b_empty);
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const float inv_c = __ldg(scales + 1), inv_t = __ldg(scales + 2);
    double* sacc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(bars) + 128) + (tid - 64) * 2;
    for (int u = 0; u < 4; ++u) sacc[u * 2 * LIN_EPI] = sacc[u * 2 * LIN_EPI + 1] = 0.0;
    uint32_t unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int u = 0; u < 4; ++u, ++unit_it) {
        const int p = unit_it & 1;
        const int o = u * 128 + m;
        const float bo = __ldg(bias_c + o);
        mbar_wait(&acc_full[p], (unit_it >> 1) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * p * NT);
        float a1 = 0.f, a2 = 0.f;
        const int j0 = ((warp - 2) >> 2) * LIN_JPW;
#pragma unroll 1
        for (int j = j0; j < j0 + LIN_JPW; ++j) {
          uint32_t rc[32], rt[32];
          tmem_ld32_async(t_addr + j * 32, rc);
          tmem_ld32_async(t_addr + NT + j * 32, rt);
          tmem_wait_ld();
          const long long g0 = (long long)tile * NT + j * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (g0 + i < num_groups) {
              const float c = fmaf(__uint_as_float(rc[i]), inv_c, bo);
              const float t = __uint_as_float(rt[i]) * inv_t;
              cbuf[(g0 + i) * 512 + o] = c;
              a1 += t + c;
              a2 = fmaf(c, fmaf(2.f, t, c), a2);
            }
        }
        sacc[u * 2 * LIN_EPI] += (double)a1;
        sacc[u * 2 * LIN_EPI + 1] += (double)a2;
        fence_before_sync();
        mbar_arrive(&acc_empty[p]);
      }
    }
    for (int u = 0; u < 4; ++u) {
      atomicAdd(stats + u * 128 + m, 32.0 * sacc[u * 2 * LIN_EPI]);
      atomicAdd(stats + 512 + u * 128 + m, 32.0 * sacc[u * 2 * LIN_EPI + 1]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}

// ======================================================================================
// pos_embed layer 1 + cls rows (models/pointbert/point_encoder.py:138-142, 241-247)
// ======================================================================================
// pos = Linear(128, 384)(GELU(Linear(3, 128)(center))).  The 3 -> 128 layer and the exact (erf) GELU run here
// in fp32 on CUDA cores, one thread per centre, and are written straight into the K-major operand images
// that group_linear<KCH = 2> multiplies with the packed 128 -> 384 weight; that kernel stores into rows
// 1..G of each cloud of pos_out.  Row 0 of each cloud (cls_token in x, cls_pos in pos) is written here.
// Blob: fp32 section {W1 rows {w0,w1,w2,b} [128][4] | b2 [384] | cls_token [384] | cls_pos [384] | scales [4] =
// 1/(weight scale * hidden scale), hidden scale, 0, 0} in 8192 bytes, then W2 images [3 units][2 chunks][split].
struct PosBlobLayout {
  __host__ __device__ static constexpr uint32_t w1() { return 0; }
  __host__ __device__ static constexpr uint32_t b2() { return 2048; }
  __host__ __device__ static constexpr uint32_t cls_token() { return 3584; }
  __host__ __device__ static constexpr uint32_t cls_pos() { return 5120; }
  __host__ __device__ static constexpr uint32_t scales() { return 6656; }
  __host__ __device__ static constexpr uint32_t W2() { return 8192; }
  __host__ __device__ static constexpr uint32_t total(uint32_t split) { return 8192 + 3 * 2 * split * IMG; }
};

template <uint32_t FMT, int SPLIT>
__global__ void __launch_bounds__(512)
pos_hidden_kernel(const float* __restrict__ center, const unsigned char* __restrict__ pblob,
                  unsigned char* __restrict__ h_img, float* __restrict__ x_out, float* __restrict__ pos_out,
                  long long num_groups, int rows_per_cloud) {
  // 512 threads per 128-centre tile: thread = (centre row, quarter of the 128 hidden channels)
  __shared__ float4 w1s[128];
  const int tid = threadIdx.x, row = tid & 127, quarter = tid >> 7;
  const float hscale = __ldg(reinterpret_cast<const float*>(pblob + PosBlobLayout::scales()) + 1);
  if (tid < 128) w1s[tid] = __ldg(reinterpret_cast<const float4*>(pblob + PosBlobLayout::w1()) + tid);
  __syncthreads();
  const long long tiles = (num_groups + 127) / 128;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long g = tile * 128 + row;
    if (g < num_groups) {
      const float x = __ldg(center + g * 3), y = __ldg(center + g * 3 + 1), z = __ldg(center + g * 3 + 2);
      unsigned char* img = h_img + (size_t)tile * 2 * SPLIT * IMG;
#pragma unroll 2
      for (int c8 = quarter * 32; c8 < quarter * 32 + 32; c8 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float v[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 w = w1s[c8 + 2 * t + h];
            const float a = fmaf(w.z, z, fmaf(w.y, y, fmaf(w.x, x, w.w)));
            v[h] = 0.5f * a * (1.f + erff(a * 0.70710678118654752440f)) * hscale;  // nn.GELU() (erf form)
          }
          hi[t] = pack2<FMT, false>(v[0], v[1]);
          if (SPLIT == 2) {
            const float2 r = unpack2<FMT>(hi[t]);
            lo[t] = pack2<FMT, false>(v[0] - r.x, v[1] - r.y);
          }
        }
        unsigned char* dst = img + (size_t)(c8 >> 6) * SPLIT * IMG + sw128_kmajor_off(row, c8 & 63);
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (SPLIT == 2) *reinterpret_cast<uint4*>(dst + IMG) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  // row 0 of every cloud
  const long long clouds = num_groups / rows_per_cloud;
  const float* cls_token = reinterpret_cast<const float*>(pblob + PosBlobLayout::cls_token());
  const float* cls_pos = reinterpret_cast<const float*>(pblob + PosBlobLayout::cls_pos());
  for (long long i = (long long)blockIdx.x * 512 + tid; i < clouds * 384; i += (long long)gridDim.x * 512) {
    const long long b = i / 384;
    const int o = (int)(i - b * 384);
    const long long dst = b * (rows_per_cloud + 1) * 384 + o;
    if (x_out) x_out[dst] = __ldg(cls_token + o);
    pos_out[dst] = __ldg(cls_pos + o);
  }
}

// ======================================================================================
// train-mode BatchNorm (models/pointbert/dvae.py:190,196 under model.train(), main_cls.py:169)
// ======================================================================================
// first_conv.1 normalises y = W x + b (3 -> 128), which is affine in the coordinates: its batch mean and biased
// variance follow exactly from the mean and covariance of the points, mean_y = W mu + b, var_y = W Sigma W^T.
// bn_moments_kernel reduces the nine first and second moments in fp64; bn_fold1_kernel folds the batch
// statistics into W1' (fp32 rows for stage 2 and the K = 16 layer-1 operand image for stage 1, the layout of
// encoder_pack.layer1_image) inside the caller's mutable blob and applies the momentum update to the
// running statistics.  second_conv.1 normalises y = W32 h1 + c_group behind a ReLU.  Its statistics need no pass of
// their own either: with G = sum_p h1 h1^T (accumulated on the tensor core inside stage 1) and the per-group
// means of h1,   sum y = 32 sum_g (W32 mean_g + c_g),   sum y^2 = w^T G w + sum_g (64 c_g (W32 mean_g) + 32 c_g^2)
// per channel (group_c_stats_kernel / group_linear_kernel<STATS>, bn_gram_reduce_kernel); bn_fold2_kernel turns the
// sums into per-channel scale / shift.  (Round 1 ran the four W32 units over every point a second time: +0.41 ms
// per 128-cloud step; this is +0.13 ms.)
__global__ void __launch_bounds__(256)
bn_moments_kernel(const float* __restrict__ nbhd, long long npoints, double* __restrict__ mom) {
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npoints; i += (long long)gridDim.x * blockDim.x) {
    const double x = __ldg(nbhd + i * 3), y = __ldg(nbhd + i * 3 + 1), z = __ldg(nbhd + i * 3 + 2);
    a[0] += x; a[1] += y; a[2] += z;
    a[3] += x * x; a[4] += x * y; a[5] += x * z; a[6] += y * y; a[7] += y * z; a[8] += z * z;
  }
  __shared__ double part[8][9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double v = a[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += part[w][threadIdx.x];
    atomicAdd(mom + threadIdx.x, v);
  }
}

struct BnDevice {  // device pointers of the two BatchNorm1d modules and the convolution in front of the first
  const float *conv1_w, *conv1_b;
  const float *bn1_w, *bn1_b;
  float *bn1_rm, *bn1_rv;
  long long* bn1_nbt;
  const float *bn2_w, *bn2_b;
  float *bn2_rm, *bn2_rv;
  long long* bn2_nbt;
  float momentum, eps;
};

template <uint32_t FMT>
__global__ void __launch_bounds__(128)
bn_fold1_kernel(BnDevice bn, const double* __restrict__ mom, long long npoints, unsigned char* __restrict__ blob,
                uint32_t split) {
  const BlobLayout L{split};
  const int ch = threadIdx.x;
  const double n = (double)npoints;
  const double mx = mom[0] / n, my = mom[1] / n, mz = mom[2] / n;
  const double cxx = mom[3] / n - mx * mx, cxy = mom[4] / n - mx * my, cxz = mom[5] / n - mx * mz;
  const double cyy = mom[6] / n - my * my, cyz = mom[7] / n - my * mz, czz = mom[8] / n - mz * mz;
  const double w0 = bn.conv1_w[ch * 3], w1 = bn.conv1_w[ch * 3 + 1], w2 = bn.conv1_w[ch * 3 + 2], b = bn.conv1_b[ch];
  const double mean = w0 * mx + w1 * my + w2 * mz + b;
  double var = w0 * w0 * cxx + w1 * w1 * cyy + w2 * w2 * czz + 2.0 * (w0 * w1 * cxy + w0 * w2 * cxz + w1 * w2 * cyz);
  var = var < 0.0 ? 0.0 : var;
  const double s = (double)bn.bn1_w[ch] / sqrt(var + (double)bn.eps);
  const float f[4] = {(float)(w0 * s), (float)(w1 * s), (float)(w2 * s), (float)((b - mean) * s + (double)bn.bn1_b[ch])};
  reinterpret_cast<float4*>(blob + L.w1())[ch] = make_float4(f[0], f[1], f[2], f[3]);
  // layer-1 operand image, row = channel: [W_hi(3) | W_hi(3) | W_lo(3) | b_hi | b_lo | 0 ...] of act_scale * W1'
  const float act = reinterpret_cast<const float*>(blob + L.scales())[5];
  uint16_t hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = f[k] * act;
    hi[k] = to_operand<FMT>(v);
    lo[k] = to_operand<FMT>(v - from_operand<FMT>(hi[k]));
  }
  unsigned char* img = blob + L.W1T();
  auto put = [&](int k, uint16_t v) { *reinterpret_cast<uint16_t*>(img + sw128_kmajor_off(ch, k)) = v; };
#pragma unroll
  for (int k = 0; k < 3; ++k) { put(k, hi[k]); put(3 + k, hi[k]); put(6 + k, lo[k]); }
  put(9, hi[3]);
  put(10, lo[3]);
  // running statistics: momentum update with the unbiased variance (torch.nn.BatchNorm1d)
  const double unbiased = npoints > 1 ? var * n / (n - 1.0) : var;
  bn.bn1_rm[ch] = (float)((1.0 - bn.momentum) * (double)bn.bn1_rm[ch] + (double)bn.momentum * mean);
  bn.bn1_rv[ch] = (float)((1.0 - bn.momentum) * (double)bn.bn1_rv[ch] + (double)bn.momentum * unbiased);
  if (ch == 0 && bn.bn1_nbt) *bn.bn1_nbt += 1;
}

// Sum of the per-CTA Gram matrices (fp32 partials of encoder_stage1_tc_kernel<TRAIN>) in fp64, scaled back from the
// activation scale: gram[i][j] = sum over every point of h1_i h1_j.
__global__ void __launch_bounds__(256)
bn_gram_reduce_kernel(const float* __restrict__ parts, int nparts, double inv_act2, double* __restrict__ gram) {
  const int e = blockIdx.x * 256 + threadIdx.x;  // 0 .. 16383
  double a = 0.0;
  for (int p = 0; p < nparts; ++p) a += (double)parts[(size_t)p * 16384 + e];
  gram[e] = a * inv_act2;
}

// One block per channel k of second_conv.1: sum_p y_k^2 = w_k^T G w_k (fp64, w = the exact fp32 W32 row) + the c terms
// already in stats[512 + k] (group_linear_kernel<STATS>); then scale / shift and the running statistics.
__global__ void __launch_bounds__(128)
bn_fold2_kernel(BnDevice bn, const double* __restrict__ stats, const double* __restrict__ gram,
                const float* __restrict__ w32f, long long npoints, float* __restrict__ bn_vec) {
  const int ch = blockIdx.x, i = threadIdx.x;
  const float* w = w32f + (size_t)ch * 128;
  double r = 0.0;
  for (int j = 0; j < 128; ++j) r += gram[j * 128 + i] * (double)w[j];  // G is symmetric: read it column-wise (coalesced)
  r *= (double)w[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  __shared__ double part[4];
  if ((i & 31) == 0) part[i >> 5] = r;
  __syncthreads();
  if (i != 0) return;
  const double quad = part[0] + part[1] + part[2] + part[3];
  const double n = (double)npoints;
  const double mean = stats[ch] / n;
  double var = (stats[512 + ch] + quad) / n - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const double s = (double)bn.bn2_w[ch] / sqrt(var + (double)bn.eps);
  bn_vec[ch] = (float)s;
  bn_vec[512 + ch] = (float)((double)bn.bn2_b[ch] - mean * s);
  const double unbiased = npoints > 1 ? var * n / (n - 1.0) : var;
  bn.bn2_rm[ch] = (float)((1.0 - bn.momentum) * (double)bn.bn2_rm[ch] + (double)bn.momentum * mean);
  bn.bn2_rv[ch] = (float)((1.0 - bn.momentum) * (double)bn.bn2_rv[ch] + (double)bn.momentum * unbiased);
  if (ch == 0 && bn.bn2_nbt) *bn.bn2_nbt += 1;
}

// ======================================================================================
// host side
// ======================================================================================
// Stage 2 runs 8 epilogue warps (64 accumulator columns per thread, 152 registers).  16 warps (32 columns each) were
// measured in round 2: the 576-thread CTA caps a thread at 96 registers, the epilogue spills, and the kernel slows from
// 0.56 to 0.75 ms -- the relu epilogues need registers more than they need warps.
template <int NT>
constexpr int stage2_epilogue_warps() { return 8; }
// Stage 1 is paced by its epilogue chain (layer-1 drain -> h1, two max epilogues per tile): 16 epilogue warps (32
// accumulator columns per thread, 79 registers) measured 0.126 ms against 0.136 ms with 8.
template <int NT>
constexpr int stage1_epilogue_warps() { return NT >= 128 ? 16 : 8; }

template <int SPLIT, int NT, int STAGE>
constexpr size_t stage_smem_bytes() {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  return (size_t)(STAGE == 1 ? 2 : 1) * SPLIT * (2u * NT * 128u) + (STAGE == 2 ? (size_t)SPLIT * (NT / 64) * MNBLK : 0) +
         (size_t)NSTAGE * SPLIT * IMG + 256 + 2048;
}
template <int SPLIT, int NT>
constexpr size_t stage1_tc_smem_bytes() {
  return (size_t)IMG + 2 * (size_t)NT * 128 + 2 * (size_t)SPLIT * (NT / 64) * 16384 + (size_t)4 * SPLIT * IMG + 256;
}
template <int SPLIT>
constexpr size_t linear_smem_bytes() {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  // operand tile + ring + barriers (128 B) + the STATS accumulators (4 units x 128 threads x 2 doubles)
  return (size_t)4 * SPLIT * IMG + (size_t)NSTAGE * SPLIT * IMG + 128 + 4 * 2 * LIN_EPI * sizeof(double);
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

struct Workspace {
  size_t g_img, c_buf, t_img, total, h_img, total_tokenizer;
  Workspace(long long groups, int split) {
    const size_t tiles128 = (size_t)((groups + 127) / 128);
    const size_t img = tiles128 * 4 * (size_t)split * IMG;
    g_img = 0;
    t_img = img;
    c_buf = 2 * img;
    total = c_buf + tiles128 * 128 * 512 * sizeof(float);
    h_img = total;  // pos_embed hidden layer, K = 128: 2 chunks per 128-centre tile (ppt_tokenizer_forward only)
    total_tokenizer = h_img + tiles128 * 2 * (size_t)split * IMG;
    // train mode: 16 doubles of point moments, 1024 doubles of channel sums, 1024 floats of scale / shift
    bn_mom = total_tokenizer;
    bn_stats = bn_mom + 16 * sizeof(double);
    bn_vec = bn_stats + 1024 * sizeof(double);
    // group means of h1 as operand images (K = 128: 2 chunks per 128-group tile), the Gram matrix in fp64 and one
    // fp32 partial per CTA (at most MAX_CTAS)
    s_img = (bn_vec + 1024 * sizeof(float) + 1023) & ~(size_t)1023;
    gram = s_img + tiles128 * 2 * (size_t)split * IMG;
    gram_parts = gram + 16384 * sizeof(double);
    total_train = gram_parts + (size_t)MAX_CTAS * 16384 * sizeof(float);
  }
  static constexpr int MAX_CTAS = 256;
  size_t bn_mom, bn_stats, bn_vec, s_img, gram, gram_parts, total_train;
};

// phases: bit 0 stage1, bit 1 group_linear(c), bit 2 stage2, bit 3 group_linear(tokens)
template <uint32_t FMT, int SPLIT, int NT>
int run_encoder(const float* nbhd, const unsigned char* blob, unsigned char* ws, float* features_out,
                float* tokens_out, long long groups, int phases, cudaStream_t st, int rows_per_cloud = 0,
                int tokens_f16 = 0, long long* clock_acc = nullptr) {
  const BlobLayout L{(uint32_t)SPLIT};
  const Workspace W(groups, SPLIT);
  constexpr int EPW2 = stage2_epilogue_warps<NT>();
  constexpr int EPW1 = stage1_epilogue_warps<NT>();
  auto k1tc = encoder_stage1_tc_kernel<FMT, SPLIT, NT, EPW1, false>;
  constexpr size_t s1tc = stage1_tc_smem_bytes<SPLIT, NT>();
  static_assert(s1tc <= 232448, "shared memory budget (227 KB per CTA)");
  auto k2 = encoder_stage_kernel<FMT, SPLIT, NT, 2, EPW2>;
  auto k2clk = encoder_stage_kernel<FMT, SPLIT, NT, 2, EPW2, BN_EVAL, true>;
  auto kb = group_linear_kernel<FMT, SPLIT, 4>;
  auto kd = group_linear_kernel<FMT, SPLIT, 3>;
  auto kda = group_linear_kernel<FMT, SPLIT, 3, 4, true>;
  constexpr size_t s2 = stage_smem_bytes<SPLIT, NT, 2>(), sl = linear_smem_bytes<SPLIT>();
  static_assert(s2 <= 232448 && sl <= 232448, "shared memory budget (227 KB per CTA)");
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k1tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1tc));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k2clk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kda, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
  }
  const long long points = groups * 32;
  const int tiles = (int)((points + NT - 1) / NT);
  const int tiles128 = (int)((groups + 127) / 128);
  const int sms = num_sms();
  const int grid_t = tiles < sms ? tiles : sms, grid_g = tiles128 < sms ? tiles128 : sms;
  float* cbuf = reinterpret_cast<float*>(ws + W.c_buf);
  const float* scales = reinterpret_cast<const float*>(blob + L.scales());
  // Rows of the last operand-image tile beyond `groups` are never written; they only feed accumulator
  // columns that are never stored.
  if (phases & 1) k1tc<<<grid_t, (EPW1 + 2) * 32, s1tc, st>>>(nbhd, blob, ws + W.g_img, groups, tiles, nullptr, nullptr);
  if (phases & 2)
    kb<<<grid_g, LIN_THREADS, sl, st>>>(ws + W.g_img, blob + L.W3A(), reinterpret_cast<const float*>(blob + L.bias_c()),
                                        scales + 1, cbuf, groups, tiles128, 0, 0);
  if (phases & 4)
    (clock_acc ? k2clk : k2)<<<grid_t, (EPW2 + 2) * 32, s2, st>>>(nbhd, blob, cbuf, ws + W.t_img, features_out, groups,
                                                                  tiles, nullptr, nullptr, clock_acc);
  if ((phases & 8) && tokens_out)
    (rows_per_cloud > 0 ? kda : kd)<<<grid_g, LIN_THREADS, sl, st>>>(
        ws + W.t_img, blob + L.WR(), reinterpret_cast<const float*>(blob + L.bias_tok()), scales + 4, tokens_out,
        groups, tiles128, rows_per_cloud, rows_per_cloud > 0 ? 0 : tokens_f16);
  return ppt_launch_status();
}

// Encoder + reduce_dim into rows 1..G of x_out, pos_embed(center) into rows 1..G of pos_out, cls rows.
template <uint32_t FMT, int SPLIT, int NT>
int run_tokenizer(const float* nbhd, const float* center, const unsigned char* blob, const unsigned char* pblob,
                  unsigned char* ws, float* x_out, float* pos_out, long long groups, int rows_per_cloud,
                  cudaStream_t st) {
  const Workspace W(groups, SPLIT);
  auto kp = group_linear_kernel<FMT, SPLIT, 3, 2, true>;
  constexpr size_t sl = linear_smem_bytes<SPLIT>();
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
  }
  const int tiles128 = (int)((groups + 127) / 128);
  const int sms = num_sms();
  const int grid_g = tiles128 < sms ? tiles128 : sms;
  // pos path first: it only depends on the centres, and its hidden images do not alias the encoder's workspace
  pos_hidden_kernel<FMT, SPLIT><<<tiles128 < 4 * sms ? tiles128 : 4 * sms, 512, 0, st>>>(
      center, pblob, ws + W.h_img, x_out, pos_out, groups, rows_per_cloud);
  kp<<<grid_g, LIN_THREADS, sl, st>>>(ws + W.h_img, pblob + PosBlobLayout::W2(),
                                      reinterpret_cast<const float*>(pblob + PosBlobLayout::b2()),
                                      reinterpret_cast<const float*>(pblob + PosBlobLayout::scales()), pos_out, groups,
                                      tiles128, rows_per_cloud, 0);
  if (!x_out) return ppt_launch_status();
  return run_encoder<FMT, SPLIT, NT>(nbhd, blob, ws, nullptr, x_out, groups, 15, st, rows_per_cloud);
}

// Encoder.forward under model.train(): batch-statistics BatchNorm, forward only.  `blob` is the caller's MUTABLE
// copy packed with both BatchNorms as identities (raw convolution weights); its W1' sections are rewritten here.
template <uint32_t FMT, int SPLIT, int NT>
int run_encoder_train(const float* nbhd, unsigned char* blob, const BnDevice& bn, unsigned char* ws,
                      float* features_out, float* tokens_out, long long groups, cudaStream_t st) {
  const BlobLayout L{(uint32_t)SPLIT};
  const Workspace W(groups, SPLIT);
  constexpr int EPW1 = stage1_epilogue_warps<NT>();
  auto k1t = encoder_stage1_tc_kernel<FMT, SPLIT, NT, EPW1, true>;
  auto kst = group_linear_kernel<FMT, SPLIT, 4, 2, false, true>;
  constexpr int EPW2 = stage2_epilogue_warps<NT>();
  auto k2a = encoder_stage_kernel<FMT, SPLIT, NT, 2, EPW2, BN_APPLY>;
  constexpr size_t s1tc = stage1_tc_smem_bytes<SPLIT, NT>(), s2 = stage_smem_bytes<SPLIT, NT, 2>(),
                   sl = linear_smem_bytes<SPLIT>();
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k1t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1tc));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kst, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k2a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  }
  const long long points = groups * 32;
  const int tiles = (int)((points + NT - 1) / NT);
  const int tiles128 = (int)((groups + 127) / 128);
  const int sms = num_sms() < Workspace::MAX_CTAS ? num_sms() : Workspace::MAX_CTAS;
  const int grid_t = tiles < sms ? tiles : sms, grid_g = tiles128 < sms ? tiles128 : sms;
  double* mom = reinterpret_cast<double*>(ws + W.bn_mom);
  double* stats = reinterpret_cast<double*>(ws + W.bn_stats);
  float* bn_vec = reinterpret_cast<float*>(ws + W.bn_vec);
  float* cbuf = reinterpret_cast<float*>(ws + W.c_buf);
  double* gram = reinterpret_cast<double*>(ws + W.gram);
  float* gram_parts = reinterpret_cast<float*>(ws + W.gram_parts);
  const float* scales = reinterpret_cast<const float*>(blob + L.scales());
  PPT_RETURN_IF_CUDA(cudaMemsetAsync(mom, 0, (16 + 1024) * sizeof(double), st));
  // first_conv.1: exact batch statistics from the point moments, folded into W1'
  const long long want = (points + 255) / 256;
  bn_moments_kernel<<<(int)(want < 4 * sms ? want : 4 * sms), 256, 0, st>>>(nbhd, points, mom);
  bn_fold1_kernel<FMT><<<1, 128, 0, st>>>(bn, mom, points, blob, (uint32_t)SPLIT);
  // stage 1 (raw weights) + the Gram matrix and group means of h1; then c = W3a g + bias (raw weights)
  k1t<<<grid_t, (EPW1 + 2) * 32, s1tc, st>>>(nbhd, blob, ws + W.g_img, groups, tiles, ws + W.s_img, gram_parts);
  // second_conv.1: sum y and the c-dependent part of sum y^2 (per-group GEMM W32 . mean h1 on the tensor core),
  // the quadratic part from the Gram matrix; scale / shift; running statistics
  if (SPLIT == 1) {
    auto kcs = group_c_stats_kernel<FMT>;
    constexpr size_t scs = (size_t)6 * IMG + 4 * IMG + 128 + 4 * 2 * LIN_EPI * sizeof(double);
    static PptOncePerDevice cs_configured;
    if (cs_configured.need())
      PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kcs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scs));
    kcs<<<grid_g, LIN_THREADS, scs, st>>>(ws + W.g_img, ws + W.s_img, blob + L.W3A(), blob + L.W32(),
                                          reinterpret_cast<const float*>(blob + L.bias_c()), scales, cbuf, stats, groups,
                                          tiles128);
  } else {
    int rc = run_encoder<FMT, SPLIT, NT>(nbhd, blob, ws, nullptr, nullptr, groups, 2, st);
    if (rc) return rc;
    kst<<<grid_g, LIN_THREADS, sl, st>>>(ws + W.s_img, blob + L.W32(), cbuf, scales + 2, reinterpret_cast<float*>(stats),
                                         groups, tiles128, 0, 0);
  }
  const float act = SPLIT == 2 ? 64.f : 1.f;  // encoder_pack.ACT_SCALE (h1 operands are stored times this)
  bn_gram_reduce_kernel<<<64, 256, 0, st>>>(gram_parts, grid_t, 1.0 / ((double)act * act), gram);
  bn_fold2_kernel<<<512, 128, 0, st>>>(bn, stats, gram, reinterpret_cast<const float*>(blob + L.W32F()), points, bn_vec);
  k2a<<<grid_t, (EPW2 + 2) * 32, s2, st>>>(nbhd, blob, cbuf, ws + W.t_img, features_out, groups, tiles, nullptr, bn_vec,
                                           nullptr);
  if (tokens_out) return run_encoder<FMT, SPLIT, NT>(nbhd, blob, ws, nullptr, tokens_out, groups, 8, st);
  return ppt_launch_status();
}

}  // namespace

extern "C" PPT_EXPORT int64_t ppt_encoder_packed_bytes(int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_FP16X3) return PPT_EINVAL;
  return (int64_t)BlobLayout{mode == PPT_ENC_FP16X3 ? 2u : 1u}.total();
}

extern "C" PPT_EXPORT int64_t ppt_encoder_workspace_bytes(int64_t num_groups, int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_FP16X3 || num_groups < 1) return PPT_EINVAL;
  return (int64_t)Workspace(num_groups, mode == PPT_ENC_FP16X3 ? 2 : 1).total;
}

extern "C" PPT_EXPORT int ppt_encoder_forward_ex(const float* neighborhood, const void* packed, void* workspace,
                                                 float* features_out, void* tokens_out_v, int64_t num_groups, int mode,
                                                 int phases, int flags, void* clock_acc_v, void* stream) {
  float* tokens_out = static_cast<float*>(tokens_out_v);
  const int tokens_f16 = (flags & PPT_TOKENS_F16) ? 1 : 0;
  long long* clock_acc = static_cast<long long*>(clock_acc_v);
  if (flags & ~PPT_TOKENS_F16) return PPT_EINVAL;
  if (!neighborhood || !packed || !workspace || (!tokens_out && !features_out) || num_groups < 1) return PPT_EINVAL;
  if (num_groups > (1ll << 31) / 32) return PPT_ERANGE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PPT_EINVAL;
  const unsigned char* blob = static_cast<const unsigned char*>(packed);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case PPT_ENC_FP16:
      return run_encoder<tc05::FMT_F16, 1, 128>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                                st, 0, tokens_f16, clock_acc);
    case PPT_ENC_BF16:
      return run_encoder<tc05::FMT_BF16, 1, 128>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                                 st, 0, tokens_f16, clock_acc);
    case PPT_ENC_FP16X3:
      return run_encoder<tc05::FMT_F16, 2, 64>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                               st, 0, tokens_f16, clock_acc);
    default:
      return PPT_EINVAL;
  }
}

extern "C" PPT_EXPORT int ppt_encoder_forward_phases(const float* neighborhood, const void* packed, void* workspace,
                                                     float* features_out, float* tokens_out, int64_t num_groups,
                                                     int mode, int phases, void* stream) {
  return ppt_encoder_forward_ex(neighborhood, packed, workspace, features_out, tokens_out, num_groups, mode, phases, 0,
                                nullptr, stream);
}

extern "C" PPT_EXPORT int ppt_encoder_forward(const float* neighborhood, const void* packed, void* workspace,
                                              float* features_out, float* tokens_out, int64_t num_groups, int mode,
                                              void* stream) {
  return ppt_encoder_forward_phases(neighborhood, packed, workspace, features_out, tokens_out, num_groups, mode, 15,
                                    stream);
}

extern "C" PPT_EXPORT int64_t ppt_posembed_packed_bytes(int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_FP16X3) return PPT_EINVAL;
  return (int64_t)PosBlobLayout::total(mode == PPT_ENC_FP16X3 ? 2u : 1u);
}

extern "C" PPT_EXPORT int64_t ppt_tokenizer_workspace_bytes(int64_t num_groups, int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_FP16X3 || num_groups < 1) return PPT_EINVAL;
  return (int64_t)Workspace(num_groups, mode == PPT_ENC_FP16X3 ? 2 : 1).total_tokenizer;
}

extern "C" PPT_EXPORT int ppt_tokenizer_forward(const float* neighborhood, const float* center,
                                                const void* encoder_packed, const void* posembed_packed,
                                                void* workspace, float* x_out, float* pos_out, int64_t num_groups,
                                                int groups_per_cloud, int mode, void* stream) {
  if (!center || !posembed_packed || !workspace || !pos_out || num_groups < 1) return PPT_EINVAL;
  if (x_out && (!neighborhood || !encoder_packed)) return PPT_EINVAL;
  if (groups_per_cloud < 32 || num_groups % groups_per_cloud != 0) return PPT_EINVAL;
  if (num_groups > (1ll << 31) / 32) return PPT_ERANGE;
  if ((reinterpret_cast<uintptr_t>(posembed_packed) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15) ||
      (reinterpret_cast<uintptr_t>(encoder_packed) & 15))
    return PPT_EINVAL;
  const unsigned char* blob = static_cast<const unsigned char*>(encoder_packed);
  const unsigned char* pblob = static_cast<const unsigned char*>(posembed_packed);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case PPT_ENC_FP16:
      return run_tokenizer<tc05::FMT_F16, 1, 128>(neighborhood, center, blob, pblob, ws, x_out, pos_out, num_groups,
                                                  groups_per_cloud, st);
    case PPT_ENC_BF16:
      return run_tokenizer<tc05::FMT_BF16, 1, 128>(neighborhood, center, blob, pblob, ws, x_out, pos_out, num_groups,
                                                   groups_per_cloud, st);
    case PPT_ENC_FP16X3:
      return run_tokenizer<tc05::FMT_F16, 2, 64>(neighborhood, center, blob, pblob, ws, x_out, pos_out, num_groups,
                                                 groups_per_cloud, st);
    default:
      return PPT_EINVAL;
  }
}

extern "C" PPT_EXPORT int64_t ppt_encoder_train_workspace_bytes(int64_t num_groups, int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_FP16X3 || num_groups < 1) return PPT_EINVAL;
  return (int64_t)Workspace(num_groups, mode == PPT_ENC_FP16X3 ? 2 : 1).total_train;
}

extern "C" PPT_EXPORT int ppt_encoder_forward_train(const float* neighborhood, void* packed_train,
                                                    const ppt_encoder_bn_t* bn, void* workspace, float* features_out,
                                                    float* tokens_out, int64_t num_groups, int mode, void* stream) {
  if (!neighborhood || !packed_train || !bn || !workspace || (!tokens_out && !features_out) || num_groups < 1)
    return PPT_EINVAL;
  if (!bn->conv1_weight || !bn->conv1_bias || !bn->bn1_weight || !bn->bn1_bias || !bn->bn1_running_mean ||
      !bn->bn1_running_var || !bn->bn2_weight || !bn->bn2_bias || !bn->bn2_running_mean || !bn->bn2_running_var)
    return PPT_EINVAL;
  if (!(bn->momentum >= 0.f && bn->momentum <= 1.f) || !(bn->eps > 0.f)) return PPT_EINVAL;
  if (num_groups > (1ll << 31) / 32) return PPT_ERANGE;
  if ((reinterpret_cast<uintptr_t>(packed_train) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return PPT_EINVAL;
  BnDevice d;
  d.conv1_w = bn->conv1_weight; d.conv1_b = bn->conv1_bias;
  d.bn1_w = bn->bn1_weight; d.bn1_b = bn->bn1_bias; d.bn1_rm = bn->bn1_running_mean; d.bn1_rv = bn->bn1_running_var;
  d.bn1_nbt = reinterpret_cast<long long*>(bn->bn1_num_batches_tracked);
  d.bn2_w = bn->bn2_weight; d.bn2_b = bn->bn2_bias; d.bn2_rm = bn->bn2_running_mean; d.bn2_rv = bn->bn2_running_var;
  d.bn2_nbt = reinterpret_cast<long long*>(bn->bn2_num_batches_tracked);
  d.momentum = bn->momentum; d.eps = bn->eps;
  unsigned char* blob = static_cast<unsigned char*>(packed_train);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case PPT_ENC_FP16:
      return run_encoder_train<tc05::FMT_F16, 1, 128>(neighborhood, blob, d, ws, features_out, tokens_out, num_groups, st);
    case PPT_ENC_BF16:
      return run_encoder_train<tc05::FMT_BF16, 1, 128>(neighborhood, blob, d, ws, features_out, tokens_out, num_groups, st);
    case PPT_ENC_FP16X3:
      return run_encoder_train<tc05::FMT_F16, 2, 64>(neighborhood, blob, d, ws, features_out, tokens_out, num_groups, st);
    default:
      return PPT_EINVAL;
  }
}

