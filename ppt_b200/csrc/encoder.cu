// mini-PointNet patch Encoder + reduce_dim on tcgen05 tensor cores (sm_100a).
//
// Replaces Encoder.forward in eval mode (models/pointbert/dvae.py:201-215) and
// reduce_dim (models/pointbert/point_encoder.py:133,239).  The host folds BatchNorm,
// composes first_conv.3 with the per-point half of second_conv.0 and packs every
// weight into swizzled operand images (ppt_b200/encoder_pack.py).  Four launches:
//
//   stage1        per 128-point tile: h1 = relu(W1'x + b1') (K = 3, CUDA cores) ->
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
//
// Pipeline per CTA (persistent over tiles): warp 0 streams 16 KB weight images from
// L2 with 1-D bulk async copies into a ring (full/empty mbarriers); warp 1 issues
// tcgen05.mma into two alternating TMEM accumulators; warps 2-5 build h1, drain the
// accumulators (tcgen05.ld), apply bias/ReLU/max and write the next operand.
#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int ENC_THREADS = 192;      // producer warp, MMA warp, 4 epilogue warps
constexpr int EPI_THREADS = 128;
constexpr uint32_t IMG = 16384;       // one operand image: 128 rows x 64 K x 2 B

// ---- packed weight blob (ppt_b200/encoder_pack.py) -----------------------------------
struct BlobLayout {
  uint32_t split;
  __host__ __device__ uint32_t w1() const { return 0; }          // [128][4] fp32
  __host__ __device__ uint32_t bias_c() const { return 2048; }   // [512]
  __host__ __device__ uint32_t b4() const { return 4096; }       // [256]
  __host__ __device__ uint32_t bias_tok() const { return 5120; } // [384]
  __host__ __device__ uint32_t W2() const { return 8192; }                       // 2 units x 2 chunks
  __host__ __device__ uint32_t W3A() const { return W2() + 4 * split * IMG; }    // 4 x 4
  __host__ __device__ uint32_t W32() const { return W3A() + 16 * split * IMG; }  // 4 x 2
  __host__ __device__ uint32_t W4() const { return W32() + 8 * split * IMG; }    // 2 x 8
  __host__ __device__ uint32_t WR() const { return W4() + 16 * split * IMG; }    // 3 x 4
  __host__ __device__ uint32_t total() const { return WR() + 12 * split * IMG; }
};

struct Ring {  // position in a ring of mbarrier-guarded stages
  uint32_t it = 0;
  template <int NSTAGE> __device__ uint32_t stage() const { return it % NSTAGE; }
  template <int NSTAGE> __device__ uint32_t parity() const { return (it / NSTAGE) & 1u; }
};

template <uint32_t FMT, int SPLIT>
__device__ __forceinline__ void issue_k64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t b_split_bytes,
                                          uint32_t idesc, bool first) {
  // One 64-wide K chunk = four K=16 instructions (x3 passes in the hi/lo split mode).
#pragma unroll
  for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
    for (int pass = 0; pass < (SPLIT == 2 ? 3 : 1); ++pass) {
      const uint32_t sa = pass == 2 ? 1 : 0, sb = pass == 1 ? 1 : 0;  // hi*hi, hi*lo, lo*hi
      const uint64_t ad = make_sdesc(a_addr + sa * IMG + k16 * 32u, 16u, 1024u);
      const uint64_t bd = make_sdesc(b_addr + sb * b_split_bytes + k16 * 32u, 16u, 1024u);
      umma_f16(d_tmem, ad, bd, idesc, (first && k16 == 0 && pass == 0) ? 0u : 1u);
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

// ======================================================================================
// stage kernels
// ======================================================================================
template <uint32_t FMT, int SPLIT, int NT, int STAGE>
__global__ void __launch_bounds__(ENC_THREADS, 1)
encoder_stage_kernel(const float* __restrict__ nbhd, const unsigned char* __restrict__ blob,
                     const float* __restrict__ cbuf,          // stage 2: [groups_pad, 512]
                     unsigned char* __restrict__ out_img,     // stage 1: g images, stage 2: t images
                     float* __restrict__ features_out,        // stage 2, nullable: [groups, 256]
                     long long num_groups, int num_tiles) {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  constexpr int GPT = NT / 32;                       // groups per tile
  constexpr int NUNITS = STAGE == 1 ? 2 : 6;
  constexpr uint32_t H1_BYTES = 2u * NT * 128u;      // per split copy: 2 chunks
  constexpr uint32_t H3_BYTES = STAGE == 2 ? 8u * NT * 128u : 0u;
  constexpr uint32_t STAGE_BYTES = SPLIT * IMG;
  constexpr int TCOLS = 2 * NT;

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* h1buf = smem;                                   // [SPLIT][2][NT x 128 B]
  unsigned char* h3buf = h1buf + SPLIT * H1_BYTES;               // [SPLIT][8][NT x 128 B]
  unsigned char* ring = h3buf + SPLIT * H3_BYTES;                // [NSTAGE][SPLIT][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE_BYTES);
  uint64_t* full = bars;                   // [NSTAGE]
  uint64_t* empty = full + NSTAGE;         // [NSTAGE]
  uint64_t* acc_full = empty + NSTAGE;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* h1_ready = acc_empty + 2;      // [1]
  uint64_t* h3_ready = h1_ready + 1;       // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h3_ready + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BlobLayout L{(uint32_t)SPLIT};

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS); }
    mbar_init(h1_ready, EPI_THREADS);
    for (int i = 0; i < 4; ++i) mbar_init(&h3_ready[i], EPI_THREADS);
    mbar_fence_init();
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
            mbar_wait(&empty[s], r.parity<NSTAGE>() ^ 1u);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(ring + s * STAGE_BYTES, blob + sec + (size_t)(blk * nkc + kc) * STAGE_BYTES, STAGE_BYTES,
                     &full[s]);
            ++r.it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    Ring r;
    const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
    const uint32_t ring_addr = smem_u32(ring), h1_addr = smem_u32(h1buf), h3_addr = smem_u32(h3buf);
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(h1_ready, tile_it & 1u);
#pragma unroll 1
      for (int u = 0; u < NUNITS; ++u) {
        const bool g3 = STAGE == 2 && u >= 4;
        const int nkc = g3 ? 8 : 2;
        const int buf = u & 1;
        // accumulator `buf` is used NUNITS/2 times per tile: its n-th use has parity n & 1
        const uint32_t use = tile_it * (NUNITS / 2) + (uint32_t)(u >> 1);
        mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tbase + (uint32_t)(buf * NT);
        for (int kc = 0; kc < nkc; ++kc) {
          if (STAGE == 2 && u == 4 && (kc & 1) == 0) {  // h3 K-chunks 2i, 2i+1 come from unit i's epilogue
            mbar_wait(&h3_ready[kc >> 1], tile_it & 1u);
          }
          const uint32_t s = r.stage<NSTAGE>();
          mbar_wait(&full[s], r.parity<NSTAGE>());
          fence_after_sync();
          if (lane == 0) {
            const uint32_t b_addr = (g3 ? h3_addr : h1_addr) + (uint32_t)kc * (NT * 128u);
            issue_k64<FMT, SPLIT>(d_tmem, ring_addr + s * STAGE_BYTES, b_addr, g3 ? H3_BYTES : H1_BYTES, idesc,
                                  kc == 0);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          ++r.it;
        }
        if (lane == 0) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int e = tid - 64;                 // 0..127
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int m = quad * 32 + lane;         // output-channel row inside a 128-row unit
    const float4* w1 = reinterpret_cast<const float4*>(blob + L.w1());
    const float* bias_b4 = reinterpret_cast<const float*>(blob + L.b4());

    auto build_h1 = [&](int tile) {
      // h1[p][ch] = relu(w.x + b), K = 3 on CUDA cores; written as the K-major B operand.
      constexpr int TPP = EPI_THREADS / NT;          // threads per point
      constexpr int CH = 128 / TPP;                  // channels per thread
      const int p = e % NT, ch0 = (e / NT) * CH;
      const long long gp = (long long)tile * NT + p;
      float x = 0.f, y = 0.f, z = 0.f;
      if (gp < num_groups * 32) {
        const float* src = nbhd + gp * 3;
        x = src[0]; y = src[1]; z = src[2];
      }
#pragma unroll 2
      for (int c8 = 0; c8 < CH; c8 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          float v[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 w = __ldg(w1 + ch0 + c8 + t + q);
            v[q] = fmaxf(fmaf(w.z, z, fmaf(w.y, y, fmaf(w.x, x, w.w))), 0.f);
          }
          const uint16_t h0 = to_operand<FMT>(v[0]), h1v = to_operand<FMT>(v[1]);
          hi[t >> 1] = (uint32_t)h0 | ((uint32_t)h1v << 16);
          if (SPLIT == 2)
            lo[t >> 1] = (uint32_t)to_operand<FMT>(v[0] - from_operand<FMT>(h0)) |
                         ((uint32_t)to_operand<FMT>(v[1] - from_operand<FMT>(h1v)) << 16);
        }
        const int ch = ch0 + c8;
        const uint32_t off = (uint32_t)(ch >> 6) * (NT * 128u) + sw128_kmajor_off(p, ch & 63);
        *reinterpret_cast<uint4*>(h1buf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (SPLIT == 2) *reinterpret_cast<uint4*>(h1buf + H1_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(h1_ready);
    };

    if ((int)blockIdx.x < num_tiles) build_h1(blockIdx.x);

    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const long long g0 = (long long)tile * GPT;  // first group of this tile
#pragma unroll 1
      for (int u = 0; u < NUNITS; ++u) {
        const int buf = u & 1;
        const bool relu_unit = STAGE == 2 && u < 4;
        float cval[GPT];
        if (relu_unit) {
#pragma unroll
          for (int j = 0; j < GPT; ++j)
            cval[j] = (g0 + j < num_groups) ? __ldg(cbuf + (g0 + j) * 512 + u * 128 + m) : 0.f;
        }
        mbar_wait(&acc_full[buf], (tile_it * (NUNITS / 2) + (uint32_t)(u >> 1)) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * NT);

        if (relu_unit) {
          // h3[p][ch] = relu(acc + c[group][ch]); ch = u*128 + m is this thread's K index.
          const int ch = u * 128 + m;
          const uint32_t kbase = (uint32_t)(ch >> 6) * (NT * 128u) + (uint32_t)(ch & 7) * 2u;
          const uint32_t piece = (uint32_t)((ch & 63) >> 3);
#pragma unroll
          for (int j = 0; j < GPT; ++j) {
            float v[32];
            tmem_ld32(t_addr + j * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int p = j * 32 + i;
              const uint32_t off = kbase + (uint32_t)p * 128u + ((piece ^ (uint32_t)(i & 7)) << 4);
              store_operand<FMT, SPLIT>(h3buf, off, H3_BYTES, fmaxf(v[i] + cval[j], 0.f));
            }
          }
          fence_proxy_async_smem();
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          mbar_arrive(&h3_ready[u]);
        } else {
          // per-group max over the 32 points (columns) this thread holds for its channel
          const int blk = STAGE == 1 ? u : u - 4;
          const int ch = blk * 128 + m;  // 0..255
#pragma unroll
          for (int j = 0; j < GPT; ++j) {
            float v[32];
            tmem_ld32(t_addr + j * 32, v);
            float mx = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
            const long long g = g0 + j;
            if (g < num_groups) {
              // operand image for group_linear: tile of 128 groups, K = 256 -> 4 chunks
              const size_t img = ((size_t)(g >> 7) * 4 + (size_t)(ch >> 6)) * (SPLIT * IMG);
              store_operand<FMT, SPLIT>(out_img + img, sw128_kmajor_off((int)(g & 127), ch & 63), IMG, mx);
              if (STAGE == 2 && features_out) features_out[g * 256 + ch] = mx + __ldg(bias_b4 + ch);
            }
          }
          fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
        }

        // every MMA that reads h1 has completed once the last h1-consuming unit's accumulator is full
        if (u == (STAGE == 1 ? 1 : 3)) {
          const int next = tile + gridDim.x;
          if (next < num_tiles) build_h1(next);
        }
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

// ======================================================================================
// group_linear: out[g][o] = sum_k W[o][k] * act[g][k] + bias[o], act given as operand images
// ======================================================================================
template <uint32_t FMT, int SPLIT, int NUNITS>
__global__ void __launch_bounds__(ENC_THREADS, 1)
group_linear_kernel(const unsigned char* __restrict__ act_img, const unsigned char* __restrict__ wsec,
                    const float* __restrict__ bias, float* __restrict__ out, long long num_groups, int num_tiles) {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  constexpr int NT = 128;                      // groups per tile (MMA N)
  constexpr int NOUT = NUNITS * 128;
  constexpr uint32_t STAGE_BYTES = SPLIT * IMG;
  constexpr uint32_t B_BYTES = 4u * SPLIT * IMG;  // [kc][split][16 KB]
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS); }
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
        mbar_wait(b_empty, (tile_it & 1u) ^ 1u);
        mbar_arrive_expect_tx(b_full, B_BYTES);
        for (int kc = 0; kc < 4; ++kc)
          bulk_g2s(bbuf + kc * STAGE_BYTES, act_img + ((size_t)tile * 4 + kc) * STAGE_BYTES, STAGE_BYTES, b_full);
#pragma unroll 1
        for (int u = 0; u < NUNITS; ++u)
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t s = r.stage<NSTAGE>();
            mbar_wait(&empty[s], r.parity<NSTAGE>() ^ 1u);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(ring + s * STAGE_BYTES, wsec + (size_t)(u * 4 + kc) * STAGE_BYTES, STAGE_BYTES, &full[s]);
            ++r.it;
          }
      }
    }
  } else if (warp == 1) {
    Ring r;
    const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
    const uint32_t ring_addr = smem_u32(ring), b_addr0 = smem_u32(bbuf);
    uint32_t tile_it = 0, unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(b_full, tile_it & 1u);
#pragma unroll 1
      for (int u = 0; u < NUNITS; ++u, ++unit_it) {
        const int buf = unit_it & 1;
        mbar_wait(&acc_empty[buf], ((unit_it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        for (int kc = 0; kc < 4; ++kc) {
          const uint32_t s = r.stage<NSTAGE>();
          mbar_wait(&full[s], r.parity<NSTAGE>());
          fence_after_sync();
          if (lane == 0) {
            // activation image order is [kc][split]: the lo copy sits IMG bytes after the hi copy
            issue_k64<FMT, SPLIT>(tbase + (uint32_t)(buf * NT), ring_addr + s * STAGE_BYTES,
                                  b_addr0 + kc * STAGE_BYTES, IMG, idesc, kc == 0);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          ++r.it;
        }
        if (lane == 0) {
          umma_commit(&acc_full[buf]);
          if (u == NUNITS - 1) umma_commit(b_empty);  // all MMAs reading this tile's activations are done
        }
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint32_t unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int u = 0; u < NUNITS; ++u, ++unit_it) {
        const int buf = unit_it & 1;
        const int o = u * 128 + m;
        const float bo = __ldg(bias + o);
        mbar_wait(&acc_full[buf], (unit_it >> 1) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * NT);
#pragma unroll
        for (int j = 0; j < NT / 32; ++j) {
          float v[32];
          tmem_ld32(t_addr + j * 32, v);
          const long long g0 = (long long)tile * NT + j * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (g0 + i < num_groups) out[(g0 + i) * NOUT + o] = v[i] + bo;
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[buf]);
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

// ======================================================================================
// host side
// ======================================================================================
template <int SPLIT, int NT, int STAGE>
constexpr size_t stage_smem_bytes() {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  return (size_t)SPLIT * (2u * NT * 128u) + (STAGE == 2 ? (size_t)SPLIT * (8u * NT * 128u) : 0) +
         (size_t)NSTAGE * SPLIT * IMG + (2 * NSTAGE + 9) * 8 + 16;
}
template <int SPLIT>
constexpr size_t linear_smem_bytes() {
  constexpr int NSTAGE = SPLIT == 2 ? 2 : 4;
  return (size_t)4 * SPLIT * IMG + (size_t)NSTAGE * SPLIT * IMG + (2 * NSTAGE + 6) * 8 + 16;
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
  size_t g_img, c_buf, t_img, total;
  Workspace(long long groups, int split) {
    const size_t tiles128 = (size_t)((groups + 127) / 128);
    const size_t img = tiles128 * 4 * (size_t)split * IMG;
    g_img = 0;
    t_img = img;
    c_buf = 2 * img;
    total = c_buf + tiles128 * 128 * 512 * sizeof(float);
  }
};

// phases: bit 0 stage1, bit 1 group_linear(c), bit 2 stage2, bit 3 group_linear(tokens)
template <uint32_t FMT, int SPLIT, int NT>
int run_encoder(const float* nbhd, const unsigned char* blob, unsigned char* ws, float* features_out,
                float* tokens_out, long long groups, int phases, cudaStream_t st) {
  const BlobLayout L{(uint32_t)SPLIT};
  const Workspace W(groups, SPLIT);
  auto k1 = encoder_stage_kernel<FMT, SPLIT, NT, 1>;
  auto k2 = encoder_stage_kernel<FMT, SPLIT, NT, 2>;
  auto kb = group_linear_kernel<FMT, SPLIT, 4>;
  auto kd = group_linear_kernel<FMT, SPLIT, 3>;
  constexpr size_t s1 = stage_smem_bytes<SPLIT, NT, 1>(), s2 = stage_smem_bytes<SPLIT, NT, 2>(),
                   sl = linear_smem_bytes<SPLIT>();
  static bool configured = false;
  if (!configured) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    configured = true;
  }
  const long long points = groups * 32;
  const int tiles = (int)((points + NT - 1) / NT);
  const int tiles128 = (int)((groups + 127) / 128);
  const int sms = num_sms();
  const int grid_t = tiles < sms ? tiles : sms, grid_g = tiles128 < sms ? tiles128 : sms;
  float* cbuf = reinterpret_cast<float*>(ws + W.c_buf);
  // The tail tile of the operand images is only partly written by the stage kernels; the unwritten rows
  // feed MMA columns that are never stored, but they must not hold NaN patterns that trap nothing -- any
  // bit pattern is fine for unused columns, so no clearing is needed.
  if (phases & 1) k1<<<grid_t, ENC_THREADS, s1, st>>>(nbhd, blob, nullptr, ws + W.g_img, nullptr, groups, tiles);
  if (phases & 2)
    kb<<<grid_g, ENC_THREADS, sl, st>>>(ws + W.g_img, blob + L.W3A(),
                                        reinterpret_cast<const float*>(blob + L.bias_c()), cbuf, groups, tiles128);
  if (phases & 4) k2<<<grid_t, ENC_THREADS, s2, st>>>(nbhd, blob, cbuf, ws + W.t_img, features_out, groups, tiles);
  if ((phases & 8) && tokens_out)
    kd<<<grid_g, ENC_THREADS, sl, st>>>(ws + W.t_img, blob + L.WR(),
                                        reinterpret_cast<const float*>(blob + L.bias_tok()), tokens_out, groups,
                                        tiles128);
  return ppt_launch_status();
}

}  // namespace

extern "C" PPT_EXPORT int64_t ppt_encoder_packed_bytes(int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_BF16X3) return PPT_EINVAL;
  return (int64_t)BlobLayout{mode == PPT_ENC_BF16X3 ? 2u : 1u}.total();
}

extern "C" PPT_EXPORT int64_t ppt_encoder_workspace_bytes(int64_t num_groups, int mode) {
  if (mode < PPT_ENC_FP16 || mode > PPT_ENC_BF16X3 || num_groups < 1) return PPT_EINVAL;
  return (int64_t)Workspace(num_groups, mode == PPT_ENC_BF16X3 ? 2 : 1).total;
}

extern "C" PPT_EXPORT int ppt_encoder_forward_phases(const float* neighborhood, const void* packed, void* workspace,
                                                     float* features_out, float* tokens_out, int64_t num_groups,
                                                     int mode, int phases, void* stream) {
  if (!neighborhood || !packed || !workspace || (!tokens_out && !features_out) || num_groups < 1) return PPT_EINVAL;
  if (num_groups > (1ll << 31) / 32) return PPT_ERANGE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PPT_EINVAL;
  const unsigned char* blob = static_cast<const unsigned char*>(packed);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case PPT_ENC_FP16:
      return run_encoder<tc05::FMT_F16, 1, 128>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                                st);
    case PPT_ENC_BF16:
      return run_encoder<tc05::FMT_BF16, 1, 128>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                                 st);
    case PPT_ENC_BF16X3:
      return run_encoder<tc05::FMT_BF16, 2, 64>(neighborhood, blob, ws, features_out, tokens_out, num_groups, phases,
                                                st);
    default:
      return PPT_EINVAL;
  }
}

extern "C" PPT_EXPORT int ppt_encoder_forward(const float* neighborhood, const void* packed, void* workspace,
                                              float* features_out, float* tokens_out, int64_t num_groups, int mode,
                                              void* stream) {
  return ppt_encoder_forward_phases(neighborhood, packed, workspace, features_out, tokens_out, num_groups, mode, 15,
                                    stream);
}
