// PointNet++ set-abstraction shared MLP + max-pool on tcgen05, sm_100a.
//
// Reference: PointNetSetAbstraction / PointNetSetAbstractionMsg.forward after the grouping,
// models/pointnet2/pointnet2_utils.py:196-201, 256-261 (eval mode):
//     new_points [B, C0, nsample, S]  ->  3 x (Conv2d 1x1 + BatchNorm2d + ReLU)  ->  max over nsample  ->  [B, C3, S]
// Here (BatchNorm folded on the host, ppt_b200/encoder_pack.py:pack_sa_mlp):
//   sa_fused_kernel            the default: gather + centre subtraction + concat -> three tensor-core layers -> max
//                              over nsample in ONE kernel per 128-column tile, activations resident in shared memory;
//   sa_gather_image_kernel,    the first version and the fallback when the activation regions do not fit
//   pointwise_linear_kernel    (PPT_SA_PER_LAYER in `mode` selects it): the gather writes the fp16 K-major operand images of
//                              layer 1 to HBM, then one kernel per layer = relu(W' act + b') on the tensor core (the
//                              group_linear pipeline of encoder.cu); layers 1, 2 store the next layer's operand images,
//                              the last one max-pools over each group's accumulator columns and stores [B, C3, S].
// In both, the fp32 [B, S, nsample, C0] tensor of sample_and_group never exists.  DESIGN.md section 9.
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr uint32_t IMG = 16384;   // one operand image: 128 rows x 64 K x 2 bytes, K-major, 128-byte swizzle
constexpr int SA_THREADS = 192;   // producer, MMA issuer, 4 epilogue warps
constexpr int SA_EPI = 128;
constexpr int SA_NSTAGE = 4;
constexpr int SA_MAX_KC = 8;      // <= 512 input channels per layer

// ---- grouping -> layer-1 operand images ---------------------------------------------------------------
// Column = one (b, s, j) neighbour; channel order [features (D) | xyz - centre (3) | zero padding] (the weight
// columns are permuted to match on the host, so the SSG [xyz, feats] and MSG [feats, xyz] orders share it).
template <uint32_t FMT>
__global__ void __launch_bounds__(256)
sa_gather_image_kernel(const float* __restrict__ xyz, const float* __restrict__ feats,
                       const float* __restrict__ new_xyz, const int64_t* __restrict__ idx,
                       unsigned char* __restrict__ img, int N, int S, int ns, int D, int KC, long long total) {
  const int r = threadIdx.x & 127, hh = threadIdx.x >> 7;
  const long long tiles = (total + 127) / 128;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long col = tile * 128 + r;
    const bool ok = col < total;
    long long n = 0;
    int b = 0;
    float cx = 0.f, cy = 0.f, cz = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    if (ok) {
      const long long g = col / ns;  // b * S + s
      b = (int)(g / S);
      n = __ldg(idx + col);
      n = n < 0 ? 0 : (n >= N ? N - 1 : n);  // ball query never returns its "nothing in range" sentinel for centres of the cloud
      const float* c = new_xyz + g * 3;
      const float* p = xyz + ((long long)b * N + n) * 3;
      cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
      px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
    }
    const float rel[3] = {__fsub_rn(px, cx), __fsub_rn(py, cy), __fsub_rn(pz, cz)};
    const float* frow = feats ? feats + ((long long)b * N + n) * D : nullptr;
    unsigned char* base = img + (size_t)tile * KC * IMG;
    for (int c8 = hh * 8; c8 < KC * 64; c8 += 16) {
      float v[8];
      if (ok && frow && c8 + 8 <= D && (D & 3) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(frow + c8));
        const float4 e = __ldg(reinterpret_cast<const float4*>(frow + c8 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = e.x; v[5] = e.y; v[6] = e.z; v[7] = e.w;
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int c = c8 + t;
          v[t] = !ok ? 0.f : (c < D ? __ldg(frow + c) : (c < D + 3 ? rel[c - D] : 0.f));
        }
      }
      uint4 w;
      w.x = pack2<FMT, false>(v[0], v[1]); w.y = pack2<FMT, false>(v[2], v[3]);
      w.z = pack2<FMT, false>(v[4], v[5]); w.w = pack2<FMT, false>(v[6], v[7]);
      *reinterpret_cast<uint4*>(base + (size_t)(c8 >> 6) * IMG + sw128_kmajor_off(r, c8 & 63)) = w;
    }
  }
}

// ---- one layer ---------------------------------------------------------------------------------------------
// out = relu(W act + bias); act as operand images [tile][KC][16 KB]; W as images [U][KC][16 KB] (rows padded to
// 128 U with zeros, K padded to 64 KC).  POOL = false: store as the next layer's images [tile][2 U][16 KB];
// POOL = true: max over each group's `ns` columns (ns in {16, 32, 64, 128}), out [B, c_out, S] fp32.
template <uint32_t FMT, bool POOL>
__global__ void __launch_bounds__(SA_THREADS, 1)
pointwise_linear_kernel(const unsigned char* __restrict__ act_img, const unsigned char* __restrict__ wimg,
                        const float* __restrict__ bias, unsigned char* __restrict__ out_img, float* __restrict__ out,
                        int KC, int U, int c_out, int ns, int S, long long num_groups, long long total_cols,
                        int num_tiles) {
  constexpr int NT = 128;
  constexpr int TCOLS = 2 * NT;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* bbuf = smem;                         // [KC][16 KB] activations of the tile
  unsigned char* ring = bbuf + (size_t)KC * IMG;      // [SA_NSTAGE][16 KB] weights
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + SA_NSTAGE * IMG);
  uint64_t* full = bars;
  uint64_t* empty = full + SA_NSTAGE;
  uint64_t* acc_full = empty + SA_NSTAGE;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint64_t* b_empty = b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < SA_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], SA_EPI); }
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
      uint32_t it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait_relaxed(b_empty, (tile_it & 1u) ^ 1u);
        mbar_arrive_expect_tx(b_full, (uint32_t)KC * IMG);
        for (int kc = 0; kc < KC; ++kc)
          bulk_g2s(bbuf + (size_t)kc * IMG, act_img + ((size_t)tile * KC + kc) * IMG, IMG, b_full);
        for (int u = 0; u < U; ++u)
          for (int kc = 0; kc < KC; ++kc, ++it) {
            const uint32_t s = it % SA_NSTAGE;
            mbar_wait_relaxed(&empty[s], ((it / SA_NSTAGE) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&full[s], IMG);
            bulk_g2s(ring + s * IMG, wimg + ((size_t)u * KC + kc) * IMG, IMG, &full[s]);
          }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
    constexpr uint32_t HI = sdesc_hi(1024u);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u), b_lo0 = sdesc_lo(smem_u32(bbuf), 16u);
    uint32_t it = 0, tile_it = 0, unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(b_full, tile_it & 1u);
      for (int u = 0; u < U; ++u, ++unit_it) {
        const uint32_t buf = unit_it & 1u;
        mbar_wait(&acc_empty[buf], ((unit_it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const uint32_t s = it % SA_NSTAGE;
          mbar_wait(&full[s], (it / SA_NSTAGE) & 1u);
          fence_after_sync();
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16)
            umma_f16_elect(tbase + buf * NT, sdesc_join(a_lo0 + s * (IMG >> 4) + (uint32_t)k16 * 2u, HI),
                           sdesc_join(b_lo0 + (uint32_t)kc * (IMG >> 4) + (uint32_t)k16 * 2u, HI), idesc,
                           (kc == 0 && k16 == 0) ? 0u : 1u);
          umma_commit_elect(&empty[s]);
        }
        umma_commit_elect(&acc_full[buf]);
        if (u == U - 1) umma_commit_elect(b_empty);  // every MMA that reads this tile's activations is complete
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint32_t unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int u = 0; u < U; ++u, ++unit_it) {
        const uint32_t buf = unit_it & 1u;
        const int o = u * 128 + m;
        const float bo = __ldg(bias + o);
        mbar_wait(&acc_full[buf], (unit_it >> 1) & 1u);
        fence_after_sync();
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + buf * NT;
        float run = 0.f;  // POOL: running max of the current group (values are >= 0 after the ReLU)
#pragma unroll 1
        for (int j = 0; j < NT / 32; ++j) {
          float v[32];
          tmem_ld32(t_addr + j * 32, v);
          const long long col0 = (long long)tile * NT + j * 32;
          if (!POOL) {
            unsigned char* dst = out_img + ((size_t)tile * (2 * U) + (size_t)(o >> 6)) * IMG;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              *reinterpret_cast<uint16_t*>(dst + sw128_kmajor_off(j * 32 + i, o & 63)) =
                  to_operand<FMT>(fmaxf(v[i] + bo, 0.f));
          } else {
            // relu(max(x) + b) == max(relu(x + b)): pool first, one bias add per group
            if (ns == 16) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float mx = v[16 * h];
#pragma unroll
                for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[16 * h + i]);
                const long long g = (col0 + 16 * h) / 16;
                if (g < num_groups && o < c_out) out[((g / S) * c_out + o) * S + g % S] = fmaxf(mx + bo, 0.f);
              }
            } else {
              float mx = v[0];
#pragma unroll
              for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
              const int per = ns / 32;  // 32-column chunks per group: 1, 2 or 4
              run = (j % per) == 0 ? mx : fmaxf(run, mx);
              if ((j % per) == per - 1) {
                const long long g = col0 / ns;
                if (g < num_groups && o < c_out) out[((g / S) * c_out + o) * S + g % S] = fmaxf(run + bo, 0.f);
              }
            }
          }
        }
        fence_before_sync();
        mbar_arrive(&acc_empty[buf]);
      }
    }
    (void)total_cols;
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

// ---- one layer with MORE input channels than fit in shared memory at once (K-blocked) ---------------------------
// out = relu(W act + bias) for K up to 64 x 64 channels: the activation tile streams through two 4-chunk (64 KB)
// buffers while ALL output units (U <= 4, c_out <= 512) keep their accumulators in tensor memory (U x 128 columns),
// so every activation chunk is read once and multiplied with U weight chunks from the ring.  Used for the second
// layer of the feature-propagation MLP (1536 -> 384, models/pointbert/point_encoder.py:300-302).
//   OUT_F32 = false: store as the next layer's operand images [tile][2 U][16 KB];
//   OUT_F32 = true : store fp32 channel-first, out[(b * c_out + o) * N + n] for column b * N + n.
constexpr int KB_CHUNKS = 4;

template <uint32_t FMT, bool OUT_F32>
__global__ void __launch_bounds__(SA_THREADS, 1)
pointwise_linear_kblock_kernel(const unsigned char* __restrict__ act_img, const unsigned char* __restrict__ wimg,
                               const float* __restrict__ bias, unsigned char* __restrict__ out_img,
                               float* __restrict__ out, int KC, int U, int c_out, int N, long long total_cols,
                               int num_tiles) {
  constexpr int NT = 128;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* bbuf = smem;                                   // [2][KB_CHUNKS][16 KB]
  unsigned char* ring = bbuf + 2 * KB_CHUNKS * IMG;             // [SA_NSTAGE][16 KB] weights
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + SA_NSTAGE * IMG);
  uint64_t* full = bars;
  uint64_t* empty = full + SA_NSTAGE;
  uint64_t* b_full = empty + SA_NSTAGE;    // [2]
  uint64_t* b_empty = b_full + 2;          // [2]
  uint64_t* acc_full = b_empty + 2;        // [1] all U accumulators of the tile complete
  uint64_t* acc_empty = acc_full + 1;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  const int nkb = (KC + KB_CHUNKS - 1) / KB_CHUNKS;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < SA_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, SA_EPI);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, blk = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++blk) {
          const uint32_t slot = blk & 1u;
          const int nch = KC - kb * KB_CHUNKS < KB_CHUNKS ? KC - kb * KB_CHUNKS : KB_CHUNKS;
          mbar_wait_relaxed(&b_empty[slot], ((blk >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&b_full[slot], (uint32_t)nch * IMG);
          for (int kc = 0; kc < nch; ++kc)
            bulk_g2s(bbuf + (size_t)(slot * KB_CHUNKS + kc) * IMG,
                     act_img + ((size_t)tile * KC + kb * KB_CHUNKS + kc) * IMG, IMG, &b_full[slot]);
          for (int u = 0; u < U; ++u)
            for (int kc = 0; kc < nch; ++kc, ++it) {
              const uint32_t s = it % SA_NSTAGE;
              mbar_wait_relaxed(&empty[s], ((it / SA_NSTAGE) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&full[s], IMG);
              bulk_g2s(ring + s * IMG, wimg + ((size_t)u * KC + kb * KB_CHUNKS + kc) * IMG, IMG, &full[s]);
            }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(FMT, 128, NT, 0);
    constexpr uint32_t HI = sdesc_hi(1024u);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u), b_lo0 = sdesc_lo(smem_u32(bbuf), 16u);
    uint32_t it = 0, blk = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(acc_empty, (tile_it & 1u) ^ 1u);  // the previous tile's accumulators are drained
      fence_after_sync();
      for (int kb = 0; kb < nkb; ++kb, ++blk) {
        const uint32_t slot = blk & 1u;
        const int nch = KC - kb * KB_CHUNKS < KB_CHUNKS ? KC - kb * KB_CHUNKS : KB_CHUNKS;
        mbar_wait(&b_full[slot], (blk >> 1) & 1u);
        fence_after_sync();
        for (int u = 0; u < U; ++u)
          for (int kc = 0; kc < nch; ++kc, ++it) {
            const uint32_t s = it % SA_NSTAGE;
            mbar_wait(&full[s], (it / SA_NSTAGE) & 1u);
            fence_after_sync();
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              umma_f16_elect(tbase + (uint32_t)(u * NT), sdesc_join(a_lo0 + s * (IMG >> 4) + (uint32_t)k16 * 2u, HI),
                             sdesc_join(b_lo0 + (slot * KB_CHUNKS + (uint32_t)kc) * (IMG >> 4) + (uint32_t)k16 * 2u, HI),
                             idesc, (kb == 0 && kc == 0 && k16 == 0) ? 0u : 1u);
            umma_commit_elect(&empty[s]);
          }
        umma_commit_elect(&b_empty[slot]);  // every MMA that reads this activation block is complete
      }
      umma_commit_elect(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(acc_full, tile_it & 1u);
      fence_after_sync();
      for (int u = 0; u < U; ++u) {
        const int o = u * 128 + m;
        const float bo = __ldg(bias + o);
        const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + (uint32_t)(u * NT);
#pragma unroll 1
        for (int j = 0; j < NT / 32; ++j) {
          float v[32];
          tmem_ld32(t_addr + j * 32, v);
          const long long col0 = (long long)tile * NT + j * 32;
          if (!OUT_F32) {
            unsigned char* dst = out_img + ((size_t)tile * (2 * U) + (size_t)(o >> 6)) * IMG;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              *reinterpret_cast<uint16_t*>(dst + sw128_kmajor_off(j * 32 + i, o & 63)) =
                  to_operand<FMT>(fmaxf(v[i] + bo, 0.f));
          } else if (o < c_out) {
            const long long b0 = col0 / N;
            const int n0 = (int)(col0 - b0 * N);
            if (col0 + 32 <= total_cols && n0 + 32 <= N && (N & 3) == 0 && (n0 & 3) == 0) {
              float4* dst = reinterpret_cast<float4*>(out + (b0 * c_out + o) * N + n0);  // 32 consecutive points
#pragma unroll
              for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(fmaxf(v[4 * q] + bo, 0.f), fmaxf(v[4 * q + 1] + bo, 0.f),
                                     fmaxf(v[4 * q + 2] + bo, 0.f), fmaxf(v[4 * q + 3] + bo, 0.f));
            } else {
#pragma unroll 1
              for (int i = 0; i < 32; ++i) {
                const long long col = col0 + i;
                if (col < total_cols) {
                  const long long b = col / N;
                  out[(b * c_out + o) * N + (col - b * N)] = fmaxf(v[i] + bo, 0.f);
                }
              }
            }
          }
        }
      }
      fence_before_sync();
      mbar_arrive(acc_empty);
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}

// ---- feature propagation: three_interpolate + concat -> layer-1 operand images ------------------------------------
// PointNetFeaturePropagation.forward (models/pointnet2/pointnet2_utils.py:297-313) up to the MLP: column = one point
// (b, n); channel order [interpolated (D2) | points1 (D1) | zero padding] (the reference concatenates
// [points1, interpolated]; layer 1's weight columns are permuted to match on the host, which keeps the 16-byte loads of
// the three neighbour rows aligned).  The interpolation is the arithmetic of ppt_three_interpolate;
// the fp32 [B, N, D2] interpolated tensor and the concatenated [B, D1 + D2, N] tensor never exist.
template <uint32_t FMT>
__global__ void __launch_bounds__(256)
fp_build_image_kernel(const float* __restrict__ points1, const float* __restrict__ feats2,
                      const int64_t* __restrict__ idx, const float* __restrict__ dist, unsigned char* __restrict__ img,
                      int N, int S, int D1, int D2, int KC, long long total) {
  const int r = threadIdx.x & 127, hh = threadIdx.x >> 7;
  const long long tiles = (total + 127) / 128;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long col = tile * 128 + r;
    const bool ok = col < total;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    const float *f0 = feats2, *f1 = feats2, *f2 = feats2;
    long long b = 0;
    int n = 0;
    if (ok) {
      b = col / N;
      n = (int)(col - b * N);
      const float r0 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + col * 3), 1e-8f));
      const float r1 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + col * 3 + 1), 1e-8f));
      const float r2 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + col * 3 + 2), 1e-8f));
      const float nrm = __fadd_rn(__fadd_rn(r0, r1), r2);
      w0 = __fdiv_rn(r0, nrm); w1 = __fdiv_rn(r1, nrm); w2 = __fdiv_rn(r2, nrm);
      auto row = [&](int j) {
        long long s = __ldg(idx + col * 3 + j);
        s = s < 0 ? 0 : (s >= S ? S - 1 : s);
        return feats2 + (b * S + s) * D2;
      };
      f0 = row(0); f1 = row(1); f2 = row(2);
    }
    unsigned char* base = img + (size_t)tile * KC * IMG;
    for (int c8 = hh * 8; c8 < KC * 64; c8 += 16) {
      float v[8];
      if (ok && c8 + 8 <= D2 && (D2 & 3) == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(f0 + c8 + 4 * h));
          const float4 e = __ldg(reinterpret_cast<const float4*>(f1 + c8 + 4 * h));
          const float4 g = __ldg(reinterpret_cast<const float4*>(f2 + c8 + 4 * h));
          v[4 * h + 0] = __fadd_rn(__fadd_rn(__fmul_rn(w0, a.x), __fmul_rn(w1, e.x)), __fmul_rn(w2, g.x));
          v[4 * h + 1] = __fadd_rn(__fadd_rn(__fmul_rn(w0, a.y), __fmul_rn(w1, e.y)), __fmul_rn(w2, g.y));
          v[4 * h + 2] = __fadd_rn(__fadd_rn(__fmul_rn(w0, a.z), __fmul_rn(w1, e.z)), __fmul_rn(w2, g.z));
          v[4 * h + 3] = __fadd_rn(__fadd_rn(__fmul_rn(w0, a.w), __fmul_rn(w1, e.w)), __fmul_rn(w2, g.w));
        }
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int c = c8 + t;
          float x = 0.f;
          if (ok && c < D2)
            x = __fadd_rn(__fadd_rn(__fmul_rn(w0, __ldg(f0 + c)), __fmul_rn(w1, __ldg(f1 + c))), __fmul_rn(w2, __ldg(f2 + c)));
          else if (ok && c < D2 + D1)
            x = __ldg(points1 + (b * D1 + (c - D2)) * N + n);  // channel-first: consecutive threads, consecutive points
          v[t] = x;
        }
      }
      uint4 w;
      w.x = pack2<FMT, false>(v[0], v[1]); w.y = pack2<FMT, false>(v[2], v[3]);
      w.z = pack2<FMT, false>(v[4], v[5]); w.w = pack2<FMT, false>(v[6], v[7]);
      *reinterpret_cast<uint4*>(base + (size_t)(c8 >> 6) * IMG + sw128_kmajor_off(r, c8 & 63)) = w;
    }
  }
}

// ---- the three layers in ONE kernel: activations never leave shared memory ---------------------------------
// Tile = 128 (group, sample) columns.  Per tile: the 8 epilogue warps gather [features | xyz - centre] into the
// K-major input operand (region X); layer 1 -> ReLU -> MN-major operand in region Y; layer 2 -> ReLU -> MN-major
// operand in region X (the input is dead by then: every layer-1 instruction has completed once its accumulators
// were drained); layer 3 -> max over each group's columns -> out.  Weights = A operand (channels on the TMEM lanes,
// so bias, ReLU and the pooling are per-thread), streamed through an mbarrier ring of `nstage` x 16 KB; two
// alternating 128-column accumulators.  Layers of one tile run back to back (no cross-tile overlap yet: a variant
// with 4 epilogue warps and two resident CTAs per SM measured slower, 179 us against 134 us on SSG level 2).
// Shared memory: X = max(16 KB kc0, 32 KB u2), Y = 32 KB u1, ring, barriers -- run_sa_mlp picks this kernel when that
// fits (all PointNet++ levels of models/pointnet2/pointnet2.py except MSG level 3's 643 input channels).
constexpr int SAF_THREADS = 320;  // producer, MMA issuer, 8 epilogue warps
constexpr int SAF_EPI = 256;

template <uint32_t FMT>
__global__ void __launch_bounds__(SAF_THREADS, 1)
sa_fused_kernel(const float* __restrict__ xyz, const float* __restrict__ feats, const float* __restrict__ new_xyz,
                const int64_t* __restrict__ idx, const unsigned char* __restrict__ blob, float* __restrict__ out,
                int N, int S, int ns, int D, int kc0, int u1, int u2, int u3, int c3, int nstage, uint32_t x_bytes,
                uint32_t z_off /* 0: layer-2 output aliases the input; else its own region (gather prefetch) */,
                uint32_t w1_off, uint32_t w2_off, uint32_t w3_off, long long num_groups, long long total, int num_tiles) {
  constexpr int NT = 128;
  constexpr int TCOLS = 2 * NT;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* xbuf = smem;                               // input (K-major) / layer-2 output (MN-major)
  unsigned char* ybuf = xbuf + x_bytes;                     // layer-1 output (MN-major), 32 KB u1
  unsigned char* zbuf = z_off ? smem + z_off : xbuf;        // layer-2 output (MN-major)
  unsigned char* ring = (z_off ? zbuf + (size_t)u2 * 32768 : ybuf + (size_t)u1 * 32768);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)nstage * IMG);
  uint64_t* full = bars;                // [4]
  uint64_t* empty = full + 4;           // [4]
  uint64_t* acc_full = empty + 4;       // [2]
  uint64_t* acc_empty = acc_full + 2;   // [2]
  uint64_t* in_ready = acc_empty + 2;   // [1] gathered input written
  uint64_t* act_ready = in_ready + 1;   // [2] layer 1 / layer 2 output written (all units)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 2);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], SAF_EPI); }
    mbar_init(in_ready, SAF_EPI);
    mbar_init(&act_ready[0], (uint32_t)(SAF_EPI * u1));
    mbar_init(&act_ready[1], (uint32_t)(SAF_EPI * u2));
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const int units[3] = {u1, u2, u3};
  const int chunks[3] = {kc0, 2 * u1, 2 * u2};
  const uint32_t woff[3] = {w1_off, w2_off, w3_off};

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int l = 0; l < 3; ++l)
          for (int u = 0; u < units[l]; ++u)
            for (int kc = 0; kc < chunks[l]; ++kc, ++it) {
              const uint32_t s = it % (uint32_t)nstage;
              mbar_wait_relaxed(&empty[s], ((it / (uint32_t)nstage) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&full[s], IMG);
              bulk_g2s(ring + s * IMG, blob + woff[l] + ((size_t)u * chunks[l] + kc) * IMG, IMG, &full[s]);
            }
    }
  } else if (warp == 1) {
    const uint32_t idesc_k = make_idesc(FMT, 128, NT, 0), idesc_mn = make_idesc(FMT, 128, NT, 1);
    constexpr uint32_t HI = sdesc_hi(1024u);
    const uint32_t a_lo0 = sdesc_lo(smem_u32(ring), 16u);
    const uint32_t b_lo[3] = {sdesc_lo(smem_u32(xbuf), 16u), sdesc_lo(smem_u32(ybuf), (uint32_t)u1 * 16384u),
                              sdesc_lo(smem_u32(zbuf), (uint32_t)u2 * 16384u)};
    uint32_t it = 0, tile_it = 0, unit_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      for (int l = 0; l < 3; ++l) {
        mbar_wait(l == 0 ? in_ready : &act_ready[l - 1], tile_it & 1u);
        fence_after_sync();
        for (int u = 0; u < units[l]; ++u, ++unit_it) {
          const uint32_t buf = unit_it & 1u;
          mbar_wait(&acc_empty[buf], ((unit_it >> 1) & 1u) ^ 1u);
          fence_after_sync();
          for (int kc = 0; kc < chunks[l]; ++kc, ++it) {
            const uint32_t s = it % (uint32_t)nstage;
            mbar_wait(&full[s], (it / (uint32_t)nstage) & 1u);
            fence_after_sync();
            // K-major input: chunk = one 16 KB image, +32 B per K=16 slice; MN-major activations: chunk = 8 K-atoms of
            // 1 KB, +2 KB per K=16 slice (tc05.cuh: make_sdesc)
            const uint32_t bk = l == 0 ? b_lo[0] + (uint32_t)kc * (IMG >> 4) : b_lo[l] + (uint32_t)kc * 512u;
            const uint32_t bstep = l == 0 ? 2u : 128u;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              umma_f16_elect(tbase + buf * NT, sdesc_join(a_lo0 + s * (IMG >> 4) + (uint32_t)k16 * 2u, HI),
                             sdesc_join(bk + (uint32_t)k16 * bstep, HI), l == 0 ? idesc_k : idesc_mn,
                             (kc == 0 && k16 == 0) ? 0u : 1u);
            umma_commit_elect(&empty[s]);
          }
          umma_commit_elect(&acc_full[buf]);
        }
      }
    }
  } else {
    const int e = tid - 64;
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;      // channel row inside a 128-row unit
    const int col0 = half * 64;          // this thread's 64 accumulator columns
    const int r = e & 127, hh = e >> 7;  // gather: column r, chunk parity hh
    const float* biasv = reinterpret_cast<const float*>(blob);
    uint32_t unit_it = 0;
    // ---- gather: [features (D) | xyz - centre (3) | 0] of column r of `tile` into the K-major input operand ----
    auto gather = [&](int tile) {
      {
        const long long col = (long long)tile * NT + r;
        const bool ok = col < total;
        long long n = 0;
        int b = 0;
        float cx = 0.f, cy = 0.f, cz = 0.f, px = 0.f, py = 0.f, pz = 0.f;
        if (ok) {
          const long long g = col / ns;
          b = (int)(g / S);
          n = __ldg(idx + col);
          n = n < 0 ? 0 : (n >= N ? N - 1 : n);
          const float* c = new_xyz + g * 3;
          const float* p = xyz + ((long long)b * N + n) * 3;
          cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
          px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
        }
        const float rel[3] = {__fsub_rn(px, cx), __fsub_rn(py, cy), __fsub_rn(pz, cz)};
        const float* frow = feats ? feats + ((long long)b * N + n) * D : nullptr;
        // four chunks per pass: all eight 16-byte loads of a pass are issued before the first conversion, so the
        // (L2-latency-bound) gather keeps 128 bytes per thread in flight instead of 32
        for (int c0 = hh * 8; c0 < kc0 * 64; c0 += 64) {
          float4 ld[4][2];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c8 = c0 + 16 * k;
            const bool vec = ok && frow && c8 + 8 <= D && (D & 3) == 0;
            ld[k][0] = vec ? __ldg(reinterpret_cast<const float4*>(frow + c8)) : make_float4(0.f, 0.f, 0.f, 0.f);
            ld[k][1] = vec ? __ldg(reinterpret_cast<const float4*>(frow + c8 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c8 = c0 + 16 * k;
            if (c8 >= kc0 * 64) break;
            float v[8] = {ld[k][0].x, ld[k][0].y, ld[k][0].z, ld[k][0].w, ld[k][1].x, ld[k][1].y, ld[k][1].z, ld[k][1].w};
            if (ok && !(frow && c8 + 8 <= D && (D & 3) == 0)) {
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const int c = c8 + t;
                v[t] = c < D ? __ldg(frow + c) : (c < D + 3 ? rel[c - D] : 0.f);
              }
            }
            uint4 w;
            w.x = pack2<FMT, false>(v[0], v[1]); w.y = pack2<FMT, false>(v[2], v[3]);
            w.z = pack2<FMT, false>(v[4], v[5]); w.w = pack2<FMT, false>(v[6], v[7]);
            *reinterpret_cast<uint4*>(xbuf + (size_t)(c8 >> 6) * IMG + sw128_kmajor_off(r, c8 & 63)) = w;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(in_ready);
      }
    };
    // With its own region for the layer-2 output, the input of the NEXT tile is gathered as soon as layer 1 of this
    // tile is complete, under the tensor work of layers 2 and 3.
    if (z_off && (int)blockIdx.x < num_tiles) gather(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (!z_off) gather(tile);
      int boff = 0;
      for (int l = 0; l < 3; ++l) {
        unsigned char* dst = l == 0 ? ybuf : zbuf;
        const uint32_t mnblk = (uint32_t)units[l] * 16384u;  // bytes between 64-column blocks of this layer's output
        for (int u = 0; u < units[l]; ++u, ++unit_it) {
          const uint32_t buf = unit_it & 1u;
          const int ch = u * 128 + m;
          const float bo = __ldg(biasv + boff + ch);
          mbar_wait(&acc_full[buf], (unit_it >> 1) & 1u);
          fence_after_sync();
          const uint32_t t_addr = tbase + ((uint32_t)(quad * 32) << 16) + buf * NT + (uint32_t)col0;
          if (l < 2) {
            const uint32_t krow = (uint32_t)(ch >> 3) * 1024u + (uint32_t)(ch & 7) * 128u, sw = (uint32_t)(ch & 7);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              float v[32];
              tmem_ld32(t_addr + jj * 32, v);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int n = col0 + jj * 32 + q * 8;
                const uint32_t off = (uint32_t)(n >> 6) * mnblk + krow + ((((uint32_t)(n & 63) >> 3) ^ sw) << 4);
                const float* w = v + q * 8;
                *reinterpret_cast<uint4*>(dst + off) =
                    make_uint4(pack2<FMT, true>(w[0] + bo, w[1] + bo), pack2<FMT, true>(w[2] + bo, w[3] + bo),
                               pack2<FMT, true>(w[4] + bo, w[5] + bo), pack2<FMT, true>(w[6] + bo, w[7] + bo));
              }
            }
            fence_proxy_async_smem();
            fence_before_sync();
            mbar_arrive(&acc_empty[buf]);
            mbar_arrive(&act_ready[l]);
          } else {
            // relu(max(x) + b) == max(relu(x + b)): pool first, one bias add per group
            float run = 0.f;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              float v[32];
              tmem_ld32(t_addr + jj * 32, v);
              const long long cbase = (long long)tile * NT + col0 + jj * 32;
              if (ns == 16) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  float mx = v[16 * h];
#pragma unroll
                  for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[16 * h + i]);
                  const long long g = (cbase + 16 * h) / 16;
                  if (g < num_groups && ch < c3) out[((g / S) * c3 + ch) * S + g % S] = fmaxf(mx + bo, 0.f);
                }
              } else {
                float mx = v[0];
#pragma unroll
                for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
                const long long g = cbase / ns;
                if (ns == 32) {
                  if (g < num_groups && ch < c3) out[((g / S) * c3 + ch) * S + g % S] = fmaxf(mx + bo, 0.f);
                } else {
                  run = jj == 0 ? mx : fmaxf(run, mx);
                  if (jj == 1 && g < num_groups && ch < c3) {
                    const float y = fmaxf(run + bo, 0.f);
                    float* o = out + ((g / S) * c3 + ch) * S + g % S;
                    if (ns == 64) *o = y;
                    else atomicMax(reinterpret_cast<int*>(o), __float_as_int(y));  // ns == 128: two halves; y >= 0, out zeroed
                  }
                }
              }
            }
            fence_before_sync();
            mbar_arrive(&acc_empty[buf]);
          }
        }
        boff += units[l] * 128;
        if (l == 0 && z_off && tile + (int)gridDim.x < num_tiles) gather(tile + gridDim.x);
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tbase);
}

int num_sms_sa() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

struct SaDims {
  int kc0, u1, u2, u3;  // K chunks of the input, 128-row units of the three layers
  __host__ SaDims(int c0, int c1, int c2, int c3)
      : kc0((c0 + 63) / 64), u1((c1 + 127) / 128), u2((c2 + 127) / 128), u3((c3 + 127) / 128) {}
  // blob: fp32 biases [128 (u1 + u2 + u3)] padded to 1 KB, then W1 [u1][kc0], W2 [u2][2 u1], W3 [u3][2 u2] images
  size_t bias_bytes() const { return (((size_t)(u1 + u2 + u3) * 128 * 4) + 1023) & ~(size_t)1023; }
  size_t w1() const { return bias_bytes(); }
  size_t w2() const { return w1() + (size_t)u1 * kc0 * IMG; }
  size_t w3() const { return w2() + (size_t)u2 * 2 * u1 * IMG; }
  size_t total() const { return w3() + (size_t)u3 * 2 * u2 * IMG; }
  // layer 1 may be wider than shared memory holds (MSG level 3: 643 channels) when its <= 4 output units fit in
  // tensor memory: the K-blocked kernel streams it
  bool ok() const {
    return (kc0 <= SA_MAX_KC || (kc0 <= 64 && u1 <= 4)) && 2 * u1 <= SA_MAX_KC && 2 * u2 <= SA_MAX_KC && u3 <= 8;
  }
  // workspace: the three activation image sets
  size_t ws0(long long tiles) const { (void)tiles; return 0; }
  size_t ws1(long long tiles) const { return (size_t)tiles * kc0 * IMG; }
  size_t ws2(long long tiles) const { return ws1(tiles) + (size_t)tiles * 2 * u1 * IMG; }
  size_t ws_total(long long tiles) const { return ws2(tiles) + (size_t)tiles * 2 * u2 * IMG; }
};

template <uint32_t FMT>
int run_sa_mlp(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
               const unsigned char* blob, unsigned char* ws, float* out, int B, int N, int S, int ns, int D,
               const SaDims& d, int c3, bool use_fused, cudaStream_t st) {
  const long long groups = (long long)B * S, total = groups * ns;
  const long long tiles_ll = (total + 127) / 128;
  if (tiles_ll > 0x7fffffffll) return PPT_ERANGE;
  const int tiles = (int)tiles_ll;
  auto kl = pointwise_linear_kernel<FMT, false>;
  auto kp = pointwise_linear_kernel<FMT, true>;
  const size_t smem_max = (size_t)SA_MAX_KC * IMG + SA_NSTAGE * IMG + 256;
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  }
  const int sms = num_sms_sa();
  const int grid = tiles < sms ? tiles : sms;
  const float* bias = reinterpret_cast<const float*>(blob);
  // single-kernel path: activations stay in shared memory; the per-layer kernels are the fallback for layer widths
  // whose activation regions do not fit (and can be forced with PPT_SA_PER_LAYER in `mode`, for cross-checks)
  // separate regions for the input and the layer-2 output (-> gather prefetch) when they fit with a 3-stage ring
  const size_t sep_fixed = (size_t)d.kc0 * IMG + (size_t)d.u1 * 32768 + (size_t)d.u2 * 32768 + 256;
  const bool sep = sep_fixed + 3 * (size_t)IMG <= 232448;
  const size_t x_bytes = sep ? (size_t)d.kc0 * IMG
                             : ((size_t)d.kc0 * IMG > (size_t)d.u2 * 32768 ? (size_t)d.kc0 * IMG : (size_t)d.u2 * 32768);
  const size_t fixed = sep ? sep_fixed : x_bytes + (size_t)d.u1 * 32768 + 256;
  const uint32_t z_off = sep ? (uint32_t)(x_bytes + (size_t)d.u1 * 32768) : 0u;
  int nstage = 0;
  for (int n = 4; n >= 2 && !nstage; --n)
    if (fixed + (size_t)n * IMG <= 232448) nstage = n;
  if (use_fused && nstage) {
    auto kf = sa_fused_kernel<FMT>;
    static PptOncePerDevice fused_configured;
    if (fused_configured.need()) {
      PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    if (ns == 128) PPT_RETURN_IF_CUDA(cudaMemsetAsync(out, 0, (size_t)groups * c3 * sizeof(float), st));
    kf<<<grid, SAF_THREADS, fixed + (size_t)nstage * IMG, st>>>(xyz, feats, new_xyz, idx, blob, out, N, S, ns, D, d.kc0,
                                                              d.u1, d.u2, d.u3, c3, nstage, (uint32_t)x_bytes, z_off,
                                                              (uint32_t)d.w1(), (uint32_t)d.w2(), (uint32_t)d.w3(),
                                                              groups, total, tiles);
    return ppt_launch_status();
  }
  sa_gather_image_kernel<FMT><<<tiles < 8 * sms ? tiles : 8 * sms, 256, 0, st>>>(xyz, feats, new_xyz, idx, ws + d.ws0(tiles),
                                                                               N, S, ns, D, d.kc0, total);
  auto smem = [](int kc) { return (size_t)kc * IMG + SA_NSTAGE * IMG + 256; };
  if (d.kc0 > SA_MAX_KC) {  // wide input (MSG level 3): K-blocked layer 1
    auto kb = pointwise_linear_kblock_kernel<FMT, false>;
    const size_t smem_kb = (size_t)2 * KB_CHUNKS * IMG + SA_NSTAGE * IMG + 256;
    static PptOncePerDevice kb_configured;
    if (kb_configured.need())
      PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kb));
    kb<<<grid, SA_THREADS, smem_kb, st>>>(ws + d.ws0(tiles), blob + d.w1(), bias, ws + d.ws1(tiles), nullptr, d.kc0, d.u1, 0,
                                          1, total, tiles);
  } else
  kl<<<grid, SA_THREADS, smem(d.kc0), st>>>(ws + d.ws0(tiles), blob + d.w1(), bias, ws + d.ws1(tiles), nullptr, d.kc0,
                                            d.u1, 0, ns, S, groups, total, tiles);
  kl<<<grid, SA_THREADS, smem(2 * d.u1), st>>>(ws + d.ws1(tiles), blob + d.w2(), bias + d.u1 * 128, ws + d.ws2(tiles),
                                               nullptr, 2 * d.u1, d.u2, 0, ns, S, groups, total, tiles);
  kp<<<grid, SA_THREADS, smem(2 * d.u2), st>>>(ws + d.ws2(tiles), blob + d.w3(), bias + (d.u1 + d.u2) * 128, nullptr,
                                               out, 2 * d.u2, d.u3, c3, ns, S, groups, total, tiles);
  return ppt_launch_status();
}



// ---- feature-propagation MLP (two layers) -------------------------------------------------------------------------
struct FpDims {
  int kc0, u1, u2;
  __host__ FpDims(int c0, int c1, int c2) : kc0((c0 + 63) / 64), u1((c1 + 127) / 128), u2((c2 + 127) / 128) {}
  // blob: fp32 biases [128 (u1 + u2)] padded to 1 KB, then W1 [u1][kc0], W2 [u2][2 u1] images
  size_t bias_bytes() const { return (((size_t)(u1 + u2) * 128 * 4) + 1023) & ~(size_t)1023; }
  size_t w1() const { return bias_bytes(); }
  size_t w2() const { return w1() + (size_t)u1 * kc0 * IMG; }
  size_t total() const { return w2() + (size_t)u2 * 2 * u1 * IMG; }
  bool ok() const { return kc0 >= 1 && kc0 <= SA_MAX_KC && u1 >= 1 && u2 >= 1 && u2 <= 4 && 2 * u1 <= 64; }
  size_t ws_total(long long tiles) const { return (size_t)tiles * (kc0 + 2 * u1) * IMG; }
};

template <uint32_t FMT>
int run_fp_mlp(const float* points1, const float* feats2, const int64_t* idx, const float* dist,
               const unsigned char* blob, unsigned char* ws, float* out, int B, int N, int S, int D1, int D2,
               const FpDims& d, int c2, cudaStream_t st) {
  const long long total = (long long)B * N;
  const long long tiles_ll = (total + 127) / 128;
  if (tiles_ll > 0x7fffffffll) return PPT_ERANGE;
  const int tiles = (int)tiles_ll;
  auto k1 = pointwise_linear_kernel<FMT, false>;
  auto k2 = pointwise_linear_kblock_kernel<FMT, true>;
  const size_t smem1 = (size_t)d.kc0 * IMG + SA_NSTAGE * IMG + 256;
  const size_t smem2 = (size_t)2 * KB_CHUNKS * IMG + SA_NSTAGE * IMG + 256;
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)((size_t)SA_MAX_KC * IMG + SA_NSTAGE * IMG + 256)));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  }
  const int sms = num_sms_sa();
  const int grid = tiles < sms ? tiles : sms;
  const float* bias = reinterpret_cast<const float*>(blob);
  unsigned char* img0 = ws;
  unsigned char* img1 = ws + (size_t)tiles * d.kc0 * IMG;
  fp_build_image_kernel<FMT><<<tiles < 8 * sms ? tiles : 8 * sms, 256, 0, st>>>(points1, feats2, idx, dist, img0, N, S, D1,
                                                                               D2, d.kc0, total);
  k1<<<grid, SA_THREADS, smem1, st>>>(img0, blob + d.w1(), bias, img1, nullptr, d.kc0, d.u1, 0, 32, 1, 0, total, tiles);
  k2<<<grid, SA_THREADS, smem2, st>>>(img1, blob + d.w2(), bias + d.u1 * 128, nullptr, out, 2 * d.u1, d.u2, c2, N, total,
                                      tiles);
  return ppt_launch_status();
}

}  // namespace

extern "C" PPT_EXPORT int64_t ppt_fp_mlp_packed_bytes(int c0, int c1, int c2) {
  if (c0 < 1 || c1 < 1 || c2 < 1) return PPT_EINVAL;
  const FpDims d(c0, c1, c2);
  return d.ok() ? (int64_t)d.total() : PPT_ERANGE;
}

extern "C" PPT_EXPORT int64_t ppt_fp_mlp_workspace_bytes(int64_t num_points, int c0, int c1, int c2) {
  if (num_points < 1 || c0 < 1 || c1 < 1 || c2 < 1) return PPT_EINVAL;
  const FpDims d(c0, c1, c2);
  return d.ok() ? (int64_t)d.ws_total((num_points + 127) / 128) : PPT_ERANGE;
}

extern "C" PPT_EXPORT int ppt_fp_mlp_forward(const float* points1, const float* feats2, const int64_t* idx,
                                             const float* dist, const void* packed, void* workspace, float* out, int B,
                                             int N, int S, int D1, int D2, int c1, int c2, int mode, void* stream) {
  if (!feats2 || !idx || !dist || !packed || !workspace || !out || B < 1 || N < 1 || S < 1 || D1 < 0 || D2 < 1)
    return PPT_EINVAL;
  if (D1 > 0 && !points1) return PPT_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PPT_EINVAL;
  const FpDims d(D1 + D2, c1, c2);
  if (c1 < 1 || c2 < 1 || !d.ok()) return PPT_ERANGE;
  const unsigned char* blob = static_cast<const unsigned char*>(packed);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  if (mode == PPT_ENC_FP16)
    return run_fp_mlp<tc05::FMT_F16>(points1, feats2, idx, dist, blob, ws, out, B, N, S, D1, D2, d, c2, (cudaStream_t)stream);
  if (mode == PPT_ENC_BF16)
    return run_fp_mlp<tc05::FMT_BF16>(points1, feats2, idx, dist, blob, ws, out, B, N, S, D1, D2, d, c2, (cudaStream_t)stream);
  return PPT_EINVAL;
}

namespace {
}  // namespace

extern "C" PPT_EXPORT int64_t ppt_sa_mlp_packed_bytes(int c0, int c1, int c2, int c3) {
  if (c0 < 1 || c1 < 1 || c2 < 1 || c3 < 1) return PPT_EINVAL;
  const SaDims d(c0, c1, c2, c3);
  return d.ok() ? (int64_t)d.total() : PPT_ERANGE;
}

extern "C" PPT_EXPORT int64_t ppt_sa_mlp_workspace_bytes(int64_t num_columns, int c0, int c1, int c2, int c3) {
  if (num_columns < 1 || c0 < 1 || c1 < 1 || c2 < 1 || c3 < 1) return PPT_EINVAL;
  const SaDims d(c0, c1, c2, c3);
  return d.ok() ? (int64_t)d.ws_total((num_columns + 127) / 128) : PPT_ERANGE;
}

extern "C" PPT_EXPORT int ppt_sa_mlp_forward(const float* xyz, const float* feats, const float* new_xyz,
                                             const int64_t* idx, const void* packed, void* workspace, float* out,
                                             int B, int N, int S, int nsample, int D, int c1, int c2, int c3, int mode,
                                             void* stream) {
  if (!xyz || !new_xyz || !idx || !packed || !workspace || !out || B < 1 || N < 1 || S < 1 || D < 0) return PPT_EINVAL;
  if (D > 0 && !feats) return PPT_EINVAL;
  if (nsample != 16 && nsample != 32 && nsample != 64 && nsample != 128) return PPT_ERANGE;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PPT_EINVAL;
  const SaDims d(D + 3, c1, c2, c3);
  if (c1 < 1 || c2 < 1 || c3 < 1 || !d.ok()) return PPT_ERANGE;
  const unsigned char* blob = static_cast<const unsigned char*>(packed);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const bool use_fused = !(mode & PPT_SA_PER_LAYER);
  mode &= ~PPT_SA_PER_LAYER;
  if (mode == PPT_ENC_FP16)
    return run_sa_mlp<tc05::FMT_F16>(xyz, D > 0 ? feats : nullptr, new_xyz, idx, blob, ws, out, B, N, S, nsample, D, d, c3,
                                     use_fused, (cudaStream_t)stream);
  if (mode == PPT_ENC_BF16)
    return run_sa_mlp<tc05::FMT_BF16>(xyz, D > 0 ? feats : nullptr, new_xyz, idx, blob, ws, out, B, N, S, nsample, D, d,
                                      c3, use_fused, (cudaStream_t)stream);
  return PPT_EINVAL;
}
