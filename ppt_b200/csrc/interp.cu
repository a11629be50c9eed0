// three_nn / three_interpolate (+ gradient) for sm_100a.
//
// Replaces the interpolation part of PointNetFeaturePropagation.forward
// (models/pointnet2/pointnet2_utils.py:297-307): the reference sorts the whole
// (B,N,S) distance matrix to keep three columns; here each thread keeps a
// running best-3 while the known points stream through shared memory.
#include "common.cuh"

namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_CHUNK = 4096;  // known points resident per pass (64 KB as float4)

// Order is (distance, index): strict '<' while scanning ascending indices.
__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, float* __restrict__ dist_out,
                int64_t* __restrict__ idx_out, int N, int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw);
  const int b = blockIdx.y;
  const int n = blockIdx.x * NN_THREADS + threadIdx.x;
  const bool live = n < N;
  const float* q = unknown + ((size_t)b * N + (live ? n : N - 1)) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float qn = ppt_sqnorm3(qx, qy, qz);
  const float inf = __int_as_float(0x7f800000);
  float d0 = inf, d1 = inf, d2 = inf;
  int i0 = 0, i1 = 0, i2 = 0;
  const float* kb = known + (size_t)b * S * 3;
  for (int c0 = 0; c0 < S; c0 += NN_CHUNK) {
    const int cn = min(NN_CHUNK, S - c0);
    if (c0) __syncthreads();
    for (int s = threadIdx.x; s < cn; s += NN_THREADS) {
      const float* p = kb + (size_t)(c0 + s) * 3;
      float4 v;
      v.x = p[0]; v.y = p[1]; v.z = p[2];
      v.w = ppt_sqnorm3(v.x, v.y, v.z);
      pts[s] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < cn; ++s) {
      const float4 p = pts[s];  // same address across the warp: broadcast
      const float d = ppt_pair_sqdist(qx, qy, qz, qn, p.x, p.y, p.z, p.w);
      const int i = c0 + s;
      if (d < d2) {
        if (d < d1) {
          d2 = d1; i2 = i1;
          if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
          else { d1 = d; i1 = i; }
        } else { d2 = d; i2 = i; }
      }
    }
  }
  if (live) {
    const size_t o = ((size_t)b * N + n) * 3;
    dist_out[o] = d0; dist_out[o + 1] = d1; dist_out[o + 2] = d2;
    idx_out[o] = i0; idx_out[o + 1] = i1; idx_out[o + 2] = i2;
  }
}

// pointnet2_utils.py:304-306, SURVEY.md F8: r = 1/(d + 1e-8), norm = (r0+r1)+r2, w = r/norm.
__device__ __forceinline__ void interp_weights(const float* __restrict__ dist, float& w0, float& w1, float& w2) {
  const float r0 = __fdiv_rn(1.0f, __fadd_rn(dist[0], 1e-8f));
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(dist[1], 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(dist[2], 1e-8f));
  const float nrm = __fadd_rn(__fadd_rn(r0, r1), r2);
  w0 = __fdiv_rn(r0, nrm); w1 = __fdiv_rn(r1, nrm); w2 = __fdiv_rn(r2, nrm);
}

__device__ __forceinline__ float interp3(float w0, float a, float w1, float b, float w2, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b)), __fmul_rn(w2, c));
}

// One warp per FOUR unknown points; lanes stride the channels (coalesced rows of feats / out).  Lanes 0-3 each
// fetch one point's three (distance, index) pairs and run its six IEEE divisions -- once per point instead of once
// per lane -- and broadcast the weights; the twelve row loads of a channel chunk are issued before the four stores.
constexpr int TI_PTS = 4;

__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float* __restrict__ feats, const int64_t* __restrict__ idx,
                         const float* __restrict__ dist, float* __restrict__ out, long long total, int N, int S, int D) {
  const int lane = threadIdx.x & 31;
  const long long p0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * TI_PTS;  // global point index b * N + n
  if (p0 >= total) return;
  float mw0 = 0.f, mw1 = 0.f, mw2 = 0.f;
  long long mr0 = 0, mr1 = 0, mr2 = 0;  // feats row (b * S + idx) of the three neighbours
  if (lane < TI_PTS && p0 + lane < total) {
    const size_t o = (size_t)(p0 + lane) * 3;
    interp_weights(dist + o, mw0, mw1, mw2);
    const long long base = ((p0 + lane) / N) * S;
    mr0 = base + idx[o]; mr1 = base + idx[o + 1]; mr2 = base + idx[o + 2];
  }
  float w0[TI_PTS], w1[TI_PTS], w2[TI_PTS];
  const float *f0[TI_PTS], *f1[TI_PTS], *f2[TI_PTS];
#pragma unroll
  for (int j = 0; j < TI_PTS; ++j) {
    w0[j] = __shfl_sync(PPT_FULL_MASK, mw0, j); w1[j] = __shfl_sync(PPT_FULL_MASK, mw1, j);
    w2[j] = __shfl_sync(PPT_FULL_MASK, mw2, j);
    f0[j] = feats + (size_t)__shfl_sync(PPT_FULL_MASK, mr0, j) * D;
    f1[j] = feats + (size_t)__shfl_sync(PPT_FULL_MASK, mr1, j) * D;
    f2[j] = feats + (size_t)__shfl_sync(PPT_FULL_MASK, mr2, j) * D;
  }
  float* y = out + (size_t)p0 * D;
  const int npts = total - p0 < TI_PTS ? (int)(total - p0) : TI_PTS;
  if ((D & 3) == 0) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 a[TI_PTS], bb[TI_PTS], cc[TI_PTS];
#pragma unroll
      for (int j = 0; j < TI_PTS; ++j) {
        a[j] = __ldg(reinterpret_cast<const float4*>(f0[j] + c));
        bb[j] = __ldg(reinterpret_cast<const float4*>(f1[j] + c));
        cc[j] = __ldg(reinterpret_cast<const float4*>(f2[j] + c));
      }
#pragma unroll
      for (int j = 0; j < TI_PTS; ++j) {
        if (j >= npts) break;
        float4 r;
        r.x = interp3(w0[j], a[j].x, w1[j], bb[j].x, w2[j], cc[j].x);
        r.y = interp3(w0[j], a[j].y, w1[j], bb[j].y, w2[j], cc[j].y);
        r.z = interp3(w0[j], a[j].z, w1[j], bb[j].z, w2[j], cc[j].z);
        r.w = interp3(w0[j], a[j].w, w1[j], bb[j].w, w2[j], cc[j].w);
        __stcs(reinterpret_cast<float4*>(y + (size_t)j * D + c), r);  // written once: streaming store
      }
    }
  } else {
    for (int j = 0; j < npts; ++j)
      for (int c = lane; c < D; c += 32)
        y[(size_t)j * D + c] = interp3(w0[j], __ldg(f0[j] + c), w1[j], __ldg(f1[j] + c), w2[j], __ldg(f2[j] + c));
  }
}

// d out / d feats: scatter-add of w_i * grad_out rows (red.global.add.f32).
__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                              const float* __restrict__ dist, float* __restrict__ grad_feats, int N, int S, int D) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const size_t o = ((size_t)b * N + n) * 3;
  float w0, w1, w2;
  interp_weights(dist + o, w0, w1, w2);
  float* g0 = grad_feats + ((size_t)b * S + idx[o]) * D;
  float* g1 = grad_feats + ((size_t)b * S + idx[o + 1]) * D;
  float* g2 = grad_feats + ((size_t)b * S + idx[o + 2]) * D;
  const float* gy = grad_out + ((size_t)b * N + n) * D;
  for (int c = lane; c < D; c += 32) {
    const float g = __ldg(gy + c);
    atomicAdd(g0 + c, __fmul_rn(w0, g));
    atomicAdd(g1 + c, __fmul_rn(w1, g));
    atomicAdd(g2 + c, __fmul_rn(w2, g));
  }
}

}  // namespace

extern "C" PPT_EXPORT int ppt_three_nn(const float* unknown, const float* known, float* dist_out, int64_t* idx_out, int B, int N,
                            int S, void* stream) {
  if (!unknown || !known || !dist_out || !idx_out || B < 0 || N < 1) return PPT_EINVAL;
  if (S < 3) return PPT_ERANGE;
  if (B == 0) return 0;
  if (B > 65535) return PPT_ERANGE;
  const size_t smem = (size_t)(S < NN_CHUNK ? S : NN_CHUNK) * sizeof(float4);
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(NN_CHUNK * sizeof(float4))));
  }
  dim3 grid((N + NN_THREADS - 1) / NN_THREADS, B);
  three_nn_kernel<<<grid, NN_THREADS, smem, (cudaStream_t)stream>>>(unknown, known, dist_out, idx_out, N, S);
  return ppt_launch_status();
}

extern "C" PPT_EXPORT int ppt_three_interpolate(const float* feats, const int64_t* idx, const float* dist, float* out, int B,
                                     int N, int S, int D, void* stream) {
  if (!feats || !idx || !dist || !out || B < 0 || N < 1 || S < 1 || D < 1) return PPT_EINVAL;
  if (B == 0) return 0;
  const long long total = (long long)B * N;
  const long long blocks = (total + 8 * TI_PTS - 1) / (8 * TI_PTS);
  if (blocks > 0x7fffffffll) return PPT_ERANGE;
  three_interpolate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(feats, idx, dist, out, total, N, S, D);
  return ppt_launch_status();
}

extern "C" PPT_EXPORT int ppt_three_interpolate_grad(const float* grad_out, const int64_t* idx, const float* dist,
                                          float* grad_feats, int B, int N, int S, int D, void* stream) {
  if (!grad_out || !idx || !dist || !grad_feats || B < 0 || N < 1 || S < 1 || D < 1) return PPT_EINVAL;
  if (B == 0) return 0;
  if (B > 65535) return PPT_ERANGE;
  dim3 grid((N + 7) / 8, B);
  three_interpolate_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, dist, grad_feats, N, S, D);
  return ppt_launch_status();
}
