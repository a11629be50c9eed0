// DGCNN_Propagation layer (models/pointbert/pointnet2_utils.py:382-390, 444-467) behind its two per-point GEMMs, sm_100a.
//
// The reference evaluates  Conv2d(2C -> Co, 1x1, no bias) on cat(x_k[idx] - x_q, x_q)  for every (query, neighbour)
// pair, then GroupNorm, LeakyReLU and a max over the k neighbours.  The convolution is linear, so with W = [Wa | Wb]
//     y[b, :, q, j] = Wa x_k[b, :, idx[b,q,j]] + (Wb - Wa) x_q[b, :, q] = U[b, :, idx[b,q,j]] + V[b, :, q]
// where U = Wa x_k and V = (Wb - Wa) x_q are two plain per-point GEMMs (k times fewer FLOPs than the pairwise
// convolution, and the [B, 2C, Nq, k] edge tensor never exists).  This file is everything after those GEMMs:
//   edge_gn_stats_kernel   sum and sum of squares of y per (sample, channel group)          -> GroupNorm statistics
//   edge_gn_max_kernel     out[b, c, q] = max_j LeakyReLU(gamma_c (y - mean) rstd + beta_c)
// Both read U rows through the index (a 4 Nk-byte row per (b, c): L1 / L2 resident), V and the indices coalesced.
// One thread per (b, c, q); four indices per thread are loaded once (k <= EC_MAX_K).
#include "common.cuh"

namespace {

constexpr int EC_MAX_K = 16;

template <bool STATS>
__global__ void __launch_bounds__(256)
edge_gn_kernel(const float* __restrict__ U, const float* __restrict__ V, const int64_t* __restrict__ idx,
               double* __restrict__ sums,              // [B, G, 2]: STATS accumulates, else read
               const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out, int C, int Nq,
               int Nk, int k, int G, float eps, float slope) {
  // grid: (ceil(Nq / 256), C, B): a block lies inside one (b, c) row, hence inside one GroupNorm group
  const int q = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  const int cg = C / G, g = c / cg;
  const long long bc = (long long)b * C + c;
  const float* urow = U + bc * Nk;
  float y[EC_MAX_K];
  float s1 = 0.f, s2 = 0.f;
  if (q < Nq) {
    const float v = __ldg(V + bc * Nq + q);
    const int64_t* ip = idx + ((long long)b * Nq + q) * k;
#pragma unroll
    for (int j = 0; j < EC_MAX_K; ++j) {
      if (j < k) {
        const int64_t n = __ldg(ip + j);
        y[j] = ((uint64_t)n < (uint64_t)Nk ? __ldg(urow + n) : __int_as_float(0x7fc00000)) + v;
        s1 += y[j];
        s2 = fmaf(y[j], y[j], s2);
      }
    }
  }
  if (STATS) {
    __shared__ float red[2][8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s1 += __shfl_xor_sync(PPT_FULL_MASK, s1, off);
      s2 += __shfl_xor_sync(PPT_FULL_MASK, s2, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += (double)red[threadIdx.x][w];
      atomicAdd(sums + ((long long)b * G + g) * 2 + threadIdx.x, t);
    }
  } else if (q < Nq) {
    const double n = (double)cg * (double)Nq * (double)k;
    const double mean = sums[((long long)b * G + g) * 2] / n;
    double var = sums[((long long)b * G + g) * 2 + 1] / n - mean * mean;  // biased, like nn.GroupNorm
    var = var < 0.0 ? 0.0 : var;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = __ldg(gamma + c) * rstd, d = __ldg(beta + c) - (float)mean * a;
    float best = -__int_as_float(0x7f800000);
#pragma unroll
    for (int j = 0; j < EC_MAX_K; ++j) {
      if (j < k) {
        const float z = fmaf(y[j], a, d);
        best = fmaxf(best, z > 0.f ? z : z * slope);
      }
    }
    out[bc * Nq + q] = best;
  }
}

}  // namespace

extern "C" PPT_EXPORT int64_t ppt_edge_gn_workspace_bytes(int B, int G) {
  if (B < 1 || G < 1) return PPT_EINVAL;
  return (int64_t)B * G * 2 * (int64_t)sizeof(double);
}

extern "C" PPT_EXPORT int ppt_edge_gn_max_forward(const float* U, const float* V, const int64_t* idx, const float* gamma,
                                                  const float* beta, void* workspace, float* out, int B, int C, int Nq,
                                                  int Nk, int k, int G, float eps, float slope, void* stream) {
  if (!U || !V || !idx || !gamma || !beta || !workspace || !out || B < 0 || C < 1 || Nq < 1 || Nk < 1) return PPT_EINVAL;
  if (k < 1 || k > EC_MAX_K || G < 1 || C % G != 0 || !(eps > 0.f)) return PPT_ERANGE;
  if (B == 0) return 0;
  if (B > 65535 || C > 65535) return PPT_ERANGE;
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = static_cast<double*>(workspace);
  PPT_RETURN_IF_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * G * 2 * sizeof(double), st));
  dim3 grid((Nq + 255) / 256, C, B);
  edge_gn_kernel<true><<<grid, 256, 0, st>>>(U, V, idx, sums, gamma, beta, out, C, Nq, Nk, k, G, eps, slope);
  edge_gn_kernel<false><<<grid, 256, 0, st>>>(U, V, idx, sums, gamma, beta, out, C, Nq, Nk, k, G, eps, slope);
  return ppt_launch_status();
}
