// DGCNN_Propagation.get_graph_feature (models/pointbert/pointnet2_utils.py:392-442) behind the kNN, sm_100a:
//     out[b, c,     q, j] = x_k[b, c, idx[b, q, j]] - x_q[b, c, q]          c < C
//     out[b, C + c, q, j] = x_q[b, c, q]
// The reference builds this with a flat advanced-indexing gather, a permute + contiguous, an expand and a cat
// (about four passes over the [B, 2C, Nq, k] tensor); here it is one pass, bound by the HBM write.  One thread
// per (b, c, q, j): consecutive threads write consecutive floats of both halves, read consecutive idx entries and
// share the x_q value; the x_k gather hits a 4*Nk-byte row that stays in L1/L2.  The subtraction is a single
// rounded fp32 operation in both implementations: results are bit-identical.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
graph_feature_kernel(const float* __restrict__ x_q, const float* __restrict__ x_k, const int64_t* __restrict__ idx,
                     float* __restrict__ out, int C, int Nq, int Nk, int k, long long total /* B*C*Nq*k */) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long per_c = (long long)Nq * k;
  const long long bc = e / per_c;           // b * C + c
  const long long r = e - bc * per_c;       // q * k + j
  const int b = (int)(bc / C), c = (int)(bc - (long long)b * C);
  const int q = (int)(r / k);
  const long long n = __ldg(idx + (long long)b * per_c + r);
  const float xq = __ldg(x_q + bc * Nq + q);
  const float xk = __ldg(x_k + bc * Nk + n);
  out[((long long)b * 2 * C + c) * per_c + r] = __fsub_rn(xk, xq);
  out[((long long)b * 2 * C + C + c) * per_c + r] = xq;
}

// k == 4 (the part-seg head, point_encoder.py:303-304): one thread per (b, c, q), 16-byte loads of the four
// indices' int64 pairs and 16-byte stores of both output halves.
__global__ void __launch_bounds__(256)
graph_feature_k4_kernel(const float* __restrict__ x_q, const float* __restrict__ x_k, const int64_t* __restrict__ idx,
                        float* __restrict__ out, int C, int Nq, int Nk, long long total /* B*C*Nq */) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long bc = e / Nq;
  const int q = (int)(e - bc * Nq);
  const int b = (int)(bc / C), c = (int)(bc - (long long)b * C);
  const longlong2* ip = reinterpret_cast<const longlong2*>(idx + ((long long)b * Nq + q) * 4);
  const longlong2 i01 = __ldg(ip), i23 = __ldg(ip + 1);
  const float xq = __ldg(x_q + e);
  const float* row = x_k + bc * Nk;
  const float4 d = make_float4(__fsub_rn(__ldg(row + i01.x), xq), __fsub_rn(__ldg(row + i01.y), xq),
                               __fsub_rn(__ldg(row + i23.x), xq), __fsub_rn(__ldg(row + i23.y), xq));
  const long long per_c = (long long)Nq * 4;
  *reinterpret_cast<float4*>(out + ((long long)b * 2 * C + c) * per_c + (long long)q * 4) = d;
  *reinterpret_cast<float4*>(out + ((long long)b * 2 * C + C + c) * per_c + (long long)q * 4) = make_float4(xq, xq, xq, xq);
}

// grad_xq[b,c,q] = sum_j (g[b,C+c,q,j] - g[b,c,q,j]);  grad_xk[b,c,idx[b,q,j]] += g[b,c,q,j]  (zero-filled by the caller)
__global__ void __launch_bounds__(256)
graph_feature_grad_kernel(const float* __restrict__ gout, const int64_t* __restrict__ idx, float* __restrict__ grad_xq,
                          float* __restrict__ grad_xk, int C, int Nq, int Nk, int k, long long total /* B*C*Nq */) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long bc = e / Nq;
  const int q = (int)(e - bc * Nq);
  const int b = (int)(bc / C), c = (int)(bc - (long long)b * C);
  const long long per_c = (long long)Nq * k;
  const float* g0 = gout + ((long long)b * 2 * C + c) * per_c + (long long)q * k;
  const float* g1 = gout + ((long long)b * 2 * C + C + c) * per_c + (long long)q * k;
  const int64_t* ib = idx + (long long)b * per_c + (long long)q * k;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const float a = __ldg(g0 + j);
    acc += __ldg(g1 + j) - a;
    atomicAdd(grad_xk + bc * Nk + __ldg(ib + j), a);
  }
  grad_xq[e] = acc;
}

}  // namespace

extern "C" PPT_EXPORT int ppt_graph_feature(const float* x_q, const float* x_k, const int64_t* idx, float* out, int B,
                                            int C, int Nq, int Nk, int k, void* stream) {
  if (!x_q || !x_k || !idx || !out || B < 1 || C < 1 || Nq < 1 || Nk < 1 || k < 1) return PPT_EINVAL;
  const long long total = (long long)B * C * Nq * k;
  if ((total + 255) / 256 > 0x7fffffffll) return PPT_ERANGE;
  if (k == 4 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(idx) & 15) == 0) {
    const long long t4 = total / 4;
    graph_feature_k4_kernel<<<(unsigned)((t4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_q, x_k, idx, out, C, Nq,
                                                                                           Nk, t4);
    return ppt_launch_status();
  }
  graph_feature_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_q, x_k, idx, out, C, Nq, Nk,
                                                                                         k, total);
  return ppt_launch_status();
}

extern "C" PPT_EXPORT int ppt_graph_feature_grad(const float* grad_out, const int64_t* idx, float* grad_xq,
                                                 float* grad_xk, int B, int C, int Nq, int Nk, int k, void* stream) {
  if (!grad_out || !idx || !grad_xq || !grad_xk || B < 1 || C < 1 || Nq < 1 || Nk < 1 || k < 1) return PPT_EINVAL;
  const long long total = (long long)B * C * Nq;
  if ((total + 255) / 256 > 0x7fffffffll) return PPT_ERANGE;
  graph_feature_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, grad_xq,
                                                                                              grad_xk, C, Nq, Nk, k, total);
  return ppt_launch_status();
}
