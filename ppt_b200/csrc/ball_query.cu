// Ball query for sm_100a: tiled scan in index order with warp-ballot compaction.
//
// Replaces query_ball_point (models/pointnet2/pointnet2_utils.py:87-107 and its
// copies).  The reference materialises a (B,S,N) int64 tensor and fully sorts it;
// here each warp scans the cloud once per four queries and stops as soon as all
// four have `nsample` members.  Membership is !(d > radius^2) with d from the
// reference's matmul-form distance, members keep ascending
// index order, the tail is padded with the first member, and a query without any
// member yields N in every slot (the reference's sentinel).
#include "common.cuh"

namespace {

constexpr int BQ_THREADS = 256;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_QW = 4;
constexpr int BQ_QPB = BQ_WARPS * BQ_QW;  // 32 queries per CTA
constexpr int BQ_CHUNK = 8192;

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int64_t* __restrict__ idx_out,
                  float thr, int N, int S, int nsample, int tiles_per_cloud) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x - b * tiles_per_cloud) * BQ_QPB + warp * BQ_QW;
  const float* cloud = xyz + (size_t)b * N * 3;
  const unsigned lt_mask = (1u << lane) - 1u;

  float qx[BQ_QW], qy[BQ_QW], qz[BQ_QW], qn[BQ_QW];
  int cnt[BQ_QW], first[BQ_QW];
  int64_t* out[BQ_QW];
#pragma unroll
  for (int u = 0; u < BQ_QW; ++u) {
    const int q = min(q0 + u, S - 1);
    const float* p = new_xyz + ((size_t)b * S + q) * 3;
    qx[u] = p[0]; qy[u] = p[1]; qz[u] = p[2];
    qn[u] = ppt_sqnorm3(qx[u], qy[u], qz[u]);
    cnt[u] = q0 + u < S ? 0 : nsample;  // surplus queries are born complete
    first[u] = N;
    out[u] = idx_out + ((size_t)b * S + q) * nsample;
  }

  for (int c0 = 0; c0 < N; c0 += BQ_CHUNK) {
    const int cn = min(BQ_CHUNK, N - c0);
    const int rows = (cn + 31) >> 5;
    if (c0) __syncthreads();
    for (int n = tid; n < rows * 32; n += BQ_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // padding is excluded by the n < N test below
      if (n < cn) {
        const float* p = cloud + (size_t)(c0 + n) * 3;
        v.x = p[0]; v.y = p[1]; v.z = p[2];
        v.w = ppt_sqnorm3(v.x, v.y, v.z);
      }
      pts[n] = v;
    }
    __syncthreads();

    for (int r = 0; r < rows; ++r) {
      bool open = false;
#pragma unroll
      for (int u = 0; u < BQ_QW; ++u) open |= cnt[u] < nsample;
      if (!open) break;  // warp-uniform: all four queries are full
      const float4 p = pts[r * 32 + lane];
      const int n = c0 + r * 32 + lane;
#pragma unroll
      for (int u = 0; u < BQ_QW; ++u) {
        const float d = ppt_pair_sqdist(qx[u], qy[u], qz[u], qn[u], p.x, p.y, p.z, p.w);
        const bool member = n < N && !(d > thr);
        const unsigned bal = __ballot_sync(PPT_FULL_MASK, member);
        if (bal && cnt[u] < nsample) {
          if (cnt[u] == 0) first[u] = c0 + r * 32 + __ffs(bal) - 1;
          const int slot = cnt[u] + __popc(bal & lt_mask);
          if (member && slot < nsample) out[u][slot] = (int64_t)n;
          cnt[u] += __popc(bal);
        }
      }
    }
  }

#pragma unroll
  for (int u = 0; u < BQ_QW; ++u) {
    if (q0 + u >= S) continue;
    for (int slot = cnt[u] + lane; slot < nsample; slot += 32) out[u][slot] = (int64_t)first[u];
  }
}

}  // namespace

extern "C" PPT_EXPORT int ppt_ball_query(const float* xyz, const float* new_xyz, int64_t* idx_out, float radius2, int B, int N,
                              int S, int nsample, void* stream) {
  if (!xyz || !new_xyz || !idx_out || B < 0 || N < 1 || S < 1 || nsample < 1) return PPT_EINVAL;
  if (B == 0) return 0;
  const int resident = N < BQ_CHUNK ? ((N + 31) / 32) * 32 : BQ_CHUNK;
  const size_t smem = (size_t)resident * sizeof(float4);
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(BQ_CHUNK * sizeof(float4))));
  }
  const int tiles = (S + BQ_QPB - 1) / BQ_QPB;
  ball_query_kernel<<<(unsigned)(B * tiles), BQ_THREADS, smem, (cudaStream_t)stream>>>(xyz, new_xyz, idx_out, radius2,
                                                                                      N, S, nsample, tiles);
  return ppt_launch_status();
}
