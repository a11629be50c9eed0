// kNN and the fused Group tail (kNN -> gather -> centre) for sm_100a.
//
// Replaces knn_point / square_distance+topk (models/pointbert/dvae.py:116-149)
// and the gather + centring of Group.forward (dvae.py:174-180).  The S x N
// distance matrix of the reference (537 MB at B=32) is never materialised.
//
// CTA = (cloud, tile of 64 queries).  The cloud is staged once per CTA in
// shared memory as float4 {x, y, z, |p|^2} (128 KB for 8192 points; larger
// clouds are streamed in 8192-point chunks).  Each warp owns four queries at a
// time so that one conflict-free LDS.128 per lane feeds four distance
// evaluations (one query per warp would be shared-memory-bandwidth bound).
// Per query the warp keeps the k best (distance, index) pairs sorted across its
// lanes; a row of 32 points is tested against the current k-th distance with
// one compare per lane and a warp vote, and the rare survivors are inserted
// with shuffles.  Points are visited in ascending index order, so a later
// point with an equal distance never displaces an earlier one: the result is
// the k smallest under (distance, index), ascending -- the oracle's rule
// (SURVEY.md F6).  Distances use the reference's exact formula (F2) and may be
// negative (F3).
#include "common.cuh"
#include "spatial_index.cuh"

namespace {

constexpr int KNN_THREADS = 512;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int KNN_QW = 4;                       // queries per warp
constexpr int KNN_QPB = KNN_WARPS * KNN_QW;     // queries per CTA = 64
constexpr int KNN_CHUNK = 8192;                 // points resident in shared memory

struct TopK {  // one query's running result, lane i holds the i-th smallest
  float d;
  int i;
  float tau;  // distance in lane k-1 (warp-uniform)
};

__device__ __forceinline__ void topk_insert_row(TopK& t, float d, bool pred, int rowbase, int k, int lane) {
  unsigned bal = __ballot_sync(PPT_FULL_MASK, pred);
  while (bal) {
    const int src = __ffs(bal) - 1;
    bal &= bal - 1;
    const float cd = __shfl_sync(PPT_FULL_MASK, d, src);
    if (!(cd < t.tau)) continue;  // tau may have dropped since the vote (warp-uniform branch)
    // Entries <= cd stay in front: equal distances keep their (lower) indices first.
    const int pos = __popc(__ballot_sync(PPT_FULL_MASK, t.d <= cd));
    const float ud = __shfl_up_sync(PPT_FULL_MASK, t.d, 1);
    const int ui = __shfl_up_sync(PPT_FULL_MASK, t.i, 1);
    if (lane == pos) { t.d = cd; t.i = rowbase + src; }
    else if (lane > pos) { t.d = ud; t.i = ui; }
    t.tau = __shfl_sync(PPT_FULL_MASK, t.d, k - 1);
  }
}

template <bool GROUP>
__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_kernel(const float* __restrict__ xyz, const float* __restrict__ query, int64_t* __restrict__ idx_out,
           float* __restrict__ dist_out, float* __restrict__ nb_out, int N, int S, int k, int tiles_per_cloud) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / tiles_per_cloud;
  const int q0 = (blockIdx.x - b * tiles_per_cloud) * KNN_QPB + warp * KNN_QW;
  const float* cloud = xyz + (size_t)b * N * 3;

  float qx[KNN_QW], qy[KNN_QW], qz[KNN_QW], qn[KNN_QW];
  TopK top[KNN_QW];
#pragma unroll
  for (int u = 0; u < KNN_QW; ++u) {
    const int q = min(q0 + u, S - 1);  // clamp: surplus lanes recompute a valid query, never stored
    const float* p = query + ((size_t)b * S + q) * 3;
    qx[u] = p[0]; qy[u] = p[1]; qz[u] = p[2];
    qn[u] = ppt_sqnorm3(qx[u], qy[u], qz[u]);
    top[u].d = __int_as_float(0x7f800000);
    top[u].i = 0x7fffffff;
    top[u].tau = __int_as_float(0x7f800000);
  }

  for (int c0 = 0; c0 < N; c0 += KNN_CHUNK) {
    const int cn = min(KNN_CHUNK, N - c0);
    const int rows = (cn + 31) >> 5;
    if (c0) __syncthreads();  // previous chunk fully consumed
    for (int n = tid; n < rows * 32; n += KNN_THREADS) {
      float4 v;
      if (n < cn) {
        const float* p = cloud + (size_t)(c0 + n) * 3;
        v.x = p[0]; v.y = p[1]; v.z = p[2];
        v.w = ppt_sqnorm3(v.x, v.y, v.z);
      } else {
        v = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));  // d = +inf: never below tau
      }
      pts[n] = v;
    }
    __syncthreads();

    for (int r = 0; r < rows; ++r) {
      const float4 p = pts[r * 32 + lane];
      float d[KNN_QW];
      bool pr[KNN_QW];
      bool any = false;
#pragma unroll
      for (int u = 0; u < KNN_QW; ++u) {
        d[u] = ppt_pair_sqdist(qx[u], qy[u], qz[u], qn[u], p.x, p.y, p.z, p.w);
        pr[u] = d[u] < top[u].tau;
        any |= pr[u];
      }
      if (__any_sync(PPT_FULL_MASK, any)) {
        const int rowbase = c0 + r * 32;
#pragma unroll
        for (int u = 0; u < KNN_QW; ++u) topk_insert_row(top[u], d[u], pr[u], rowbase, k, lane);
      }
    }
  }

#pragma unroll
  for (int u = 0; u < KNN_QW; ++u) {
    const int q = q0 + u;
    if (q >= S || lane >= k) continue;
    const size_t o = ((size_t)b * S + q) * k + lane;
    if (idx_out) idx_out[o] = (int64_t)top[u].i;
    if (dist_out) dist_out[o] = top[u].d;
    if (GROUP) {
      // dvae.py:177-180: neighborhood = xyz[idx] - center  (one fp32 subtract each)
      // NaN coordinates leave unfilled slots (index 0x7fffffff): NaN rows, not an out-of-bounds read
      const bool ok = (unsigned)top[u].i < (unsigned)N;
      const float* p = cloud + (size_t)(ok ? top[u].i : 0) * 3;
      const float nanv = __int_as_float(0x7fc00000);
      nb_out[o * 3 + 0] = ok ? __fsub_rn(p[0], qx[u]) : nanv;
      nb_out[o * 3 + 1] = ok ? __fsub_rn(p[1], qy[u]) : nanv;
      nb_out[o * 3 + 2] = ok ? __fsub_rn(p[2], qz[u]) : nanv;
    }
  }
}

}  // namespace

namespace {

template <bool GROUP>
int launch_knn(const float* xyz, const float* query, int64_t* idx_out, float* dist_out, float* nb_out,
               const void* index, int B, int N, int S, int k, cudaStream_t st) {
  if (!xyz || !query || B < 0 || N < 1 || S < 1) return PPT_EINVAL;
  if (k < 1 || k > 32 || k > N) return PPT_ERANGE;
  if (B == 0) return 0;
  if (index && spidx::supported(N))
    return ppt_knn_grid_search(xyz, query, index, idx_out, dist_out, nb_out, B, N, S, k, GROUP, st);
  auto kern = knn_kernel<GROUP>;
  const int resident = N < KNN_CHUNK ? ((N + 31) / 32) * 32 : KNN_CHUNK;
  const size_t smem = (size_t)resident * sizeof(float4);
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(KNN_CHUNK * sizeof(float4))));
  }
  const int tiles = (S + KNN_QPB - 1) / KNN_QPB;
  kern<<<(unsigned)(B * tiles), KNN_THREADS, smem, st>>>(xyz, query, idx_out, dist_out, nb_out, N, S, k, tiles);
  return ppt_launch_status();
}

__global__ void square_distance_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                       float* __restrict__ out, int S, int N) {
  // grid: (ceil(N/256), S, B); one query row per blockIdx.y
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y, b = blockIdx.z;
  const float* q = src + ((size_t)b * S + s) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float qn = ppt_sqnorm3(qx, qy, qz);
  if (n < N) {
    const float* p = dst + ((size_t)b * N + n) * 3;
    const float x = p[0], y = p[1], z = p[2];
    out[((size_t)b * S + s) * N + n] = ppt_pair_sqdist(qx, qy, qz, qn, x, y, z, ppt_sqnorm3(x, y, z));
  }
}

}  // namespace

extern "C" PPT_EXPORT int64_t ppt_spatial_index_bytes(int B, int N) {
  if (B < 0 || N < 1) return PPT_EINVAL;
  return (int64_t)ppt_index_bytes(B, N);
}

extern "C" PPT_EXPORT int ppt_spatial_index_build(const float* xyz, void* index, int B, int N, void* stream) {
  if (!xyz || !index || B < 0 || N < 1) return PPT_EINVAL;
  if (!spidx::supported(N)) return PPT_ERANGE;
  if (B == 0) return 0;
  return ppt_index_build(xyz, index, B, N, (cudaStream_t)stream);
}

extern "C" PPT_EXPORT int ppt_knn(const float* xyz, const float* query, int64_t* idx_out, float* dist_out,
                                  const void* index, int B, int N, int S, int k, void* stream) {
  if (!idx_out) return PPT_EINVAL;
  return launch_knn<false>(xyz, query, idx_out, dist_out, nullptr, index, B, N, S, k, (cudaStream_t)stream);
}

extern "C" PPT_EXPORT int ppt_knn_group(const float* xyz, const float* center, float* neighborhood_out,
                                        int64_t* idx_out, const void* index, int B, int N, int G, int k,
                                        void* stream) {
  if (!neighborhood_out) return PPT_EINVAL;
  return launch_knn<true>(xyz, center, idx_out, nullptr, neighborhood_out, index, B, N, G, k, (cudaStream_t)stream);
}

extern "C" PPT_EXPORT int ppt_square_distance(const float* src, const float* dst, float* out, int B, int S, int N, void* stream) {
  if (!src || !dst || !out || B < 0 || S < 1 || N < 1) return PPT_EINVAL;
  if (B == 0) return 0;
  if (S > 65535 || B > 65535) return PPT_ERANGE;
  dim3 grid((N + 255) / 256, S, B);
  square_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, out, S, N);
  return ppt_launch_status();
}
