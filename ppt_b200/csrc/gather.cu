// Index gathers for sm_100a: index_points and the grouping tail of
// sample_and_group / PointNetSetAbstractionMsg.  Pure HBM/L2 traffic: every
// thread produces four consecutive output floats (one 16-byte store when the
// output is 16-byte aligned) and reads the gathered rows with coalesced loads.
#include "common.cuh"

namespace {

// index_points (models/pointbert/misc.py:26-42): out[b,m,:] = points[b,idx[b,m],:]
__global__ void gather_kernel(const float* __restrict__ points, const int64_t* __restrict__ idx,
                              float* __restrict__ out, int N, int C, int M) {
  const int b = blockIdx.y;
  const size_t total = (size_t)M * C;
  const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e0 >= total) return;
  const float* src = points + (size_t)b * N * C;
  const int64_t* ib = idx + (size_t)b * M;
  float* dst = out + (size_t)b * total;
  float v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const size_t e = e0 + t;
    if (e < total) {
      const int m = (int)(e / C), c = (int)(e - (size_t)m * C);
      const int64_t n = ib[m];
      // an index outside [0, N) (e.g. query_ball_point's empty-ball sentinel N, which makes the reference raise
      // an IndexError) yields NaN instead of an out-of-bounds read
      v[t] = (uint64_t)n < (uint64_t)N ? __ldg(src + (size_t)n * C + c) : __int_as_float(0x7fc00000);
    } else {
      v[t] = 0.f;
    }
  }
  if (e0 + 3 < total && ((reinterpret_cast<uintptr_t>(dst + e0) & 15) == 0)) {
    *reinterpret_cast<float4*>(dst + e0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (e0 + t < total) dst[e0 + t] = v[t];
  }
}

// models/pointnet2/pointnet2_utils.py:127-134 (SSG: [xyz - centre, feats]) and
// :244-254 (MSG: [feats, xyz - centre]).  Row = one (b, s, j) neighbour, 3 + D floats.
__global__ void group_concat_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                    const float* __restrict__ points, const int64_t* __restrict__ idx,
                                    float* __restrict__ out, int N, int S, int K, int D, int xyz_first) {
  const int b = blockIdx.y;
  const int C = 3 + D;
  const size_t total = (size_t)S * K * C;
  const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e0 >= total) return;
  const float* cloud = xyz + (size_t)b * N * 3;
  const float* ctr = new_xyz + (size_t)b * S * 3;
  const float* feat = points ? points + (size_t)b * N * D : nullptr;
  const int64_t* ib = idx + (size_t)b * S * K;
  float* dst = out + (size_t)b * total;
  const int xyz_lo = xyz_first ? 0 : D;  // first column of the coordinate block
  float v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const size_t e = e0 + t;
    if (e < total) {
      const int row = (int)(e / C), c = (int)(e - (size_t)row * C);
      const int64_t n = ib[row];
      const int cx = c - xyz_lo;
      if ((uint64_t)n >= (uint64_t)N) {
        v[t] = __int_as_float(0x7fc00000);  // out-of-range index: NaN, never an out-of-bounds read
      } else if (cx >= 0 && cx < 3) {
        const int s = row / K;
        v[t] = __fsub_rn(__ldg(cloud + (size_t)n * 3 + cx), __ldg(ctr + (size_t)s * 3 + cx));
      } else {
        const int cf = xyz_first ? c - 3 : c;
        v[t] = __ldg(feat + (size_t)n * D + cf);
      }
    } else {
      v[t] = 0.f;
    }
  }
  if (e0 + 3 < total && ((reinterpret_cast<uintptr_t>(dst + e0) & 15) == 0)) {
    *reinterpret_cast<float4*>(dst + e0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (e0 + t < total) dst[e0 + t] = v[t];
  }
}

// Wide rows (C >= 32).  The element-wise kernels above spend a divide, an 8-byte index load and a scattered 4-byte
// load per element and reached 18 % of the HBM peak at C = 131; a warp-per-row version with direct 4-byte stores
// reached 27 % (rows of 131 floats start at odd 4-byte offsets, so every store straddles sectors).  Here a block
// assembles GR_ROWS consecutive output rows in shared memory -- warps fetch whole source rows with 16-byte loads --
// and then writes the block's contiguous slice of the output with 16-byte ALIGNED stores (GR_ROWS * C * 4 bytes is a
// multiple of 16 for every C), so the HBM sees full-sector streaming writes.
constexpr int GR_ROWS = 32;

template <bool CONCAT>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const float* __restrict__ points,
                   const int64_t* __restrict__ idx, float* __restrict__ out, long long total_rows, int rows_per_cloud,
                   int N, int K, int D, int xyz_first) {
  // CONCAT: out row = [xyz - centre | feats] (or feats first), C = 3 + D; else: out row = points row, C = D
  extern __shared__ __align__(16) float stage[];  // [GR_ROWS][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = CONCAT ? D + 3 : D;
  const int xyz_lo = xyz_first ? 0 : D, feat_lo = (CONCAT && xyz_first) ? 3 : 0;
  const float nanv = __int_as_float(0x7fc00000);
  const bool vec = (D & 3) == 0;
  // (cloud, row inside the cloud) of the step's first row, advanced incrementally: two 64-bit divisions per thread in
  // the whole kernel instead of one per step
  const long long step = (long long)gridDim.x * GR_ROWS;
  const long long step_b = step / rows_per_cloud;
  const unsigned step_in = (unsigned)(step - step_b * rows_per_cloud);
  long long b0 = ((long long)blockIdx.x * GR_ROWS) / rows_per_cloud;
  unsigned in0 = (unsigned)((long long)blockIdx.x * GR_ROWS - b0 * rows_per_cloud);
  for (long long row0 = (long long)blockIdx.x * GR_ROWS; row0 < total_rows; row0 += step) {
    const int nrows = total_rows - row0 < GR_ROWS ? (int)(total_rows - row0) : GR_ROWS;
    // Index arithmetic once per row, not once per row and lane: lane r of every warp locates row r of the step (one
    // 64-bit division per step, 32-bit ones per row), the four rows a warp then copies get their source row, centre
    // row and validity by shuffle.  (A first version did this per lane and per row and was issue-bound: 680 warp
    // instructions per 32-row step, 46 per float.)
    const unsigned in = in0 + (unsigned)lane;
    const unsigned over = in / (unsigned)rows_per_cloud;  // 0 unless the step crosses into the next cloud(s)
    const long long bl = b0 + over;
    const long long nl = lane < nrows ? __ldg(idx + row0 + lane) : 0;
    const bool okl = (unsigned long long)nl < (unsigned long long)N;  // out-of-range index: NaN row, no wild read
    const long long src_l = bl * N + (okl ? nl : 0);                  // row of `points` / `xyz`
    long long ctr_l = 0;
    if (CONCAT) ctr_l = bl * (rows_per_cloud / K) + (in - over * (unsigned)rows_per_cloud) / (unsigned)K;
#pragma unroll
    for (int rr = 0; rr < GR_ROWS / 8; ++rr) {
      const int r = warp + 8 * rr;
      if (r >= nrows) continue;  // warp-uniform
      const long long src = __shfl_sync(PPT_FULL_MASK, src_l, r);
      const bool ok = __shfl_sync(PPT_FULL_MASK, (int)okl, r) != 0;
      const float* frow = points + (size_t)src * D;
      float* srow = stage + r * C;
      if (CONCAT) {
        const long long ctr = __shfl_sync(PPT_FULL_MASK, ctr_l, r);
        if (lane < 3) {
          const float pv = __ldg(xyz + (size_t)src * 3 + lane);
          const float cv = __ldg(new_xyz + (size_t)ctr * 3 + lane);
          srow[xyz_lo + lane] = ok ? __fsub_rn(pv, cv) : nanv;
        }
      }
      if (vec) {
        for (int c = lane * 4; c < D; c += 128) {
          float4 v = __ldg(reinterpret_cast<const float4*>(frow + c));
          if (!ok) v = make_float4(nanv, nanv, nanv, nanv);
          float* d = srow + feat_lo + c;  // 4-byte aligned only (C is odd in the concat case)
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
      } else {
        for (int c = lane; c < D; c += 32) srow[feat_lo + c] = ok ? __ldg(frow + c) : nanv;
      }
    }
    __syncthreads();
    const int total = nrows * C;
    float* dst = out + (size_t)row0 * C;  // 16-byte aligned: row0 is a multiple of GR_ROWS
    const int total4 = total & ~3;
    for (int i = threadIdx.x * 4; i < total4; i += 1024)
      __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(stage + i));  // written once: streaming
    if (threadIdx.x < total - total4) dst[total4 + threadIdx.x] = stage[total4 + threadIdx.x];
    __syncthreads();
    b0 += step_b;
    in0 += step_in;
    if (in0 >= (unsigned)rows_per_cloud) { in0 -= (unsigned)rows_per_cloud; ++b0; }
  }
}

}  // namespace

static int gather_rows_launch_dims(long long total_rows, int C, size_t* smem) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  *smem = (size_t)GR_ROWS * C * sizeof(float);
  const long long want = (total_rows + GR_ROWS - 1) / GR_ROWS;
  int per_sm = (int)((200 * 1024) / (*smem + 1024));
  per_sm = per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm);
  const long long cap = (long long)sms * per_sm;
  return (int)(want < cap ? want : cap);
}

extern "C" PPT_EXPORT int ppt_gather(const float* points, const int64_t* idx, float* out, int B, int N, int C, int M,
                          void* stream) {
  if (!points || !idx || !out || B < 0 || N < 1 || C < 1 || M < 0) return PPT_EINVAL;
  if (B == 0 || M == 0) return 0;
  if (C >= 32 && C <= 1024) {
    const long long rows = (long long)B * M;
    size_t smem;
    const int grid = gather_rows_launch_dims(rows, C, &smem);
    static PptOncePerDevice configured;
    if (configured.need())
      PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(gather_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              GR_ROWS * 1024 * 4));
    gather_rows_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(nullptr, nullptr, points, idx, out, rows, M, N, 1,
                                                                         C, 1);
    return ppt_launch_status();
  }
  if (B > 65535) return PPT_ERANGE;
  const size_t quads = ((size_t)M * C + 3) / 4;
  dim3 grid((unsigned)((quads + 255) / 256), B);
  gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, out, N, C, M);
  return ppt_launch_status();
}

extern "C" PPT_EXPORT int ppt_group_concat(const float* xyz, const float* new_xyz, const float* points, const int64_t* idx,
                                float* out, int B, int N, int S, int K, int D, int xyz_first, void* stream) {
  if (!xyz || !new_xyz || !idx || !out || B < 0 || N < 1 || S < 1 || K < 1 || D < 0) return PPT_EINVAL;
  if (D > 0 && !points) return PPT_EINVAL;
  if (B == 0) return 0;
  if (D >= 32 && D + 3 <= 1024) {
    const long long rows = (long long)B * S * K;
    size_t smem;
    const int grid = gather_rows_launch_dims(rows, D + 3, &smem);
    static PptOncePerDevice configured;
    if (configured.need())
      PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(gather_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              GR_ROWS * 1024 * 4));
    gather_rows_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(xyz, new_xyz, points, idx, out, rows, S * K, N, K,
                                                                        D, xyz_first);
    return ppt_launch_status();
  }
  if (B > 65535) return PPT_ERANGE;
  const size_t quads = ((size_t)S * K * (3 + D) + 3) / 4;
  dim3 grid((unsigned)((quads + 255) / 256), B);
  group_concat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, new_xyz, D ? points : nullptr, idx, out, N, S, K, D,
                                                             xyz_first);
  return ppt_launch_status();
}
