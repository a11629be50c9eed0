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
// load per element and reached 18 % of the HBM peak at C = 131.  Here a WARP owns GR_ROWS = 4 consecutive output rows:
// 4 * C * 4 bytes is a multiple of 16 for every C, so the warp's slice of the output starts 16-byte aligned whatever
// the row length.  Source rows are read with 16-byte loads; a row's feature span starts s = (r C + offset) mod 4
// floats into an aligned 16-byte slot, so each aligned output slot is assembled from the lane's own float4 and its
// left neighbour's (one shuffle, s is warp-uniform) and stored with one aligned 16-byte streaming store; only the
// first and last slot of a span (shared with the coordinates / the neighbouring row) are written with scalar stores.
// No shared memory, no barrier.  Row bookkeeping (cloud, centre row, validity) is done by lanes 0-3 for their row and
// shuffled; the position inside the cloud advances incrementally (one 64-bit division per warp in the whole kernel).
// History at C = 131 (fraction of the HBM peak): element-wise 18 %, warp-per-row with 4-byte stores 27 % (every store
// straddles sectors), staged through shared memory 31-44 % (bank conflicts of the 4-byte staging stores, barriers),
// per-lane 64-bit divides made the early versions issue-bound (46 instructions per float).
constexpr int GR_ROWS = 4;
constexpr int GR_WARPS = 8;

__device__ __forceinline__ float4 shfl_up4(float4 v) {
  return make_float4(__shfl_up_sync(PPT_FULL_MASK, v.x, 1), __shfl_up_sync(PPT_FULL_MASK, v.y, 1),
                     __shfl_up_sync(PPT_FULL_MASK, v.z, 1), __shfl_up_sync(PPT_FULL_MASK, v.w, 1));
}
__device__ __forceinline__ float4 shfl_from4(float4 v, int src) {
  return make_float4(__shfl_sync(PPT_FULL_MASK, v.x, src), __shfl_sync(PPT_FULL_MASK, v.y, src),
                     __shfl_sync(PPT_FULL_MASK, v.z, src), __shfl_sync(PPT_FULL_MASK, v.w, src));
}

template <bool CONCAT>
__global__ void __launch_bounds__(GR_WARPS * 32)
gather_rows_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const float* __restrict__ points,
                   const int64_t* __restrict__ idx, float* __restrict__ out, long long total_rows, int rows_per_cloud,
                   int N, int K, int D, int xyz_first) {
  // CONCAT: out row = [xyz - centre | feats] (or feats first), C = 3 + D; else: out row = points row, C = D
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // shuffle: provably warp-uniform
  const int C = CONCAT ? D + 3 : D;
  const int xyz_lo = xyz_first ? 0 : D, feat_lo = (CONCAT && xyz_first) ? 3 : 0;
  const float nanv = __int_as_float(0x7fc00000);
  const bool vec = (D & 3) == 0;
  const long long nwarps = (long long)gridDim.x * GR_WARPS;
  const long long step = nwarps * GR_ROWS;
  const long long step_b = step / rows_per_cloud;
  const unsigned step_in = (unsigned)(step - step_b * rows_per_cloud);
  long long row0 = ((long long)blockIdx.x * GR_WARPS + warp) * GR_ROWS;
  long long b0 = row0 / rows_per_cloud;
  unsigned in0 = (unsigned)(row0 - b0 * rows_per_cloud);
  for (; row0 < total_rows; row0 += step) {
    const int nrows = total_rows - row0 < GR_ROWS ? (int)(total_rows - row0) : GR_ROWS;
    // lanes 8r .. 8r+7 locate row r (source row, validity); lanes 8r .. 8r+2 also write its three centred coordinates
    const int rl = lane >> 3, jl = lane & 7;
    float* dst = out + (size_t)row0 * C;  // 16-byte aligned: row0 is a multiple of 4
    long long src_l = 0;
    int ok_l = 0;
    if (rl < nrows) {
      const unsigned in = in0 + (unsigned)rl;
      const unsigned over = in >= (unsigned)rows_per_cloud ? in / (unsigned)rows_per_cloud : 0u;
      const long long bl = b0 + over;
      const long long nl = __ldg(idx + row0 + rl);
      ok_l = (unsigned long long)nl < (unsigned long long)N;  // out-of-range index: NaN row, no wild read
      src_l = bl * N + (ok_l ? nl : 0);
      if (CONCAT && jl < 3) {
        const long long ctr = bl * (rows_per_cloud / K) + (in - over * (unsigned)rows_per_cloud) / (unsigned)K;
        const float pv = __ldg(xyz + (size_t)src_l * 3 + jl);
        const float cv = __ldg(new_xyz + (size_t)ctr * 3 + jl);
        dst[rl * C + xyz_lo + jl] = ok_l ? __fsub_rn(pv, cv) : nanv;
      }
    }
#pragma unroll
    for (int r = 0; r < GR_ROWS; ++r) {
      if (r >= nrows) break;  // warp-uniform
      const long long src = __shfl_sync(PPT_FULL_MASK, src_l, r * 8);
      const bool ok = __shfl_sync(PPT_FULL_MASK, ok_l, r * 8) != 0;
      const float* frow = points + (size_t)src * D;
      const int base = r * C + feat_lo;  // float offset of the feature span inside the slice
      if (!vec) {
        for (int c = lane; c < D; c += 32) dst[base + c] = ok ? __ldg(frow + c) : nanv;
        continue;
      }
      const int sft = base & 3, nf = D >> 2;   // warp-uniform
      float4* slot = reinterpret_cast<float4*>(dst) + (base >> 2);  // slot[f] holds span elements 4f - sft .. 4f - sft + 3
      float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int f0 = 0; f0 < nf; f0 += 32) {
        const int f = f0 + lane;
        float4 v = make_float4(nanv, nanv, nanv, nanv);
        if (f < nf && ok) v = __ldg(reinterpret_cast<const float4*>(frow) + f);
        if (sft == 0) {
          if (f < nf) __stcs(slot + f, v);  // written once: streaming store
          continue;
        }
        float4 up = shfl_up4(v);
        if (nf > 32) {  // rows longer than one warp pass: lane 0 continues from lane 31 of the previous pass
          if (lane == 0) up = carry;
          if (f0 + 32 < nf) carry = shfl_from4(v, 31);
        }
        float4 o;
        if (sft == 1) o = make_float4(up.w, v.x, v.y, v.z);
        else if (sft == 2) o = make_float4(up.z, up.w, v.x, v.y);
        else o = make_float4(up.y, up.z, up.w, v.x);
        if (f >= 1 && f < nf) {
          __stcs(slot + f, o);
        } else if (f == 0) {  // first slot of the span: its leading sft floats belong to the coordinates / the row before
          float* p = reinterpret_cast<float*>(slot);
          p[sft] = v.x;
          if (sft <= 2) p[sft + 1] = v.y;
          if (sft <= 1) p[sft + 2] = v.z;
        }
        if (f == nf - 1) {  // the span's last sft floats spill into the next slot
          float* p = reinterpret_cast<float*>(slot + nf);
          if (sft == 1) p[0] = v.w;
          else if (sft == 2) { p[0] = v.z; p[1] = v.w; }
          else { p[0] = v.y; p[1] = v.z; p[2] = v.w; }
        }
      }
    }
    b0 += step_b;
    in0 += step_in;
    if (in0 >= (unsigned)rows_per_cloud) { in0 -= (unsigned)rows_per_cloud; ++b0; }
  }
}

}  // namespace

static int gather_rows_grid(long long total_rows) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (total_rows + GR_WARPS * GR_ROWS - 1) / (GR_WARPS * GR_ROWS);
  const long long cap = (long long)sms * 8;  // 8 resident blocks of 8 warps per SM
  return (int)(want < cap ? want : cap);
}

extern "C" PPT_EXPORT int ppt_gather(const float* points, const int64_t* idx, float* out, int B, int N, int C, int M,
                          void* stream) {
  if (!points || !idx || !out || B < 0 || N < 1 || C < 1 || M < 0) return PPT_EINVAL;
  if (B == 0 || M == 0) return 0;
  if (C >= 32) {
    const long long rows = (long long)B * M;
    gather_rows_kernel<false><<<gather_rows_grid(rows), GR_WARPS * 32, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, points, idx, out, rows, M, N, 1, C, 1);
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
  if (D >= 32) {
    const long long rows = (long long)B * S * K;
    gather_rows_kernel<true><<<gather_rows_grid(rows), GR_WARPS * 32, 0, (cudaStream_t)stream>>>(
        xyz, new_xyz, points, idx, out, rows, S * K, N, K, D, xyz_first);
    return ppt_launch_status();
  }
  if (B > 65535) return PPT_ERANGE;
  const size_t quads = ((size_t)S * K * (3 + D) + 3) / 4;
  dim3 grid((unsigned)((quads + 255) / 256), B);
  group_concat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, new_xyz, D ? points : nullptr, idx, out, N, S, K, D,
                                                             xyz_first);
  return ppt_launch_status();
}
