// Exact bucketed farthest point sampling on the spatial index (512 <= N <= 8192), sm_100a.
//
// Same result as fps.cu (and therefore as the reference), but an iteration only
// recomputes the rows of 32 points that the new centre can still affect: for a row with bounding box
// [lo, hi] and centre c, the per-axis gaps g = max(fl(lo - c), fl(c - hi), 0) satisfy |fl(x - c)| >= g
// for every point of the row (rounding is monotone), hence
//     lb = fl(fl(fl(gx*gx) + fl(gy*gy)) + fl(gz*gz))  <=  the reference distance of every point,
// evaluated with the very same rounded operations.  If lb >= the row's largest running min-distance,
// min(mind, d) leaves every point of the row unchanged and the row (and its cached argmax) is skipped.
//
// One CTA per cloud, 32 warps; warp w owns rows w, w+32, ... (a centre's neighbourhood is a run of
// consecutive Morton rows, so interleaving spreads the affected rows over the warps).  Lane j of a warp
// also keeps row (w + 32 j)'s box, max and argmax, so the skip test of all rows of a warp is ONE pass
// over 8 lanes.  Ties: (max value, smallest ORIGINAL index), as torch.max.
#include "common.cuh"
#include "spatial_index.cuh"

namespace {

constexpr int FG_THREADS = 1024;
constexpr int FG_WARPS = 32;

template <int RPW>  // rows per warp: rows <= 32 * RPW
__global__ void __launch_bounds__(FG_THREADS, 1)
fps_grid_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                const unsigned char* __restrict__ index, int64_t* __restrict__ idx_out,
                float* __restrict__ centers_out, int N, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);   // unsorted cloud, SoA: centroid lookup by original index
  float* sy = sx + N;
  float* sz = sy + N;
  int2* slot = reinterpret_cast<int2*>(sz + N);     // [2][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  const spidx::Layout L(N);
  const unsigned char* rec = index + (size_t)b * L.total;
  const float4* pts = reinterpret_cast<const float4*>(rec + L.pts);
  const int* sidx = reinterpret_cast<const int*>(rec + L.idx);
  const spidx::RowBox* boxes = reinterpret_cast<const spidx::RowBox*>(rec + L.boxes);
  const int rows = (N + 31) / 32;

  for (int i = tid; i < 3 * N; i += FG_THREADS) {
    const float v = cloud[i];
    const int n = i / 3, c = i - n * 3;
    (c == 0 ? sx : c == 1 ? sy : sz)[n] = v;
  }

  // this lane's point of each owned row
  float px[RPW], py[RPW], pz[RPW], mind[RPW];
  unsigned oid[RPW];
#pragma unroll
  for (int j = 0; j < RPW; ++j) {
    const int row = warp + FG_WARPS * j;
    px[j] = py[j] = pz[j] = 0.f;
    mind[j] = -1.0f;
    oid[j] = 0xffffffffu;
    if (row < rows) {
      const float4 p = __ldg(pts + row * 32 + lane);
      const int o = __ldg(sidx + row * 32 + lane);
      px[j] = p.x; py[j] = p.y; pz[j] = p.z;
      if (o != 0x7fffffff) { mind[j] = 1e10f; oid[j] = (unsigned)o; }
    }
  }
  // lane j < RPW: state of row (warp + 32 j)
  float blo0 = 0.f, blo1 = 0.f, blo2 = 0.f, bhi0 = 0.f, bhi1 = 0.f, bhi2 = 0.f;
  float rmax = -1.0f;         // largest running min-distance in the row (-1: no real point)
  unsigned rarg = 0xffffffffu;
  {
    const int myrow = warp + FG_WARPS * lane;
    if (lane < RPW && myrow < rows) {
      const spidx::RowBox bx = boxes[myrow];
      blo0 = bx.lo[0]; blo1 = bx.lo[1]; blo2 = bx.lo[2];
      bhi0 = bx.hi[0]; bhi1 = bx.hi[1]; bhi2 = bx.hi[2];
      rmax = 1e10f;  // forces the first iteration to visit the row
    }
  }
  __syncthreads();

  unsigned far = (unsigned)start[b];
  float cx = sx[far], cy = sy[far], cz = sz[far];
  int64_t* out = idx_out + (size_t)b * G;
  float* cout = centers_out ? centers_out + (size_t)b * G * 3 : nullptr;
  const int neg1 = __float_as_int(-1.0f);

  for (int g = 0; g < G; ++g) {
    if (tid == 0) {
      out[g] = (int64_t)far;
      if (cout) { cout[g * 3 + 0] = cx; cout[g * 3 + 1] = cy; cout[g * 3 + 2] = cz; }
    }
    if (g == G - 1) break;

    // which of this warp's rows can the new centre still affect?
    const float gx = fmaxf(fmaxf(__fsub_rn(blo0, cx), __fsub_rn(cx, bhi0)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(blo1, cy), __fsub_rn(cy, bhi1)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(blo2, cz), __fsub_rn(cz, bhi2)), 0.f);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    const unsigned mask = __ballot_sync(PPT_FULL_MASK, lane < RPW && lb < rmax);

#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      if (mask & (1u << j)) {  // warp-uniform
        const float d = ppt_fps_dist(px[j], py[j], pz[j], cx, cy, cz);
        const float m = fminf(mind[j], d);  // torch.min(distance, dist)
        mind[j] = m;
        const int vb = __float_as_int(m);
        const int wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
        const unsigned widx = __reduce_min_sync(PPT_FULL_MASK, vb == wmax ? oid[j] : 0xffffffffu);
        if (lane == j) { rmax = __int_as_float(wmax); rarg = widx; }
      }
    }

    // best row of this warp, then of the block (one barrier per iteration, double-buffered slots)
    const int vb = lane < RPW ? __float_as_int(rmax) : neg1;
    const int wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
    const unsigned widx = __reduce_min_sync(PPT_FULL_MASK, (lane < RPW && vb == wmax) ? rarg : 0xffffffffu);
    const int par = g & 1;
    if (lane == 0) slot[par * 32 + warp] = make_int2(wmax, (int)widx);
    __syncthreads();
    const int2 s = slot[par * 32 + lane];
    const int cmax = __reduce_max_sync(PPT_FULL_MASK, s.x);
    far = __reduce_min_sync(PPT_FULL_MASK, s.x == cmax ? (unsigned)s.y : 0xffffffffu);
    cx = sx[far]; cy = sy[far]; cz = sz[far];
  }
}

template <int RPW>
int launch(const float* xyz, const int64_t* start, const unsigned char* index, int64_t* idx_out, float* centers_out,
           int B, int N, int G, cudaStream_t st) {
  auto kern = fps_grid_kernel<RPW>;
  static bool configured = false;
  if (!configured) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(spidx::MAX_N * 12 + 64 * sizeof(int2))));
    configured = true;
  }
  const size_t smem = (size_t)N * 12 + 64 * sizeof(int2);
  kern<<<B, FG_THREADS, smem, st>>>(xyz, start, index, idx_out, centers_out, N, G);
  return ppt_launch_status();
}

}  // namespace

int ppt_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                 int B, int N, int G, cudaStream_t st) {
  const unsigned char* ix = static_cast<const unsigned char*>(index);
  const int rows = (N + 31) / 32;
  if (rows <= 64) return launch<2>(xyz, start, ix, idx_out, centers_out, B, N, G, st);
  if (rows <= 128) return launch<4>(xyz, start, ix, idx_out, centers_out, B, N, G, st);
  return launch<8>(xyz, start, ix, idx_out, centers_out, B, N, G, st);
}
