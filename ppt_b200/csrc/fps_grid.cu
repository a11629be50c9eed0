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
// One CTA per cloud, 32 warps.  The kernel is bound by the dependent-latency chain of one iteration
// (skip test -> row updates -> warp argmax -> barrier -> block argmax -> centroid), not by work, so:
//  * the sorted cloud, running min-distances and original indices live in shared memory (one point
//    per lane per row: conflict-free) and the rows a centre touches are visited with a loop over the
//    set bits of a ballot; warp w owns rows w, w+32, ... (a centre's neighbourhood is a run of
//    consecutive Morton rows, so interleaving leaves at most one or two per warp);
//  * lane j keeps row (w + 32 j)'s box, max and argmax POSITION, so the skip test of all rows of a
//    warp is one pass and every argmax level is ONE max-reduction plus a ballot; the original-index
//    tie-break (torch.max keeps the first index, F4) runs only when the ballot shows an actual tie.
#include "common.cuh"
#include "spatial_index.cuh"

namespace {

template <int FG_WARPS>
__global__ void __launch_bounds__(FG_WARPS * 32, 1)
fps_grid_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                const unsigned char* __restrict__ index, int64_t* __restrict__ idx_out,
                float* __restrict__ centers_out, int N, int G) {
  constexpr int FG_THREADS = FG_WARPS * 32;
  constexpr int FG_RPW = spidx::MAX_N / 32 / FG_WARPS;  // rows per warp at most
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = (N + 31) & ~31, rows = np / 32;
  float4* spts = reinterpret_cast<float4*>(smem_raw);       // [np] sorted {x,y,z,|p|^2}
  float* smind = reinterpret_cast<float*>(spts + np);       // [np] running min-distance (-1: padding)
  unsigned* soid = reinterpret_cast<unsigned*>(smind + np); // [np] original index
  int2* slot = reinterpret_cast<int2*>(soid + np);          // [2][FG_WARPS] (value bits, sorted position)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  const spidx::Layout L(N);
  const unsigned char* rec = index + (size_t)b * L.total;
  const float4* pts = reinterpret_cast<const float4*>(rec + L.pts);
  const int* sidx = reinterpret_cast<const int*>(rec + L.idx);
  const spidx::RowBox* boxes = reinterpret_cast<const spidx::RowBox*>(rec + L.boxes);

  for (int i = tid; i < np; i += FG_THREADS) {
    spts[i] = __ldg(pts + i);
    const int o = __ldg(sidx + i);
    soid[i] = (unsigned)o;
    smind[i] = o != 0x7fffffff ? 1e10f : -1.0f;
  }

  // lane j: state of row (warp + FG_WARPS * j)
  const int myrow = warp + FG_WARPS * lane;
  const bool owner = lane < FG_RPW && myrow < rows;
  float blo0 = 0.f, blo1 = 0.f, blo2 = 0.f, bhi0 = 0.f, bhi1 = 0.f, bhi2 = 0.f;
  float rmax = -1.0f;          // largest running min-distance in the row (-1: nothing to pick)
  int rpos = 0;                // sorted position of that point (smallest original index among equals)
  if (owner) {
    const spidx::RowBox bx = boxes[myrow];
    blo0 = bx.lo[0]; blo1 = bx.lo[1]; blo2 = bx.lo[2];
    bhi0 = bx.hi[0]; bhi1 = bx.hi[1]; bhi2 = bx.hi[2];
    rmax = 1e10f;  // forces the first iteration to visit the row
  }
  __syncthreads();

  const int64_t s0 = start[b];  // the reference indexes xyz[start] (raises when out of range): clamp instead
  unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
  float cx = cloud[far * 3 + 0], cy = cloud[far * 3 + 1], cz = cloud[far * 3 + 2];
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
    unsigned mask = __ballot_sync(PPT_FULL_MASK, owner && lb < rmax);

    // Two rows per pass: their load -> distance -> reduce chains are independent and overlap.
    while (mask) {  // warp-uniform
      const int j0 = __ffs(mask) - 1;
      mask &= mask - 1;
      const bool two = mask != 0;
      const int j1 = two ? __ffs(mask) - 1 : j0;
      mask &= mask - 1;  // no-op when it is already 0
      const int pos0 = (warp + FG_WARPS * j0) * 32 + lane, pos1 = (warp + FG_WARPS * j1) * 32 + lane;
      const float4 p0 = spts[pos0], p1 = spts[pos1];
      const float m0 = fminf(smind[pos0], ppt_fps_dist(p0.x, p0.y, p0.z, cx, cy, cz));  // torch.min(distance, dist)
      const float m1 = fminf(smind[pos1], ppt_fps_dist(p1.x, p1.y, p1.z, cx, cy, cz));  // padding stays -1
      smind[pos0] = m0;
      if (two) smind[pos1] = m1;
      const int vb0 = __float_as_int(m0), vb1 = __float_as_int(m1);
      const int w0 = __reduce_max_sync(PPT_FULL_MASK, vb0), w1 = __reduce_max_sync(PPT_FULL_MASK, vb1);
      unsigned eq0 = __ballot_sync(PPT_FULL_MASK, vb0 == w0), eq1 = __ballot_sync(PPT_FULL_MASK, vb1 == w1);
      if (eq0 & (eq0 - 1)) {  // several lanes hold the maximum: the smallest original index wins
        const unsigned cand = vb0 == w0 ? soid[pos0] : 0xffffffffu;
        eq0 = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
      }
      if (eq1 & (eq1 - 1)) {
        const unsigned cand = vb1 == w1 ? soid[pos1] : 0xffffffffu;
        eq1 = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
      }
      if (lane == j0) { rmax = __int_as_float(w0); rpos = pos0 - lane + __ffs(eq0) - 1; }
      if (two && lane == j1) { rmax = __int_as_float(w1); rpos = pos1 - lane + __ffs(eq1) - 1; }
    }

    // best row of this warp, then of the block (one barrier per iteration, double-buffered slots)
    const int vb = owner ? __float_as_int(rmax) : neg1;
    const int wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
    unsigned eq = __ballot_sync(PPT_FULL_MASK, owner && vb == wmax);
    if (eq & (eq - 1)) {
      const unsigned cand = (owner && vb == wmax) ? soid[rpos] : 0xffffffffu;
      eq = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
    }
    const int wpos = __shfl_sync(PPT_FULL_MASK, rpos, eq ? __ffs(eq) - 1 : 0);
    const int par = g & 1;
    if (lane == 0) slot[par * FG_WARPS + warp] = make_int2(wmax, wpos);
    __syncthreads();
    const int2 s = lane < FG_WARPS ? slot[par * FG_WARPS + lane] : make_int2(neg1, 0);
    const int cmax = __reduce_max_sync(PPT_FULL_MASK, s.x);
    unsigned ceq = __ballot_sync(PPT_FULL_MASK, s.x == cmax);
    if (ceq & (ceq - 1)) {
      const unsigned cand = s.x == cmax ? soid[s.y] : 0xffffffffu;
      ceq = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
    }
    const int cpos = __shfl_sync(PPT_FULL_MASK, s.y, __ffs(ceq) - 1);
    far = soid[cpos];
    const float4 c = spts[cpos];
    cx = c.x; cy = c.y; cz = c.z;
  }
}

template <int W>
int launch_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                    int B, int N, int G, cudaStream_t st) {
  static PptOncePerDevice configured;
  const size_t slots = 2 * W * sizeof(int2);
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(fps_grid_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)((size_t)spidx::MAX_N * 24 + slots)));
  }
  const size_t smem = (size_t)((N + 31) & ~31) * 24 + slots;
  fps_grid_kernel<W><<<B, W * 32, smem, st>>>(xyz, start, static_cast<const unsigned char*>(index), idx_out,
                                             centers_out, N, G);
  return ppt_launch_status();
}

}  // namespace

int ppt_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                 int B, int N, int G, cudaStream_t st) {
  // 16 warps: what measured best on B200 (8 and 32 were within 5 % and slower)
  return launch_fps_grid<16>(xyz, start, index, idx_out, centers_out, B, N, G, st);
}
