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
#include <type_traits>

#ifdef FPS_TRACE
// Measurement build only (tools/build_variant.py ... -DFPS_TRACE): warp 0 of cloud 0 accumulates the cycles of each
// phase of an iteration; never compiled into the shipped library.
__device__ unsigned long long fps_trace_phase[8];
__device__ unsigned fps_trace_rows[1024];   // rows touched per iteration, whole CTA
__device__ unsigned fps_trace_rowsmax[1024];  // ... most rows any one warp had
__device__ unsigned fps_trace_rows0[1024];  // ... by warp 0
__device__ unsigned fps_trace_iter[1024];   // cycles per iteration (warp 0)
#define FPS_T(k)                                                                         \
  do {                                                                                   \
    if (b == 0 && tid == 0) {                                                            \
      const long long now = clock64();                                                   \
      t_acc[k] += now - t_prev;                                                          \
      if (k == 4) { fps_trace_iter[g] = (unsigned)(now - t_iter); t_iter = now; }        \
      t_prev = now;                                                                      \
    }                                                                                    \
  } while (0)
extern "C" PPT_EXPORT int ppt_debug_fps_trace(void* dst, void* stream) {  // dst: 8 u64 + 4 x 1024 u32, device memory
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* d = static_cast<unsigned char*>(dst);
  cudaMemcpyFromSymbolAsync(d, fps_trace_phase, 64, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyFromSymbolAsync(d + 64, fps_trace_rows, 4096, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyFromSymbolAsync(d + 64 + 4096, fps_trace_rows0, 4096, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyFromSymbolAsync(d + 64 + 8192, fps_trace_iter, 4096, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyFromSymbolAsync(d + 64 + 12288, fps_trace_rowsmax, 4096, 0, cudaMemcpyDeviceToDevice, st);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbolAsync(fps_trace_phase, z, 64, 0, cudaMemcpyHostToDevice, st);
  static unsigned zr[1024];
  cudaMemcpyToSymbolAsync(fps_trace_rows, zr, 4096, 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(fps_trace_rowsmax, zr, 4096, 0, cudaMemcpyHostToDevice, st);
  return (int)cudaGetLastError();
}
#else
#define FPS_T(k) do { } while (0)
#endif

namespace {

#ifndef FPS_GRID_MICRO
#define FPS_GRID_MICRO 1
#endif
#ifndef FPS_GRID_HYBRID
#define FPS_GRID_HYBRID 1
#endif
// index of a set bit of x != 0: every mask this is applied to either holds a single bit or may be walked in any order
#if FPS_GRID_MICRO
#define FPS_BIT(x) (31 - __clz((int)(x)))   // one FLO
#else
#define FPS_BIT(x) (__ffs((int)(x)) - 1)    // BREV + FLO
#endif

template <int FG_WARPS, int RPP>
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

  // lane j: state of row row_of(j).  Rows are dealt to the warps in groups of FG_WARPS consecutive rows, each group
  // rotated by its own number: a run of consecutive Morton rows still lands on distinct warps, and so do rows that
  // are a power of two apart (the Morton neighbours across a y / z cell boundary), which "row % FG_WARPS" would
  // all hand to the same warp.
#ifndef FPS_GRID_ROT
#define FPS_GRID_ROT 1
#endif
  auto row_of = [&](int j) { return FG_WARPS * j + (FPS_GRID_ROT ? ((warp - j) & (FG_WARPS - 1)) : warp); };
  const int myrow = row_of(lane);
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
  int wmax = neg1, wpos = 0;  // this warp's best row: (value bits, sorted position)

#ifdef FPS_TRACE
  long long t_prev = clock64(), t_iter = t_prev, t_acc[5] = {0, 0, 0, 0, 0};
#endif
  for (int g = 0; g < G; ++g) {
    if (tid == 0) {
      out[g] = (int64_t)far;
      if (cout) { cout[g * 3 + 0] = cx; cout[g * 3 + 1] = cy; cout[g * 3 + 2] = cz; }
    }
    if (g == G - 1) {
#ifdef FPS_TRACE
      if (b == 0 && tid == 0)
        for (int k = 0; k < 5; ++k) fps_trace_phase[k] = (unsigned long long)t_acc[k];
#endif
      break;
    }

    // which of this warp's rows can the new centre still affect?
    const float gx = fmaxf(fmaxf(__fsub_rn(blo0, cx), __fsub_rn(cx, bhi0)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(blo1, cy), __fsub_rn(cy, bhi1)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(blo2, cz), __fsub_rn(cz, bhi2)), 0.f);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    unsigned mask = __ballot_sync(PPT_FULL_MASK, owner && lb < rmax);
    FPS_T(0);
#ifdef FPS_TRACE
    if (lane == 0 && b == 0) atomicAdd(&fps_trace_rows[g], (unsigned)__popc(mask));
    if (lane == 0 && b == 0) atomicMax(&fps_trace_rowsmax[g], (unsigned)__popc(mask));
    if (b == 0 && warp == 0 && lane == 0) fps_trace_rows0[g] = (unsigned)__popc(mask);
#endif

    // R rows per pass: their load -> distance -> reduce chains are independent and overlap.
    const bool touched = mask != 0;  // warp-uniform
    auto pass = [&](auto rc) {
      constexpr int R = decltype(rc)::value;
      int j[R], pos[R], vb[R], w[R];
      bool on[R];
      unsigned eq[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        on[r] = mask != 0;  // r == 0: always; a slot past the last set bit repeats row j[0] and stores nothing
        j[r] = on[r] ? FPS_BIT(mask) : j[0];
        mask = on[r] ? mask ^ (1u << j[r]) : 0u;
        pos[r] = row_of(j[r]) * 32 + lane;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = spts[pos[r]];
        const float m = fminf(smind[pos[r]], ppt_fps_dist(p.x, p.y, p.z, cx, cy, cz));  // torch.min(distance, dist)
        if (on[r]) smind[pos[r]] = m;                                                    // padding stays -1
        vb[r] = __float_as_int(m);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = __reduce_max_sync(PPT_FULL_MASK, vb[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) eq[r] = __ballot_sync(PPT_FULL_MASK, vb[r] == w[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (eq[r] & (eq[r] - 1)) {  // several lanes hold the maximum: the smallest original index wins
          const unsigned cand = vb[r] == w[r] ? soid[pos[r]] : 0xffffffffu;
          eq[r] = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
        }
        const bool me = on[r] && lane == j[r];
        rmax = me ? __int_as_float(w[r]) : rmax;
        rpos = me ? pos[r] - lane + FPS_BIT(eq[r]) : rpos;
      }
    };
    while (mask) {  // warp-uniform
#if FPS_GRID_HYBRID
      const int nrow = __popc(mask);
      if (nrow >= 3) pass(std::integral_constant<int, 4>{});
      else if (nrow == 2) pass(std::integral_constant<int, 2>{});
      else pass(std::integral_constant<int, 1>{});
#else
      pass(std::integral_constant<int, RPP>{});
#endif
    }
    FPS_T(1);

    // best row of this warp, then of the block (one barrier per iteration, double-buffered slots)
    // (a warp none of whose rows was touched keeps last iteration's answer)
    if (!FPS_GRID_MICRO || touched) {
      const int vb = owner ? __float_as_int(rmax) : neg1;
      wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
      unsigned eq = __ballot_sync(PPT_FULL_MASK, owner && vb == wmax);
      if (eq & (eq - 1)) {
        const unsigned cand = (owner && vb == wmax) ? soid[rpos] : 0xffffffffu;
        eq = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
      }
      wpos = __shfl_sync(PPT_FULL_MASK, rpos, eq ? FPS_BIT(eq) : 0);
    }
    const int par = g & 1;
    if (lane == 0) slot[par * FG_WARPS + warp] = make_int2(wmax, wpos);
    FPS_T(2);
    __syncthreads();
    FPS_T(3);
    const int2 s = lane < FG_WARPS ? slot[par * FG_WARPS + lane] : make_int2(neg1, 0);
    const int cmax = __reduce_max_sync(PPT_FULL_MASK, s.x);
    unsigned ceq = __ballot_sync(PPT_FULL_MASK, s.x == cmax);
    if (ceq & (ceq - 1)) {
      const unsigned cand = s.x == cmax ? soid[s.y] : 0xffffffffu;
      ceq = __ballot_sync(PPT_FULL_MASK, cand == __reduce_min_sync(PPT_FULL_MASK, cand));
    }
    const int cpos = __shfl_sync(PPT_FULL_MASK, s.y, FPS_BIT(ceq));
    far = soid[cpos];
    const float4 c = spts[cpos];
    cx = c.x; cy = c.y; cz = c.z;
    FPS_T(4);
  }
}

template <int W, int RPP>
int launch_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                    int B, int N, int G, cudaStream_t st) {
  static PptOncePerDevice configured;
  const size_t slots = 2 * W * sizeof(int2);
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(fps_grid_kernel<W, RPP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)((size_t)spidx::MAX_N * 24 + slots)));
  }
  const size_t smem = (size_t)((N + 31) & ~31) * 24 + slots;
  fps_grid_kernel<W, RPP><<<B, W * 32, smem, st>>>(xyz, start, static_cast<const unsigned char*>(index), idx_out,
                                             centers_out, N, G);
  return ppt_launch_status();
}

}  // namespace

int ppt_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                 int B, int N, int G, cudaStream_t st) {
  // 16 warps: what measured best on B200 (8 and 32 were within 5 % and slower)
#ifndef FPS_GRID_WARPS
#define FPS_GRID_WARPS 16
#endif
#ifndef FPS_GRID_RPP
#define FPS_GRID_RPP 2
#endif
  return launch_fps_grid<FPS_GRID_WARPS, FPS_GRID_RPP>(xyz, start, index, idx_out, centers_out, B, N, G, st);
}
