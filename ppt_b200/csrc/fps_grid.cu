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
// One CTA per cloud, 16 warps.  128 clouds occupy 128 SMs for 511 dependent iterations, so the kernel's time IS the
// latency of one iteration (skip test -> row passes of the busiest warp -> warp argmax -> barrier -> block argmax ->
// centre), about half of it dependent-instruction latency and half issue contention among the 16 warps that all run
// the same phase at the same time.  What that led to (each step measured on B200 at 128 x 8192 -> 512):
//  * the sorted cloud, running min-distances and original indices live in shared memory (one point per lane per
//    row: conflict-free); lane j of a warp keeps the box, max and argmax POSITION of one of the warp's rows, so the
//    skip test of all rows of a warp is one pass and a ballot;
//  * rows are dealt to the warps in rotated groups (row_of), touched rows are visited 1, 2 or 4 per pass depending
//    on how many the warp has (independent load -> distance -> reduce chains overlap; padding a pass with repeated
//    rows costs issue slots that other warps need);
//  * every argmax level is a max-reduction of the value bits, a second max-reduction of the position over the
//    lanes that hold the maximum, and a ballot that only feeds the tie test; the original-index tie-break
//    (torch.max keeps the first index, F4) runs only on an actual tie;
//  * min-distances only decrease, so a warp recomputes its best row only when that very row was touched and
//    otherwise re-publishes last iteration's (max, position);
//  * long-latency bit scans are avoided (one FLO per visited row; __ffs would be BREV + FLO.SH) and nothing is
//    stored to global memory inside the loop: the winners are remembered as 16-bit sorted positions in shared
//    memory and written out by the whole block at the end.
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

// index of the highest set bit of x != 0 (the row masks may be walked in any order).  Measured: this (FLO + 2 ALU)
// 282 us, lowest-bit-first (__ffs: BREV + FLO.SH, two dependent long-latency ops per use) 340 us; a bfind.u32 in
// inline asm makes ptxas guard every *_sync of the loop with WARPSYNC.
#define FPS_BIT(x) (31 - __clz((int)(x)))

template <int FG_WARPS>
__global__ void __launch_bounds__(FG_WARPS * 32, 1)
fps_grid_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                const unsigned char* __restrict__ index, int64_t* __restrict__ idx_out,
                float* __restrict__ centers_out, int N, int G) {
  constexpr int FG_THREADS = FG_WARPS * 32;
  constexpr int FG_RPW = spidx::MAX_N / 32 / FG_WARPS;  // rows per warp at most
  static_assert(spidx::MIN_N / 32 >= FG_WARPS, "every warp must own a row through lane 0 (wlane starts at 0)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = (N + 31) & ~31, rows = np / 32;
  int2* slot = reinterpret_cast<int2*>(smem_raw);           // [2][32] (value bits, sorted position); unused: (-1, 0)
  float4* spts = reinterpret_cast<float4*>(slot + 64);      // [np] sorted {x,y,z,|p|^2}
  float* smind = reinterpret_cast<float*>(spts + np);       // [np] running min-distance (-1: padding)
  unsigned* soid = reinterpret_cast<unsigned*>(smind + np); // [np] original index
  unsigned short* sel = reinterpret_cast<unsigned short*>(soid + np);  // [G] sorted position of sample g >= 1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  const spidx::Layout L(N);
  const unsigned char* rec = index + (size_t)b * L.total;
  const float4* pts = reinterpret_cast<const float4*>(rec + L.pts);
  const int* sidx = reinterpret_cast<const int*>(rec + L.idx);
  const spidx::RowBox* boxes = reinterpret_cast<const spidx::RowBox*>(rec + L.boxes);

  if (tid < 64) slot[tid] = make_int2(__float_as_int(-1.0f), 0);
  for (int i = tid; i < np; i += FG_THREADS) {
    spts[i] = __ldg(pts + i);
    const int o = __ldg(sidx + i);
    soid[i] = (unsigned)o;
    smind[i] = o != 0x7fffffff ? 1e10f : -1.0f;
  }

  // lane j: state of row row_of(j).  Rows are dealt to the warps in groups of FG_WARPS consecutive rows, each group
  // rotated by its own number: a run of consecutive rows (neighbours along the Hilbert curve) still lands on distinct warps, and so do rows that
  // are a power of two apart (where the curve comes back next to itself), which "row % FG_WARPS" would
  // all hand to the same warp.
  auto row_of = [&](int j) { return FG_WARPS * j + ((warp - j) & (FG_WARPS - 1)); };
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

  // shared-window addresses, taken once (common.cuh)
  // (through a shuffle: otherwise ptxas treats the base as a constant and re-derives it at every use all the same)
  const uint32_t a_slot = __shfl_sync(PPT_FULL_MASK, ppt_smem_addr(smem_raw), 0);
  const uint32_t a_pts = a_slot + 64 * 8, a_mind = a_pts + np * 16, a_oid = a_mind + np * 4,
                 a_sel = a_oid + np * 4;

  const int64_t s0 = start[b];  // the reference indexes xyz[start] (raises when out of range): clamp instead
  unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
  float cx = cloud[far * 3 + 0], cy = cloud[far * 3 + 1], cz = cloud[far * 3 + 2];
  int64_t* out = idx_out + (size_t)b * G;
  float* cout = centers_out ? centers_out + (size_t)b * G * 3 : nullptr;
  const int neg1 = __float_as_int(-1.0f);
  int wmax = neg1, wpos = 0;  // this warp's best row: (value bits, sorted position)
  int wlane = 0;              // ... and the lane that owns it (lane 0 owns a row in every warp: rows >= FG_WARPS)
  int wr_slot = warp, rd_slot = lane;  // double-buffered by iteration parity (one barrier each)

#ifdef FPS_TRACE
  long long t_prev = clock64(), t_iter = t_prev, t_acc[5] = {0, 0, 0, 0, 0};
#endif
  // Sample 0 is the start point; samples 1 .. G-1 are remembered as sorted positions in shared memory and written out
  // by the whole block after the loop (the sorted copy holds the cloud's own bits), so that an iteration carries no
  // global store and no thread-0 branch.
  if (tid == 0) {
    out[0] = (int64_t)far;
    if (cout) { cout[0] = cx; cout[1] = cy; cout[2] = cz; }
  }
  for (int g = 0; g < G - 1; ++g) {
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
    // Running min-distances only decrease, so the warp's best row changes only when that very row is touched.
    const bool redo_best = (mask >> wlane) & 1u;  // warp-uniform
    auto pass = [&](auto rc) {
      constexpr int R = decltype(rc)::value;
      int j[R], pos[R], vb[R], w[R], best[R];
      bool on[R];
      unsigned eq[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        on[r] = mask != 0;  // r == 0: always; a slot past the last set bit repeats row j[0] and stores nothing
        j[r] = on[r] ? FPS_BIT(mask) : j[0];
        if (on[r]) mask ^= 1u << j[r];
        pos[r] = row_of(j[r]) * 32 + lane;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = ppt_lds128(a_pts + pos[r] * 16);
        const float m = fminf(ppt_lds_f32(a_mind + pos[r] * 4), ppt_fps_dist(p.x, p.y, p.z, cx, cy, cz));  // torch.min
        if (on[r]) ppt_sts_f32(a_mind + pos[r] * 4, m);                                  // padding stays -1
        vb[r] = __float_as_int(m);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = __reduce_max_sync(PPT_FULL_MASK, vb[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) eq[r] = __ballot_sync(PPT_FULL_MASK, vb[r] == w[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) best[r] = __reduce_max_sync(PPT_FULL_MASK, vb[r] == w[r] ? pos[r] : 0);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (eq[r] & (eq[r] - 1)) {  // several lanes hold the maximum: the smallest original index wins
          const unsigned cand = vb[r] == w[r] ? ppt_lds_u32(a_oid + pos[r] * 4) : 0xffffffffu;
          const unsigned mn = __reduce_min_sync(PPT_FULL_MASK, cand);
          best[r] = __reduce_max_sync(PPT_FULL_MASK, cand == mn ? pos[r] : 0);
        }
        const bool me = on[r] && lane == j[r];
        rpos = me ? best[r] : rpos;
        rmax = me ? __int_as_float(w[r]) : rmax;
      }
    };
    while (mask) {  // warp-uniform
      const unsigned m2 = mask & (mask - 1);  // at least two / three rows left?  (plain ALU: no POPC)
      if (m2 & (m2 - 1)) pass(std::integral_constant<int, 4>{});
      else if (m2) pass(std::integral_constant<int, 2>{});
      else pass(std::integral_constant<int, 1>{});
    }
    FPS_T(1);

    // best row of this warp (only when it may have changed), then of the block (one barrier per iteration,
    // double-buffered slots)
    if (redo_best) {
      const int vb = owner ? __float_as_int(rmax) : neg1;
      wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
      const bool hit = owner && vb == wmax;
      const unsigned eq = __ballot_sync(PPT_FULL_MASK, hit);
      wpos = __reduce_max_sync(PPT_FULL_MASK, hit ? rpos : 0);
      if (eq & (eq - 1)) {
        const unsigned cand = hit ? ppt_lds_u32(a_oid + rpos * 4) : 0xffffffffu;
        const unsigned mn = __reduce_min_sync(PPT_FULL_MASK, cand);
        wpos = __reduce_max_sync(PPT_FULL_MASK, cand == mn ? rpos : 0);
      }
      wlane = (wpos >> 5) / FG_WARPS;  // the lane that owns that row (row_of)
    }
    if (lane == 0) ppt_sts64(a_slot + wr_slot * 8, wmax, wpos);
    FPS_T(2);
    __syncthreads();
    FPS_T(3);
    const int2 s = ppt_lds64(a_slot + rd_slot * 8);
    wr_slot ^= 32;  // the other parity
    rd_slot ^= 32;
    const int cmax = __reduce_max_sync(PPT_FULL_MASK, s.x);
    unsigned ceq = __ballot_sync(PPT_FULL_MASK, s.x == cmax);
    int cpos = __reduce_max_sync(PPT_FULL_MASK, s.x == cmax ? s.y : 0);
    if (ceq & (ceq - 1)) {
      const unsigned cand = s.x == cmax ? ppt_lds_u32(a_oid + s.y * 4) : 0xffffffffu;
      const unsigned mn = __reduce_min_sync(PPT_FULL_MASK, cand);
      cpos = __reduce_max_sync(PPT_FULL_MASK, cand == mn ? s.y : 0);
    }
    if (tid == 0) ppt_sts_u16(a_sel + (g + 1) * 2, (unsigned short)cpos);
    const float4 c = ppt_lds128(a_pts + cpos * 16);
    cx = c.x; cy = c.y; cz = c.z;
    FPS_T(4);
  }
#ifdef FPS_TRACE
  if (b == 0 && tid == 0)
    for (int k = 0; k < 5; ++k) fps_trace_phase[k] = (unsigned long long)t_acc[k];
#endif
  __syncthreads();
  for (int g = 1 + tid; g < G; g += FG_THREADS) {
    const int p = sel[g];
    out[g] = (int64_t)soid[p];
    if (cout) {
      const float4 c = spts[p];
      cout[g * 3 + 0] = c.x; cout[g * 3 + 1] = c.y; cout[g * 3 + 2] = c.z;
    }
  }
}

template <int W>
int launch_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                    int B, int N, int G, cudaStream_t st) {
  static PptOncePerDevice configured;
  const size_t slots = 2 * 32 * sizeof(int2);
  const size_t sel = ((size_t)G * sizeof(unsigned short) + 15) & ~(size_t)15;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(fps_grid_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)((size_t)spidx::MAX_N * 26 + slots + 16)));
  }
  const size_t smem = (size_t)((N + 31) & ~31) * 24 + slots + sel;  // G <= N
  fps_grid_kernel<W><<<B, W * 32, smem, st>>>(xyz, start, static_cast<const unsigned char*>(index), idx_out,
                                             centers_out, N, G);
  return ppt_launch_status();
}

}  // namespace

int ppt_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                 int B, int N, int G, cudaStream_t st) {
  // 16 warps: what measured best on B200 (8 and 32 warps: 5-15 % slower, before and after the round-2 rework)
#ifndef FPS_GRID_WARPS
#define FPS_GRID_WARPS 16
#endif
  return launch_fps_grid<FPS_GRID_WARPS>(xyz, start, index, idx_out, centers_out, B, N, G, st);
}
