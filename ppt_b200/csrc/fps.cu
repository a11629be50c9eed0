// Farthest point sampling for sm_100a: one cloud per CTA (or per thread-block
// cluster for N > 8192), coordinates and running min-distances in registers,
// one __syncthreads per iteration.
//
// Replaces farthest_point_sample (models/pointbert/misc.py:44-69 and its three
// copies).  Arithmetic: F1 (un-fused (dx^2+dy^2)+dz^2), running
// minimum starts at 1e10, argmax breaks ties on the first index (F4).
//
// Work split: thread t owns points t, t+T, t+2T, ... (T = THREADS), so within a
// thread ascending slot = ascending point index and a strict '>' keeps the first
// maximum; across threads the first index is recovered with a min-reduction over
// the lanes that hold the maximum.  All distances are >= +0, so their bit
// patterns order like signed ints and both reductions are single CREDUX
// instructions; padding slots carry -1.0f and never win.
#include <cooperative_groups.h>

#include "common.cuh"
#include "spatial_index.cuh"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// round(a*a) per lane.  ptxas 12.9 contracts an f32x2 multiply feeding an f32x2 add into FFMA2 even
// with explicit .rn and --fmad=false (checked in SASS), which would change the reference's rounding
// (SURVEY.md F1).  So only the subtractions and the squares are packed; the two additions are
// scalar add.rn.f32 (__fadd_rn), which ptxas never contracts.
__device__ __forceinline__ unsigned long long sqr_f32x2(unsigned long long a) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(a));
  return r;
}
__device__ __forceinline__ float max3_f32(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

struct FpsRecord {  // one per CTA of a cluster, exchanged through DSMEM
  int val;          // float bits of the CTA-local maximum
  unsigned idx;     // its global point index
  float x, y, z;    // its coordinates
};

template <int THREADS, int PPT, int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start, int64_t* __restrict__ idx_out,
           float* __restrict__ centers_out, int N, int G) {
  constexpr int NW = THREADS / 32;
  constexpr int LOCAL = THREADS * PPT;  // points held by this CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + LOCAL;
  float* sz = sy + LOCAL;
  int2* slot = reinterpret_cast<int2*>(sz + LOCAL);                 // [2][32] (val, idx) per warp
  FpsRecord* rec = reinterpret_cast<FpsRecord*>(slot + 64);         // [2][CLUSTER]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = CLUSTER > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int b = blockIdx.x / CLUSTER;
  const int base = rank * LOCAL;  // first global point index of this CTA
  const float* cloud = xyz + (size_t)b * N * 3;

  // Coalesced AoS read, SoA in shared memory (also serves the centroid lookup).
  const int lo = base * 3, hi = min(N, base + LOCAL) * 3;
  for (int i = lo + tid; i < hi; i += THREADS) {
    const float v = cloud[i];
    const int n = i / 3 - base, c = i - (i / 3) * 3;
    (c == 0 ? sx : c == 1 ? sy : sz)[n] = v;
  }
  __syncthreads();

  // Coordinates live in registers as packed pairs (slot j, slot j+1) for the f32x2 pipe:
  // sub/mul/add.rn.f32x2 round each lane like the scalar instruction, so the distance is still
  // exactly (dx*dx + dy*dy) + dz*dz -- two points per instruction.
  static_assert(PPT % 2 == 0, "points per thread must be even");
  unsigned long long px[PPT / 2], py[PPT / 2], pz[PPT / 2];
  float mind[PPT];
#pragma unroll
  for (int j = 0; j < PPT; j += 2) {
    const int l0 = j * THREADS + tid, l1 = (j + 1) * THREADS + tid;
    const bool ok0 = base + l0 < N, ok1 = base + l1 < N;
    px[j / 2] = pack_f32x2(ok0 ? sx[l0] : 0.f, ok1 ? sx[l1] : 0.f);
    py[j / 2] = pack_f32x2(ok0 ? sy[l0] : 0.f, ok1 ? sy[l1] : 0.f);
    pz[j / 2] = pack_f32x2(ok0 ? sz[l0] : 0.f, ok1 ? sz[l1] : 0.f);
    mind[j] = ok0 ? 1e10f : -1.0f;
    mind[j + 1] = ok1 ? 1e10f : -1.0f;
  }

  const int64_t s0 = start[b];  // the reference indexes xyz[start] (raises when out of range): clamp instead
  unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
  float cx = cloud[far * 3 + 0], cy = cloud[far * 3 + 1], cz = cloud[far * 3 + 2];
  int64_t* out = idx_out + (size_t)b * G;
  float* cout = centers_out ? centers_out + (size_t)b * G * 3 : nullptr;

  for (int g = 0; g < G; ++g) {
    if (tid == 0 && rank == 0) {
      out[g] = (int64_t)far;
      if (cout) { cout[g * 3 + 0] = cx; cout[g * 3 + 1] = cy; cout[g * 3 + 2] = cz; }
    }
    if (g == G - 1) break;

    const unsigned long long cx2 = pack_f32x2(cx, cx), cy2 = pack_f32x2(cy, cy), cz2 = pack_f32x2(cz, cz);
#pragma unroll
    for (int j = 0; j < PPT; j += 2) {
      const unsigned long long dx = sub_f32x2(px[j / 2], cx2), dy = sub_f32x2(py[j / 2], cy2),
                               dz = sub_f32x2(pz[j / 2], cz2);
      float x0, x1, y0, y1, z0, z1;
      unpack_f32x2(sqr_f32x2(dx), x0, x1);
      unpack_f32x2(sqr_f32x2(dy), y0, y1);
      unpack_f32x2(sqr_f32x2(dz), z0, z1);
      const float d0 = __fadd_rn(__fadd_rn(x0, y0), z0), d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
      mind[j] = fminf(mind[j], d0);        // torch.min(distance, dist) for finite input
      mind[j + 1] = fminf(mind[j + 1], d1);
    }
    // Thread-local maximum with 3-input max, then the first slot that holds it (ascending slot =
    // ascending point index inside a thread), instead of carrying an index through every compare.
    float best = mind[0];
#pragma unroll
    for (int j = 1; j + 1 < PPT; j += 2) best = max3_f32(best, mind[j], mind[j + 1]);
    if ((PPT & 1) == 0) best = fmaxf(best, mind[PPT - 1]);
    int bj = PPT - 1;
#pragma unroll
    for (int j = PPT - 2; j >= 0; --j) bj = mind[j] == best ? j : bj;
    const int vb = __float_as_int(best);
    const int wmax = __reduce_max_sync(PPT_FULL_MASK, vb);
    const unsigned cand = vb == wmax ? (unsigned)(base + bj * THREADS + tid) : 0xffffffffu;
    const unsigned widx = __reduce_min_sync(PPT_FULL_MASK, cand);
    const int par = g & 1;
    if (lane == 0) slot[par * 32 + warp] = make_int2(wmax, (int)widx);
    __syncthreads();
    const int2 s = lane < NW ? slot[par * 32 + lane] : make_int2(__float_as_int(-1.0f), -1);
    const int cmax = __reduce_max_sync(PPT_FULL_MASK, s.x);
    const unsigned cidx = __reduce_min_sync(PPT_FULL_MASK, s.x == cmax ? (unsigned)s.y : 0xffffffffu);

    if (CLUSTER == 1) {
      far = cidx;
      cx = sx[far]; cy = sy[far]; cz = sz[far];
    } else {
      cg::cluster_group cluster = cg::this_cluster();
      // One thread per destination CTA publishes this CTA's winner (value, index,
      // coordinates) into that CTA's record table, then the cluster meets once.
      if (tid < CLUSTER) {
        const int l = (int)cidx - base;
        FpsRecord r;
        r.val = cmax; r.idx = cidx; r.x = sx[l]; r.y = sy[l]; r.z = sz[l];
        FpsRecord* remote = cluster.map_shared_rank(rec, tid);
        remote[par * CLUSTER + rank] = r;
      }
      cluster.sync();
      int bv = __float_as_int(-1.0f);
      unsigned bi = 0xffffffffu;
      int br = 0;
#pragma unroll
      for (int r = 0; r < CLUSTER; ++r) {
        const int v = rec[par * CLUSTER + r].val;
        const unsigned i = rec[par * CLUSTER + r].idx;
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; br = r; }
      }
      far = bi;
      cx = rec[par * CLUSTER + br].x; cy = rec[par * CLUSTER + br].y; cz = rec[par * CLUSTER + br].z;
    }
  }
  if (CLUSTER > 1) cg::this_cluster().sync();  // no CTA may exit while peers can still write its smem
}

template <int THREADS, int PPT, int CLUSTER>
int launch_fps(const float* xyz, const int64_t* start, int64_t* idx_out, float* centers_out, int B, int N, int G,
               cudaStream_t st) {
  auto kern = fps_kernel<THREADS, PPT, CLUSTER>;
  const size_t smem = (size_t)THREADS * PPT * 12 + 64 * sizeof(int2) + 2 * CLUSTER * sizeof(FpsRecord);
  static PptOncePerDevice configured;  // per instantiation
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * CLUSTER);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CLUSTER > 1 ? 1 : 0;
  PPT_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, start, idx_out, centers_out, N, G));
  return ppt_launch_status();
}

}  // namespace

extern "C" PPT_EXPORT int ppt_fps(const float* xyz, const int64_t* start, int64_t* idx_out, float* centers_out,
                                  const void* index, int B, int N, int G, void* stream) {
  if (!xyz || !start || !idx_out || B < 0 || N < 1 || G < 1) return PPT_EINVAL;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (index && spidx::resident(N) && G <= N) return ppt_fps_grid(xyz, start, index, idx_out, centers_out, B, N, G, st);
  if (N <= 512) return launch_fps<128, 4, 1>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 1024) return launch_fps<256, 4, 1>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 2048) return launch_fps<512, 4, 1>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 4096) return launch_fps<512, 8, 1>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 8192) return launch_fps<1024, 8, 1>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 16384) return launch_fps<512, 8, 4>(xyz, start, idx_out, centers_out, B, N, G, st);
  // 8 CTAs of 512 threads per cloud: 0.577 ms for 8 x 32768 -> 512 against 0.684 ms with 4 CTAs of 1024 threads
  if (N <= 32768) return launch_fps<512, 8, 8>(xyz, start, idx_out, centers_out, B, N, G, st);
  if (N <= 65536) return launch_fps<1024, 8, 8>(xyz, start, idx_out, centers_out, B, N, G, st);
  return PPT_ERANGE;
}
