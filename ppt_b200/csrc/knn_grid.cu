// Exact kNN with spatial pruning for clouds of 512 to 32768 points (sm_100a).
//
// Same contract as the brute-force scan in knn.cu (the k nearest under (distance, index) with the
// reference's fp32 distance formula) -- but most of the cloud is never touched:
//
//   knn_prepare_kernel  one CTA per cloud: bounding box -> 16x16x16 cells in Hilbert order -> counting sort in
//                       shared memory -> the cloud in cell order as float4 {x,y,z,|p|^2} + original
//                       indices, one bounding box (+ max |p|^2) per row of 32 consecutive sorted points,
//                       and the first sorted position of every cell.
//   knn_search_kernel   CTA = (cloud, 128 queries), the sorted cloud resident in shared memory; a warp
//                       takes one query at a time: seeds its sorted top-k list from the three rows
//                       around the query's own cell, then visits only rows whose box can still hold a
//                       point closer than the current k-th distance.
//
// Pruning is conservative in floating point: the reference distance d_ref = fl(fl(-2 q.p + |q|^2) + |p|^2)
// differs from the exact squared distance by at most ~10 ulp of (|q|^2 + |p|^2); a row is skipped only
// if its box distance exceeds tau + 2e-6 (|q|^2 + max|p|^2 + |tau|), a bound three times wider than
// that.  Hence the result is bit-identical to the full scan; the order inside a cell (atomics) cannot
// matter because candidates are ordered by (distance, original index) explicitly.
#include "common.cuh"
#include "spatial_index.cuh"

namespace {

using spidx::CELLS;
using spidx::CloudHeader;
using spidx::RowBox;
using GridLayout = spidx::Layout;
constexpr int GRID_MAX_N = spidx::MAX_N;
constexpr int PREP_THREADS = 1024;
constexpr int SEARCH_THREADS = 1024;     // 32 warps: the per-query work is latency-bound (shuffles), so
constexpr int SEARCH_WARPS = SEARCH_THREADS / 32;  // occupancy is what hides it

// Cell of a point: 16 x 16 x 16 grid over the cloud's bounding box, as the linear id x | y << 4 | z << 8.
__device__ __forceinline__ int cell_of(float x, float y, float z, const float* lo, const float* inv) {
  const int cx = min(15, max(0, (int)((x - lo[0]) * inv[0])));
  const int cy = min(15, max(0, (int)((y - lo[1]) * inv[1])));
  const int cz = min(15, max(0, (int)((z - lo[2]) * inv[2])));
  return cx | (cy << 4) | (cz << 8);
}
// Position of cell (x, y, z) along the 3-D Hilbert curve of order 4 (Skilling's transpose algorithm).  The cloud is
// sorted in this order: consecutive cells are always face neighbours, so a row of 32 consecutive points is a compact
// blob.  (Round 1 used Morton order, whose jumps at power-of-two boundaries give the rows that straddle them large
// bounding boxes: more rows touched per FPS iteration, more rows scanned per kNN query.)
__device__ __forceinline__ int hilbert4(int lin) {
  unsigned X[3] = {(unsigned)lin & 15u, ((unsigned)lin >> 4) & 15u, ((unsigned)lin >> 8) & 15u};
#pragma unroll
  for (unsigned Q = 8; Q > 1; Q >>= 1) {
    const unsigned P = Q - 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) X[0] ^= P;
      else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0]; X[2] ^= X[1];
  unsigned t = 0;
#pragma unroll
  for (unsigned Q = 8; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  unsigned h = 0;
#pragma unroll
  for (int bit = 3; bit >= 0; --bit)
#pragma unroll
    for (int i = 0; i < 3; ++i) h = (h << 1) | ((X[i] >> bit) & 1u);
  return (int)h;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(PPT_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PPT_FULL_MASK, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PREP_THREADS, 1)
knn_prepare_kernel(const float* __restrict__ xyz, unsigned char* __restrict__ workspace, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = ((N + 31) / 32) * 32, rows = np / 32;
  // clouds of up to 8192 points are sorted in shared memory and copied out; larger ones (up to 32768) are scattered
  // straight into the record in global memory (`direct`), which then also feeds the row boxes
  const bool direct = N > spidx::MAX_N;
  const int nps = direct ? 0 : np;
  float4* spts = reinterpret_cast<float4*>(smem_raw);                  // [np]   (direct: re-pointed below)
  int* sidx = reinterpret_cast<int*>(spts + nps);                      // [np]
  int* hist = sidx + nps;                                              // [CELLS]  counts, then cursors
  int* warp_tot = hist + CELLS;                                        // [32]
  float* red = reinterpret_cast<float*>(warp_tot + 32);                // [6][32]
  float* box = red + 6 * 32;                                           // lo[3], inv[3]
  unsigned short* lut = reinterpret_cast<unsigned short*>(box + 8);    // [CELLS] linear cell id -> Hilbert position

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  const GridLayout L(N);
  unsigned char* rec = workspace + (size_t)b * L.total;
  if (direct) {
    spts = reinterpret_cast<float4*>(rec + L.pts);
    sidx = reinterpret_cast<int*>(rec + L.idx);
  }

  // 1. bounding box of the cloud
  const float inf = __int_as_float(0x7f800000);
  float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
  for (int n = tid; n < N; n += PREP_THREADS) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = cloud[(size_t)n * 3 + c];
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = warp_min(mn[c]), z = warp_max(mx[c]);
    if (lane == 0) { red[c * 32 + warp] = a; red[(3 + c) * 32 + warp] = z; }
  }
  for (int i = tid; i < CELLS; i += PREP_THREADS) { hist[i] = 0; lut[i] = (unsigned short)hilbert4(i); }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = warp_min(red[c * 32 + lane]), z = warp_max(red[(3 + c) * 32 + lane]);
      if (lane == 0) {
        const float ext = z - a;
        box[c] = a;
        box[3 + c] = ext > 0.f ? 16.0f / ext : 0.f;
      }
    }
  }
  __syncthreads();
  const float lo[3] = {box[0], box[1], box[2]}, inv[3] = {box[3], box[4], box[5]};

  // 2. histogram of cells
  for (int n = tid; n < N; n += PREP_THREADS) {
    const float x = cloud[(size_t)n * 3], y = cloud[(size_t)n * 3 + 1], z = cloud[(size_t)n * 3 + 2];
    atomicAdd(&hist[lut[cell_of(x, y, z, lo, inv)]], 1);
  }
  __syncthreads();

  // 3. exclusive scan of the 4096 counts (4 per thread); cell starts go to the workspace
  int c4[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { c4[i] = hist[tid * 4 + i]; sum += c4[i]; }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(PPT_FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(PPT_FULL_MASK, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive
  }
  __syncthreads();
  int run = warp_tot[warp] + incl - sum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hist[tid * 4 + i] = run;       // first sorted position of Hilbert cell tid * 4 + i; becomes the scatter cursor
    run += c4[i];
  }
  __syncthreads();
  // the searches look a cell up by its linear id
  int* cells_out = reinterpret_cast<int*>(rec + L.cells);
  for (int i = tid; i < CELLS; i += PREP_THREADS) cells_out[i] = hist[lut[i]];
  if (tid == 0) cells_out[CELLS] = N;
  __syncthreads();

  // 4. scatter into cell order (order inside a cell is whatever the atomics give; see header comment)
  for (int n = tid; n < N; n += PREP_THREADS) {
    const float x = cloud[(size_t)n * 3], y = cloud[(size_t)n * 3 + 1], z = cloud[(size_t)n * 3 + 2];
    const int pos = atomicAdd(&hist[lut[cell_of(x, y, z, lo, inv)]], 1);
    spts[pos] = make_float4(x, y, z, ppt_sqnorm3(x, y, z));
    sidx[pos] = n;
  }
  for (int n = N + tid; n < np; n += PREP_THREADS) {
    spts[n] = make_float4(0.f, 0.f, 0.f, inf);  // padding: distance +inf, never selected
    sidx[n] = 0x7fffffff;
  }
  __syncthreads();

  // 5. sorted cloud, indices, per-row boxes and the header to the workspace
  float4* pts_out = reinterpret_cast<float4*>(rec + L.pts);
  int* idx_out = reinterpret_cast<int*>(rec + L.idx);
  if (!direct)
    for (int n = tid; n < np; n += PREP_THREADS) { pts_out[n] = spts[n]; idx_out[n] = sidx[n]; }
  RowBox* boxes = reinterpret_cast<RowBox*>(rec + L.boxes);
  for (int r = warp; r < rows; r += PREP_THREADS / 32) {
    const float4 p = spts[r * 32 + lane];
    const bool ok = r * 32 + lane < N;
    const float bx0 = warp_min(ok ? p.x : inf), by0 = warp_min(ok ? p.y : inf), bz0 = warp_min(ok ? p.z : inf);
    const float bx1 = warp_max(ok ? p.x : -inf), by1 = warp_max(ok ? p.y : -inf), bz1 = warp_max(ok ? p.z : -inf);
    const float nmax = warp_max(ok ? p.w : 0.f);
    if (lane == 0) {
      RowBox rb;
      rb.lo[0] = bx0; rb.lo[1] = by0; rb.lo[2] = bz0; rb.npmax = nmax;
      rb.hi[0] = bx1; rb.hi[1] = by1; rb.hi[2] = bz1; rb.unused = 0.f;
      boxes[r] = rb;
    }
  }
  if (tid == 0) {
    CloudHeader h;
    for (int c = 0; c < 3; ++c) { h.lo[c] = lo[c]; h.inv[c] = inv[c]; }
    h.rows = rows;
    h.pad = 0;
    *reinterpret_cast<CloudHeader*>(rec) = h;
  }
}

// ---------------------------------------------------------------------------------------------
// (distance, original index) as ONE 64-bit key whose unsigned order is the lexicographic order: the
// high word is the distance with its bits flipped into unsigned order (negative distances exist, F3;
// -0 is canonicalised to +0 first), the low word the index.  A compare-exchange is then a 64-bit
// compare and a select instead of two float compares, an int compare and predicate logic.
__device__ __forceinline__ unsigned long long make_key(float d, int idx) {
  const unsigned b = __float_as_uint(d + 0.0f);
  const unsigned f = b ^ ((unsigned)((int)b >> 31) | 0x80000000u);  // negative: all bits flipped; else: the sign bit
  return ((unsigned long long)f << 32) | (unsigned)idx;
}
__device__ __forceinline__ float key_dist(unsigned long long key) {
  const unsigned f = (unsigned)(key >> 32);
  return __uint_as_float(f ^ ((f >> 31) ? 0x80000000u : 0xffffffffu));
}
constexpr unsigned long long KEY_INF = 0xff8000007fffffffull;  // (+inf, INT_MAX)

struct TopList {  // lane i holds the i-th smallest key; tau = entry k-1 (warp-uniform)
  unsigned long long key;
  unsigned long long tau;
};

// One compare-exchange stage of a bitonic network across the warp: partner = lane ^ j.
__device__ __forceinline__ void bitonic_stage(unsigned long long& key, int j, bool want_min) {
  const unsigned long long o = __shfl_xor_sync(PPT_FULL_MASK, key, j);
  if ((o < key) == want_min) key = o;
}
// Full ascending sort of one key per lane: 15 stages.
__device__ __forceinline__ void bitonic_sort32(unsigned long long& key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) bitonic_stage(key, j, ((lane & j) == 0) == ((lane & k) == 0));
}
// list := the 32 smallest of (list U row), ascending.  Both inputs one key per lane; the row need not be sorted.
__device__ __forceinline__ void merge_row(TopList& t, unsigned long long key, int k, int lane) {
  bitonic_sort32(key, lane);
  const unsigned long long r = __shfl_sync(PPT_FULL_MASK, key, 31 - lane);  // descending copy of the row
  if (r < t.key) t.key = r;                                                  // element-wise min: a bitonic sequence
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) bitonic_stage(t.key, j, (lane & j) == 0);
  t.tau = __shfl_sync(PPT_FULL_MASK, t.key, k - 1);
}

// Candidate buffer (one per warp, CAND_CAP keys in shared memory): a scanned row only FILTERS -- lanes whose key beats
// the current k-th best (tau) append it with a ballot compaction, four instructions -- and the expensive part, a
// 32-key bitonic sort plus merge into the list (~170 instructions), runs once per 32 collected candidates instead of
// once per candidate-rich row.  tau is then older than it could be, which only makes the row pruning slightly less
// sharp (a stale tau is still an upper bound of the final k-th distance, so the result stays exact).
constexpr int CAND_CAP = 64;  // fewer than 32 pending + at most 32 from one row

__device__ __forceinline__ void flush32(TopList& t, unsigned long long* __restrict__ buf, int& cnt, int k, int lane) {
  __syncwarp();
  if (t.tau == KEY_INF && __all_sync(PPT_FULL_MASK, t.key == KEY_INF)) {  // empty list (first seed row): a sort is enough
    unsigned long long key = buf[lane];
    bitonic_sort32(key, lane);
    t.key = key;
    t.tau = __shfl_sync(PPT_FULL_MASK, t.key, k - 1);
  } else {
    merge_row(t, buf[lane], k, lane);          // the first 32 pending keys
  }
  const int rest = cnt - 32;                   // 0 .. 31 stay pending
  const unsigned long long x = lane < rest ? buf[32 + lane] : 0ull;
  __syncwarp();
  if (lane < rest) buf[lane] = x;
  cnt = rest;
}

// measured at 128 x 8192 / 512 queries / k = 32 (Hilbert order): 4 and 10 equal (0.121 ms), 20: 0.127; seeding with 1 row 0.148,
// 2 rows 0.126, 3 rows 0.121, 5 rows 0.132
constexpr int INSERT_MAX = 10;  // pending keys up to which serial insertion (~12 instructions each) beats sort + merge (~210)

__device__ __forceinline__ void flush_all(TopList& t, unsigned long long* __restrict__ buf, int& cnt, int k, int lane) {
  if (cnt >= 32) flush32(t, buf, cnt, k, lane);
  if (cnt > INSERT_MAX) {
    __syncwarp();
    merge_row(t, lane < cnt ? buf[lane] : KEY_INF, k, lane);
  } else if (cnt > 0) {
    __syncwarp();
    const unsigned long long mine = lane < cnt ? buf[lane] : KEY_INF;
    for (int i = 0; i < cnt; ++i) {  // warp-uniform
      const unsigned long long c = __shfl_sync(PPT_FULL_MASK, mine, i);
      if (!(c < t.tau)) continue;  // tau tightened since the key was buffered
      const int pos = __popc(__ballot_sync(PPT_FULL_MASK, t.key < c));
      const unsigned long long up = __shfl_up_sync(PPT_FULL_MASK, t.key, 1);
      if (lane == pos) t.key = c;
      else if (lane > pos) t.key = up;
      t.tau = __shfl_sync(PPT_FULL_MASK, t.key, k - 1);
    }
  }
  cnt = 0;
}

__device__ __forceinline__ void scan_row(TopList& t, unsigned long long* __restrict__ buf, int& cnt,
                                         const float4* __restrict__ pts, const int* __restrict__ sidx, int row, float qx,
                                         float qy, float qz, float qn, int k, int lane) {
  const float4 p = pts[row * 32 + lane];
  const unsigned long long key =
      make_key(ppt_pair_sqdist(qx, qy, qz, qn, p.x, p.y, p.z, p.w), sidx[row * 32 + lane]);
  const bool c = key < t.tau;
  const unsigned bal = __ballot_sync(PPT_FULL_MASK, c);
  if (bal) {
    if (c) buf[cnt + __popc(bal & ((1u << lane) - 1u))] = key;
    cnt += __popc(bal);
    if (cnt >= 32) flush32(t, buf, cnt, k, lane);
  }
}

// RESIDENT = false (8192 < N <= 32768): the sorted points and their indices stay in the record (655 KB at 32768 points:
// L2-resident) and rows are read with global loads; only the row and batch boxes live in shared memory.
template <bool GROUP, bool RESIDENT>
__global__ void __launch_bounds__(SEARCH_THREADS, 1)
knn_search_kernel(const float* __restrict__ xyz, const float* __restrict__ query,
                  const unsigned char* __restrict__ workspace, int64_t* __restrict__ idx_out,
                  float* __restrict__ dist_out, float* __restrict__ nb_out, int N, int S, int k,
                  long long total_queries) {
  // Persistent, balanced: the B * S queries are cut into gridDim.x equal contiguous ranges (one CTA per SM), a CTA
  // (re)loads the sorted cloud whenever its range crosses into the next cloud -- at most twice for ranges shorter
  // than S.  (One CTA per (cloud, 128 queries) left the last of 3.46 waves at BASELINE configs[1] 54 % empty.)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = ((N + 31) / 32) * 32, rows = np / 32;
  const int nps = RESIDENT ? np : 0;
  const float4* pts = reinterpret_cast<const float4*>(smem_raw);     // [np]  (not RESIDENT: re-pointed per cloud)
  const int* sidx = reinterpret_cast<const int*>(pts + nps);         // [np]
  RowBox* boxes = reinterpret_cast<RowBox*>(smem_raw + (size_t)nps * 20);  // [rows]
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(boxes + rows) + (threadIdx.x >> 5) * CAND_CAP;
  // one box per BATCH of 32 consecutive rows (1024 consecutive points of the Hilbert order): a query tests these first and only looks
  // at the row boxes of batches its tau-ball reaches
  RowBox* bbox = reinterpret_cast<RowBox*>(reinterpret_cast<unsigned long long*>(boxes + rows) + SEARCH_WARPS * CAND_CAP);
  const int nbatch = (rows + 31) >> 5;  // <= 8 (RESIDENT) / <= 32

  const int tid = threadIdx.x, lane = tid & 31;
  // the warp index through a shuffle: ptxas then knows it is warp-uniform and drops the BRA.DIV guard it otherwise
  // puts in front of every vote / shuffle of the per-query loop
  const int warp = __shfl_sync(PPT_FULL_MASK, tid >> 5, 0);
  const GridLayout L(N);
  const float inf = __int_as_float(0x7f800000);
  const long long per = (total_queries + gridDim.x - 1) / gridDim.x;
  const long long q_lo = (long long)blockIdx.x * per;
  const long long q_hi = q_lo + per < total_queries ? q_lo + per : total_queries;
  for (long long base = q_lo; base < q_hi;) {
  const int b = (int)(base / S);
  const long long cloud_end = (long long)(b + 1) * S;
  const long long seg_end = cloud_end < q_hi ? cloud_end : q_hi;
  const unsigned char* rec = workspace + (size_t)b * L.total;
  const float* cloud = xyz + (size_t)b * N * 3;
  if (base != q_lo) __syncthreads();  // every warp is done with the previous cloud
  if (RESIDENT) {  // contiguous in the record: pts | idx | boxes
    const uint4* src = reinterpret_cast<const uint4*>(rec + L.pts);
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    const int n16 = (int)((L.cells - L.pts) / 16);
    for (int i = tid; i < n16; i += SEARCH_THREADS) dst[i] = __ldg(src + i);
  } else {
    pts = reinterpret_cast<const float4*>(rec + L.pts);
    sidx = reinterpret_cast<const int*>(rec + L.idx);
    const uint4* src = reinterpret_cast<const uint4*>(rec + L.boxes);
    uint4* dst = reinterpret_cast<uint4*>(boxes);
    const int n16 = (int)((L.cells - L.boxes) / 16);
    for (int i = tid; i < n16; i += SEARCH_THREADS) dst[i] = __ldg(src + i);
  }
  const CloudHeader hdr = *reinterpret_cast<const CloudHeader*>(rec);
  const int* cell_start = reinterpret_cast<const int*>(rec + L.cells);
  __syncthreads();
  if (warp < nbatch) {
    const int r = warp * 32 + lane;
    RowBox bx;
    bx.lo[0] = bx.lo[1] = bx.lo[2] = inf; bx.hi[0] = bx.hi[1] = bx.hi[2] = -inf; bx.npmax = 0.f; bx.unused = 0.f;
    if (r < rows) bx = boxes[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) { bx.lo[c] = warp_min(bx.lo[c]); bx.hi[c] = warp_max(bx.hi[c]); }
    bx.npmax = warp_max(bx.npmax);
    if (lane == 0) bbox[warp] = bx;
  }
  __syncthreads();

  for (long long qg = base + warp; qg < seg_end; qg += SEARCH_WARPS) {
    const int q = (int)(qg - (long long)b * S);
    const float* qp = query + ((size_t)b * S + q) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    const float qn = ppt_sqnorm3(qx, qy, qz);
    TopList t;
    t.key = KEY_INF; t.tau = KEY_INF;

    // seed: the rows around the query's own cell
    const int r0 = min(rows - 1, __ldg(cell_start + cell_of(qx, qy, qz, hdr.lo, hdr.inv)) >> 5);
    const int ra = max(0, r0 - 1), rz = min(rows - 1, r0 + 1);
    int cnt = 0;
    scan_row(t, cand, cnt, pts, sidx, r0, qx, qy, qz, qn, k, lane);  // the query's own row first: the tightest first tau
    if (ra < r0) scan_row(t, cand, cnt, pts, sidx, ra, qx, qy, qz, qn, k, lane);
    if (rz > r0) scan_row(t, cand, cnt, pts, sidx, rz, qx, qy, qz, qn, k, lane);
    flush_all(t, cand, cnt, k, lane);  // a tight tau before the pruning pass

    // batches of 32 rows the tau-ball reaches (tau only shrinks from here on, so a batch rejected now stays rejected)
    unsigned bmask;
    {
      float lbb = inf, slb = 0.f;
      if (lane < nbatch) {
        const RowBox bx = bbox[lane];
        const float dx = fmaxf(fmaxf(bx.lo[0] - qx, qx - bx.hi[0]), 0.f);
        const float dy = fmaxf(fmaxf(bx.lo[1] - qy, qy - bx.hi[1]), 0.f);
        const float dz = fmaxf(fmaxf(bx.lo[2] - qz, qz - bx.hi[2]), 0.f);
        lbb = dx * dx + dy * dy + dz * dz;
        slb = 2e-6f * (qn + bx.npmax);
      }
      const float tau0 = key_dist(t.tau);
      bmask = __ballot_sync(PPT_FULL_MASK, !(lbb > tau0 + slb + 2e-6f * fabsf(tau0)));
    }
    // every other row (of those batches) whose box may still contain a closer point
    while (bmask) {
      const int rb = (__ffs(bmask) - 1) << 5;
      bmask &= bmask - 1;
      const int r = rb + lane;
      float lb = inf, slack = 0.f;
      if (r < rows && (r < ra || r > rz)) {
        const RowBox bx = boxes[r];
        const float dx = fmaxf(fmaxf(bx.lo[0] - qx, qx - bx.hi[0]), 0.f);
        const float dy = fmaxf(fmaxf(bx.lo[1] - qy, qy - bx.hi[1]), 0.f);
        const float dz = fmaxf(fmaxf(bx.lo[2] - qz, qz - bx.hi[2]), 0.f);
        lb = dx * dx + dy * dy + dz * dz;
        slack = 2e-6f * (qn + bx.npmax);
      }
      // conservative: keep the row unless lb > tau + 2e-6 (|q|^2 + max|p|^2 + |tau|).  tau may tighten while the
      // batch's rows are scanned; re-testing every row against the newer tau (two shuffles + the bound, 12
      // instructions per row) rejected 5 % of them -- scanning those costs less than asking.
      const float tau_d = key_dist(t.tau);
      unsigned bal = __ballot_sync(PPT_FULL_MASK, !(lb > tau_d + slack + 2e-6f * fabsf(tau_d)));
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        scan_row(t, cand, cnt, pts, sidx, rb + src, qx, qy, qz, qn, k, lane);
      }
    }
    flush_all(t, cand, cnt, k, lane);

    if (lane < k) {
      const size_t o = ((size_t)b * S + q) * k + lane;
      const int ti = (int)(unsigned)(t.key & 0xffffffffull);
      if (idx_out) idx_out[o] = (int64_t)ti;
      if (dist_out) dist_out[o] = key_dist(t.key);
      if (GROUP) {
        const bool ok = (unsigned)ti < (unsigned)N;  // NaN coordinates leave unfilled slots: NaN rows, no wild read
        const float* p = cloud + (size_t)(ok ? ti : 0) * 3;
        const float nanv = __int_as_float(0x7fc00000);
        nb_out[o * 3 + 0] = ok ? __fsub_rn(p[0], qx) : nanv;
        nb_out[o * 3 + 1] = ok ? __fsub_rn(p[1], qy) : nanv;
        nb_out[o * 3 + 2] = ok ? __fsub_rn(p[2], qz) : nanv;
      }
    }
  }
  base = seg_end;
  }
}

size_t prep_smem(int N) {
  const size_t np = N > spidx::MAX_N ? 0 : (size_t)((N + 31) / 32) * 32;
  return np * 20 + (CELLS + 32) * sizeof(int) + (6 * 32 + 8) * sizeof(float) + CELLS * sizeof(unsigned short);
}
size_t search_smem(int N) {
  const size_t np = (size_t)((N + 31) / 32) * 32;
  return (N > spidx::MAX_N ? 0 : np * 20) + (np / 32) * sizeof(RowBox) +
         (size_t)SEARCH_WARPS * CAND_CAP * sizeof(unsigned long long) + 32 * sizeof(RowBox);
}

}  // namespace

size_t ppt_index_bytes(int B, int N) {
  if (!spidx::supported(N)) return 0;
  return (size_t)B * GridLayout(N).total;
}

int ppt_index_build(const float* xyz, void* index, int B, int N, cudaStream_t st) {
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(knn_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)prep_smem(GRID_MAX_N)));
  }
  knn_prepare_kernel<<<B, PREP_THREADS, prep_smem(N), st>>>(xyz, static_cast<unsigned char*>(index), N);
  return ppt_launch_status();
}

int ppt_knn_grid_search(const float* xyz, const float* query, const void* index, int64_t* idx_out, float* dist_out,
                        float* nb_out, int B, int N, int S, int k, bool group, cudaStream_t st) {
  static PptOncePerDevice configured;
  if (configured.need()) {
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(knn_search_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)search_smem(GRID_MAX_N)));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(knn_search_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)search_smem(GRID_MAX_N)));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(knn_search_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)search_smem(spidx::MAX_N_INDEX)));
    PPT_RETURN_IF_CUDA(cudaFuncSetAttribute(knn_search_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)search_smem(spidx::MAX_N_INDEX)));
  }
  const unsigned char* ws = static_cast<const unsigned char*>(index);
  const long long total = (long long)B * S;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one CTA per SM at N = 8192 (the sorted cloud fills most of its shared memory); never fewer than 32 queries each
  const long long want = (total + SEARCH_WARPS - 1) / SEARCH_WARPS;
  const int grid = (int)(want < (long long)sms ? want : (long long)sms);  // 1024 threads x 64 registers: one CTA per SM
  const bool res = N <= spidx::MAX_N;
  auto kern = group ? (res ? knn_search_kernel<true, true> : knn_search_kernel<true, false>)
                    : (res ? knn_search_kernel<false, true> : knn_search_kernel<false, false>);
  kern<<<grid, SEARCH_THREADS, search_smem(N), st>>>(xyz, query, ws, idx_out, dist_out, nb_out, N, S, k, total);
  return ppt_launch_status();
}
