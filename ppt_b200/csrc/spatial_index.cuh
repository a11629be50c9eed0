// Per-cloud spatial index shared by the pruned kNN search and the bucketed FPS (sm_100a).
//
// Built by spatial_index_build (knn_grid.cu): the cloud sorted by 16x16x16 cell along a Hilbert curve, as float4
// {x,y,z,|p|^2} with the original indices, one bounding box per ROW of 32 consecutive sorted points,
// and the first sorted position of every cell.  One record per cloud in a caller-owned buffer.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace spidx {

constexpr int MAX_N = 8192;         // up to here the consumers keep a whole cloud in shared memory (bucketed FPS, kNN)
constexpr int MAX_N_INDEX = 32768;  // up to here an index can be built; the kNN search then reads the sorted cloud from L2
constexpr int MIN_N = 512;    // below this the plain kernels are already cheap
constexpr int CELLS = 4096;   // 16^3

struct CloudHeader {  // 32 bytes at the start of each cloud's record
  float lo[3];
  float inv[3];       // cells per unit length
  int rows;
  int pad;
};

struct RowBox {       // 32 bytes
  float lo[3];
  float npmax;        // max |p|^2 in the row
  float hi[3];
  float unused;
};

__host__ __device__ inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Layout {       // byte offsets inside one cloud's record
  size_t pts, idx, boxes, cells, total;
  __host__ __device__ explicit Layout(int N) {
    const size_t np = (size_t)((N + 31) / 32) * 32;
    pts = align256(sizeof(CloudHeader));
    idx = pts + np * 16;
    boxes = idx + np * sizeof(int);
    cells = boxes + (np / 32) * sizeof(RowBox);
    total = align256(cells + (CELLS + 1) * sizeof(int));
  }
};

inline bool supported(int N) { return N >= MIN_N && N <= MAX_N_INDEX; }
inline bool resident(int N) { return N >= MIN_N && N <= MAX_N; }

}  // namespace spidx

// Host-side launchers (internal, not part of the C ABI).
size_t ppt_index_bytes(int B, int N);
int ppt_index_build(const float* xyz, void* index, int B, int N, cudaStream_t st);
int ppt_knn_grid_search(const float* xyz, const float* query, const void* index, int64_t* idx_out, float* dist_out,
                        float* nb_out, int B, int N, int S, int k, bool group, cudaStream_t st);
int ppt_fps_grid(const float* xyz, const int64_t* start, const void* index, int64_t* idx_out, float* centers_out,
                 int B, int N, int G, cudaStream_t st);
