// Shared helpers for the sm_100a kernels of libppt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ppt_b200.h"

#define PPT_FULL_MASK 0xffffffffu
#define PPT_EXPORT __attribute__((visibility("default")))

// Launch-site error capture: the C ABI returns the cudaError_t, never exits
// (contrast the reference's vendored kernels, which exit(-1):
// models/pointnext/PointNeXt/openpoints/cpp/pointnet2_batch/src/sampling_gpu.cu:255-259).
#define PPT_RETURN_IF_CUDA(expr)                    \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) return (int)_e;          \
  } while (0)

// cudaFuncSetAttribute (dynamic shared-memory opt-in) is per device: a call site configures its kernel once per
// device it is used on (one process normally drives one GPU, but nothing here should depend on that).
struct PptOncePerDevice {
  unsigned long long done = 0;
  bool need() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

static inline int ppt_launch_status() { return (int)cudaPeekAtLastError(); }

// ---- exact fp32 arithmetic of the reference's CPU path ----
// Intrinsics with explicit rounding are never contracted into FMAs by nvcc,
// whatever -fmad says.

// models/pointbert/misc.py:65  torch.sum((xyz - centroid) ** 2, -1) = (dx*dx + dy*dy) + dz*dz
__device__ __forceinline__ float ppt_fps_dist(float x, float y, float z, float cx, float cy, float cz) {
  const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// torch.sum(p ** 2, -1) for a 3-vector
__device__ __forceinline__ float ppt_sqnorm3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// models/pointbert/dvae.py:146-148: K=3 sgemm is an FMA chain in k order, then
// (-2*dot + |src|^2) + |dst|^2.
__device__ __forceinline__ float ppt_pair_sqdist(float sx, float sy, float sz, float ns,
                                                 float dx, float dy, float dz, float nd) {
  const float dot = __fmaf_rn(sz, dz, __fmaf_rn(sy, dy, __fmul_rn(sx, dx)));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), ns), nd);
}

__device__ __forceinline__ int ppt_lane() { return threadIdx.x & 31; }

// ---- shared memory through explicit 32-bit shared-window addresses ----
// With pointers derived from `extern __shared__`, ptxas re-derives the window base (S2R SR_CgaCtaId; MOV 0x400; LEA)
// in front of shared accesses inside loops instead of keeping it in a register: three extra instructions and an S2R
// latency per access in latency-bound loops (bucketed FPS: three times per iteration).  Taking the address once and
// issuing ld/st.shared on integers keeps it out of the loops.  (volatile: ordered among themselves and against
// barriers, like the pointer accesses they replace.)
__device__ __forceinline__ uint32_t ppt_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 ppt_lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int2 ppt_lds64(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float ppt_lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t ppt_lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void ppt_sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void ppt_sts64(uint32_t a, int x, int y) {
  asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y));
}
__device__ __forceinline__ void ppt_sts_u16(uint32_t a, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v)); }
