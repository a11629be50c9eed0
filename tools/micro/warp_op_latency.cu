// Dependent-chain latency of the warp-level operations the FPS / kNN loops are made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_op_latency warp_op_latency.cu && ./warp_op_latency
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

template <int OP>
__global__ void chain(int* out, long long* cyc, int seed) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + seed) & 1023;
  __syncthreads();
  int v = threadIdx.x * 2654435 + seed;
  unsigned u = (unsigned)v | 1u;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < ITER; ++i) {
    if (OP == 0) v = __reduce_max_sync(0xffffffffu, v ^ threadIdx.x) + (int)threadIdx.x;   // REDUX + 2 ALU
    if (OP == 1) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;                                // SHFL + 1 ALU
    if (OP == 2) v = (int)__ballot_sync(0xffffffffu, v & 1) + (int)threadIdx.x;             // VOTE + ALU (+ISETP)
    if (OP == 3) { u = (unsigned)(31 - __clz((int)u)) | 0x10000u; u += threadIdx.x; }       // FLO + ALU
    if (OP == 4) { u = (unsigned)(__ffs((int)u)) | 0x10000u; u += threadIdx.x; }            // BREV + FLO + ALU
    if (OP == 5) v = sm[v & 1023];                                                          // LDS + LOP
    if (OP == 6) v = v * 3 + 1;                                                             // IMAD
    if (OP == 7) { u = (unsigned)__popc(u) | 0x10000u; u += threadIdx.x; }                  // POPC + ALU
    if (OP == 8) v = __reduce_max_sync(0xffffffffu, __reduce_max_sync(0xffffffffu, v) == v ? (int)threadIdx.x : 0) + v;  // 2 REDUX
    if (OP == 9) { __syncthreads(); v += 1; }                                               // BAR
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + (int)u;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"REDUX.MAX (+2 ALU)", "SHFL (+1 ALU)", "VOTE (+ISETP+ALU)", "FLO (+2 ALU)", "BREV+FLO (+2 ALU)",
                         "LDS (+LOP)", "IMAD", "POPC (+2 ALU)", "REDUX -> cmp/sel -> REDUX (+ALU)", "BAR.SYNC (+ALU)"};
  for (int threads : {32, 512}) {
    printf("block of %d threads, one block: cycles per dependent step\n", threads);
    for (int op = 0; op < 10; ++op) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, cyc, rep); break;
          case 1: chain<1><<<1, threads>>>(out, cyc, rep); break;
          case 2: chain<2><<<1, threads>>>(out, cyc, rep); break;
          case 3: chain<3><<<1, threads>>>(out, cyc, rep); break;
          case 4: chain<4><<<1, threads>>>(out, cyc, rep); break;
          case 5: chain<5><<<1, threads>>>(out, cyc, rep); break;
          case 6: chain<6><<<1, threads>>>(out, cyc, rep); break;
          case 7: chain<7><<<1, threads>>>(out, cyc, rep); break;
          case 8: chain<8><<<1, threads>>>(out, cyc, rep); break;
          case 9: chain<9><<<1, threads>>>(out, cyc, rep); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("  %-36s %7.1f\n", names[op], (double)h / ITER);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
