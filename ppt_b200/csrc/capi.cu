// ABI bookkeeping entry points of libppt_b200.so.
#include "common.cuh"

extern "C" PPT_EXPORT int ppt_abi_version(void) { return PPT_B200_ABI_VERSION; }

extern "C" PPT_EXPORT const char* ppt_strerror(int code) {
  if (code == 0) return "success";
  if (code == PPT_EINVAL) return "ppt_b200: invalid argument (null pointer or bad shape)";
  if (code == PPT_ERANGE) return "ppt_b200: size outside the supported range of this kernel";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "ppt_b200: unknown error code";
}

// ---- measurement aid: the SM clock actually seen while other kernels run ------------------------
// One thread samples (globaltimer ns, clock64 cycles) every `period_ns`.  Launched on a side stream next to
// the kernels under test (it needs no shared memory and a handful of registers, so it co-resides with the
// persistent one-CTA-per-SM kernels), it gives the SM clock INSIDE a kernel: nvidia-smi's 200 ms samples
// cannot see that a B200 drops from 1965 MHz to ~1650 MHz within a tensor-heavy 0.5 ms kernel.
__global__ void clock_probe_kernel(long long* __restrict__ out, int samples, long long period_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < samples; ++i) {
    unsigned long long now;
    do {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now < t0 + (unsigned long long)i * (unsigned long long)period_ns) __nanosleep(200);
    } while (now < t0 + (unsigned long long)i * (unsigned long long)period_ns);
    out[2 * i] = (long long)now;
    out[2 * i + 1] = clock64();
  }
}

extern "C" PPT_EXPORT int ppt_clock_probe(void* out, int samples, int64_t period_ns, void* stream) {
  if (!out || samples < 2 || period_ns < 1000) return PPT_EINVAL;
  clock_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<long long*>(out), samples, period_ns);
  return ppt_launch_status();
}
