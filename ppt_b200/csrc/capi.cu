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
