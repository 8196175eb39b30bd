// Version and error strings of the C ABI (include/fragnet_b200.h).
#include "common.cuh"

unsigned long long g_fnb_launches = 0;

extern "C" int fnb_version(void) { return FNB_ABI_VERSION; }

extern "C" uint64_t fnb_launch_count(void) { return __atomic_load_n(&g_fnb_launches, __ATOMIC_RELAXED); }

extern "C" const char *fnb_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case FNB_ERR_NULL: return "required pointer is NULL";
    case FNB_ERR_SIZE: return "invalid size (negative, zero where positive is required, or >= 2^31 for an int32 index space)";
    case FNB_ERR_MODE: return "unknown mode or unsupported width";
    case FNB_ERR_WORKSPACE: return "workspace too small";
    case FNB_ERR_ALIGN: return "pointer or stride not 16-byte aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}
