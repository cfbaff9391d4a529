// posetraj_b200 — library-level C ABI: error reporting, launch accounting, TMA descriptor encoding.
#include <atomic>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "launch.h"
#include "../../include/posetraj_b200.h"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int pt_fail(int code, const char* what) {
  if (code == (int)cudaErrorInvalidValue) {
    snprintf(g_err, sizeof(g_err), "%s", what);
  } else {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
  }
  return code;
}

int pt_launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pt_fail((int)e, what);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int pt_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < PT_MAX_DEVICES ? dev : PT_MAX_DEVICES - 1;
}

int pt_num_sms() {
  static int sms[PT_MAX_DEVICES] = {0};
  const int slot = pt_device_slot();
  if (sms[slot] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot) != cudaSuccess || n <= 0) n = 148;
    sms[slot] = n;
  }
  return sms[slot];
}

bool pt_pdl_enabled() {
#ifndef PT_ENABLE_PDL
  return false;   // the device-side griddepcontrol instructions are compiled out: the attribute must stay off
#endif
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PT_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;  // measured slower inside the captured step (profiles/r1h): opt-in
  }
  return on != 0;
}

extern "C" int pt_pdl(void) { return pt_pdl_enabled() ? 1 : 0; }
extern "C" const char* pt_last_error(void) { return g_err; }
extern "C" int pt_version(void) { return 100; }
extern "C" int64_t pt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

extern "C" int pt_tensormap_encode_bf16(PtTensorMap* out, const void* base, int rank, const uint64_t* dims,
                                        const uint64_t* strides_bytes, const uint32_t* box) {
  PT_CHECK_ARG(out != nullptr && base != nullptr && dims != nullptr && box != nullptr, "pt_tensormap_encode_bf16: null argument");
  PT_CHECK_ARG(rank == 2 || rank == 3, "pt_tensormap_encode_bf16: rank must be 2 or 3");
  PT_CHECK_ARG(box[0] == 64, "pt_tensormap_encode_bf16: box[0] must be 64 (128-byte swizzle row)");
  PT_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "pt_tensormap_encode_bf16: base must be 16-byte aligned");
  PT_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 63u) == 0, "pt_tensormap_encode_bf16: out must be 64-byte aligned");
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return pt_fail(cudaErrorInvalidValue, "pt_tensormap_encode_bf16: cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    PT_CHECK_ARG(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "pt_tensormap_encode_bf16: bad dim/box");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    PT_CHECK_ARG((strides_bytes[i] & 15u) == 0, "pt_tensormap_encode_bf16: strides must be multiples of 16 bytes");
  }
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                  const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof(msg), "pt_tensormap_encode_bf16: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return pt_fail(cudaErrorInvalidValue, msg);
  }
  return 0;
}

// ABI self-check: lets the Python binding assert that its ctypes mirrors match the C structs.
extern "C" int pt_sizeof(const char* name) {
  if (name == nullptr) return -1;
  if (!strcmp(name, "PtTensorMap")) return (int)sizeof(PtTensorMap);
  if (!strcmp(name, "PtCfgEulerArgs")) return (int)sizeof(PtCfgEulerArgs);
  if (!strcmp(name, "PtGemmArgs")) return (int)sizeof(PtGemmArgs);
  if (!strcmp(name, "PtMlpArgs")) return (int)sizeof(PtMlpArgs);
  if (!strcmp(name, "PtGroupNormArgs")) return (int)sizeof(PtGroupNormArgs);
  if (!strcmp(name, "PtLayerNormArgs")) return (int)sizeof(PtLayerNormArgs);
  if (!strcmp(name, "PtAttnSpatialArgs")) return (int)sizeof(PtAttnSpatialArgs);
  if (!strcmp(name, "PtAttnTemporalArgs")) return (int)sizeof(PtAttnTemporalArgs);
  if (!strcmp(name, "PtSmallLinearArgs")) return (int)sizeof(PtSmallLinearArgs);
  if (!strcmp(name, "PtSinCosArgs")) return (int)sizeof(PtSinCosArgs);
  if (!strcmp(name, "PtUpsampleArgs")) return (int)sizeof(PtUpsampleArgs);
  if (!strcmp(name, "PtConvDirectArgs")) return (int)sizeof(PtConvDirectArgs);
  if (!strcmp(name, "PtLayoutArgs")) return (int)sizeof(PtLayoutArgs);
  if (!strcmp(name, "PtRasterArgs")) return (int)sizeof(PtRasterArgs);
  if (!strcmp(name, "PtRowBlockCopyArgs")) return (int)sizeof(PtRowBlockCopyArgs);
  if (!strcmp(name, "PtAxpyArgs")) return (int)sizeof(PtAxpyArgs);
  if (!strcmp(name, "PtSoftmaxArgs")) return (int)sizeof(PtSoftmaxArgs);
  if (!strcmp(name, "PtBlurArgs")) return (int)sizeof(PtBlurArgs);
  if (!strcmp(name, "PtBicubicArgs")) return (int)sizeof(PtBicubicArgs);
  if (!strcmp(name, "PtAttnSmallArgs")) return (int)sizeof(PtAttnSmallArgs);
  if (!strcmp(name, "PtTimeConvArgs")) return (int)sizeof(PtTimeConvArgs);
  if (!strcmp(name, "PtEdmLossArgs")) return (int)sizeof(PtEdmLossArgs);
  if (!strcmp(name, "PtGroupNormBwdArgs")) return (int)sizeof(PtGroupNormBwdArgs);
  if (!strcmp(name, "PtLayerNormBwdArgs")) return (int)sizeof(PtLayerNormBwdArgs);
  if (!strcmp(name, "PtColsumArgs")) return (int)sizeof(PtColsumArgs);
  if (!strcmp(name, "PtAdamWArgs")) return (int)sizeof(PtAdamWArgs);
  if (!strcmp(name, "PtWgradArgs")) return (int)sizeof(PtWgradArgs);
  if (!strcmp(name, "PtAttnSpatialBwdArgs")) return (int)sizeof(PtAttnSpatialBwdArgs);
  if (!strcmp(name, "PtAttnTemporalBwdArgs")) return (int)sizeof(PtAttnTemporalBwdArgs);
  if (!strcmp(name, "PtSmallLinearBwdArgs")) return (int)sizeof(PtSmallLinearBwdArgs);
  if (!strcmp(name, "PtColsumGroupedArgs")) return (int)sizeof(PtColsumGroupedArgs);
  return -1;
}
