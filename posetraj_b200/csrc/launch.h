// posetraj_b200 — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

// records a thread-local error message and returns the code (api.cu)
int pt_fail(int code, const char* what);
// to be called right after a kernel launch: checks cudaGetLastError, counts the launch
int pt_launched(const char* what);
int pt_num_sms();
// index of the calling thread's current device, clamped to [0, PT_MAX_DEVICES): per-device caches (function attributes,
// SM counts, occupancy) are keyed by it, so one process may drive several GPUs
constexpr int PT_MAX_DEVICES = 64;
int pt_device_slot();

#define PT_CHECK_ARG(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) return pt_fail(cudaErrorInvalidValue, msg);      \
  } while (0)

// Programmatic dependent launch (PDL), opt-in: build with -DPT_ENABLE_PDL and run with PT_PDL=1; kernels are then launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a grid may be scheduled while its predecessor in the
// stream is still draining; each kernel runs its private prologue (barrier init, TMEM allocation, descriptor
// prefetch, parameter fetch), then `griddepcontrol.wait`s before touching anything another kernel produced.
// Off by default: inside the captured denoise step it measured 1 % SLOWER than plain graph edges
// (profiles/r1h_pdl_gn.md); the device-side waits are no-ops for grids launched without the attribute.
bool pt_pdl_enabled();

// set for the NEXT pt_launch of the calling thread only: launch cooperatively, i.e. the driver guarantees that every CTA
// of the grid is co-resident or fails the launch (kernels with an in-kernel grid rendezvous: gn_fused_kernel)
inline bool& pt_next_launch_cooperative() {
  static thread_local bool flag = false;
  return flag;
}

template <typename... KArgs, typename... Args>
inline cudaError_t pt_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, int cluster_x,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (pt_next_launch_cooperative()) {
    pt_next_launch_cooperative() = false;
    attr[n].id = cudaLaunchAttributeCooperative;
    attr[n].val.cooperative = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pt_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
