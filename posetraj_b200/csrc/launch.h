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

#define PT_CHECK_ARG(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) return pt_fail(cudaErrorInvalidValue, msg);      \
  } while (0)
