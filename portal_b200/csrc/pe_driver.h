// Lazily bound CUDA driver entry points (module load + kernel launch of the NVRTC-built scene
// program).  libcuda.so.1 is dlopen'ed at pe_create(device >= 0) so that the library itself
// loads -- and program generation / NVRTC compilation work -- on a machine without a GPU driver.
#pragma once
#include <cstddef>
#include <string>

namespace pe_host {

typedef int CUresult_t;
typedef struct CUmod_st* CUmodule_t;
typedef struct CUfunc_st* CUfunction_t;
typedef unsigned long long CUdeviceptr_t;
typedef struct CUstream_st* CUstream_t;

struct DriverApi {
    CUresult_t (*cuInit)(unsigned) = nullptr;
    CUresult_t (*cuModuleLoadData)(CUmodule_t*, const void*) = nullptr;
    CUresult_t (*cuModuleUnload)(CUmodule_t) = nullptr;
    CUresult_t (*cuModuleGetFunction)(CUfunction_t*, CUmodule_t, const char*) = nullptr;
    CUresult_t (*cuModuleGetGlobal)(CUdeviceptr_t*, size_t*, CUmodule_t, const char*) = nullptr;
    CUresult_t (*cuLaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                 CUstream_t, void**, void**) = nullptr;
    CUresult_t (*cuMemcpyHtoDAsync)(CUdeviceptr_t, const void*, size_t, CUstream_t) = nullptr;
    CUresult_t (*cuMemsetD32Async)(CUdeviceptr_t, unsigned, size_t, CUstream_t) = nullptr;
    CUresult_t (*cuStreamWaitValue32)(CUstream_t, CUdeviceptr_t, unsigned, unsigned) = nullptr;
    CUresult_t (*cuGetErrorString)(CUresult_t, const char**) = nullptr;
    CUresult_t (*cuFuncGetAttribute)(int*, int, CUfunction_t) = nullptr;
    CUresult_t (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction_t, int, size_t) = nullptr;
    bool loaded = false;
};

// Returns nullptr and fills `why` if libcuda cannot be bound.
const DriverApi* driver_api(std::string& why);
std::string driver_error(const DriverApi* api, CUresult_t r);

}  // namespace pe_host
