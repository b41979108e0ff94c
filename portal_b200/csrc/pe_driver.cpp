#include "pe_driver.h"

#include <dlfcn.h>

#include <mutex>

namespace pe_host {

static DriverApi g_api;
static std::string g_why;
static std::once_flag g_once;

template <class F>
static bool bind(void* h, const char* name, F& fn) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    if (!fn) g_why = std::string("libcuda: missing symbol ") + name;
    return fn != nullptr;
}

const DriverApi* driver_api(std::string& why) {
    std::call_once(g_once, [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            g_why = std::string("cannot load libcuda.so.1 (no NVIDIA driver on this machine?): ") + dlerror();
            return;
        }
        bool ok = bind(h, "cuInit", g_api.cuInit) && bind(h, "cuModuleLoadData", g_api.cuModuleLoadData) &&
                  bind(h, "cuModuleUnload", g_api.cuModuleUnload) &&
                  bind(h, "cuModuleGetFunction", g_api.cuModuleGetFunction) &&
                  bind(h, "cuModuleGetGlobal_v2", g_api.cuModuleGetGlobal) &&
                  bind(h, "cuLaunchKernel", g_api.cuLaunchKernel) &&
                  bind(h, "cuMemcpyHtoDAsync_v2", g_api.cuMemcpyHtoDAsync) &&
                  bind(h, "cuMemsetD32Async", g_api.cuMemsetD32Async) &&
                  bind(h, "cuStreamWaitValue32_v2", g_api.cuStreamWaitValue32) &&
                  bind(h, "cuGetErrorString", g_api.cuGetErrorString) &&
                  bind(h, "cuFuncGetAttribute", g_api.cuFuncGetAttribute) &&
                  bind(h, "cuOccupancyMaxActiveBlocksPerMultiprocessor", g_api.cuOccupancyMaxActiveBlocksPerMultiprocessor);
        if (ok && g_api.cuInit(0) != 0) {
            g_why = "cuInit failed";
            ok = false;
        }
        g_api.loaded = ok;
    });
    if (!g_api.loaded) {
        why = g_why;
        return nullptr;
    }
    return &g_api;
}

std::string driver_error(const DriverApi* api, CUresult_t r) {
    const char* s = nullptr;
    if (api && api->cuGetErrorString) api->cuGetErrorString(r, &s);
    return std::string(s ? s : "unknown CUDA driver error") + " (" + std::to_string(r) + ")";
}

}  // namespace pe_host
