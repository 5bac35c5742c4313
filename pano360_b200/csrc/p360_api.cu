// Library-level entry points of the C ABI: version, error text, device facts.
#include "p360_common.cuh"

extern "C" int p360_version(void) { return P360_VERSION; }

extern "C" int p360_last_error(char *buf, int n) {
    if (!buf || n <= 0) return P360_EINVAL;
    strncpy(buf, p360::err_buf(), (size_t)n - 1);
    buf[n - 1] = 0;
    return 0;
}

extern "C" int p360_device_info(int device, int32_t out_host[4]) {
    const char *where = "p360_device_info";
    P360_REQUIRE(out_host != nullptr, where);
    cudaDeviceProp prop;
    P360_CUDA(cudaGetDeviceProperties(&prop, device), where);
    out_host[0] = prop.multiProcessorCount;
    out_host[1] = prop.major;
    out_host[2] = prop.minor;
    out_host[3] = prop.l2CacheSize;
    return 0;
}
