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

// Rectangle copy between any two of {device, peer device, page-locked host}: `rows` runs of `width_bytes`
// bytes, `src_pitch` / `dst_pitch` bytes apart (cudaMemcpy2DAsync, direction inferred from the
// addresses).  The column windows of the compositor upload sub-rectangles of the images and
// download / push sub-rectangles of the mosaic with it.
extern "C" int p360_copy_rect(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
                              int64_t width_bytes, int64_t rows, void *stream) {
    const char *where = "p360_copy_rect";
    P360_REQUIRE(dst && src && width_bytes >= 0 && rows >= 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes, where);
    if (width_bytes == 0 || rows == 0) return 0;
    P360_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)rows,
                                cudaMemcpyDefault, (cudaStream_t)stream), where);
    return 0;
}
