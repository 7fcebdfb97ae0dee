// api.cu -- version / status strings and the two device-function probes used by the parity tests.
#include "common.cuh"

using namespace mrcnn;

MRCNN_EXPORT const char* mrcnn_roi_b200_version(void) { return "mrcnn_roi_b200 0.1.0 (sm_100a)"; }

MRCNN_EXPORT const char* mrcnn_status_string(int rc) {
    switch (rc) {
        case MRCNN_OK: return "ok";
        case MRCNN_ERR_NULL: return "required pointer is NULL";
        case MRCNN_ERR_RANGE: return "size or threshold outside the supported range";
        case MRCNN_ERR_WORKSPACE: return "workspace too small";
        case MRCNN_ERR_ALIGN: return "pointer not 16-byte aligned";
        default: return rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "unknown status";
    }
}

namespace {
__global__ void test_expf_kernel(const float* x, float* y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = det_expf(x[i]);
}
__global__ void test_logf_kernel(const float* x, float* y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = det_logf(x[i]);
}
}  // namespace

MRCNN_EXPORT int mrcnn_test_expf(const float* x, float* y, int n, void* stream) {
    if (!x || !y) return MRCNN_ERR_NULL;
    if (n < 1) return MRCNN_ERR_RANGE;
    test_expf_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return last_error();
}

MRCNN_EXPORT int mrcnn_test_logf(const float* x, float* y, int n, void* stream) {
    if (!x || !y) return MRCNN_ERR_NULL;
    if (n < 1) return MRCNN_ERR_RANGE;
    test_logf_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return last_error();
}
