// api.cu -- version / status strings and the two device-function probes used by the parity tests.
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace mrcnn;

namespace mrcnn {
namespace cache_detail {
std::mutex g_cache_mutex;
constexpr int kMaxDevices = 64;
DeviceProps g_props[kMaxDevices];
bool g_props_ok[kMaxDevices];
struct OccKey { int dev; const void* kernel; int threads, cs; size_t smem; int active; };
std::vector<OccKey> g_occ;
struct SmemKey { int dev; const void* kernel; size_t bytes; };
std::vector<SmemKey> g_static_smem;
int current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); d = 0; } return d; }
}  // namespace cache_detail
using namespace cache_detail;

DeviceProps device_props() {
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    if (dev >= 0 && dev < kMaxDevices && g_props_ok[dev]) return g_props[dev];
    DeviceProps p{148, 227 * 1024};
    if (cudaDeviceGetAttribute(&p.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) p.sms = 148;
    if (cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        p.smem_optin = 227 * 1024;
    (void)cudaGetLastError();
    if (dev >= 0 && dev < kMaxDevices) { g_props[dev] = p; g_props_ok[dev] = true; }
    return p;
}

int max_active_clusters(const void* kernel, int threads, int cs, size_t smem) {
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (const OccKey& k : g_occ)
        if (k.dev == dev && k.kernel == kernel && k.threads == threads && k.cs == cs && k.smem == smem) return k.active;
    int active = -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * 64));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if ((cs > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&active, kernel, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        active = -1;
    }
    g_occ.push_back(OccKey{dev, kernel, threads, cs, smem, active});
    return active;
}

size_t static_smem_bytes(const void* kernel) {
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (const SmemKey& k : g_static_smem)
        if (k.dev == dev && k.kernel == kernel) return k.bytes;
    cudaFuncAttributes fa;
    size_t bytes = 32 * 1024;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) bytes = fa.sharedSizeBytes;
    (void)cudaGetLastError();
    g_static_smem.push_back(SmemKey{dev, kernel, bytes});
    return bytes;
}

// process-wide, read once (C++11 static initialisation is thread-safe): an A/B switch for measurements, not state
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("MRCNN_PDL"); return !(e && strcmp(e, "0") == 0); }();
    return on;
}
int tuning_knob(const char* name, int fallback) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : fallback;
}
}  // namespace mrcnn

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
