// runtime.cu — device/memory/stream/event entry points of the C ABI.
#include <mutex>

#include "common.cuh"

namespace rrc {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
std::atomic<uint64_t> g_launches{0};

static int g_sm[64];
static int g_smem[64];
static std::once_flag g_once[64];

static void probe(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) g_sm[device] = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess) g_smem[device] = v;
}
int sm_count(int device) {
    if (device < 0 || device >= 64) return 0;
    std::call_once(g_once[device], probe, device);
    return g_sm[device];
}
int max_smem_optin(int device) {
    if (device < 0 || device >= 64) return 0;
    std::call_once(g_once[device], probe, device);
    return g_smem[device];
}

__global__ void synth_kernel(uint64_t seed, uint64_t first, float* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = synth_value(seed, first + i);
}

}  // namespace rrc

using namespace rrc;

extern "C" {

int rrc_abi_version(void) { return RRC_ABI_VERSION; }
const char* rrc_last_error(void) { return err_buf(); }

int rrc_device_count(int* count) {
    if (!count) return fail(RRC_ERR_INVALID, "count is NULL");
    RRC_CUDA(cudaGetDeviceCount(count));
    return RRC_OK;
}
int rrc_device_name(int device, char* buf, size_t buflen) {
    if (!buf || !buflen) return fail(RRC_ERR_INVALID, "buf is NULL");
    cudaDeviceProp p;
    RRC_CUDA(cudaGetDeviceProperties(&p, device));
    snprintf(buf, buflen, "%s", p.name);
    return RRC_OK;
}
int rrc_device_sm_count(int device, int* sms) {
    if (!sms) return fail(RRC_ERR_INVALID, "sms is NULL");
    RRC_CUDA(cudaSetDevice(device));
    *sms = sm_count(device);
    return RRC_OK;
}

int rrc_malloc_device(int device, size_t bytes, void** p) {
    if (!p) return fail(RRC_ERR_INVALID, "ptr is NULL");
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    return RRC_OK;
}
int rrc_free_device(int device, void* p) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaFree(p));
    return RRC_OK;
}
int rrc_malloc_pinned(size_t bytes, void** p) {
    if (!p) return fail(RRC_ERR_INVALID, "ptr is NULL");
    RRC_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return RRC_OK;
}
int rrc_free_pinned(void* p) {
    RRC_CUDA(cudaFreeHost(p));
    return RRC_OK;
}
int rrc_host_register(void* p, size_t bytes) {
    RRC_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return RRC_OK;
}
int rrc_host_unregister(void* p) {
    RRC_CUDA(cudaHostUnregister(p));
    return RRC_OK;
}
int rrc_memset_device(int device, void* p, int value, size_t bytes, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMemsetAsync(p, value, bytes, as_stream(stream)));
    return RRC_OK;
}
int rrc_memcpy_h2d(int device, void* d, const void* h, size_t bytes, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return RRC_OK;
}
int rrc_memcpy_d2h(int device, void* h, const void* d, size_t bytes, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return RRC_OK;
}
int rrc_memcpy_d2d(int device, void* dst, const void* src, size_t bytes, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return RRC_OK;
}
int rrc_stream_create(int device, void** stream) {
    if (!stream) return fail(RRC_ERR_INVALID, "stream is NULL");
    RRC_CUDA(cudaSetDevice(device));
    cudaStream_t s;
    RRC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return RRC_OK;
}
int rrc_stream_destroy(int device, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaStreamDestroy(as_stream(stream)));
    return RRC_OK;
}
int rrc_stream_sync(int device, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return RRC_OK;
}
int rrc_device_sync(int device) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaDeviceSynchronize());
    return RRC_OK;
}
int rrc_event_create(int device, void** ev) {
    if (!ev) return fail(RRC_ERR_INVALID, "event is NULL");
    RRC_CUDA(cudaSetDevice(device));
    cudaEvent_t e;
    RRC_CUDA(cudaEventCreate(&e));
    *ev = e;
    return RRC_OK;
}
int rrc_event_destroy(int device, void* ev) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return RRC_OK;
}
int rrc_event_record(int device, void* ev, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaEventRecord((cudaEvent_t)ev, as_stream(stream)));
    return RRC_OK;
}
int rrc_event_sync(int device, void* ev) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
    return RRC_OK;
}
int rrc_event_elapsed_ms(int device, void* a, void* b, float* ms) {
    if (!ms) return fail(RRC_ERR_INVALID, "ms is NULL");
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return RRC_OK;
}
int rrc_synth_f32(int device, uint64_t seed, uint64_t first, float* out, size_t n, void* stream) {
    if (!out && n) return fail(RRC_ERR_INVALID, "out is NULL");
    RRC_CUDA(cudaSetDevice(device));
    if (n == 0) return RRC_OK;
    int sms = sm_count(device);
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)sms * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(seed, first, out, n);
    RRC_CHECK_LAUNCH();
    count_launch();
    return RRC_OK;
}
int rrc_launch_count(uint64_t* n) {
    if (!n) return fail(RRC_ERR_INVALID, "n is NULL");
    *n = g_launches.load();
    return RRC_OK;
}

}  // extern "C"
