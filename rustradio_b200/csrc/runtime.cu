// runtime.cu — device/memory/stream/event entry points of the C ABI.
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdlib>
#include <map>
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

static std::mutex g_near_mu;
static std::map<void*, size_t> g_near;       // rrc_malloc_pinned_near mappings: base -> length

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
    {
        std::lock_guard<std::mutex> g(g_near_mu);
        auto it = g_near.find(p);
        if (it != g_near.end()) {                      // an rrc_malloc_pinned_near mapping
            const size_t len = it->second;
            g_near.erase(it);
            cudaHostUnregister(p);
            munmap(p, len);
            return RRC_OK;
        }
    }
    RRC_CUDA(cudaFreeHost(p));
    return RRC_OK;
}
int rrc_device_numa_node(int device, int* node) {
    if (!node) return fail(RRC_ERR_INVALID, "node is NULL");
    *node = -1;
    char bus[32] = {0};
    RRC_CUDA(cudaDeviceGetPCIBusId(bus, sizeof bus, device));
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE* f = fopen(path, "r")) {
        int v = -1;
        if (fscanf(f, "%d", &v) == 1) *node = v;
        fclose(f);
    }
    return RRC_OK;
}
int rrc_malloc_pinned_near(int device, size_t bytes, void** p) {
    if (!p) return fail(RRC_ERR_INVALID, "ptr is NULL");
    int node = -1;
    rrc_device_numa_node(device, &node);
    if (node < 0 || node >= 1024) return rrc_malloc_pinned(bytes, p);       // no NUMA information: plain pinned allocation
    const size_t page = 2u << 20;
    const size_t len = ((bytes ? bytes : 1) + page - 1) / page * page;
    void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return fail(RRC_ERR_NOMEM, "mmap of %zu bytes failed", len);
    unsigned long mask[16] = {0};
    mask[node / 64] = 1ul << (node % 64);
    // MPOL_PREFERRED = 1: pages come from the GPU's node when it has room (never fails the allocation)
    syscall(SYS_mbind, m, len, 1, mask, 1024ul, 0u);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaHostRegister(m, len, cudaHostRegisterPortable);     // faults the pages in, on the preferred node
    if (e != cudaSuccess) {
        munmap(m, len);
        return fail(RRC_ERR_CUDA, "cudaHostRegister of %zu bytes failed: %s", len, cudaGetErrorString(e));
    }
    {
        std::lock_guard<std::mutex> g(g_near_mu);
        g_near[m] = len;
    }
    *p = m;
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
/* ---- peer access (halos over NVLink without NCCL) ---- */
int rrc_peer_enable(int device, int peer) {
    if (device == peer) return RRC_OK;
    int can = 0;
    RRC_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return fail(RRC_ERR_UNSUPPORTED, "device %d cannot access device %d", device, peer);
    RRC_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    RRC_CUDA(e);
    return RRC_OK;
}
int rrc_memcpy_peer(int dst_device, void* dst, int src_device, const void* src, size_t bytes, void* stream) {
    RRC_CUDA(cudaSetDevice(dst_device));
    RRC_CUDA(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, bytes, as_stream(stream)));
    return RRC_OK;
}
int rrc_ipc_export(int device, void* dev_ptr, unsigned char handle[64]) {
    if (!dev_ptr || !handle) return fail(RRC_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    RRC_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t hd;
    RRC_CUDA(cudaIpcGetMemHandle(&hd, dev_ptr));
    memcpy(handle, &hd, 64);
    return RRC_OK;
}
int rrc_ipc_open(int device, const unsigned char handle[64], void** dev_ptr) {
    if (!dev_ptr || !handle) return fail(RRC_ERR_INVALID, "NULL argument");
    RRC_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    RRC_CUDA(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return RRC_OK;
}
int rrc_ipc_close(int device, void* dev_ptr) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaIpcCloseMemHandle(dev_ptr));
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
int rrc_event_wait(int device, void* ev, void* stream) {
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaStreamWaitEvent(as_stream(stream), (cudaEvent_t)ev, 0));
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
