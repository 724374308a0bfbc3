// blocks.cu — implementation of blocks.hpp (host-side Block/Stream contract over the C ABI)
// plus the rrb_* C entry points used by the block-level parity tests.
#include "blocks.hpp"

#include <cuda.h>          // driver-API TYPES only; entry points are resolved at run time
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstring>

#include "common.cuh"

namespace rr {

using rrc::fail;

// ------------------------------------------------------------ driver API (VMM) ----
// libcuda is NOT linked (the library must load on machines without a driver); the few
// driver entry points the double-mapped device ring needs are fetched through the runtime.
namespace drv {
typedef CUresult (*pfn_cuMemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*pfn_cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*pfn_cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*pfn_cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*pfn_cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*pfn_cuMemUnmap)(CUdeviceptr, size_t);
typedef CUresult (*pfn_cuMemRelease)(CUmemGenericAllocationHandle);
typedef CUresult (*pfn_cuMemAddressFree)(CUdeviceptr, size_t);

struct Api {
    pfn_cuMemGetAllocationGranularity granularity = nullptr;
    pfn_cuMemAddressReserve reserve = nullptr;
    pfn_cuMemCreate create = nullptr;
    pfn_cuMemMap map = nullptr;
    pfn_cuMemSetAccess set_access = nullptr;
    pfn_cuMemUnmap unmap = nullptr;
    pfn_cuMemRelease release = nullptr;
    pfn_cuMemAddressFree addr_free = nullptr;
    bool ok = false;
};

static const Api& api() {
    static Api a = [] {
        Api x;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        x.ok = get("cuMemGetAllocationGranularity", (void**)&x.granularity) && get("cuMemAddressReserve", (void**)&x.reserve) &&
               get("cuMemCreate", (void**)&x.create) && get("cuMemMap", (void**)&x.map) &&
               get("cuMemSetAccess", (void**)&x.set_access) && get("cuMemUnmap", (void**)&x.unmap) &&
               get("cuMemRelease", (void**)&x.release) && get("cuMemAddressFree", (void**)&x.addr_free);
        return x;
    }();
    return a;
}
}  // namespace drv

static std::atomic<size_t> g_next_stream_id{1};    // NEXT_STREAM_ID, src/lib.rs:273-274

void* graph_stream(int device) {
    static std::mutex mu;
    static cudaStream_t streams[64] = {nullptr};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return nullptr;
    if (!streams[device]) {
        cudaSetDevice(device);
        cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking);
    }
    return streams[device];
}

// --------------------------------------------------------------------- Buffer -----
std::shared_ptr<Buffer> Buffer::create(size_t elem_size, size_t bytes, Residency res, int device, std::string* err) {
    auto set_err = [&](const std::string& m) { if (err) *err = m; return std::shared_ptr<Buffer>(); };
    if (elem_size == 0 || bytes < elem_size) return set_err("stream size too small");
    std::shared_ptr<Buffer> b(new Buffer());
    b->id_ = g_next_stream_id.fetch_add(1);
    const bool pin = res == Residency::HostPinned;
    if (pin) res = Residency::Host;
    b->elem_ = elem_size; b->res_ = res; b->device_ = device;
    if (res == Residency::Host) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        if (bytes % page) return set_err("host stream size must be a multiple of the page size (src/stream.rs:96)");
        if (bytes % elem_size) return set_err("stream size must be a multiple of the element size");
        int fd = memfd_create("rrb_stream", 0);
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { if (fd >= 0) close(fd); return set_err("memfd_create/ftruncate failed"); }
        void* base = mmap(nullptr, 2 * bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (base == MAP_FAILED) { close(fd); return set_err("mmap reserve failed"); }
        void* m1 = mmap(base, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
        void* m2 = mmap((char*)base + bytes, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
        close(fd);
        if (m1 == MAP_FAILED || m2 == MAP_FAILED) { munmap(base, 2 * bytes); return set_err("double mmap failed"); }
        b->base_ = (char*)base; b->map_bytes_ = bytes;
        b->cap_ = bytes / elem_size;
        if (pin) {
            // Page-lock the ring (the analogue of the reference's double mmap, src/nowasm/circular_buffer.rs:96-128,
            // made DMA-able): first the whole doubled range in one registration, else its two halves (both map the
            // same memfd pages).
            if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); munmap(base, 2 * bytes); b->base_ = nullptr; return set_err("no CUDA device for a pinned host stream"); }
            memset(base, 0, bytes);                                          // fault the pages in before locking them
            if (cudaHostRegister(base, 2 * bytes, cudaHostRegisterPortable) == cudaSuccess) {
                b->pinned_parts_ = 1;
            } else {
                cudaGetLastError();
                const cudaError_t e1 = cudaHostRegister(base, bytes, cudaHostRegisterPortable);
                const cudaError_t e2 = e1 == cudaSuccess ? cudaHostRegister((char*)base + bytes, bytes, cudaHostRegisterPortable) : e1;
                if (e1 != cudaSuccess || e2 != cudaSuccess) {
                    cudaGetLastError();
                    if (e1 == cudaSuccess) cudaHostUnregister(base);
                    munmap(base, 2 * bytes); b->base_ = nullptr;
                    return set_err(std::string("cudaHostRegister of the host ring failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
                }
                b->pinned_parts_ = 2;
            }
            b->pinned_ = true;
        }
    } else {
        const auto& d = drv::api();
        if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return set_err("no CUDA device for a device-resident stream");
        if (!d.ok) return set_err("CUDA VMM driver entry points unavailable");
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        size_t gran = 0;
        if (d.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return set_err("cuMemGetAllocationGranularity failed");
        // capacity: round up to the VMM granularity AND to a whole number of elements
        size_t sz = (bytes + gran - 1) / gran * gran;
        while (sz % elem_size) sz += gran;
        CUdeviceptr va = 0;
        if (d.reserve(&va, 2 * sz, gran, 0, 0) != CUDA_SUCCESS) return set_err("cuMemAddressReserve failed");
        CUmemGenericAllocationHandle h = 0;
        if (d.create(&h, sz, &prop, 0) != CUDA_SUCCESS) { d.addr_free(va, 2 * sz); return set_err("cuMemCreate failed"); }
        CUmemAccessDesc acc{};
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (d.map(va, sz, 0, h, 0) != CUDA_SUCCESS || d.map(va + sz, sz, 0, h, 0) != CUDA_SUCCESS ||
            d.set_access(va, 2 * sz, &acc, 1) != CUDA_SUCCESS) {
            d.unmap(va, 2 * sz); d.release(h); d.addr_free(va, 2 * sz);
            return set_err("cuMemMap/cuMemSetAccess failed");
        }
        b->base_ = (char*)va; b->map_bytes_ = sz; b->vmm_handle_ = h; b->vmm_va_ = va;
        b->cap_ = sz / elem_size;
        // zero-initialised like the reference's tempfile (Appendix B.8).  The blocks run on a non-blocking stream, which
        // does not wait for the legacy default stream: fill on that stream and wait, or a source's first H2D copy can
        // be overwritten by a still-running cudaMemset (seen once as an all-zero RtlSdrDecode input in the GPU suite).
        cudaStream_t zs = (cudaStream_t)graph_stream(device);
        cudaMemsetAsync((void*)va, 0, sz, zs);
        cudaStreamSynchronize(zs);
    }
    return b;
}

Buffer::~Buffer() {
    if (!base_) return;
    if (res_ == Residency::Host) {
        if (pinned_) {
            cudaSetDevice(device_);
            cudaStreamSynchronize((cudaStream_t)graph_stream(device_));   // no DMA may still target the ring
            cudaHostUnregister(base_);
            if (pinned_parts_ == 2) cudaHostUnregister(base_ + map_bytes_);
        }
        munmap(base_, 2 * map_bytes_);
    } else {
        const auto& d = drv::api();
        cudaSetDevice(device_);
        cudaDeviceSynchronize();
        d.unmap((CUdeviceptr)vmm_va_, 2 * map_bytes_);
        d.release((CUmemGenericAllocationHandle)vmm_handle_);
        d.addr_free((CUdeviceptr)vmm_va_, 2 * map_bytes_);
    }
}

size_t Buffer::used() { std::lock_guard<std::mutex> lk(mu_); return used_; }
size_t Buffer::free_space() { std::lock_guard<std::mutex> lk(mu_); return cap_ - used_; }

void Buffer::write_window(char** ptr, size_t* len) {
    std::lock_guard<std::mutex> lk(mu_);
    *ptr = base_ + wpos_ * elem_;
    *len = cap_ - used_;
}

void Buffer::produce(size_t n, const std::vector<Tag>& tags) {
    if (n == 0) return;                                       // tags on an empty produce are dropped (:528-533)
    std::lock_guard<std::mutex> lk(mu_);
    for (const Tag& t : tags) {
        const size_t pos = (t.pos + wpos_) % cap_;
        Tag c = t; c.pos = pos;
        tags_[pos].push_back(std::move(c));
    }
    wpos_ = (wpos_ + n) % cap_;
    used_ += n;
    cv_.notify_all();
}

void Buffer::read_window(const char** ptr, size_t* len, std::vector<Tag>* tags) {
    std::lock_guard<std::mutex> lk(mu_);
    const size_t start = rpos_, end = rpos_ + used_;
    *ptr = base_ + rpos_ * elem_;
    *len = used_;
    if (!tags) return;
    tags->clear();
    for (const auto& kv : tags_) {
        const size_t m = kv.first % cap_;
        if (end < cap_ && start < cap_) {
            if (m < start || m >= end) continue;
        } else {
            if (m >= (end % cap_) && m < start) continue;
        }
        for (const Tag& t : kv.second) {
            Tag c = t; c.pos = (t.pos + cap_ - start) % cap_;
            tags->push_back(std::move(c));
        }
    }
    std::stable_sort(tags->begin(), tags->end(), [](const Tag& a, const Tag& b) { return a.pos < b.pos; });
}

void Buffer::consume(size_t n) {
    if (n == 0) return;
    std::lock_guard<std::mutex> lk(mu_);
    const size_t newpos = (rpos_ + n) % cap_;
    if (newpos > rpos_) {
        tags_.erase(tags_.lower_bound(rpos_), tags_.lower_bound(newpos));
    } else {
        tags_.erase(tags_.lower_bound(rpos_), tags_.end());
        tags_.erase(tags_.begin(), tags_.lower_bound(newpos));
    }
    rpos_ = newpos;
    used_ -= n;
    cv_.notify_all();
}

size_t Buffer::wait_for_read(size_t need) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait_for(lk, std::chrono::milliseconds(100), [&] { return used_ >= need; });
    return used_;
}
size_t Buffer::wait_for_write(size_t need) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait_for(lk, std::chrono::milliseconds(100), [&] { return cap_ - used_ >= need; });
    return cap_ - used_;
}
void Buffer::writer_dropped() { std::lock_guard<std::mutex> lk(mu_); writer_alive_ = false; cv_.notify_all(); }
void Buffer::reader_dropped() { std::lock_guard<std::mutex> lk(mu_); reader_alive_ = false; cv_.notify_all(); }
bool Buffer::writer_alive() { std::lock_guard<std::mutex> lk(mu_); return writer_alive_; }
bool Buffer::reader_alive() { std::lock_guard<std::mutex> lk(mu_); return reader_alive_; }

StreamPair new_stream(size_t elem_size, size_t bytes, Residency res, int device, std::string* err) {
    StreamPair p;
    auto b = Buffer::create(elem_size, bytes, res, device, err);
    if (!b) return p;
    p.w.reset(new WriteStream(b));
    p.r.reset(new ReadStream(b));
    return p;
}

// -------------------------------------------------------------------- Scratch -----
Scratch::~Scratch() { if (ptr) { cudaSetDevice(device_); cudaFree(ptr); } }
int Scratch::reserve(int device, size_t bytes) {
    if (bytes <= cap_) return RRC_OK;
    if (ptr) { RRC_CUDA(cudaSetDevice(device_)); RRC_CUDA(cudaFree(ptr)); ptr = nullptr; cap_ = 0; }
    device_ = device;
    RRC_CUDA(cudaSetDevice(device));
    RRC_CUDA(cudaMalloc((void**)&ptr, bytes));
    cap_ = bytes;
    return RRC_OK;
}

// cudaMemcpyAsync between a host ring window (`ring_ptr`, inside `ring`) and device memory.  A pinned ring whose
// halves are two page-locked registrations cannot be crossed by ONE copy (invalid argument): split at the seam.
static int ring_copy(Buffer& ring, void* dst, const void* src, const char* ring_ptr, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    if (!bytes) return RRC_OK;
    const char* seam = ring.dma_seam();
    if (seam && ring_ptr < seam && ring_ptr + bytes > seam) {
        const size_t first = (size_t)(seam - ring_ptr);
        RRC_CUDA(cudaMemcpyAsync(dst, src, first, kind, st));
        RRC_CUDA(cudaMemcpyAsync((char*)dst + first, (const char*)src + first, bytes - first, kind, st));
        return RRC_OK;
    }
    RRC_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, st));
    return RRC_OK;
}

// Device view of an input window: the ring itself if it is device resident, else an H2D copy.
static int stage_input(Buffer& b, const char* win, size_t bytes, Scratch& s, int device, const char** dev) {
    if (b.residency() == Residency::Device) { *dev = win; return RRC_OK; }
    RRC_TRY(s.reserve(device, std::max<size_t>(bytes, 16)));
    RRC_TRY(ring_copy(b, s.ptr, win, win, bytes, cudaMemcpyHostToDevice, (cudaStream_t)graph_stream(device)));
    *dev = s.ptr;
    return RRC_OK;
}
// Device view of an output window; finish_output() copies back and synchronises for host rings
// (a CPU consumer must see the data before produce(), SURVEY section 7 "Hard parts").
static int stage_output(Buffer& b, char* win, size_t bytes, Scratch& s, int device, char** dev) {
    if (b.residency() == Residency::Device) { *dev = win; return RRC_OK; }
    RRC_TRY(s.reserve(device, std::max<size_t>(bytes, 16)));
    *dev = s.ptr;
    return RRC_OK;
}
static int finish_output(Buffer& b, char* win, size_t bytes, Scratch& s, int device) {
    if (b.residency() == Residency::Device) return RRC_OK;
    cudaStream_t st = (cudaStream_t)graph_stream(device);
    RRC_TRY(ring_copy(b, win, s.ptr, win, bytes, cudaMemcpyDeviceToHost, st));
    RRC_CUDA(cudaStreamSynchronize(st));
    return RRC_OK;
}

// A device-resident input ring must live on the block's own device: the VMM mapping only has
// access rights for its device and ordering between blocks relies on the per-device graph stream.
int check_src_device(ReadStream& src, int device, const char* who) {
    if (src.buffer().residency() == Residency::Device && src.buffer().device() != device)
        return fail(RRC_ERR_INVALID, "%s: input ring lives on device %d, block runs on device %d (cross-device chains need a host edge)",
                    who, src.buffer().device(), device);
    return RRC_OK;
}

int make_output_stream(size_t elem, const StreamOpts& o, std::unique_ptr<WriteStream>* w, std::unique_ptr<ReadStream>* r);
static int make_output(size_t elem, const StreamOpts& o, std::unique_ptr<WriteStream>* w, std::unique_ptr<ReadStream>* r) {
    return make_output_stream(elem, o, w, r);
}
int make_output_stream(size_t elem, const StreamOpts& o, std::unique_ptr<WriteStream>* w, std::unique_ptr<ReadStream>* r) {
    std::string err;
    StreamPair p = new_stream(elem, o.bytes, o.res, o.device, &err);
    if (!p.w) return fail(RRC_ERR_CUDA, "new_stream failed: %s", err.c_str());
    *w = std::move(p.w); *r = std::move(p.r);
    return RRC_OK;
}

// ------------------------------------------------------------------ FirFilter -----
int FirFilter::create(std::unique_ptr<ReadStream>& src, bool cplx, const float* taps, size_t ntaps, size_t deci,
                      bool translate, float samp_rate, float freq, unsigned flags, const StreamOpts& o,
                      std::unique_ptr<FirFilter>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    std::unique_ptr<FirFilter> b(new FirFilter());
    b->cplx_ = cplx; b->ntaps_ = ntaps; b->deci_ = deci; b->elem_ = cplx ? 8 : 4; b->device_ = o.device;
    if (src->buffer().elem() != b->elem_) return fail(RRC_ERR_INVALID, "FirFilter: stream element size mismatch");
    RRC_TRY(cplx ? rrc_fir_c32_create(o.device, taps, ntaps, deci, flags, &b->h_)
                 : rrc_fir_f32_create(o.device, taps, ntaps, deci, flags, &b->h_));
    if (translate) RRC_TRY(rrc_fir_set_translate(b->h_, samp_rate, freq));
    RRC_TRY(make_output(b->elem_, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
FirFilter::~FirFilter() { rrc_fir_destroy(h_); }

int FirFilter::work(BlockRet* ret) {          // src/fir.rs:492-550
    const char* in; size_t in_len; std::vector<Tag> tags;
    src_->buffer().read_window(&in, &in_len, &tags);
    char* outp; size_t out_free;
    dst_->buffer().write_window(&outp, &out_free);
    size_t n, need, out_n, wait_need; int wait_out;
    RRC_TRY(rrc_fir_plan(ntaps_, deci_, in_len, out_free, &n, &need, &out_n, &wait_need, &wait_out));
    if (n == 0) {
        *ret = BlockRet::wait(wait_out ? (const StreamWait*)dst_.get() : (const StreamWait*)src_.get(), wait_need);
        return RRC_OK;
    }
    const char* din; char* dout;
    RRC_TRY(stage_input(src_->buffer(), in, need * elem_, sin_, device_, &din));
    RRC_TRY(stage_output(dst_->buffer(), outp, out_n * elem_, sout_, device_, &dout));
    RRC_TRY(rrc_fir_run(h_, din, need, dout, out_n, graph_stream(device_)));
    RRC_TRY(finish_output(dst_->buffer(), outp, out_n * elem_, sout_, device_));
    // a host-resident input window was staged with an async copy: it must have left the ring before consume()
    // hands the space back to the producer (matters once the host ring is pinned)
    if (src_->buffer().residency() == Residency::Host && dst_->buffer().residency() == Residency::Device)
        RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
    tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());   // :536
    src_->buffer().consume(n);                                                                                   // :537
    if (deci_ != 1) for (Tag& t : tags) t.pos /= deci_;                                                          // :541-543
    dst_->buffer().produce(out_n, tags);
    *ret = BlockRet::again();                                                                                    // :549
    return RRC_OK;
}

// ------------------------------------------------------------------ FftFilter -----
int FftFilter::create(std::unique_ptr<ReadStream>& src, const float* taps, size_t ntaps, const StreamOpts& o,
                      std::unique_ptr<FftFilter>* out, bool real) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    const size_t elem = real ? 4 : 8;
    if (src->buffer().elem() != elem) return fail(RRC_ERR_INVALID, real ? "FftFilter(real): stream must carry f32" : "FftFilter: stream must carry Complex<f32>");
    std::unique_ptr<FftFilter> b(new FftFilter());
    b->device_ = o.device; b->ntaps_ = ntaps; b->elem_ = elem;
    RRC_TRY(real ? rrc_fftfilt_f32_create(o.device, taps, ntaps, &b->h_) : rrc_fftfilt_c32_create(o.device, taps, ntaps, &b->h_));
    size_t fft_size;
    RRC_TRY(rrc_fftfilt_ref_fft_size(ntaps, &fft_size, &b->nsamples_));
    RRC_CUDA(cudaSetDevice(o.device));
    RRC_CUDA(cudaMalloc((void**)&b->partial_, b->nsamples_ * elem));
    RRC_TRY(make_output(elem, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
FftFilter::~FftFilter() {
    rrc_fftfilt_destroy(h_);
    if (partial_) { cudaSetDevice(device_); cudaFree(partial_); }
}

int FftFilter::work(BlockRet* ret) {          // src/fft_filter.rs:290-354, whole loop in one pass
    const size_t S = nsamples_, E = elem_;
    cudaStream_t st = (cudaStream_t)graph_stream(device_);
    char* outp; size_t out_free;
    dst_->buffer().write_window(&outp, &out_free);
    const char* in; size_t in_len; std::vector<Tag> tags;
    src_->buffer().read_window(&in, &in_len, &tags);
    size_t blocks, consume, buffered_after, wait_need; int wait_out;
    RRC_TRY(rrc_fftfilt_plan(ntaps_, buffered_, in_len, out_free, &blocks, &consume, &buffered_after, &wait_need, &wait_out));
    const bool host_in = src_->buffer().residency() == Residency::Host;
    const cudaMemcpyKind in_kind = host_in ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    RRC_CUDA(cudaSetDevice(device_));
    char* dout = nullptr;
    if (blocks) RRC_TRY(stage_output(dst_->buffer(), outp, blocks * S * E, sout_, device_, &dout));
    size_t ipos = 0;       // input samples handed to the filter so far
    size_t done = 0;       // blocks done
    if (blocks && buffered_ > 0) {            // complete the block that was being accumulated (:306-308)
        const size_t add = S - buffered_;
        RRC_TRY(ring_copy(src_->buffer(), partial_ + buffered_ * E, in, in, add * E, in_kind, st));
        RRC_TRY(rrc_fftfilt_run(h_, (const float*)partial_, S, (float*)dout, st));
        ipos = add; done = 1;
    }
    if (blocks > done) {                      // whole blocks straight from the input window
        const size_t nb = blocks - done;
        const char* din;
        if (host_in) {
            RRC_TRY(sin_.reserve(device_, nb * S * E));
            RRC_TRY(ring_copy(src_->buffer(), sin_.ptr, in + ipos * E, in + ipos * E, nb * S * E, cudaMemcpyHostToDevice, st));
            din = sin_.ptr;
        } else {
            din = in + ipos * E;
        }
        RRC_TRY(rrc_fftfilt_run(h_, (const float*)din, nb * S, (float*)(dout + done * S * E), st));
        ipos += nb * S;
    }
    // trailing partial accumulation (:306-327): buf keeps `buffered_after` samples
    const size_t base = blocks ? 0 : buffered_;
    if (consume > ipos) RRC_TRY(ring_copy(src_->buffer(), partial_ + base * E, in + ipos * E, in + ipos * E, (consume - ipos) * E, in_kind, st));
    if (host_in) RRC_CUDA(cudaStreamSynchronize(st));          // the host window is released by consume()
    if (blocks) RRC_TRY(finish_output(dst_->buffer(), outp, blocks * S * E, sout_, device_));

    // tags (:309-313): absolute position in (buf ++ input) = buffered + pos; emitted with their block.
    std::vector<Tag> out_tags = std::move(pending_tags_);
    pending_tags_.clear();
    for (Tag& t : tags) {
        if (t.pos >= consume) continue;
        t.pos += buffered_;
        out_tags.push_back(std::move(t));
    }
    std::vector<Tag> emit;
    for (Tag& t : out_tags) {
        if (t.pos < blocks * S) emit.push_back(std::move(t));
        else { t.pos -= blocks * S; pending_tags_.push_back(std::move(t)); }
    }
    src_->buffer().consume(consume);
    dst_->buffer().produce(blocks * S, emit);
    buffered_ = buffered_after;
    *ret = BlockRet::wait(wait_out ? (const StreamWait*)dst_.get() : (const StreamWait*)src_.get(), wait_need);
    return RRC_OK;
}

// ------------------------------------------------------------- FftFilterFloat -----
// The reference widens the stream to Complex, runs the complex FftFilter and keeps .re
// (src/fft_filter.rs:428-470).  Here the inner filter runs the kernel's real-stream mode on f32 inner
// rings (two real blocks per complex transform, rrc_fftfilt_f32_create), so the "convert" and
// "replicate" steps are plain copies and no widened intermediate exists.  The inner rings are sized
// for the same number of SAMPLES as the reference's Complex inner streams, so counts, BlockRets and tag
// positions per work() call are unchanged.

int FftFilterFloat::create(std::unique_ptr<ReadStream>& src, const float* taps, size_t ntaps, const StreamOpts& o,
                           std::unique_ptr<FftFilterFloat>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 4) return fail(RRC_ERR_INVALID, "FftFilterFloat: stream must carry f32");
    if (!taps || ntaps == 0) return fail(RRC_ERR_INVALID, "FftFilterFloat needs at least one tap");
    std::unique_ptr<FftFilterFloat> b(new FftFilterFloat());
    b->device_ = o.device;
    StreamOpts inner = o;
    inner.res = Residency::Device;                            // inner streams never leave the device
    inner.bytes = o.bytes / 2;                                // f32 ring with the sample capacity of a Complex one
    std::string err;
    StreamPair p = new_stream(4, inner.bytes, inner.res, inner.device, &err);
    if (!p.w) return fail(RRC_ERR_CUDA, "new_stream failed: %s", err.c_str());
    b->inner_in_ = std::move(p.w);
    b->inner_in_id_ = b->inner_in_->id();
    RRC_TRY(FftFilter::create(p.r, taps, ntaps, inner, &b->complex_, /*real=*/true));
    b->inner_out_ = b->complex_->take_output();
    RRC_TRY(make_output(4, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}

int FftFilterFloat::work(BlockRet* ret) {     // src/fft_filter.rs:428-490
    cudaStream_t st = (cudaStream_t)graph_stream(device_);
    RRC_CUDA(cudaSetDevice(device_));
    {   // Convert input to Complex (:430-445)
        const char* in; size_t in_len; std::vector<Tag> tags;
        src_->buffer().read_window(&in, &in_len, &tags);
        char* to; size_t to_len;
        inner_in_->buffer().write_window(&to, &to_len);
        const size_t n = std::min(in_len, to_len);
        if (n) {
            const bool host_in = src_->buffer().residency() == Residency::Host;
            RRC_TRY(ring_copy(src_->buffer(), to, in, in, n * 4, host_in ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
            if (host_in) RRC_CUDA(cudaStreamSynchronize(st));
        }
        tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());
        inner_in_->buffer().produce(n, tags);
        src_->buffer().consume(n);
    }
    BlockRet inner;
    RRC_TRY(complex_->work(&inner));          // :450
    {   // Replicate stream write (:453-470)
        const char* from; size_t from_len; std::vector<Tag> tags;
        inner_out_->buffer().read_window(&from, &from_len, &tags);
        char* to; size_t to_len;
        dst_->buffer().write_window(&to, &to_len);
        const size_t n = std::min(from_len, to_len);
        if (n == 0 && from_len != 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
        if (n) {
            char* dout;
            RRC_TRY(stage_output(dst_->buffer(), to, n * 4, sout_, device_, &dout));
            RRC_CUDA(cudaMemcpyAsync(dout, from, n * 4, cudaMemcpyDeviceToDevice, st));
            RRC_TRY(finish_output(dst_->buffer(), to, n * 4, sout_, device_));
        }
        tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());
        inner_out_->buffer().consume(n);
        dst_->buffer().produce(n, tags);
    }
    if (inner.kind == RetKind::WaitForStream) {               // :474-489
        if (inner.stream->id() == inner_in_id_) *ret = BlockRet::wait(src_.get(), inner.need);
        else *ret = BlockRet::wait(dst_.get(), inner.need);
    } else {
        *ret = inner;
    }
    return RRC_OK;
}

// ---------------------------------------------------------- RationalResampler -----
int RationalResampler::create(std::unique_ptr<ReadStream>& src, size_t interp, size_t deci, const StreamOpts& o,
                              std::unique_ptr<RationalResampler>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    std::unique_ptr<RationalResampler> b(new RationalResampler());
    b->device_ = o.device; b->elem_ = src->buffer().elem();
    RRC_TRY(rrc_resampler_create(o.device, b->elem_, interp, deci, &b->h_));   // Err on 0 (:130-135)
    RRC_TRY(make_output(b->elem_, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
RationalResampler::~RationalResampler() { rrc_resampler_destroy(h_); }

bool RationalResampler::eof() {
    int pending = 0;
    rrc_resampler_state(h_, nullptr, nullptr, nullptr, &pending);
    return !pending && src_->eof();
}

int RationalResampler::work(BlockRet* ret) {  // src/rational_resampler.rs:155-206
    char* outp; size_t cap;
    dst_->buffer().write_window(&outp, &cap);
    if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
    const char* in; size_t in_len;
    src_->buffer().read_window(&in, &in_len, nullptr);        // tags dropped (:156,200)
    const char* din = in; char* dout;
    if (in_len) RRC_TRY(stage_input(src_->buffer(), in, in_len * elem_, sin_, device_, &din));
    RRC_TRY(stage_output(dst_->buffer(), outp, cap * elem_, sout_, device_, &dout));
    size_t consumed, produced; int wait_out;
    RRC_TRY(rrc_resampler_run(h_, din, in_len, dout, cap, &consumed, &produced, &wait_out, graph_stream(device_)));
    RRC_TRY(finish_output(dst_->buffer(), outp, produced * elem_, sout_, device_));
    if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
    src_->buffer().consume(consumed);
    dst_->buffer().produce(produced, {});
    *ret = BlockRet::wait(wait_out ? (const StreamWait*)dst_.get() : (const StreamWait*)src_.get(), 1);
    return RRC_OK;
}

// ------------------------------------------------------------ QuadratureDemod -----
int QuadratureDemod::create(std::unique_ptr<ReadStream>& src, float gain, const StreamOpts& o,
                            std::unique_ptr<QuadratureDemod>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 8) return fail(RRC_ERR_INVALID, "QuadratureDemod: stream must carry Complex<f32>");
    std::unique_ptr<QuadratureDemod> b(new QuadratureDemod());
    b->device_ = o.device; b->gain_ = gain;
    RRC_TRY(make_output(4, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}

int QuadratureDemod::work(BlockRet* ret) {    // src/quadrature_demod.rs:46-113
    for (;;) {
        const char* in; size_t in_len;
        src_->buffer().read_window(&in, &in_len, nullptr);    // tags dropped (:111)
        if (in_len < 2) { *ret = BlockRet::wait(src_.get(), 2); return RRC_OK; }
        char* outp; size_t cap;
        dst_->buffer().write_window(&outp, &cap);
        if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
        const size_t n1 = std::min(in_len - 1, cap);
        const char* din; char* dout;
        RRC_TRY(stage_input(src_->buffer(), in, (n1 + 1) * 8, sin_, device_, &din));
        RRC_TRY(stage_output(dst_->buffer(), outp, n1 * 4, sout_, device_, &dout));
        RRC_TRY(rrc_quad_demod_run(device_, (const float*)din, n1 + 1, gain_, (float*)dout, graph_stream(device_)));
        RRC_TRY(finish_output(dst_->buffer(), outp, n1 * 4, sout_, device_));
        if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
        src_->buffer().consume(n1);                           // keeps the last sample as history (:110)
        dst_->buffer().produce(n1, {});
    }
}

static Tag mk_tag_bool(size_t pos, const char* key, bool v);
static Tag mk_tag_u64(size_t pos, const char* key, uint64_t v);

// ------------------------------------------------------------------ FftStream -----
int FftStream::create(std::unique_ptr<ReadStream>& src, size_t size, const StreamOpts& o, std::unique_ptr<FftStream>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 8) return fail(RRC_ERR_INVALID, "FftStream: stream must carry Complex<f32>");
    if (size == 0) return fail(RRC_ERR_INVALID, "FFT size must be nonzero (src/fft_stream.rs:42)");
    std::unique_ptr<FftStream> b(new FftStream());
    b->device_ = o.device; b->size_ = size;
    RRC_TRY(make_output(8, o, &b->dst_, &b->out_r_));
    char* w; size_t cap;
    b->dst_->buffer().write_window(&w, &cap);
    if (size > cap) return fail(RRC_ERR_INVALID, "FFT size (%zu) must be no bigger than stream size (%zu) (src/fft_stream.rs:46-50)", size, cap);
    RRC_TRY(rrc_fft_c32_create(o.device, size, &b->h_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
FftStream::~FftStream() { rrc_fft_destroy(h_); }

int FftStream::work(BlockRet* ret) {          // src/fft_stream.rs:71-117 (one batch per call, then Again)
    const char* in; size_t in_len;
    src_->buffer().read_window(&in, &in_len, nullptr);        // input tags are not forwarded (:73)
    if (in_len < size_) { *ret = BlockRet::wait(src_.get(), size_); return RRC_OK; }
    char* outp; size_t cap;
    dst_->buffer().write_window(&outp, &cap);
    if (cap < size_) { *ret = BlockRet::wait(dst_.get(), size_); return RRC_OK; }
    size_t len = std::min(in_len, cap);
    len -= len % size_;
    const char* din; char* dout;
    RRC_TRY(stage_input(src_->buffer(), in, len * 8, sin_, device_, &din));
    RRC_TRY(stage_output(dst_->buffer(), outp, len * 8, sout_, device_, &dout));
    RRC_TRY(rrc_fft_run(h_, (const float*)din, len / size_, (float*)dout, graph_stream(device_)));
    RRC_TRY(finish_output(dst_->buffer(), outp, len * 8, sout_, device_));
    if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
    std::vector<Tag> tags;
    tags.reserve(len / size_ * 3);
    for (size_t pos = 0; pos < len; pos += size_) {           // :95-106
        tags.push_back(mk_tag_u64(pos, "FftStream::size", (uint64_t)size_));
        tags.push_back(mk_tag_bool(pos, "FftStream::frame", true));
        tags.push_back(mk_tag_bool(pos + size_ - 1, "FftStream::frame", false));
    }
    src_->buffer().consume(len);
    dst_->buffer().produce(len, tags);
    *ret = BlockRet::again();
    return RRC_OK;
}

// --------------------------------------------------------------- RtlSdrDecode -----
int RtlSdrDecode::create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<RtlSdrDecode>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 1) return fail(RRC_ERR_INVALID, "RtlSdrDecode: stream must carry u8");
    std::unique_ptr<RtlSdrDecode> b(new RtlSdrDecode());
    b->device_ = o.device;
    RRC_TRY(make_output(8, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}

int RtlSdrDecode::work(BlockRet* ret) {       // src/rtlsdr_decode.rs:18-48
    for (;;) {
        const char* in; size_t in_len;
        src_->buffer().read_window(&in, &in_len, nullptr);    // "TODO: handle tags" (:21): dropped
        size_t isamples = in_len & ~(size_t)1;                // :23
        if (isamples == 0) { *ret = BlockRet::wait(src_.get(), 2); return RRC_OK; }
        char* outp; size_t cap;
        dst_->buffer().write_window(&outp, &cap);
        if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
        isamples = std::min(isamples, cap * 2);               // :32
        const size_t osamples = isamples / 2;
        const char* din; char* dout;
        RRC_TRY(stage_input(src_->buffer(), in, isamples, sin_, device_, &din));
        RRC_TRY(stage_output(dst_->buffer(), outp, osamples * 8, sout_, device_, &dout));
        RRC_TRY(rrc_rtlsdr_decode_run(device_, (const unsigned char*)din, isamples, (float*)dout, graph_stream(device_)));
        RRC_TRY(finish_output(dst_->buffer(), outp, osamples * 8, sout_, device_));
        if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
        src_->buffer().consume(isamples);
        dst_->buffer().produce(osamples, {});
    }
}

// --------------------------------------------------------------- RtlSdrEncode -----
int RtlSdrEncode::create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<RtlSdrEncode>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 8) return fail(RRC_ERR_INVALID, "RtlSdrEncode: stream must carry Complex");
    std::unique_ptr<RtlSdrEncode> b(new RtlSdrEncode());
    b->device_ = o.device;
    RRC_TRY(make_output(1, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}

int RtlSdrEncode::work(BlockRet* ret) {       // src/rtlsdr_encode.rs:28-52
    for (;;) {
        const char* in; size_t in_len;
        src_->buffer().read_window(&in, &in_len, nullptr);    // "TODO: handle tags" (:31): dropped
        if (in_len == 0) { *ret = BlockRet::wait(src_.get(), 1); return RRC_OK; }            // :33-35
        char* outp; size_t cap;
        dst_->buffer().write_window(&outp, &cap);
        if (cap < 2) { *ret = BlockRet::wait(dst_.get(), 2); return RRC_OK; }                // :37-39
        const size_t isamples = std::min(in_len, cap / 2);    // :41
        const size_t obytes = isamples * 2;
        const char* din; char* dout;
        RRC_TRY(stage_input(src_->buffer(), in, isamples * 8, sin_, device_, &din));
        RRC_TRY(stage_output(dst_->buffer(), outp, obytes, sout_, device_, &dout));
        RRC_TRY(rrc_rtlsdr_encode_run(device_, (const float*)din, isamples, (unsigned char*)dout, graph_stream(device_)));
        RRC_TRY(finish_output(dst_->buffer(), outp, obytes, sout_, device_));
        if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
        src_->buffer().consume(isamples);
        dst_->buffer().produce(obytes, {});
    }
}

// -------------------------------------------------------------------- Hilbert -----
int Hilbert::create(std::unique_ptr<ReadStream>& src, size_t ntaps, int window_type, float window_parm, const StreamOpts& o,
                    std::unique_ptr<Hilbert>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    if (src->buffer().elem() != 4) return fail(RRC_ERR_INVALID, "Hilbert: stream must carry f32");
    if (!(ntaps > 1 && (ntaps & 1) == 1)) return fail(RRC_ERR_INVALID, "hilbert filter len must be odd and greater than 1 (src/hilbert.rs:44-47)");
    std::vector<float> win(ntaps), taps(ntaps);
    RRC_TRY(rrc_make_window(window_type, window_parm, ntaps, win.data()));
    RRC_TRY(rrc_hilbert_taps(win.data(), ntaps, taps.data()));
    std::unique_ptr<Hilbert> b(new Hilbert());
    b->device_ = o.device;
    RRC_TRY(rrc_hilbert_create(o.device, taps.data(), ntaps, &b->h_));
    RRC_TRY(make_output(8, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
Hilbert::~Hilbert() { rrc_hilbert_destroy(h_); }

int Hilbert::work(BlockRet* ret) {            // src/hilbert.rs:72-128 (one pass per call, then Again)
    const char* in; size_t in_len; std::vector<Tag> tags;
    src_->buffer().read_window(&in, &in_len, &tags);
    if (in_len == 0) { *ret = BlockRet::wait(src_.get(), 1); return RRC_OK; }            // :76-78
    char* outp; size_t cap;
    dst_->buffer().write_window(&outp, &cap);
    if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }               // :81-83
    const size_t n = std::min(in_len, cap);   // :85-87: len - ntaps with len = history.len() + inout, history.len() == ntaps
    const char* din; char* dout;
    RRC_TRY(stage_input(src_->buffer(), in, n * 4, sin_, device_, &din));
    RRC_TRY(stage_output(dst_->buffer(), outp, n * 8, sout_, device_, &dout));
    RRC_TRY(rrc_hilbert_run(h_, (const float*)din, n, (float*)dout, graph_stream(device_)));
    RRC_TRY(finish_output(dst_->buffer(), outp, n * 8, sout_, device_));
    if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
    tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());   // :117-120
    dst_->buffer().produce(n, tags);
    src_->buffer().consume(n);
    *ret = BlockRet::again();
    return RRC_OK;
}

// -------------------------------------------------------------------- SyncMap -----
int SyncMap::create(std::unique_ptr<ReadStream>& src, Op op, bool cplx, float val_re, float val_im, const StreamOpts& o,
                    std::unique_ptr<SyncMap>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    std::unique_ptr<SyncMap> b(new SyncMap());
    if (op == Op::ComplexToMag2 || op == Op::IqBalance) cplx = true;
    b->op_ = op; b->cplx_ = cplx; b->re_ = val_re; b->im_ = val_im; b->device_ = o.device;
    b->in_elem_ = cplx ? 8 : 4;
    b->out_elem_ = op == Op::ComplexToMag2 ? 4 : b->in_elem_;
    if (src->buffer().elem() != b->in_elem_) return fail(RRC_ERR_INVALID, "sync block: stream element size mismatch");
    if (op == Op::IqBalance) RRC_TRY(rrc_iq_balance_create(o.device, val_re, &b->iq_));
    RRC_TRY(make_output(b->out_elem_, o, &b->dst_, &b->out_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
SyncMap::~SyncMap() { rrc_iq_balance_destroy(iq_); }
const char* SyncMap::block_name() const {
    switch (op_) {
    case Op::MultiplyConst: return "MultiplyConst";
    case Op::AddConst: return "AddConst";
    case Op::ComplexToMag2: return "ComplexToMag2";
    default: return "IqBalance";
    }
}

int SyncMap::work(BlockRet* ret) {            // rustradio_macros_code/src/lib.rs:458-513
    for (;;) {
        const char* in; size_t in_len; std::vector<Tag> tags;
        src_->buffer().read_window(&in, &in_len, &tags);
        if (in_len == 0) { *ret = BlockRet::wait(src_.get(), 1); return RRC_OK; }
        char* outp; size_t cap;
        dst_->buffer().write_window(&outp, &cap);
        if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
        const size_t n = std::min(in_len, cap);
        const char* din; char* dout;
        void* st = graph_stream(device_);
        RRC_TRY(stage_input(src_->buffer(), in, n * in_elem_, sin_, device_, &din));
        RRC_TRY(stage_output(dst_->buffer(), outp, n * out_elem_, sout_, device_, &dout));
        switch (op_) {
        case Op::MultiplyConst:
            RRC_TRY(cplx_ ? rrc_multiply_const_c32_run(device_, (const float*)din, n, re_, im_, (float*)dout, st)
                          : rrc_multiply_const_f32_run(device_, (const float*)din, n, re_, (float*)dout, st));
            break;
        case Op::AddConst:
            RRC_TRY(cplx_ ? rrc_add_const_c32_run(device_, (const float*)din, n, re_, im_, (float*)dout, st)
                          : rrc_add_const_f32_run(device_, (const float*)din, n, re_, (float*)dout, st));
            break;
        case Op::ComplexToMag2: RRC_TRY(rrc_complex_to_mag2_run(device_, (const float*)din, n, (float*)dout, st)); break;
        case Op::IqBalance: RRC_TRY(rrc_iq_balance_run(iq_, (const float*)din, n, (float*)dout, st)); break;
        }
        RRC_TRY(finish_output(dst_->buffer(), outp, n * out_elem_, sout_, device_));
        if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)st));
        tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());
        src_->buffer().consume(n);
        dst_->buffer().produce(n, tags);
    }
}

// ------------------------------------------------------------------------ Tee -----
int Tee::create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<Tee>* out) {
    if (!src) return fail(RRC_ERR_INVALID, "src is NULL");
    RRC_TRY(check_src_device(*src, o.device, "block constructor"));
    std::unique_ptr<Tee> b(new Tee());
    b->device_ = o.device; b->elem_ = src->buffer().elem();
    RRC_TRY(make_output(b->elem_, o, &b->dst1_, &b->out_r_));
    RRC_TRY(make_output(b->elem_, o, &b->dst2_, &b->out2_r_));
    b->src_ = std::move(src);                 // last fallible step is behind us: `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}

int Tee::work(BlockRet* ret) {                // src/tee.rs:9-24 through the sync loop
    for (;;) {
        const char* in; size_t in_len; std::vector<Tag> tags;
        src_->buffer().read_window(&in, &in_len, &tags);
        if (in_len == 0) { *ret = BlockRet::wait(src_.get(), 1); return RRC_OK; }
        char *o1, *o2; size_t c1, c2;
        dst1_->buffer().write_window(&o1, &c1);
        if (c1 == 0) { *ret = BlockRet::wait(dst1_.get(), 1); return RRC_OK; }
        dst2_->buffer().write_window(&o2, &c2);
        if (c2 == 0) { *ret = BlockRet::wait(dst2_.get(), 1); return RRC_OK; }
        const size_t n = std::min(in_len, std::min(c1, c2));
        const char* din; char *d1, *d2;
        void* st = graph_stream(device_);
        RRC_TRY(stage_input(src_->buffer(), in, n * elem_, sin_, device_, &din));
        RRC_TRY(stage_output(dst1_->buffer(), o1, n * elem_, sout1_, device_, &d1));
        RRC_TRY(stage_output(dst2_->buffer(), o2, n * elem_, sout2_, device_, &d2));
        RRC_TRY(rrc_tee_run(device_, din, n * elem_, d1, d2, st));
        RRC_TRY(finish_output(dst1_->buffer(), o1, n * elem_, sout1_, device_));
        RRC_TRY(finish_output(dst2_->buffer(), o2, n * elem_, sout2_, device_));
        if (src_->buffer().residency() == Residency::Host) RRC_CUDA(cudaStreamSynchronize((cudaStream_t)st));
        tags.erase(std::remove_if(tags.begin(), tags.end(), [&](const Tag& t) { return t.pos >= n; }), tags.end());
        src_->buffer().consume(n);
        dst1_->buffer().produce(n, tags);
        dst2_->buffer().produce(n, tags);
    }
}

// --------------------------------------------------------------- VectorSource -----
int VectorSource::create(const void* data, size_t n, size_t elem_size, uint64_t repeat, const StreamOpts& o,
                         std::unique_ptr<VectorSource>* out) {
    std::unique_ptr<VectorSource> b(new VectorSource());
    b->elem_ = elem_size; b->n_ = n; b->repeat_ = repeat; b->device_ = o.device;
    b->data_.assign((const char*)data, (const char*)data + n * elem_size);
    RRC_TRY(make_output(elem_size, o, &b->dst_, &b->out_r_));
    *out = std::move(b);
    return RRC_OK;
}

static Tag mk_tag_bool(size_t pos, const char* key, bool v) { Tag t; t.pos = pos; t.key = key; t.val.kind = TagKind::Bool; t.val.b = v; return t; }
static Tag mk_tag_u64(size_t pos, const char* key, uint64_t v) { Tag t; t.pos = pos; t.key = key; t.val.kind = TagKind::U64; t.val.u = v; return t; }

int VectorSource::work(BlockRet* ret) {       // src/vector_source.rs:101-143
    if (n_ == 0 || count_ >= repeat_) { *ret = BlockRet::eof(); return RRC_OK; }
    if (!dst_) return fail(RRC_ERR_STATE, "VectorSource output dropped");
    std::vector<Tag> tags;
    if (pos_ == 0) {
        tags.push_back(mk_tag_bool(0, "VectorSource::start", true));
        tags.push_back(mk_tag_u64(0, "VectorSource::repeat", count_));
        if (count_ == 0) tags.push_back(mk_tag_bool(0, "VectorSource::first", true));
    }
    char* outp; size_t cap;
    dst_->buffer().write_window(&outp, &cap);
    if (cap == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }
    const size_t n = std::min(cap, n_ - pos_);
    if (dst_->buffer().residency() == Residency::Device) {
        RRC_CUDA(cudaSetDevice(device_));
        RRC_CUDA(cudaMemcpyAsync(outp, data_.data() + pos_ * elem_, n * elem_, cudaMemcpyHostToDevice, (cudaStream_t)graph_stream(device_)));
        RRC_CUDA(cudaStreamSynchronize((cudaStream_t)graph_stream(device_)));
    } else {
        memcpy(outp, data_.data() + pos_ * elem_, n * elem_);
    }
    dst_->buffer().produce(n, tags);
    pos_ += n;
    if (pos_ == n_) {
        ++count_;
        if (!(count_ < repeat_)) { *ret = BlockRet::eof(); return RRC_OK; }
        pos_ = 0;
    }
    *ret = BlockRet::again();
    return RRC_OK;
}

// ---------------------------------------------------------------------- Graph -----
int Graph::run() {                            // src/graph.rs:99-160
    std::vector<bool> eof(blocks_.size(), false);
    for (;;) {
        bool done = true, all_idle = true;
        for (size_t i = 0; i < blocks_.size(); ++i) {
            if (eof[i]) continue;
            BlockRet r;
            RRC_TRY(blocks_[i]->work(&r));
            switch (r.kind) {
            case RetKind::Again: done = false; all_idle = false; break;
            case RetKind::Pending: done = false; break;
            case RetKind::WaitForStream:
                if (blocks_[i]->eof() || (r.stream && r.stream->closed())) eof[i] = true;
                break;
            case RetKind::EOF_: eof[i] = true; break;
            }
        }
        if (done) break;
        if (all_idle) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    return RRC_OK;
}

}  // namespace rr
