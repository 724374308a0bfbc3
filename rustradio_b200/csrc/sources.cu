// sources.cu — capture ingest and egress for the hot path (second half): raw sample files
// (cf32 / rf32 / u8 ...) and SigMF recordings / archives streamed into a host, pinned-host or DEVICE
// ring.  Restates
//   FileSource<T>::new / builder().repeat() / work()      rustradio src/file_source.rs:11-153
//   SigMFSource<T>::new2 / from_recording / from_archive / work()   src/sigmf.rs:229-613
//   Repeat::again                                          src/lib.rs:449-506
//   FileSink<T>::new / builder().mode().flush() / work()   src/file_sink.rs:11-160
// Samples on disk are little-endian (Sample::parse, src/lib.rs:724-800), which on this platform is the
// in-memory layout, so "parsing" is a byte copy; whole samples only (a partial tail is never produced).
// The file I/O itself is host work: a read lands in a page-locked staging buffer and goes to a device
// ring with one asynchronous copy on the blocks' stream.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "blocks.hpp"
#include "common.cuh"

using rrc::fail;

namespace rr {

// Repeat::again (src/lib.rs:477-490): registers one completed pass, true = go round again.
bool Repeat::again() {
    ++count;
    if (infinite) return true;
    if (n == 0) return false;
    return count < n;
}

HostStage::~HostStage() { if (ptr) cudaFreeHost(ptr); }
int HostStage::reserve(size_t bytes) {
    if (bytes <= cap) return RRC_OK;
    if (ptr) { cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
    RRC_CUDA(cudaHostAlloc((void**)&ptr, bytes, cudaHostAllocPortable));
    cap = bytes;
    return RRC_OK;
}

// Copy `bytes` of whole samples into the output window (host memcpy, or H2D through pinned staging).
static int emit(WriteStream& dst, int device, HostStage& stage, const char* src, char* win, size_t bytes) {
    Buffer& b = dst.buffer();
    if (b.residency() == Residency::Device) {
        RRC_CUDA(cudaSetDevice(device));
        cudaStream_t st = (cudaStream_t)graph_stream(device);
        const char* from = src;
        if (src != stage.ptr) {                                   // carried bytes live in a std::vector: stage them
            RRC_TRY(stage.reserve(bytes));
            memcpy(stage.ptr, src, bytes);
            from = stage.ptr;
        }
        RRC_CUDA(cudaMemcpyAsync(win, from, bytes, cudaMemcpyHostToDevice, st));
        RRC_CUDA(cudaStreamSynchronize(st));                      // the staging buffer is reused by the next read
    } else {
        memcpy(win, src, bytes);
    }
    return RRC_OK;
}

// ----------------------------------------------------------------- FileSource -----
int FileSource::create(const char* path, size_t elem_size, Repeat repeat, const StreamOpts& o, std::unique_ptr<FileSource>* out) {
    if (!path || !elem_size) return fail(RRC_ERR_INVALID, "FileSource: bad arguments");
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return fail(RRC_ERR_INVALID, "file io on %s: %s", path, strerror(errno));      // Error::file_io, :66
    std::unique_ptr<FileSource> b(new FileSource());
    b->fd_ = fd; b->path_ = path; b->elem_ = elem_size; b->repeat_ = repeat; b->device_ = o.device;
    RRC_TRY(make_output_stream(elem_size, o, &b->dst_, &b->out_r_));
    *out = std::move(b);
    return RRC_OK;
}
FileSource::~FileSource() { if (fd_ >= 0) close(fd_); }

int FileSource::work(BlockRet* ret) {         // src/file_source.rs:91-152
    if (!dst_) return fail(RRC_ERR_STATE, "FileSource output dropped");
    char* win; size_t want;
    dst_->buffer().write_window(&win, &want);
    const size_t have0 = buf_.size() / elem_;
    if (want == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }                    // :96-99
    if (have0 < want) {
        const size_t get_bytes = (want - have0) * elem_;                                      // :102-104
        const bool dev = dst_->buffer().residency() == Residency::Device;
        char* rbuf;
        if (dev) { RRC_TRY(stage_.reserve(get_bytes)); rbuf = stage_.ptr; }
        else { scratch_.resize(get_bytes); rbuf = scratch_.data(); }
        ssize_t n;
        do { n = read(fd_, rbuf, get_bytes); } while (n < 0 && errno == EINTR);
        if (n < 0) return fail(RRC_ERR_INVALID, "file io on %s: %s", path_.c_str(), strerror(errno));
        if (n == 0) {                                                                         // :106-119
            if (repeat_.again()) {
                buf_.clear();                                                                 // partial tail discarded
                if (lseek(fd_, 0, SEEK_SET) < 0) return fail(RRC_ERR_INVALID, "seek on %s: %s", path_.c_str(), strerror(errno));
                *ret = BlockRet::again();
                return RRC_OK;
            }
            *ret = BlockRet::eof();
            return RRC_OK;
        }
        if (buf_.empty() && (size_t)n % elem_ == 0) {                                         // fast path, :120-130
            RRC_TRY(emit(*dst_, device_, stage_, rbuf, win, (size_t)n));
            dst_->buffer().produce((size_t)n / elem_, {});
            *ret = BlockRet::again();
            return RRC_OK;
        }
        buf_.insert(buf_.end(), rbuf, rbuf + n);                                              // :131
    }
    const size_t have = buf_.size() / elem_;                                                  // :134
    if (have == 0) { *ret = BlockRet::pending(); return RRC_OK; }                             // :135-138
    RRC_TRY(emit(*dst_, device_, stage_, buf_.data(), win, have * elem_));
    buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)(have * elem_));                       // :146
    dst_->buffer().produce(have, {});
    *ret = BlockRet::again();
    return RRC_OK;
}

// ---------------------------------------------------------------- SigMFSource -----
namespace {

// The two global fields the source needs (src/sigmf.rs:112-130), from the JSON text.
bool json_string_field(const std::string& s, const char* key, std::string* out) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = s.find(k);
    if (p == std::string::npos) return false;
    p = s.find(':', p + k.size());
    if (p == std::string::npos) return false;
    p = s.find('"', p);
    if (p == std::string::npos) return false;
    const size_t q = s.find('"', p + 1);
    if (q == std::string::npos) return false;
    *out = s.substr(p + 1, q - p - 1);
    return true;
}
bool json_number_field(const std::string& s, const char* key, double* out) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = s.find(k);
    if (p == std::string::npos) return false;
    p = s.find(':', p + k.size());
    if (p == std::string::npos) return false;
    const char* c = s.c_str() + p + 1;
    char* end = nullptr;
    const double v = strtod(c, &end);
    if (end == c) return false;                  // null or not a number
    *out = v;
    return true;
}

bool read_all(int fd, uint64_t off, uint64_t len, std::string* out) {
    out->resize(len);
    uint64_t got = 0;
    while (got < len) {
        const ssize_t n = pread(fd, &(*out)[got], len - got, (off_t)(off + got));
        if (n <= 0) return false;
        got += (uint64_t)n;
    }
    return true;
}

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

struct TarEntry { std::string name; uint64_t pos, size; char type; };

// ustar walk (what tar::Archive::entries_with_seek does for regular archives): name, data offset, size.
int tar_entries(int fd, std::vector<TarEntry>* out) {
    uint64_t off = 0;
    std::string long_name;
    for (;;) {
        unsigned char h[512];
        const ssize_t n = pread(fd, h, 512, (off_t)off);
        if (n == 0) break;
        if (n != 512) return fail(RRC_ERR_INVALID, "short tar header");
        bool zero = true;
        for (int i = 0; i < 512 && zero; ++i) zero = h[i] == 0;
        if (zero) break;
        uint64_t size = 0;
        if (h[124] & 0x80) { for (int i = 125; i < 136; ++i) size = (size << 8) | h[i]; }         // base-256 (GNU)
        else { for (int i = 124; i < 136 && h[i] >= '0' && h[i] <= '7'; ++i) size = size * 8 + (h[i] - '0'); }
        std::string name((const char*)h, strnlen((const char*)h, 100));
        if (memcmp(h + 257, "ustar", 5) == 0 && h[345]) name = std::string((const char*)h + 345, strnlen((const char*)h + 345, 155)) + "/" + name;
        const char type = h[156] ? (char)h[156] : '0';
        const uint64_t data = off + 512;
        if (type == 'L') {                                       // GNU long name: applies to the next entry
            if (!read_all(fd, data, size, &long_name)) return fail(RRC_ERR_INVALID, "short tar long-name entry");
            while (!long_name.empty() && long_name.back() == 0) long_name.pop_back();
        } else {
            if (!long_name.empty()) { name = long_name; long_name.clear(); }
            out->push_back({name, data, size, type});
        }
        off = data + (size + 511) / 512 * 512;
    }
    return RRC_OK;
}

}  // namespace

int SigMFSource::create(const char* path, size_t elem_size, const char* type_string, double samp_rate, bool ignore_type_error,
                        Repeat repeat, const StreamOpts& o, std::unique_ptr<SigMFSource>* out) {
    if (!path || !elem_size || !type_string) return fail(RRC_ERR_INVALID, "SigMFSource: bad arguments");
    std::unique_ptr<SigMFSource> b(new SigMFSource());
    b->elem_ = elem_size; b->repeat_ = repeat; b->device_ = o.device; b->path_ = path;
    std::string meta;
    struct stat stt;
    if (stat(path, &stt) == 0) {                                                  // archive, src/sigmf.rs:439-543
        b->fd_ = open(path, O_RDONLY | O_CLOEXEC);
        if (b->fd_ < 0) return fail(RRC_ERR_INVALID, "file io on %s: %s", path, strerror(errno));
        std::vector<TarEntry> ents;
        RRC_TRY(tar_entries(b->fd_, &ents));
        const TarEntry* m = nullptr;
        for (const TarEntry& e : ents) {
            if (!ends_with(e.name, ".sigmf-meta")) continue;
            if (e.type != '0') return fail(RRC_ERR_INVALID, "data file is of bad type %c", e.type);
            if (m) return fail(RRC_ERR_UNSUPPORTED, "sigmf doesn't yet support multiple recordings in an archive");
            m = &e;
        }
        if (!m) return fail(RRC_ERR_INVALID, "sigmf doesn't contain any recording");
        if (!read_all(b->fd_, m->pos, m->size, &meta)) return fail(RRC_ERR_INVALID, "short read of %s", m->name.c_str());
        const std::string want = m->name.substr(0, m->name.size() - 5) + "-data";
        const TarEntry* d = nullptr;
        for (const TarEntry& e : ents) {
            if (e.name != want) continue;
            if (e.type == 'S') return fail(RRC_ERR_UNSUPPORTED, "SigMF source block doesn't support sparse tar files");
            if (e.type != '0') return fail(RRC_ERR_INVALID, "data file is of bad type %c", e.type);
            if (d) return fail(RRC_ERR_INVALID, "Multiple files named '%s' in archive", want.c_str());
            d = &e;
        }
        if (!d) return fail(RRC_ERR_INVALID, "data file for base %s missing", want.c_str());
        b->range_lo_ = d->pos; b->range_len_ = d->size;
    } else {                                                                      // recording files, :416-437
        const std::string mp = std::string(path) + "-meta", dp = std::string(path) + "-data";
        const int mfd = open(mp.c_str(), O_RDONLY | O_CLOEXEC);
        if (mfd < 0)
            return fail(RRC_ERR_INVALID, "SigMF Archive '%s' doesn't exist, and trying to read separated Recording files failed too: %s: %s",
                        path, mp.c_str(), strerror(errno));
        struct stat ms;
        const bool ok = fstat(mfd, &ms) == 0 && read_all(mfd, 0, (uint64_t)ms.st_size, &meta);
        close(mfd);
        if (!ok) return fail(RRC_ERR_INVALID, "short read of %s", mp.c_str());
        b->fd_ = open(dp.c_str(), O_RDONLY | O_CLOEXEC);
        if (b->fd_ < 0)
            return fail(RRC_ERR_INVALID, "SigMF Archive '%s' doesn't exist, and trying to read separated Recording files failed too: %s: %s",
                        path, dp.c_str(), strerror(errno));
        struct stat ds;
        if (fstat(b->fd_, &ds) != 0) return fail(RRC_ERR_INVALID, "stat of %s failed", dp.c_str());
        b->range_lo_ = 0; b->range_len_ = (uint64_t)ds.st_size;
    }
    if (!json_string_field(meta, "core:datatype", &b->datatype_)) return fail(RRC_ERR_INVALID, "sigmf metadata of %s has no core:datatype", path);
    b->has_rate_ = json_number_field(meta, "core:sample_rate", &b->sample_rate_);
    if (samp_rate >= 0 && b->has_rate_ && b->sample_rate_ != samp_rate)           // :391-399
        return fail(RRC_ERR_INVALID, "sigmf file %s sample rate (%g) is not the expected %g", path, b->sample_rate_, samp_rate);
    if (!ignore_type_error) {                                                     // :401-411
        const std::string expected = std::string(type_string) + "_le";
        if (b->datatype_ != expected)
            return fail(RRC_ERR_INVALID, "sigmf file %s data type (%s) not the expected %s", path, b->datatype_.c_str(), expected.c_str());
    }
    b->left_ = b->range_len_;
    b->pos_ = b->range_lo_;
    RRC_TRY(make_output_stream(elem_size, o, &b->dst_, &b->out_r_));
    *out = std::move(b);
    return RRC_OK;
}
SigMFSource::~SigMFSource() { if (fd_ >= 0) close(fd_); }

int SigMFSource::work(BlockRet* ret) {        // src/sigmf.rs:563-612
    if (!dst_) return fail(RRC_ERR_STATE, "SigMFSource output dropped");
    if ((uint64_t)buf_.size() + left_ < (uint64_t)elem_) {                                    // :565-577
        if (repeat_.again()) {
            buf_.clear();
            pos_ = range_lo_;
            left_ = range_len_;
        } else {
            *ret = BlockRet::eof();
            return RRC_OK;
        }
    }
    char* win; size_t want;
    dst_->buffer().write_window(&win, &want);
    if (want == 0) { *ret = BlockRet::wait(dst_.get(), 1); return RRC_OK; }                    // :579-581
    if (buf_.size() / elem_ == 0) {                                                           // :584-592
        const uint64_t want_bytes = std::min<uint64_t>((uint64_t)want * elem_, left_);
        if (want_bytes == 0) return fail(RRC_ERR_STATE, "SigMFSource: nothing left to read (assert_ne, src/sigmf.rs:587)");
        const size_t old = buf_.size();
        buf_.resize(old + want_bytes);
        ssize_t n;
        do { n = pread(fd_, buf_.data() + old, want_bytes, (off_t)pos_); } while (n < 0 && errno == EINTR);
        if (n < 0) return fail(RRC_ERR_INVALID, "file io on %s: %s", path_.c_str(), strerror(errno));
        buf_.resize(old + (size_t)n);
        pos_ += (uint64_t)n;
        left_ -= (uint64_t)n;
    }
    const size_t samples = std::min(buf_.size() / elem_, want);                               // :593-594
    if (samples == 0) { *ret = BlockRet::pending(); return RRC_OK; }                          // :595-599
    RRC_TRY(emit(*dst_, device_, stage_, buf_.data(), win, samples * elem_));
    dst_->buffer().produce(samples, {});
    buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)(samples * elem_));                    // :608
    *ret = BlockRet::wait(dst_.get(), 1);                                                     // :609
    return RRC_OK;
}

// ------------------------------------------------------------------- FileSink -----
int FileSink::create(std::unique_ptr<ReadStream>& src, const char* path, int mode, bool flush, int device, std::unique_ptr<FileSink>* out) {
    if (!src || !path) return fail(RRC_ERR_INVALID, "FileSink: bad arguments");
    RRC_TRY(check_src_device(*src, device, "block constructor"));
    int flags = O_WRONLY | O_CLOEXEC;
    switch (mode) {                                                        // src/file_sink.rs:93-105
    case Create: flags |= O_CREAT | O_EXCL; break;
    case Overwrite: flags |= O_CREAT | O_TRUNC; break;
    case Append: flags |= O_CREAT | O_APPEND; break;
    default: return fail(RRC_ERR_INVALID, "FileSink: unknown mode %d", mode);
    }
    const int fd = open(path, flags, 0666);
    if (fd < 0) return fail(RRC_ERR_INVALID, "file io on %s: %s", path, strerror(errno));      // Error::file_io, :107
    std::unique_ptr<FileSink> b(new FileSink());
    b->fd_ = fd; b->path_ = path; b->flush_ = flush; b->device_ = device;
    b->src_ = std::move(src);                 // `src` is consumed iff RRC_OK
    *out = std::move(b);
    return RRC_OK;
}
FileSink::~FileSink() { if (fd_ >= 0) close(fd_); }     // writes are unbuffered here: nothing left to flush (Drop, :124-133)

int FileSink::work(BlockRet* ret) {           // src/file_sink.rs:139-159
    const char* in; size_t n;
    src_->buffer().read_window(&in, &n, nullptr);
    if (n == 0) { *ret = BlockRet::wait(src_.get(), 1); return RRC_OK; }
    const size_t bytes = n * src_->buffer().elem();
    const char* from = in;
    if (src_->buffer().residency() == Residency::Device) {
        RRC_CUDA(cudaSetDevice(device_));
        cudaStream_t st = (cudaStream_t)graph_stream(device_);
        RRC_TRY(stage_.reserve(bytes));
        RRC_CUDA(cudaMemcpyAsync(stage_.ptr, in, bytes, cudaMemcpyDeviceToHost, st));
        RRC_CUDA(cudaStreamSynchronize(st));
        from = stage_.ptr;
    }                                         // host rings: a producing block has synchronised before produce() (finish_output)
    size_t off = 0;
    while (off < bytes) {                     // write_all
        const ssize_t w = write(fd_, from + off, bytes - off);
        if (w < 0) { if (errno == EINTR) continue; return fail(RRC_ERR_INVALID, "file io on %s: %s", path_.c_str(), strerror(errno)); }
        off += (size_t)w;
    }
    if (flush_ && fsync(fd_) < 0 && errno != EINVAL && errno != EROFS) return fail(RRC_ERR_INVALID, "flush of %s: %s", path_.c_str(), strerror(errno));
    src_->buffer().consume(n);
    *ret = BlockRet::again();
    return RRC_OK;
}

}  // namespace rr
