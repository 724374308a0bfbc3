// blocks.hpp — host-side mirror of rustradio's Block / Stream contract for the
// GPU filtering blocks.
//
// rustradio is Rust and the image has no Rust toolchain, so the block contract
// the `rustradio-cuda` crate implements in Rust (rustradio_b200/rust/) is also
// restated here in C++ over the same C ABI, with the same names, constructor
// arguments, BlockRet protocol, output counts and tag propagation, so that the
// block-level parity tests read like the reference's own tests.
//
// Mirrors (paths relative to the rustradio v0.18.2 tree):
//   Tag / TagValue                      src/stream.rs:17-93
//   Buffer (ring, tags, waits)          src/nowasm/circular_buffer.rs:174-216,340-616
//   ReadStream / WriteStream / EOF      src/stream.rs:180-339
//   Block / BlockRet                    src/block.rs:12-126
//   FirFilter<T>                        src/fir.rs:303-551
//   FftFilter / FftFilterFloat          src/fft_filter.rs:210-491
//   RationalResampler<T>                src/rational_resampler.rs:94-213
//   QuadratureDemod                     src/quadrature_demod.rs:32-114
//   Hilbert                             src/hilbert.rs:22-129
//   MultiplyConst / AddConst / ComplexToMag2 / Tee / IqBalance (sync blocks)
//                                       rustradio_macros_code/src/lib.rs:436-514
//   VectorSource<T> (test fixture)      src/vector_source.rs:60-144
//   Graph::run                          src/graph.rs:99-173
//
// Stream memory: the reference double-mmaps a tempfile so every window is
// contiguous (circular_buffer.rs:96-128).  Here a stream is either
//   HOST   : the same trick with memfd_create + two MAP_FIXED mappings, or
//   DEVICE : the device analogue, one cuMemCreate allocation mapped twice
//            back-to-back into one cuMemAddressReserve range (CUDA VMM),
// so kernels never see a wrap and chained GPU blocks never round-trip through
// host memory.  Capacity is configurable (the reference's is fixed at
// 4,096,000 bytes, src/stream.rs:105).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rustradio_cuda.h"

namespace rr {

enum class TagKind : int { String = 0, Float = 1, Bool = 2, U64 = 3, I64 = 4 };

struct TagValue {
    TagKind kind = TagKind::Bool;
    std::string s;
    float f = 0.f;
    bool b = false;
    uint64_t u = 0;
    int64_t i = 0;
    bool operator==(const TagValue& o) const {
        return kind == o.kind && s == o.s && f == o.f && b == o.b && u == o.u && i == o.i;
    }
};

struct Tag {
    size_t pos = 0;
    std::string key;
    TagValue val;
};

// HostPinned: a Host ring whose doubled mapping is page-locked with cudaHostRegister, so a GPU block's copies
// from / to its windows are real asynchronous DMA (no pageable staging); residency() reports Host, pinned() true.
enum class Residency : int { Host = 0, Device = 1, HostPinned = 2 };

constexpr size_t DEFAULT_STREAM_SIZE = 4096000;   // bytes, src/stream.rs:105

// One circular buffer shared by a WriteStream and a ReadStream.
class Buffer {
public:
    static std::shared_ptr<Buffer> create(size_t elem_size, size_t bytes, Residency res, int device, std::string* err);
    ~Buffer();

    size_t id() const { return id_; }
    size_t elem() const { return elem_; }
    size_t capacity() const { return cap_; }          // samples
    Residency residency() const { return res_; }
    bool pinned() const { return pinned_; }
    // Pinned host rings whose two halves are separate cudaHostRegister ranges: one copy must not span the seam.
    const char* dma_seam() const { return (pinned_ && pinned_parts_ == 2) ? base_ + map_bytes_ : nullptr; }
    int device() const { return device_; }

    size_t used();
    size_t free_space();
    bool is_empty() { return used() == 0; }

    // Windows (contiguous thanks to the double mapping).
    void write_window(char** ptr, size_t* len);                       // circular_buffer.rs:607-615
    void produce(size_t n, const std::vector<Tag>& tags);             // :518-557
    void read_window(const char** ptr, size_t* len, std::vector<Tag>* tags);   // :572-604
    void consume(size_t n);                                           // :472-513

    size_t wait_for_read(size_t need);                                // :433-442 (<= 100 ms)
    size_t wait_for_write(size_t need);                               // :401-410

    // Peer liveness (the reference infers it from Arc::strong_count, src/stream.rs:237-246).
    void writer_dropped();
    void reader_dropped();
    bool writer_alive();
    bool reader_alive();

    char* base() { return base_; }

private:
    Buffer() = default;
    size_t id_ = 0, elem_ = 0, cap_ = 0;
    Residency res_ = Residency::Host;
    int device_ = 0;
    char* base_ = nullptr;        // start of the doubled mapping
    size_t map_bytes_ = 0;        // bytes of ONE mapping
    // DEVICE: VMM handles
    unsigned long long vmm_handle_ = 0;
    unsigned long long vmm_va_ = 0;
    std::mutex mu_;
    std::condition_variable cv_;
    size_t rpos_ = 0, wpos_ = 0, used_ = 0;
    std::map<size_t, std::vector<Tag>> tags_;
    bool writer_alive_ = true, reader_alive_ = true;
    bool pinned_ = false;
    int pinned_parts_ = 0;        // 1: the doubled range registered in one piece, 2: the two halves separately
};

class StreamWait {
public:
    virtual ~StreamWait() = default;
    virtual size_t id() const = 0;
    virtual bool wait(size_t need) const = 0;      // true: `need` will never be satisfied
    virtual bool closed() const = 0;
};

class ReadStream : public StreamWait {
public:
    explicit ReadStream(std::shared_ptr<Buffer> b) : buf_(std::move(b)) {}
    ~ReadStream() override { if (buf_) buf_->reader_dropped(); }
    ReadStream(const ReadStream&) = delete;
    Buffer& buffer() const { return *buf_; }
    std::shared_ptr<Buffer> share() const { return buf_; }
    size_t id() const override { return buf_->id(); }
    bool wait(size_t need) const override { return buf_->wait_for_read(need) < need && !buf_->writer_alive(); }
    bool closed() const override { return !buf_->writer_alive(); }
    bool eof() const { return !buf_->writer_alive() && buf_->is_empty(); }     // src/stream.rs:237-246
private:
    std::shared_ptr<Buffer> buf_;
};

class WriteStream : public StreamWait {
public:
    explicit WriteStream(std::shared_ptr<Buffer> b) : buf_(std::move(b)) {}
    ~WriteStream() override { if (buf_) buf_->writer_dropped(); }
    WriteStream(const WriteStream&) = delete;
    Buffer& buffer() const { return *buf_; }
    size_t id() const override { return buf_->id(); }
    bool wait(size_t need) const override { return buf_->wait_for_write(need) < need && !buf_->reader_alive(); }
    bool closed() const override { return !buf_->reader_alive(); }
private:
    std::shared_ptr<Buffer> buf_;
};

// new_stream::<T>() (src/stream.rs:336-339) with configurable size/residency.
struct StreamPair {
    std::unique_ptr<WriteStream> w;
    std::unique_ptr<ReadStream> r;
};
StreamPair new_stream(size_t elem_size, size_t bytes, Residency res, int device, std::string* err);

enum class RetKind : int { Again = 0, Pending = 1, WaitForStream = 2, EOF_ = 3 };

struct BlockRet {
    RetKind kind = RetKind::Again;
    const StreamWait* stream = nullptr;
    size_t need = 0;
    static BlockRet again() { return {RetKind::Again, nullptr, 0}; }
    static BlockRet eof() { return {RetKind::EOF_, nullptr, 0}; }
    static BlockRet pending() { return {RetKind::Pending, nullptr, 0}; }
    static BlockRet wait(const StreamWait* s, size_t n) { return {RetKind::WaitForStream, s, n}; }
};

struct StreamOpts {
    size_t bytes = DEFAULT_STREAM_SIZE;
    Residency res = Residency::Device;
    int device = 0;
};

// trait Block: BlockName + BlockEOF (src/block.rs:91-126).  work() returns <0 (RRC_ERR_*) on
// failure, which is fatal for the graph like Err in the reference.
class Block {
public:
    virtual ~Block() = default;
    virtual int work(BlockRet* ret) = 0;
    virtual const char* block_name() const = 0;
    virtual bool eof() = 0;
    // Output stream handed back by the constructor in the reference (`new() -> (Self, ReadStream)`).
    std::unique_ptr<ReadStream> take_output() { return std::move(out_r_); }
    std::unique_ptr<ReadStream> take_output2() { return std::move(out2_r_); }   // second output (Tee)
protected:
    std::unique_ptr<ReadStream> out_r_, out2_r_;
};

// Scratch device memory for blocks whose stream lives in host memory.
class Scratch {
public:
    ~Scratch();
    int reserve(int device, size_t bytes);
    char* ptr = nullptr;
private:
    size_t cap_ = 0;
    int device_ = 0;
};

class FirFilter : public Block {
public:
    // FirFilter::builder(taps).deci(deci).translate(samp_rate, freq).build(src); cplx = T is Complex.
    static int create(std::unique_ptr<ReadStream>& src, bool cplx, const float* taps, size_t ntaps, size_t deci,
                      bool translate, float samp_rate, float freq, unsigned flags, const StreamOpts& o,
                      std::unique_ptr<FirFilter>* out);
    ~FirFilter() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return cplx_ ? "FirFilter<Complex>" : "FirFilter<Float>"; }
    bool eof() override { return src_->eof(); }
private:
    FirFilter() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    rrc_fir_t* h_ = nullptr;
    bool cplx_ = true;
    size_t ntaps_ = 0, deci_ = 1, elem_ = 8;
    int device_ = 0;
    Scratch sin_, sout_;
};

class FftFilter : public Block {
public:
    // real = false: FftFilter (Complex stream, Complex taps).  real = true: the inner filter of
    // FftFilterFloat — f32 stream, `taps` are f32, the device runs the real-stream kernel mode.
    static int create(std::unique_ptr<ReadStream>& src, const float* taps, size_t ntaps, const StreamOpts& o,
                      std::unique_ptr<FftFilter>* out, bool real = false);
    ~FftFilter() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "FftFilter"; }
    bool eof() override { return src_->eof(); }
    size_t nsamples() const { return nsamples_; }
private:
    FftFilter() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    rrc_fftfilt_t* h_ = nullptr;
    size_t ntaps_ = 0, nsamples_ = 0, buffered_ = 0, elem_ = 8;
    char* partial_ = nullptr;                 // device: up to nsamples accumulated samples (self.buf)
    std::vector<Tag> pending_tags_;           // self.tags
    int device_ = 0;
    Scratch sin_, sout_;
};

class FftFilterFloat : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, const float* taps, size_t ntaps, const StreamOpts& o,
                      std::unique_ptr<FftFilterFloat>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "FftFilterFloat"; }
    bool eof() override { return src_->eof(); }
private:
    FftFilterFloat() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    std::unique_ptr<WriteStream> inner_in_;
    std::unique_ptr<ReadStream> inner_out_;
    std::unique_ptr<FftFilter> complex_;
    size_t inner_in_id_ = 0;
    int device_ = 0;
    Scratch sin_, sout_;
};

class RationalResampler : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, size_t interp, size_t deci, const StreamOpts& o,
                      std::unique_ptr<RationalResampler>* out);
    ~RationalResampler() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "RationalResampler"; }
    bool eof() override;                                          // src/rational_resampler.rs:209-213
private:
    RationalResampler() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    rrc_resampler_t* h_ = nullptr;
    size_t elem_ = 4;
    int device_ = 0;
    Scratch sin_, sout_;
};

class QuadratureDemod : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, float gain, const StreamOpts& o,
                      std::unique_ptr<QuadratureDemod>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "QuadratureDemod"; }
    bool eof() override { return src_->eof(); }
private:
    QuadratureDemod() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    float gain_ = 1.f;
    int device_ = 0;
    Scratch sin_, sout_;
};

// FftStream (src/fft_stream.rs:27-117): forward FFT of every `size` samples, frame tags.
class FftStream : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, size_t size, const StreamOpts& o, std::unique_ptr<FftStream>* out);
    ~FftStream() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "FftStream"; }
    bool eof() override { return src_->eof(); }
private:
    FftStream() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    rrc_fft_t* h_ = nullptr;
    size_t size_ = 0;
    int device_ = 0;
    Scratch sin_, sout_;
};

// RtlSdrDecode (src/rtlsdr_decode.rs:9-48): ReadStream<u8> -> WriteStream<Complex>.
class RtlSdrDecode : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<RtlSdrDecode>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "RtlSdrDecode"; }
    bool eof() override { return src_->eof(); }
private:
    RtlSdrDecode() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    int device_ = 0;
    Scratch sin_, sout_;
};

// RtlSdrEncode (src/rtlsdr_encode.rs:12-52): ReadStream<Complex> -> WriteStream<u8>, tags dropped.
class RtlSdrEncode : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<RtlSdrEncode>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "RtlSdrEncode"; }
    bool eof() override { return src_->eof(); }
private:
    RtlSdrEncode() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    int device_ = 0;
    Scratch sin_, sout_;
};

// Hilbert (src/hilbert.rs:22-129): ReadStream<Float> -> WriteStream<Complex>, identity tags.
class Hilbert : public Block {
public:
    // Hilbert::new(src, ntaps, &window_type): taps = fir::hilbert(window_type.make_window(ntaps)).
    static int create(std::unique_ptr<ReadStream>& src, size_t ntaps, int window_type, float window_parm, const StreamOpts& o,
                      std::unique_ptr<Hilbert>* out);
    ~Hilbert() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "Hilbert"; }
    bool eof() override { return src_->eof(); }
private:
    Hilbert() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    rrc_hilbert_t* h_ = nullptr;
    int device_ = 0;
    Scratch sin_, sout_;
};

// The macro-generated `sync` blocks (rustradio_macros_code/src/lib.rs:436-514) next to the filters:
// MultiplyConst<T>, AddConst<T>, ComplexToMag2, IqBalance — one input, one output, tags passed through.
class SyncMap : public Block {
public:
    enum class Op : int { MultiplyConst = 0, AddConst = 1, ComplexToMag2 = 2, IqBalance = 3 };
    // cplx: T = Complex (val = re + i*im) or Float (val = re); IqBalance: val_re = alpha.
    static int create(std::unique_ptr<ReadStream>& src, Op op, bool cplx, float val_re, float val_im, const StreamOpts& o,
                      std::unique_ptr<SyncMap>* out);
    ~SyncMap() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override;
    bool eof() override { return src_->eof(); }
private:
    SyncMap() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst_;
    Op op_ = Op::MultiplyConst;
    bool cplx_ = false;
    float re_ = 0.f, im_ = 0.f;
    size_t in_elem_ = 4, out_elem_ = 4;
    rrc_iq_balance_t* iq_ = nullptr;
    int device_ = 0;
    Scratch sin_, sout_;
};

// Tee<T> (src/tee.rs:9-24): every sample and every tag goes to both outputs.
class Tee : public Block {
public:
    static int create(std::unique_ptr<ReadStream>& src, const StreamOpts& o, std::unique_ptr<Tee>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "Tee"; }
    bool eof() override { return src_->eof(); }
private:
    Tee() = default;
    std::unique_ptr<ReadStream> src_;
    std::unique_ptr<WriteStream> dst1_, dst2_;
    size_t elem_ = 4;
    int device_ = 0;
    Scratch sin_, sout1_, sout2_;
};

// Test fixture: VectorSource<T> with its tags (src/vector_source.rs:97-144).
class VectorSource : public Block {
public:
    static int create(const void* data, size_t n, size_t elem_size, uint64_t repeat, const StreamOpts& o,
                      std::unique_ptr<VectorSource>* out);
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "VectorSource"; }
    bool eof() override { return false; }
    void drop_output() { dst_.reset(); }
private:
    VectorSource() = default;
    std::unique_ptr<WriteStream> dst_;
    std::vector<char> data_;
    size_t elem_ = 1, n_ = 0, pos_ = 0;
    uint64_t repeat_ = 1, count_ = 0;
    int device_ = 0;
};

// Repeat (src/lib.rs:449-506).
struct Repeat {
    bool infinite = false;
    uint64_t n = 1, count = 0;
    static Repeat finite(uint64_t n) { Repeat r; r.n = n; return r; }
    static Repeat forever() { Repeat r; r.infinite = true; return r; }
    bool again();
};

// Page-locked staging buffer for file reads that go to a device ring.
struct HostStage {
    char* ptr = nullptr;
    size_t cap = 0;
    ~HostStage();
    int reserve(size_t bytes);
};

int make_output_stream(size_t elem, const StreamOpts& o, std::unique_ptr<WriteStream>* w, std::unique_ptr<ReadStream>* r);
int check_src_device(ReadStream& src, int device, const char* who);   // a device-resident src ring must live on `device`

// FileSource<T> (src/file_source.rs:11-153): raw little-endian samples from a file (a cf32 capture,
// /dev/zero ...), whole samples only, optional repeat.  SURVEY 8f rank 1.
class FileSource : public Block {
public:
    static int create(const char* path, size_t elem_size, Repeat repeat, const StreamOpts& o, std::unique_ptr<FileSource>* out);
    ~FileSource() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "FileSource"; }
    bool eof() override { return false; }
private:
    FileSource() = default;
    std::unique_ptr<WriteStream> dst_;
    int fd_ = -1, device_ = 0;
    std::string path_;
    size_t elem_ = 1;
    Repeat repeat_;
    std::vector<char> buf_, scratch_;      // carried partial bytes (`buf`, :50) / host read buffer
    HostStage stage_;
};

// FileSink<T> (src/file_sink.rs:11-160): raw little-endian samples to a file; Mode Create (fails if the file exists) /
// Overwrite / Append; work() writes everything readable and returns Again, WaitForStream(src, 1) on an empty stream.
// A DEVICE input ring is read through a pinned staging buffer (one D2H copy per work() on the blocks' stream).
class FileSink : public Block {
public:
    enum Mode { Create = 0, Overwrite = 1, Append = 2 };
    static int create(std::unique_ptr<ReadStream>& src, const char* path, int mode, bool flush, int device, std::unique_ptr<FileSink>* out);
    ~FileSink() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "FileSink"; }
    bool eof() override { return src_->eof(); }
private:
    FileSink() = default;
    std::unique_ptr<ReadStream> src_;
    int fd_ = -1, device_ = 0;
    bool flush_ = false;
    std::string path_;
    HostStage stage_;
};

// SigMFSource<T> (src/sigmf.rs:229-613): a SigMF Archive (tar) or separate Recording files
// (<path>-meta / <path>-data); checks core:datatype == <type>_le and core:sample_rate.
class SigMFSource : public Block {
public:
    // samp_rate < 0: no expectation (None).
    static int create(const char* path, size_t elem_size, const char* type_string, double samp_rate, bool ignore_type_error,
                      Repeat repeat, const StreamOpts& o, std::unique_ptr<SigMFSource>* out);
    ~SigMFSource() override;
    int work(BlockRet* ret) override;
    const char* block_name() const override { return "SigMFSource"; }
    bool eof() override { return false; }
    bool sample_rate(double* r) const { if (has_rate_) *r = sample_rate_; return has_rate_; }   // :549-551
    const std::string& datatype() const { return datatype_; }
private:
    SigMFSource() = default;
    std::unique_ptr<WriteStream> dst_;
    int fd_ = -1, device_ = 0;
    std::string path_, datatype_;
    size_t elem_ = 1;
    uint64_t range_lo_ = 0, range_len_ = 0, left_ = 0, pos_ = 0;   // `range`, `left` (:274-279)
    bool has_rate_ = false;
    double sample_rate_ = 0;
    Repeat repeat_;
    std::vector<char> buf_;
    HostStage stage_;
};

// Graph::run (src/graph.rs:99-173): single-threaded round robin.
class Graph {
public:
    void add(Block* b) { blocks_.push_back(b); }
    int run();
private:
    std::vector<Block*> blocks_;
};

// The CUDA stream all blocks of this process enqueue on for `device` (in-order, so a
// consumer block's kernel always sees its producer's output).
void* graph_stream(int device);

}  // namespace rr
