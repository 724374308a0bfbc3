// blocks_capi.cu — rrb_* C entry points over blocks.hpp.
#include <cstring>

#include "blocks.hpp"
#include "common.cuh"

using rrc::fail;

struct rrb_rstream {
    std::unique_ptr<rr::ReadStream> s;
    std::vector<rr::Tag> tags;      // snapshot of the last read_buf()
};
struct rrb_wstream {
    std::unique_ptr<rr::WriteStream> s;
};
struct rrb_block {
    std::unique_ptr<rr::Block> b;
};

namespace {

rr::StreamOpts opts(size_t bytes, int residency, int device) {
    rr::StreamOpts o;
    o.bytes = bytes ? bytes : rr::DEFAULT_STREAM_SIZE;
    o.res = residency == RRB_HOST ? rr::Residency::Host : residency == RRB_HOST_PINNED ? rr::Residency::HostPinned : rr::Residency::Device;
    o.device = device;
    return o;
}

template <typename B>
int finish(std::unique_ptr<B> b, rrb_block_t** blk, rrb_rstream_t** out, rrb_rstream_t* src = nullptr) {
    if (src) delete src;               // its stream has been moved into the block
    auto* r = new rrb_rstream();
    r->s = b->take_output();
    auto* h = new rrb_block();
    h->b = std::move(b);
    *blk = h; *out = r;
    return RRC_OK;
}

// Ownership rule of every constructor (documented in the header): `src` is consumed iff the call
// returns RRC_OK.  create() moves the stream out of the wrapper only after its last fallible step;
// the wrapper itself is deleted by finish().

}  // namespace

extern "C" {

int rrb_stream_new(size_t elem_size, size_t bytes, int residency, int device, rrb_wstream_t** w, rrb_rstream_t** r) {
    if (!w || !r) return fail(RRC_ERR_INVALID, "NULL argument");
    std::string err;
    rr::StreamOpts o = opts(bytes, residency, device);
    rr::StreamPair p = rr::new_stream(elem_size, o.bytes, o.res, o.device, &err);
    if (!p.w) return fail(o.res == rr::Residency::Device ? RRC_ERR_CUDA : RRC_ERR_INVALID, "new_stream: %s", err.c_str());
    *w = new rrb_wstream{std::move(p.w)};
    *r = new rrb_rstream{std::move(p.r), {}};
    return RRC_OK;
}

int rrb_wstream_write(rrb_wstream_t* w, const void* data, size_t n, const rrb_tag_t* tags, size_t ntags, size_t* written) {
    if (!w || !w->s) return fail(RRC_ERR_INVALID, "stream is NULL");
    rr::Buffer& b = w->s->buffer();
    char* p; size_t len;
    b.write_window(&p, &len);
    const size_t m = n < len ? n : len;
    if (m) {
        if (b.residency() == rr::Residency::Device) {
            RRC_CUDA(cudaSetDevice(b.device()));
            cudaStream_t st = (cudaStream_t)rr::graph_stream(b.device());
            RRC_CUDA(cudaMemcpyAsync(p, data, m * b.elem(), cudaMemcpyHostToDevice, st));
            RRC_CUDA(cudaStreamSynchronize(st));
        } else {
            memcpy(p, data, m * b.elem());
        }
    }
    std::vector<rr::Tag> tv;
    for (size_t i = 0; i < ntags; ++i) {
        if (tags[i].pos >= m) continue;               // produce() requires pos < n (circular_buffer.rs:519-527)
        rr::Tag t;
        t.pos = (size_t)tags[i].pos; t.key = tags[i].key ? tags[i].key : "";
        t.val.kind = (rr::TagKind)tags[i].kind;
        if (tags[i].s) t.val.s = tags[i].s;
        t.val.f = tags[i].f; t.val.b = tags[i].b != 0; t.val.u = tags[i].u; t.val.i = tags[i].i;
        tv.push_back(std::move(t));
    }
    b.produce(m, tv);
    if (written) *written = m;
    return RRC_OK;
}
int rrb_wstream_free(rrb_wstream_t* w, size_t* n) {
    if (!w || !w->s || !n) return fail(RRC_ERR_INVALID, "NULL argument");
    *n = w->s->buffer().free_space();
    return RRC_OK;
}
int rrb_wstream_id(rrb_wstream_t* w, size_t* id) {
    if (!w || !w->s || !id) return fail(RRC_ERR_INVALID, "NULL argument");
    *id = w->s->id();
    return RRC_OK;
}
int rrb_wstream_drop(rrb_wstream_t* w) { delete w; return RRC_OK; }

int rrb_rstream_read(rrb_rstream_t* r, void* out, size_t max, size_t* window_len, size_t* ntags) {
    if (!r || !r->s) return fail(RRC_ERR_INVALID, "stream is NULL");
    rr::Buffer& b = r->s->buffer();
    const char* p; size_t len;
    b.read_window(&p, &len, &r->tags);
    const size_t m = max < len ? max : len;
    if (m && out) {
        if (b.residency() == rr::Residency::Device) {
            RRC_CUDA(cudaSetDevice(b.device()));
            cudaStream_t st = (cudaStream_t)rr::graph_stream(b.device());
            RRC_CUDA(cudaMemcpyAsync(out, p, m * b.elem(), cudaMemcpyDeviceToHost, st));
            RRC_CUDA(cudaStreamSynchronize(st));
        } else {
            memcpy(out, p, m * b.elem());
        }
    }
    if (window_len) *window_len = len;
    if (ntags) *ntags = r->tags.size();
    return RRC_OK;
}
int rrb_rstream_tag(rrb_rstream_t* r, size_t i, rrb_tag_t* t) {
    if (!r || !t || i >= r->tags.size()) return fail(RRC_ERR_INVALID, "tag index out of range");
    const rr::Tag& s = r->tags[i];
    t->pos = s.pos; t->key = s.key.c_str(); t->kind = (int)s.val.kind; t->s = s.val.s.c_str();
    t->f = s.val.f; t->b = s.val.b ? 1 : 0; t->u = s.val.u; t->i = s.val.i;
    return RRC_OK;
}
int rrb_rstream_consume(rrb_rstream_t* r, size_t n) {
    if (!r || !r->s) return fail(RRC_ERR_INVALID, "stream is NULL");
    if (n > r->s->buffer().used()) return fail(RRC_ERR_INVALID, "trying to consume %zu, but only have %zu", n, r->s->buffer().used());
    r->s->buffer().consume(n);
    return RRC_OK;
}
int rrb_rstream_id(rrb_rstream_t* r, size_t* id) {
    if (!r || !r->s || !id) return fail(RRC_ERR_INVALID, "NULL argument");
    *id = r->s->id();
    return RRC_OK;
}
int rrb_rstream_capacity(rrb_rstream_t* r, size_t* n) {
    if (!r || !r->s || !n) return fail(RRC_ERR_INVALID, "NULL argument");
    *n = r->s->buffer().capacity();
    return RRC_OK;
}
int rrb_rstream_eof(rrb_rstream_t* r, int* eof) {
    if (!r || !r->s || !eof) return fail(RRC_ERR_INVALID, "NULL argument");
    *eof = r->s->eof() ? 1 : 0;
    return RRC_OK;
}
int rrb_rstream_drop(rrb_rstream_t* r) { delete r; return RRC_OK; }

int rrb_vector_source_new(const void* data, size_t n, size_t elem_size, uint64_t repeat, size_t bytes, int res, int device,
                          rrb_block_t** blk, rrb_rstream_t** out) {
    if (!blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::VectorSource> b;
    RRC_TRY(rr::VectorSource::create(data, n, elem_size, repeat, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out);
}
static rr::Repeat repeat_of(uint64_t repeat) { return repeat == UINT64_MAX ? rr::Repeat::forever() : rr::Repeat::finite(repeat); }

int rrb_file_source_new(const char* path, size_t elem_size, uint64_t repeat, size_t bytes, int res, int device,
                        rrb_block_t** blk, rrb_rstream_t** out) {
    if (!path || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FileSource> b;
    RRC_TRY(rr::FileSource::create(path, elem_size, repeat_of(repeat), opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out);
}
int rrb_sigmf_source_new(const char* path, size_t elem_size, const char* type_string, double samp_rate, int ignore_type_error,
                         uint64_t repeat, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out,
                         double* sample_rate_out, int* has_sample_rate) {
    if (!path || !type_string || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::SigMFSource> b;
    RRC_TRY(rr::SigMFSource::create(path, elem_size, type_string, samp_rate, ignore_type_error != 0, repeat_of(repeat),
                                    opts(bytes, res, device), &b));
    double r = 0;
    const bool has = b->sample_rate(&r);
    if (sample_rate_out) *sample_rate_out = r;
    if (has_sample_rate) *has_sample_rate = has ? 1 : 0;
    return finish(std::move(b), blk, out);
}
int rrb_fir_filter_new(rrb_rstream_t* src, int cplx, const float* taps, size_t ntaps, size_t deci, int translate,
                       float samp_rate, float freq, unsigned flags, size_t bytes, int res, int device,
                       rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FirFilter> b;
    RRC_TRY(rr::FirFilter::create(src->s, cplx != 0, taps, ntaps, deci, translate != 0, samp_rate, freq, flags,
                                  opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_fft_filter_new(rrb_rstream_t* src, const float* taps, size_t ntaps, size_t bytes, int res, int device,
                       rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FftFilter> b;
    RRC_TRY(rr::FftFilter::create(src->s, taps, ntaps, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_fft_filter_float_new(rrb_rstream_t* src, const float* taps, size_t ntaps, size_t bytes, int res, int device,
                             rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FftFilterFloat> b;
    RRC_TRY(rr::FftFilterFloat::create(src->s, taps, ntaps, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_rational_resampler_new(rrb_rstream_t* src, size_t interp, size_t deci, size_t bytes, int res, int device,
                               rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    // Validate before taking ownership so the caller keeps `src` on Err (the reference returns
    // Err from new(), src/rational_resampler.rs:130-135).
    if (deci == 0) return fail(RRC_ERR_INVALID, "RationalResampler created using deci 0");
    if (interp == 0) return fail(RRC_ERR_INVALID, "RationalResampler created using interp 0");
    std::unique_ptr<rr::RationalResampler> b;
    RRC_TRY(rr::RationalResampler::create(src->s, interp, deci, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_quadrature_demod_new(rrb_rstream_t* src, float gain, size_t bytes, int res, int device,
                             rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::QuadratureDemod> b;
    RRC_TRY(rr::QuadratureDemod::create(src->s, gain, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}

int rrb_fft_stream_new(rrb_rstream_t* src, size_t size, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FftStream> b;
    RRC_TRY(rr::FftStream::create(src->s, size, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_rtlsdr_decode_new(rrb_rstream_t* src, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::RtlSdrDecode> b;
    RRC_TRY(rr::RtlSdrDecode::create(src->s, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}

int rrb_rtlsdr_encode_new(rrb_rstream_t* src, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::RtlSdrEncode> b;
    RRC_TRY(rr::RtlSdrEncode::create(src->s, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_file_sink_new(rrb_rstream_t* src, const char* path, int mode, int flush, int device, rrb_block_t** blk) {
    if (!src || !path || !blk) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::FileSink> b;
    RRC_TRY(rr::FileSink::create(src->s, path, mode, flush != 0, device, &b));
    delete src;                        // its stream has been moved into the block
    auto* h = new rrb_block();
    h->b = std::move(b);
    *blk = h;
    return RRC_OK;
}

int rrb_hilbert_new(rrb_rstream_t* src, size_t ntaps, int window_type, float window_parm, size_t bytes, int res, int device,
                    rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    // validate before taking ownership of `src` (the reference panics in new(), src/hilbert.rs:44-47)
    if (!(ntaps > 1 && (ntaps & 1) == 1)) return fail(RRC_ERR_INVALID, "hilbert filter len must be odd and greater than 1");
    if (window_type < RRC_WINDOW_HAMMING || window_type > RRC_WINDOW_HAMMING_PARM) return fail(RRC_ERR_INVALID, "unknown window type %d", window_type);
    std::unique_ptr<rr::Hilbert> b;
    RRC_TRY(rr::Hilbert::create(src->s, ntaps, window_type, window_parm, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
static int sync_new(rrb_rstream_t* src, rr::SyncMap::Op op, int cplx, float re, float im, size_t bytes, int res, int device,
                    rrb_block_t** blk, rrb_rstream_t** out) {
    if (!src || !blk || !out) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::SyncMap> b;
    RRC_TRY(rr::SyncMap::create(src->s, op, cplx != 0, re, im, opts(bytes, res, device), &b));
    return finish(std::move(b), blk, out, src);
}
int rrb_multiply_const_new(rrb_rstream_t* src, int cplx, float val_re, float val_im, size_t bytes, int res, int device,
                           rrb_block_t** blk, rrb_rstream_t** out) {
    return sync_new(src, rr::SyncMap::Op::MultiplyConst, cplx, val_re, val_im, bytes, res, device, blk, out);
}
int rrb_add_const_new(rrb_rstream_t* src, int cplx, float val_re, float val_im, size_t bytes, int res, int device,
                      rrb_block_t** blk, rrb_rstream_t** out) {
    return sync_new(src, rr::SyncMap::Op::AddConst, cplx, val_re, val_im, bytes, res, device, blk, out);
}
int rrb_complex_to_mag2_new(rrb_rstream_t* src, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out) {
    return sync_new(src, rr::SyncMap::Op::ComplexToMag2, 1, 0.f, 0.f, bytes, res, device, blk, out);
}
int rrb_iq_balance_new(rrb_rstream_t* src, float alpha, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out) {
    return sync_new(src, rr::SyncMap::Op::IqBalance, 1, alpha, 0.f, bytes, res, device, blk, out);
}
int rrb_tee_new(rrb_rstream_t* src, size_t bytes, int res, int device, rrb_block_t** blk, rrb_rstream_t** out1, rrb_rstream_t** out2) {
    if (!src || !blk || !out1 || !out2) return fail(RRC_ERR_INVALID, "NULL argument");
    std::unique_ptr<rr::Tee> b;
    RRC_TRY(rr::Tee::create(src->s, opts(bytes, res, device), &b));
    auto* r2 = new rrb_rstream();
    r2->s = b->take_output2();
    *out2 = r2;
    return finish(std::move(b), blk, out1, src);
}

int rrb_block_work(rrb_block_t* b, int* kind, size_t* stream_id, size_t* need) {
    if (!b || !b->b || !kind) return fail(RRC_ERR_INVALID, "NULL argument");
    rr::BlockRet r;
    RRC_TRY(b->b->work(&r));
    *kind = (int)r.kind;
    if (stream_id) *stream_id = r.stream ? r.stream->id() : 0;
    if (need) *need = r.need;
    return RRC_OK;
}
int rrb_block_eof(rrb_block_t* b, int* eof) {
    if (!b || !b->b || !eof) return fail(RRC_ERR_INVALID, "NULL argument");
    *eof = b->b->eof() ? 1 : 0;
    return RRC_OK;
}
const char* rrb_block_name(rrb_block_t* b) { return (b && b->b) ? b->b->block_name() : ""; }
int rrb_block_drop(rrb_block_t* b) { delete b; return RRC_OK; }

int rrb_graph_run(rrb_block_t** blocks, size_t n) {
    if (!blocks && n) return fail(RRC_ERR_INVALID, "NULL argument");
    rr::Graph g;
    for (size_t i = 0; i < n; ++i) g.add(blocks[i]->b.get());
    return g.run();
}

}  // extern "C"
