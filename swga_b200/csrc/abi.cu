// abi.cu -- context, memory helpers, host-side FASTA reader, and the whole-genome counting
// pipelines (device-resident and host-buffer) of libswga_b200.so.
#include <ctype.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"

int swga_mark_records_origin(swga_ctx *ctx, const uint64_t *d_rec_starts, uint64_t n_rec, uint64_t origin,
                             uint64_t n, uint32_t *d_brk, cudaStream_t stream);

// ---- errors ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void swga_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *swga_last_error(void) { return g_err; }
extern "C" int swga_abi_version(void) { return 2; }

static std::atomic<unsigned long long> g_launches{0};
void swga_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t swga_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- context --------------------------------------------------------------------------------
static int ctx_init_streams(swga_ctx *c)
{
    CUDA_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    return SWGA_OK;
}

extern "C" void swga_ctx_destroy(swga_ctx *c);

extern "C" int swga_ctx_create(int device, swga_ctx **out)
{
    ARG_CHECK(out);
    *out = nullptr;
    int n_dev = 0;
    CUDA_TRY(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) {
        swga_set_error("device %d not available (%d CUDA devices); libswga_b200 has no CPU path", device, n_dev);
        return SWGA_E_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        swga_set_error("device %d is sm_%d%d; libswga_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return SWGA_E_CUDA;
    }
    swga_ctx *c = new swga_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    const int rc = ctx_init_streams(c);
    if (rc != SWGA_OK) {                                       // nothing half-built is handed out or leaked
        swga_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return SWGA_OK;
}

static void free_buf(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" void swga_ctx_destroy(swga_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_buf(c->packed); free_buf(c->brk); free_buf(c->fwd); free_buf(c->canon); free_buf(c->recs);
    free_buf(c->scratch); free_buf(c->sort_hist); free_buf(c->score_tmp); free_buf(c->bucket_data); free_buf(c->bucket_meta);
    for (int i = 0; i < 2; ++i) { free_buf(c->sort_keys[i]); free_buf(c->sort_vals[i]); }
    for (int i = 0; i < 2; ++i) {
        free_buf(c->ascii[i]); free_buf(c->packed2[i]); free_buf(c->brk2[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
    for (int i = 0; i < 2; ++i) {
        if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
        if (c->h_recs[i]) cudaFreeHost(c->h_recs[i]);
    }
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    delete c;
}

extern "C" int swga_ctx_sm_count(const swga_ctx *c) { return c ? c->sm_count : 0; }
extern "C" int swga_ctx_set_count_impl(swga_ctx *c, int impl)
{
    ARG_CHECK(c && impl >= 0 && impl <= 2);
    c->count_impl = impl;
    return SWGA_OK;
}
extern "C" int swga_ctx_set_count_via_pack(swga_ctx *c, int on)
{
    ARG_CHECK(c);
    c->count_via_pack = on ? 1 : 0;
    return SWGA_OK;
}

int swga_ensure(swga_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes && b.p) return SWGA_OK;
    if (b.p) CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        swga_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return SWGA_E_NOMEM;
    }
    b.cap = want;
    return SWGA_OK;
}

// ---- memory helpers -------------------------------------------------------------------------
extern "C" int swga_malloc(swga_ctx *ctx, void **d_ptr, size_t bytes)
{
    ARG_CHECK(ctx && d_ptr);
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        swga_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SWGA_E_NOMEM;
    }
    return SWGA_OK;
}
extern "C" int swga_free(swga_ctx *ctx, void *d_ptr)
{
    ARG_CHECK(ctx);
    CUDA_TRY(cudaFree(d_ptr));
    return SWGA_OK;
}
extern "C" int swga_host_alloc(swga_ctx *ctx, void **h_ptr, size_t bytes)
{
    ARG_CHECK(ctx && h_ptr);
    cudaError_t e = cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        swga_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SWGA_E_NOMEM;
    }
    return SWGA_OK;
}
extern "C" int swga_host_free(swga_ctx *ctx, void *h_ptr)
{
    ARG_CHECK(ctx);
    CUDA_TRY(cudaFreeHost(h_ptr));
    return SWGA_OK;
}
extern "C" int swga_memset(swga_ctx *ctx, void *d_ptr, int value, size_t bytes, void *stream)
{
    ARG_CHECK(ctx && (d_ptr || bytes == 0));
    CUDA_TRY(cudaMemsetAsync(d_ptr, value, bytes, as_stream(stream)));
    return SWGA_OK;
}
extern "C" int swga_h2d(swga_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    ARG_CHECK(ctx && ((d_dst && h_src) || bytes == 0));
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return SWGA_OK;
}
extern "C" int swga_d2h(swga_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    ARG_CHECK(ctx && ((h_dst && d_src) || bytes == 0));
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return SWGA_OK;
}
extern "C" int swga_stream_sync(swga_ctx *ctx, void *stream)
{
    ARG_CHECK(ctx);
    CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    return SWGA_OK;
}

// ---- host FASTA reader ----------------------------------------------------------------------
// Line-oriented formulation of the record rules of ext/dsk/minia/Bank.cpp:150-209:
//   * the image must begin with '>' or '@' (Bank.cpp:284-299: anything else is taken for a list of
//     file names); an empty image has no records
//   * a line whose first byte is '>' or '@' (or 0xFF, which the reference's signed-char reader
//     mistakes for EOF and then resumes from as a header, Bank.cpp:66-72,170) starts a record and
//     is a header line; every other non-empty line is sequence and is appended whole
//   * after appending a line, one trailing '\r' is dropped if the record is then longer than one
//     byte (Bank.cpp:124-125)
//   * '+' at line start switches the reference to its FASTQ branch: such files go to scan_sequential below
static inline bool is_rec_start(uint8_t c) { return c == '>' || c == '@' || c == 0xFF; }

// One contiguous range of LINES of the image, [pos, end) with pos at a line start.  `rec_len` enters as the
// length of the record the range continues (only "0" versus "more" matters, for the '\r' rule).  With `fill`
// the range writes its bytes at h_seq + base and its records from index rec0 on.
struct ScanRange {
    uint64_t pos = 0, end = 0;             // in: byte range
    uint64_t rec_len_in = 0;               // in: 0 = the previous non-empty line is a header (or there is none)
    uint64_t bases = 0, recs = 0;          // out of the counting pass; in for the filling pass as prefix sums
    uint64_t base0 = 0, rec0 = 0;
    int status = SWGA_OK;
    uint64_t bad_pos = 0;
};

static void scan_range(const uint8_t *f, ScanRange &r, uint8_t *h_seq, uint64_t *h_rec_starts, uint64_t *h_hdr,
                       std::vector<uint64_t> *rec_local = nullptr)
{
    const bool fill = h_seq != nullptr || h_rec_starts != nullptr;
    uint64_t total = r.base0, rec = r.rec0, rec_len = r.rec_len_in, pos = r.pos;
    while (pos < r.end) {
        const uint8_t c = f[pos];
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(f + pos, '\n', r.end - pos));
        const uint64_t eol = nl ? (uint64_t)(nl - f) : r.end;  // index of '\n' or the end of the range (= end of file)
        if (is_rec_start(c)) {
            if (rec_local) rec_local->push_back(total - r.base0);  // record start, in bases from the start of the range
            if (fill && h_rec_starts) h_rec_starts[rec] = total;
            if (fill && h_hdr) {
                h_hdr[2 * rec] = pos + 1;
                h_hdr[2 * rec + 1] = eol - (pos + 1);
            }
            // The reference reads the header TOKEN with an isspace() delimiter and only then the rest
            // of the line: when the token's delimiter is itself '\n' the header ends there.  A '\v',
            // '\f' or '\r' delimiter is followed by "rest of line", same as here.
            rec++;
            rec_len = 0;
        } else if (c == '+') {
            r.status = SWGA_E_FORMAT;
            r.bad_pos = pos;
            return;
        } else if (eol > pos) {
            const uint64_t len = eol - pos;
            rec_len += len;
            // one trailing '\r' is dropped -- except for a last line of exactly one byte without a newline: the
            // reference has taken that byte with getc and its gets() returns at end of file before the rule
            // (Bank.cpp:80-81,124-125).  Decided BEFORE the copy: the byte after this line's share of h_seq
            // belongs to the next line, which another thread may be writing.
            const uint64_t drop = (rec_len > 1 && f[eol - 1] == '\r' && !(nl == nullptr && len == 1)) ? 1 : 0;
            if (h_seq) memcpy(h_seq + total, f + pos, len - drop);
            total += len - drop;
            rec_len -= drop;
        }
        pos = eol + 1;
    }
    r.bases = total - r.base0;
    r.recs = rec - r.rec0;
}

// Cuts the image into n_rg ranges of whole lines (cut t = the first line start at or after t * n / n_rg) and finds,
// for each, whether the record it continues is still empty: walk back over the lines before the cut -- empty
// lines do not count, a header line means "empty", any other line means "not empty".
static void cut_ranges(const uint8_t *f, uint64_t n, unsigned n_rg, std::vector<ScanRange> &rg)
{
    rg.assign(n_rg, ScanRange());
    std::vector<uint64_t> cut(n_rg + 1, n);
    cut[0] = 0;
    for (unsigned t = 1; t < n_rg; ++t) {
        uint64_t p = std::max(cut[t - 1], n / n_rg * t);
        if (p > 0 && p < n && f[p - 1] != '\n') {
            const uint8_t *nl = static_cast<const uint8_t *>(memchr(f + p, '\n', n - p));
            p = nl ? (uint64_t)(nl - f) + 1 : n;
        }
        cut[t] = std::min(p, n);
    }
    for (unsigned t = 0; t < n_rg; ++t) {
        rg[t].pos = cut[t];
        rg[t].end = cut[t + 1];
        uint64_t q = cut[t];
        rg[t].rec_len_in = 0;
        while (q > 0) {
            uint64_t e = q - 1;                                // the '\n' that ends the previous line
            uint64_t s0 = e;
            while (s0 > 0 && f[s0 - 1] != '\n') --s0;          // start of that line
            if (e > s0) {                                      // non-empty line
                rg[t].rec_len_in = is_rec_start(f[s0]) ? 0 : 2;
                break;
            }
            q = s0;
        }
    }
}

// ---- sequential reader: FASTQ and mixed files ---------------------------------------------------
// The line-parallel scanner above covers FASTA.  A '+' at a line start switches the reference to its FASTQ branch
// (Bank.cpp:192-201), whose quality lines may begin with '>' or '@' and whose end depends on the length of the
// read AND of the rest of the header line (the reference reads both into the same string): that cannot be cut
// into independent ranges, so such files -- and every file that begins with '@' -- are read by one thread with the
// reference's own byte-level rules.  Outputs may be NULL (sizes only).
namespace {
struct ByteCursor {
    const uint8_t *b;
    uint64_t n, pos;
    int getc() { return pos < n ? (int)(signed char)b[pos++] : -1; }     // 0xFF reads as -1, like buffered_getc
    bool at_end() const { return pos >= n; }
    // buffered_gets(.., allow_spaces = true): the bytes up to the next '\n' (consumed) or the end of the file
    uint64_t line(const uint8_t **start)
    {
        const uint8_t *p = b + pos;
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(p, '\n', n - pos));
        const uint64_t len = nl ? (uint64_t)(nl - p) : n - pos;
        *start = p;
        pos += len + (nl ? 1 : 0);
        return len;
    }
};
// the '\r' rule of buffered_gets on a string that grows by appending (Bank.cpp:124-125)
inline void drop_cr(std::vector<uint8_t> &s)
{
    if (s.size() > 1 && s.back() == '\r') s.pop_back();
}
}   // namespace

static void scan_sequential(const uint8_t *f, uint64_t n, uint8_t *h_seq, uint64_t *h_rec_starts, uint64_t *h_hdr,
                            uint64_t *n_bases, uint64_t *n_rec)
{
    ByteCursor cur{f, n, 0};
    std::vector<uint8_t> dummy;                                // rest of the header line, then the quality lines
    uint64_t total = 0, rec = 0;
    int last_char = 0, c;
    for (;;) {
        if (last_char == 0) {                                  // Bank.cpp:154-160: next header, byte by byte
            while ((c = cur.getc()) != -1 && c != '>' && c != '@') {}
            if (c == -1) break;
            last_char = c;
        }
        if (cur.at_end()) break;                               // nothing after the header mark: no record
        const uint64_t hdr_pos = cur.pos;
        uint64_t q = cur.pos;                                  // header token: up to the first isspace() byte
        while (q < n && !isspace(f[q])) ++q;
        dummy.clear();
        cur.pos = q < n ? q + 1 : n;
        if (q < n && f[q] != '\n' && !cur.at_end()) {
            const uint8_t *st;
            const uint64_t l = cur.line(&st);
            dummy.assign(st, st + l);
            drop_cr(dummy);
        }
        if (h_rec_starts) h_rec_starts[rec] = total;
        if (h_hdr) {
            const uint8_t *nl = static_cast<const uint8_t *>(memchr(f + hdr_pos, '\n', n - hdr_pos));
            h_hdr[2 * rec] = hdr_pos;
            h_hdr[2 * rec + 1] = (nl ? (uint64_t)(nl - f) : n) - hdr_pos;
        }
        uint64_t len = 0;                                      // this record's bases so far
        while ((c = cur.getc()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;                           // empty line
            const bool alone_at_eof = cur.at_end();            // gets() finds the end of the file: no '\r' rule
            --cur.pos;                                         // the byte just taken is the first of its line
            const uint8_t *st;
            const uint64_t l = cur.line(&st);
            // the dropped '\r' is never written: the byte after this record's share of h_seq is not ours
            const uint64_t drop = (!alone_at_eof && len + l > 1 && st[l - 1] == '\r') ? 1 : 0;
            if (h_seq) memcpy(h_seq + total + len, st, l - drop);
            len += l - drop;
        }
        if (c == '>' || c == '@') last_char = c;
        if (c == '+') {                                        // Bank.cpp:192-201
            while ((c = cur.getc()) != -1 && c != '\n') {}     // rest of the quality comment
            while (!cur.at_end()) {                            // quality, appended to the rest of the header line
                const uint8_t *st;
                const uint64_t l = cur.line(&st);
                dummy.insert(dummy.end(), st, st + l);
                drop_cr(dummy);
                if (dummy.size() >= len) break;
            }
            last_char = 0;
        }
        total += len;
        ++rec;
    }
    if (h_rec_starts) h_rec_starts[rec] = total;
    *n_bases = total;
    *n_rec = rec;
}

static bool needs_sequential(const uint8_t *f, uint64_t n) { return n > 0 && f[0] == '@'; }

// Both passes (sizes, then fill) run on up to 16 host threads: the image is cut at line starts into ranges that
// are scanned independently (a line never depends on another one except through "is the record still empty",
// which a range recovers by looking at the lines before its start), then prefix sums place every range.
extern "C" int swga_fasta_scan(const uint8_t *f, uint64_t n, uint8_t *h_seq, uint64_t *h_rec_starts,
                               uint64_t *h_hdr, uint64_t *n_bases, uint64_t *n_rec)
{
    ARG_CHECK((f || n == 0) && n_bases && n_rec);
    if (n > 0 && f[0] != '>' && f[0] != '@') {
        swga_set_error("not a FASTA image: first byte is 0x%02x, expected '>' (DSK would read it as a list of file names)",
                       f[0]);
        return SWGA_E_FORMAT;
    }
    if (needs_sequential(f, n)) {                                // FASTQ (or '@' headers): the byte-level reader
        scan_sequential(f, n, h_seq, h_rec_starts, h_hdr, n_bases, n_rec);
        return SWGA_OK;
    }
    unsigned hw = std::thread::hardware_concurrency();
    const char *env_chunk = getenv("SWGA_FASTA_CHUNK");          // bytes per range at least (tests force small ranges)
    const uint64_t per_thread_min = env_chunk && atoll(env_chunk) > 0 ? (uint64_t)atoll(env_chunk) : (8ull << 20);
    const unsigned n_thr = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(hw ? hw : 1, 16u), n / per_thread_min));
    std::vector<ScanRange> rg;
    cut_ranges(f, n, n_thr, rg);
    auto run = [&](uint8_t *seq, uint64_t *rs, uint64_t *hdr) {
        if (n_thr == 1) {
            scan_range(f, rg[0], seq, rs, hdr);
            return;
        }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < n_thr; ++t) th.emplace_back([&, t] { scan_range(f, rg[t], seq, rs, hdr); });
        for (auto &x : th) x.join();
    };
    // the sizes pass of the caller's first call (h_seq == NULL) is remembered for the filling call that follows
    static thread_local const uint8_t *last_f = nullptr;
    static thread_local uint64_t last_n = 0;
    static thread_local std::vector<ScanRange> last_rg;
    const bool filling = h_seq != nullptr || h_rec_starts != nullptr;
    if (filling && last_f == f && last_n == n && last_rg.size() == rg.size()) {
        rg = last_rg;
    } else {
        run(nullptr, nullptr, nullptr);                        // sizes
    }
    last_f = filling ? nullptr : f;
    last_n = n;
    if (!filling) last_rg = rg;
    uint64_t total = 0, rec = 0;
    for (unsigned t = 0; t < n_thr; ++t) {
        if (rg[t].status != SWGA_OK) {                           // a FASTQ record inside a '>' file
            last_f = nullptr;
            scan_sequential(f, n, h_seq, h_rec_starts, h_hdr, n_bases, n_rec);
            return SWGA_OK;
        }
        rg[t].base0 = total;
        rg[t].rec0 = rec;
        total += rg[t].bases;
        rec += rg[t].recs;
    }
    if (h_seq != nullptr || h_rec_starts != nullptr) {
        run(h_seq, h_rec_starts, h_hdr);                       // fill
        if (h_rec_starts) h_rec_starts[rec] = total;
    }
    *n_bases = total;
    *n_rec = rec;
    return SWGA_OK;
}

extern "C" int swga_dsk_mode(const uint64_t *h_rec_starts, uint64_t n_rec)
{
    if (!h_rec_starts) return 0;
    const uint64_t lim = n_rec < 10001 ? n_rec : 10001;
    uint64_t mx = 0;
    for (uint64_t r = 0; r < lim; ++r) mx = std::max(mx, h_rec_starts[r + 1] - h_rec_starts[r]);
    return mx > 1000000 ? 1 : 0;
}

// ---- whole-genome counting, device-resident input -------------------------------------------
extern "C" int swga_count_genome_device(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, const uint64_t *d_rec_starts,
                                        uint64_t n_rec, int mask_mode, int kmin, int kmax, uint64_t n_starts,
                                        int finalize, uint32_t *d_fwd, uint32_t *d_canon, void *stream)
{
    ARG_CHECK(ctx && (d_ascii || n == 0) && n_starts <= n);
    ARG_CHECK(d_fwd || finalize);
    ARG_CHECK(d_canon || !finalize);
    const uint64_t words = swga_tables_words(kmin, kmax);
    TRY(swga_ensure(ctx, ctx->packed, swga_packed_words(n) * 4));
    TRY(swga_ensure(ctx, ctx->brk, swga_brk_words(n) * 4));
    if (!d_fwd) {
        TRY(swga_ensure(ctx, ctx->fwd, words * 4));
        d_fwd = (uint32_t *)ctx->fwd.p;
    }
    uint32_t *packed = (uint32_t *)ctx->packed.p, *brk = (uint32_t *)ctx->brk.p;
    CUDA_TRY(cudaMemsetAsync(d_fwd, 0, words * 4, as_stream(stream)));
    TRY(swga_count_ascii(ctx, d_ascii, n, n_rec > 1 ? d_rec_starts + 1 : nullptr, n_rec > 1 ? n_rec - 1 : 0, 0, mask_mode,
                         n_starts, kmin, kmax, packed, brk, d_fwd, as_stream(stream)));
    if (finalize) {
        TRY(swga_count_finalize(ctx, kmin, kmax, d_fwd, stream));
        TRY(swga_fold_canonical(ctx, kmin, kmax, d_fwd, d_canon, stream));
    }
    return SWGA_OK;
}

// One chunk of the host pipelines: H2D of `len` bases from pinned memory into staging buffer b (stream s_copy),
// then pack, record marks and the count of the chunk's first `owned` start positions (stream s_comp).  `origin`
// is the genome coordinate of the chunk's first base.
static int enqueue_chunk(swga_ctx *ctx, int b, const uint8_t *src, uint64_t len, uint64_t origin, uint64_t owned,
                         const uint64_t *h_rec_starts, uint64_t n_rec, int mask_mode, int kmin, int kmax, uint32_t *d_fwd)
{
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[b], 0));   // the kernels that last read this buffer
    CUDA_TRY(cudaMemcpyAsync(ctx->ascii[b].p, src, len, cudaMemcpyHostToDevice, ctx->s_copy));
    CUDA_TRY(cudaEventRecord(ctx->ev_h2d[b], ctx->s_copy));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_h2d[b], 0));
    uint32_t *packed = (uint32_t *)ctx->packed2[b].p, *brk = (uint32_t *)ctx->brk2[b].p;
    const uint64_t *marks = nullptr;
    uint64_t n_marks = 0;
    if (n_rec > 1) {
        // records whose start s satisfies origin < s <= origin + len
        const uint64_t *lo = std::upper_bound(h_rec_starts + 1, h_rec_starts + n_rec, origin);
        const uint64_t *hi = std::upper_bound(h_rec_starts + 1, h_rec_starts + n_rec, origin + len);
        if (hi > lo) {
            marks = (const uint64_t *)ctx->recs.p + (lo - h_rec_starts);
            n_marks = (uint64_t)(hi - lo);
        }
    }
    TRY(swga_count_ascii(ctx, (const uint8_t *)ctx->ascii[b].p, len, marks, n_marks, origin, mask_mode, owned, kmin, kmax,
                         packed, brk, d_fwd, ctx->s_comp));
    CUDA_TRY(cudaEventRecord(ctx->ev_done[b], ctx->s_comp));
    return SWGA_OK;
}

static int ensure_stage(swga_ctx *ctx, size_t want)
{
    if (ctx->h_stage_cap >= want) return SWGA_OK;
    CUDA_TRY(cudaStreamSynchronize(ctx->s_copy));              // no copy may still be reading the old blocks
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_stage[i]) CUDA_TRY(cudaFreeHost(ctx->h_stage[i]));
        ctx->h_stage[i] = nullptr;
        ctx->h_stage_cap = 0;
        CUDA_TRY(cudaHostAlloc(&ctx->h_stage[i], want, cudaHostAllocDefault));
    }
    ctx->h_stage_cap = want;
    return SWGA_OK;
}

// ---- whole-genome counting, host buffers (what swga/kmers.py would call instead of `dsk`) -----
// The genome is cut into chunks of CHUNK start positions; chunk i needs bases
// [a_i, min(n, b_i + kmax - 1)).  H2D of chunk i+1 (stream s_copy) overlaps pack + count of chunk i
// (stream s_comp) through two staging buffers.  Accumulates into d_fwd (additive form) on s_comp
// and returns without synchronising s_comp.
extern "C" int swga_count_stream_host(swga_ctx *ctx, const uint8_t *h_seq, uint64_t n, const uint64_t *h_rec_starts,
                                      uint64_t n_rec, int mode_b, int kmin, int kmax, uint64_t n_starts,
                                      uint32_t *d_fwd)
{
    ARG_CHECK(ctx && (h_seq || n == 0) && d_fwd && (h_rec_starts || n_rec == 0) && n_starts <= n);
    if (kmin < 2 || kmax < kmin || kmax > SWGA_KMAX_LIMIT) {
        swga_set_error("need 2 <= kmin <= kmax <= %d", SWGA_KMAX_LIMIT);
        return SWGA_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t CHUNK = 64ull << 20;                        // start positions per chunk (multiple of 32)
    const int mask_mode = mode_b ? SWGA_MASK_NONE : SWGA_MASK_N;
    const uint64_t cap = std::min<uint64_t>(n, CHUNK + 64);
    for (int i = 0; i < 2; ++i) {
        TRY(swga_ensure(ctx, ctx->ascii[i], cap + 64));
        TRY(swga_ensure(ctx, ctx->packed2[i], swga_packed_words(cap) * 4));
        TRY(swga_ensure(ctx, ctx->brk2[i], swga_brk_words(cap) * 4));
    }
    if (n_rec > 0) {
        TRY(swga_ensure(ctx, ctx->recs, (n_rec + 1) * 8));
        CUDA_TRY(cudaMemcpyAsync(ctx->recs.p, h_rec_starts, (n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->s_comp));
    }
    // A pageable source would make every cudaMemcpyAsync a synchronous, driver-staged copy (~9 GB/s measured):
    // such buffers go through two pinned staging blocks of the context, filled by a few host threads while the
    // previous chunk is on the bus.  (Pinning the caller's buffer instead costs ~0.45 s per GB.)
    bool pinned = true;
    {
        cudaPointerAttributes attr;
        const cudaError_t e = cudaPointerGetAttributes(&attr, h_seq);
        if (e != cudaSuccess) cudaGetLastError();
        pinned = (e == cudaSuccess && attr.type == cudaMemoryTypeHost);
    }
    if (!pinned && n > 0) TRY(ensure_stage(ctx, (size_t)cap + 64));
    int it = 0;
    for (uint64_t a = 0; a < n_starts; a += CHUNK, ++it) {
        const int b = it & 1;
        const uint64_t e_starts = std::min(n_starts, a + CHUNK);   // owned starts [a, e_starts)
        const uint64_t e_bases = std::min(n, e_starts + (uint64_t)kmax - 1);
        const uint64_t len = e_bases - a;
        const uint8_t *src = h_seq + a;
        if (!pinned) {
            CUDA_TRY(cudaEventSynchronize(ctx->ev_h2d[b]));     // the copy that last read this block (this call or an earlier one)
            uint8_t *dst = (uint8_t *)ctx->h_stage[b];
            const unsigned n_thr = len >= (8u << 20) ? 4u : 1u;
            if (n_thr == 1) {
                memcpy(dst, src, len);
            } else {
                std::vector<std::thread> th;
                const uint64_t per = (len + n_thr - 1) / n_thr;
                for (unsigned t = 0; t < n_thr; ++t) {
                    const uint64_t lo = std::min<uint64_t>(len, t * per), hi = std::min<uint64_t>(len, lo + per);
                    th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
                }
                for (auto &x : th) x.join();
            }
            src = dst;
        }
        TRY(enqueue_chunk(ctx, b, src, len, a, e_starts - a, h_rec_starts, n_rec, mask_mode, kmin, kmax, d_fwd));
    }
    return SWGA_OK;
}

// Result tables to the caller's buffer at the end of the *_host calls (on s_comp, synchronised on return).  A
// pageable destination (a plain numpy array) would be a driver-staged copy with the page faults of a fresh
// buffer inside it (~3 GB/s measured): it goes through the pinned staging blocks instead, 16 MiB at a time,
// copied out by a few host threads while the next block is on the bus.
static int d2h_result(swga_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, h_dst);
    if (e != cudaSuccess) cudaGetLastError();
    if (e == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->s_comp));
        CUDA_TRY(cudaStreamSynchronize(ctx->s_comp));
        return SWGA_OK;
    }
    const size_t BLK = 16u << 20;
    TRY(ensure_stage(ctx, BLK));
    const size_t n_blk = (bytes + BLK - 1) / BLK;
    auto issue = [&](size_t i) -> int {
        const size_t off = i * BLK, len = std::min(BLK, bytes - off);
        CUDA_TRY(cudaMemcpyAsync(ctx->h_stage[i & 1], (const uint8_t *)d_src + off, len, cudaMemcpyDeviceToHost, ctx->s_comp));
        CUDA_TRY(cudaEventRecord(ctx->ev_done[i & 1], ctx->s_comp));
        return SWGA_OK;
    };
    if (n_blk) TRY(issue(0));
    for (size_t i = 0; i < n_blk; ++i) {
        if (i + 1 < n_blk) TRY(issue(i + 1));
        CUDA_TRY(cudaEventSynchronize(ctx->ev_done[i & 1]));
        const size_t off = i * BLK, len = std::min(BLK, bytes - off);
        const uint8_t *src = (const uint8_t *)ctx->h_stage[i & 1];
        uint8_t *dst = (uint8_t *)h_dst + off;
        const unsigned n_cp = 8;                               // mostly page faults of a fresh destination: they scale with threads
        const size_t per = ((len + n_cp - 1) / n_cp + 4095) & ~size_t(4095);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < n_cp; ++t) {
            const size_t lo = std::min(len, t * per), hi = std::min(len, lo + per);
            if (hi > lo) th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
        }
        for (auto &x : th) x.join();
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->s_comp));
    return SWGA_OK;
}

extern "C" void *swga_ctx_compute_stream(swga_ctx *ctx) { return ctx ? (void *)ctx->s_comp : nullptr; }

extern "C" int swga_count_genome_host(swga_ctx *ctx, const uint8_t *h_seq, uint64_t n, const uint64_t *h_rec_starts,
                                      uint64_t n_rec, int mode_b, int kmin, int kmax, uint32_t *h_canon)
{
    ARG_CHECK(ctx && h_canon);
    if (kmin < 2 || kmax < kmin || kmax > SWGA_KMAX_LIMIT) {
        swga_set_error("need 2 <= kmin <= kmax <= %d", SWGA_KMAX_LIMIT);
        return SWGA_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t words = swga_tables_words(kmin, kmax);
    TRY(swga_ensure(ctx, ctx->fwd, words * 4));
    TRY(swga_ensure(ctx, ctx->canon, words * 4));
    uint32_t *fwd = (uint32_t *)ctx->fwd.p, *canon = (uint32_t *)ctx->canon.p;
    CUDA_TRY(cudaMemsetAsync(fwd, 0, words * 4, ctx->s_comp));
    TRY(swga_count_stream_host(ctx, h_seq, n, h_rec_starts, n_rec, mode_b, kmin, kmax, n, fwd));
    TRY(swga_count_finalize(ctx, kmin, kmax, fwd, ctx->s_comp));
    TRY(swga_fold_canonical(ctx, kmin, kmax, fwd, canon, ctx->s_comp));
    TRY(d2h_result(ctx, h_canon, canon, words * 4));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_copy));
    return SWGA_OK;
}

// ---- whole-genome counting straight from a FASTA file image -----------------------------------------
// What `dsk <genome.fasta> <k>` is to swga/kmers.py:44-47, for all k at once and without a host copy of the
// joined bases.  The image is cut into ranges of whole lines (~4 MiB).  DSK's read mode is decided from the head
// of the file (decide_mode).  Then ONE pass: the ranges of a chunk (<= 64 MiB of file) are parsed by the host
// threads DIRECTLY into a pinned staging block while the previous chunk is on the bus and the one before is being
// counted.  A chunk owns the starts of all its bases but the last kmax-1, which the next chunk finds at the head
// of its device buffer (a device-to-device copy of 11 bytes).
template <class F>
static void parallel_for(unsigned n_items, unsigned n_thr, F fn)
{
    if (n_thr <= 1 || n_items <= 1) {
        for (unsigned i = 0; i < n_items; ++i) fn(i);
        return;
    }
    std::atomic<unsigned> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < std::min(n_thr, n_items); ++t)
        th.emplace_back([&] {
            for (;;) {
                const unsigned i = next.fetch_add(1);
                if (i >= n_items) break;
                fn(i);
            }
        });
    for (auto &x : th) x.join();
}

static uint64_t env_u64(const char *name, uint64_t dflt)
{
    const char *v = getenv(name);
    return v && atoll(v) > 0 ? (uint64_t)atoll(v) : dflt;
}

// DSK's read mode from the head of the file: ranges are sized batch by batch until one of the first 10,001
// records is seen to exceed 1,000,000 bases (mode B), or 10,001 records have ended without one (mode A), or the
// file ends.  For a chromosome-scale genome that is the first batch.
static int decide_mode(const uint8_t *f, std::vector<ScanRange> &rg, unsigned n_thr, int *mode)
{
    const unsigned n_rg = (unsigned)rg.size();
    uint64_t total = 0, n_seen = 0, last_start = 0;
    bool big = false;
    for (unsigned r0 = 0; r0 < n_rg; r0 += n_thr) {
        const unsigned r1 = std::min(n_rg, r0 + n_thr);
        std::vector<std::vector<uint64_t>> loc(r1 - r0);
        parallel_for(r1 - r0, n_thr, [&](unsigned j) { scan_range(f, rg[r0 + j], nullptr, nullptr, nullptr, &loc[j]); });
        for (unsigned j = 0; j < r1 - r0; ++j) {
            if (rg[r0 + j].status != SWGA_OK) {                // a FASTQ record: the caller takes the byte-level reader
                *mode = -2;
                return SWGA_OK;
            }
            for (uint64_t l : loc[j]) {
                const uint64_t s = total + l;
                if (n_seen >= 1 && n_seen <= 10001 && s - last_start > 1000000) big = true;   // record n_seen-1 has ended
                last_start = s;
                ++n_seen;
            }
            total += rg[r0 + j].bases;
        }
        if (n_seen >= 1 && n_seen <= 10001 && total - last_start > 1000000) big = true;       // the record still open
        if (big || n_seen > 10001) break;
    }
    *mode = big ? 1 : 0;
    return SWGA_OK;
}

static int ensure_recs(swga_ctx *ctx, size_t want)
{
    if (ctx->h_recs_cap >= want) return SWGA_OK;
    CUDA_TRY(cudaStreamSynchronize(ctx->s_comp));
    want = std::max<size_t>(want * 2, 1 << 16);
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_recs[i]) CUDA_TRY(cudaFreeHost(ctx->h_recs[i]));
        ctx->h_recs[i] = nullptr;
        ctx->h_recs_cap = 0;
        CUDA_TRY(cudaHostAlloc(&ctx->h_recs[i], want, cudaHostAllocDefault));
    }
    ctx->h_recs_cap = want;
    return SWGA_OK;
}

// FASTQ and mixed files (scan_sequential): read by one host thread into a host buffer, then the host-buffer pipeline.
static int count_fasta_sequential(swga_ctx *ctx, const uint8_t *f, uint64_t n_file, int mode, int kmin, int kmax,
                                  uint32_t *d_fwd, uint64_t *n_bases, uint64_t *n_rec, int *mode_used)
{
    uint64_t nb = 0, nr = 0;
    scan_sequential(f, n_file, nullptr, nullptr, nullptr, &nb, &nr);
    std::vector<uint8_t> seq(nb + 64);
    std::vector<uint64_t> rs(nr + 1);
    scan_sequential(f, n_file, seq.data(), rs.data(), nullptr, &nb, &nr);
    if (mode < 0) mode = swga_dsk_mode(rs.data(), nr);
    if (mode_used) *mode_used = mode;
    if (n_bases) *n_bases = nb;
    if (n_rec) *n_rec = nr;
    TRY(swga_count_stream_host(ctx, seq.data(), nb, rs.data(), nr, mode, kmin, kmax, nb, d_fwd));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_copy));              // `seq` and `rs` go out of scope
    return SWGA_OK;
}

static int count_fasta_stream_impl(swga_ctx *ctx, const uint8_t *f, uint64_t n_file, int mode, int kmin, int kmax,
                                   uint32_t *d_fwd, uint64_t *n_bases, uint64_t *n_rec, int *mode_used, bool *fastq_midway);

extern "C" int swga_count_fasta_stream(swga_ctx *ctx, const uint8_t *f, uint64_t n_file, int mode, int kmin, int kmax,
                                       uint32_t *d_fwd, uint64_t *n_bases, uint64_t *n_rec, int *mode_used)
{
    bool midway = false;
    return count_fasta_stream_impl(ctx, f, n_file, mode, kmin, kmax, d_fwd, n_bases, n_rec, mode_used, &midway);
}

static int count_fasta_stream_impl(swga_ctx *ctx, const uint8_t *f, uint64_t n_file, int mode, int kmin, int kmax,
                                   uint32_t *d_fwd, uint64_t *n_bases, uint64_t *n_rec, int *mode_used, bool *fastq_midway)
{
    ARG_CHECK(ctx && (f || n_file == 0) && d_fwd && mode >= -1 && mode <= 1);
    if (kmin < 2 || kmax < kmin || kmax > SWGA_KMAX_LIMIT) {
        swga_set_error("need 2 <= kmin <= kmax <= %d", SWGA_KMAX_LIMIT);
        return SWGA_E_ARG;
    }
    if (n_file > 0 && f[0] != '>' && f[0] != '@') {
        swga_set_error("not a FASTA image: first byte is 0x%02x, expected '>' (DSK would read it as a list of file names)",
                       f[0]);
        return SWGA_E_FORMAT;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned n_thr = std::min<unsigned>(hw ? hw : 1, 16u);
    const uint64_t range_bytes = env_u64("SWGA_FASTA_CHUNK", 4ull << 20);       // tests force small ranges and chunks
    const uint64_t CHUNK = std::max<uint64_t>(64, env_u64("SWGA_STREAM_CHUNK", 64ull << 20));   // file bytes per chunk
    const unsigned n_rg = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_file / range_bytes, 1u << 20));
    const bool trace = getenv("SWGA_TRACE") != nullptr;       // host-side stage times on stderr
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_fill = 0, t_wait = 0, t_enq = 0;
    const double t0 = now();
    std::vector<ScanRange> rg;
    if (needs_sequential(f, n_file))
        return count_fasta_sequential(ctx, f, n_file, mode, kmin, kmax, d_fwd, n_bases, n_rec, mode_used);
    cut_ranges(f, n_file, n_rg, rg);
    if (mode < 0) {
        TRY(decide_mode(f, rg, n_thr, &mode));
        if (mode == -2)                                        // a FASTQ record near the head of a '>' file
            return count_fasta_sequential(ctx, f, n_file, -1, kmin, kmax, d_fwd, n_bases, n_rec, mode_used);
    }
    const double t1 = now();
    if (mode_used) *mode_used = mode;
    const int mask_mode = mode ? SWGA_MASK_NONE : SWGA_MASK_N;
    const uint64_t halo = (uint64_t)kmax - 1;
    // chunks = runs of whole ranges of at most CHUNK file bytes (at least one range); the largest one sizes the blocks
    uint64_t max_ext = 0;
    for (unsigned r = 0; r < n_rg;) {
        unsigned r1 = r + 1;
        while (r1 < n_rg && rg[r1].end - rg[r].pos <= CHUNK) ++r1;
        max_ext = std::max(max_ext, rg[r1 - 1].end - rg[r].pos);
        r = r1;
    }
    if (max_ext > (2ull << 30)) {
        swga_set_error("a line of the FASTA image is longer than 2 GiB: read it with swga_fasta_scan + swga_count_genome_host");
        return SWGA_E_FORMAT;
    }
    const uint64_t cap = max_ext + halo;
    for (int i = 0; i < 2; ++i) {
        TRY(swga_ensure(ctx, ctx->ascii[i], cap + 64));
        TRY(swga_ensure(ctx, ctx->packed2[i], swga_packed_words(cap) * 4));
        TRY(swga_ensure(ctx, ctx->brk2[i], swga_brk_words(cap) * 4));
    }
    TRY(ensure_stage(ctx, (size_t)max_ext + 64));
    // ONE pass over the image: the ranges of a chunk are parsed by the host threads, each into the slot of the
    // pinned block that mirrors its place in the file (a range never yields more bases than it has bytes), and go
    // to the device with one copy per range, packed together behind the last kmax-1 bases of the previous chunk.
    std::vector<std::vector<uint64_t>> rec_local(n_thr * 64);
    std::vector<uint64_t> rs;                                  // record starts so far (genome coordinates)
    uint64_t origin = 0, carry = 0, prev_len = 0, total = 0;   // origin: genome coordinate of the block's first base
    int it = 0, prev_b = 0;
    for (unsigned r = 0; r < n_rg; ++it) {
        const int b = it & 1;
        unsigned r1 = r + 1;
        while (r1 < n_rg && rg[r1].end - rg[r].pos <= CHUNK) ++r1;
        const unsigned cnt = r1 - r;
        if (rec_local.size() < cnt) rec_local.resize(cnt);
        const double ta = now();
        CUDA_TRY(cudaEventSynchronize(ctx->ev_h2d[b]));         // the copies that last read this block
        const double tb = now();
        uint8_t *blk = (uint8_t *)ctx->h_stage[b];
        const uint64_t f0 = rg[r].pos;
        parallel_for(cnt, n_thr, [&](unsigned j) {
            ScanRange &x = rg[r + j];
            x.base0 = x.pos - f0;                              // slot of this range inside the block
            x.rec0 = 0;
            rec_local[j].clear();
            scan_range(f, x, blk, nullptr, nullptr, &rec_local[j]);
        });
        const double tc = now();
        // device layout of the chunk: [carry][range r][range r + 1] ...
        uint8_t *dst = (uint8_t *)ctx->ascii[b].p;
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[b], 0));   // the kernels that last read this buffer
        if (carry)
            CUDA_TRY(cudaMemcpyAsync(dst, (const uint8_t *)ctx->ascii[prev_b].p + (prev_len - carry), carry,
                                     cudaMemcpyDeviceToDevice, ctx->s_copy));
        uint64_t fresh = 0;
        for (unsigned j = 0; j < cnt; ++j) {
            const ScanRange &x = rg[r + j];
            if (x.status != SWGA_OK) {
                // a FASTQ record ('+' at a line start) deep inside a '>' file: chunks before it are already being
                // accumulated into d_fwd.  swga_count_fasta_host starts over with the byte-level reader; a caller
                // of the accumulating call has to do the same on a fresh block.
                *fastq_midway = true;
                swga_set_error("FASTQ record ('+' at the start of a line, byte %llu) inside a FASTA image: d_fwd is partial; "
                               "zero it and use swga_fasta_scan + swga_count_stream_host, or swga_count_fasta_host",
                               (unsigned long long)x.bad_pos);
                return SWGA_E_FORMAT;
            }
            for (uint64_t l : rec_local[j]) rs.push_back(total + fresh + l);
            if (x.bases)
                CUDA_TRY(cudaMemcpyAsync(dst + carry + fresh, blk + x.base0, x.bases, cudaMemcpyHostToDevice, ctx->s_copy));
            fresh += x.bases;
        }
        CUDA_TRY(cudaEventRecord(ctx->ev_h2d[b], ctx->s_copy));
        total += fresh;
        const uint64_t len = carry + fresh;
        const bool last = r1 == n_rg;
        const uint64_t owned = last ? len : (len > halo ? len - halo : 0);
        if (owned > 0) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_h2d[b], 0));
            uint32_t *packed = (uint32_t *)ctx->packed2[b].p, *brk = (uint32_t *)ctx->brk2[b].p;
            // record boundaries inside the chunk: starts s with origin < s <= origin + len (the first record never marks)
            const uint64_t *lo = std::upper_bound(rs.data(), rs.data() + rs.size(), origin);
            const uint64_t *hi = std::upper_bound(rs.data(), rs.data() + rs.size(), origin + len);
            if (hi > lo) {
                const size_t bytes = (size_t)(hi - lo) * 8;
                TRY(ensure_recs(ctx, bytes));
                CUDA_TRY(cudaEventSynchronize(ctx->ev_done[b]));   // the copy that last read this pinned list
                memcpy(ctx->h_recs[b], lo, bytes);
                TRY(swga_ensure(ctx, ctx->recs, bytes));
                CUDA_TRY(cudaMemcpyAsync(ctx->recs.p, ctx->h_recs[b], bytes, cudaMemcpyHostToDevice, ctx->s_comp));
            }
            TRY(swga_count_ascii(ctx, dst, len, hi > lo ? (const uint64_t *)ctx->recs.p : nullptr, (uint64_t)(hi - lo), origin,
                                 mask_mode, owned, kmin, kmax, packed, brk, d_fwd, ctx->s_comp));
            CUDA_TRY(cudaEventRecord(ctx->ev_done[b], ctx->s_comp));
        }
        t_wait += tb - ta;
        t_fill += tc - tb;
        t_enq += now() - tc;
        origin += owned;
        carry = len - owned;
        prev_len = len;
        prev_b = b;
        r = r1;
    }
    if (n_bases) *n_bases = total;
    if (n_rec) *n_rec = rs.size();
    if (trace)
        fprintf(stderr, "[swga_count_fasta_stream] %u ranges, %d chunks: mode %.1f ms, setup %.1f ms, parse %.1f ms, "
                        "waiting for blocks %.1f ms, enqueue %.1f ms\n", n_rg, it, (t1 - t0) * 1e3,
                (now() - t1) * 1e3 - (t_fill + t_wait + t_enq) * 1e3, t_fill * 1e3, t_wait * 1e3, t_enq * 1e3);
    return SWGA_OK;
}

extern "C" int swga_count_fasta_host(swga_ctx *ctx, const uint8_t *f, uint64_t n_file, int mode, int kmin, int kmax,
                                     uint32_t *h_canon, uint64_t *n_bases, uint64_t *n_rec, int *mode_used)
{
    ARG_CHECK(ctx && h_canon);
    if (kmin < 2 || kmax < kmin || kmax > SWGA_KMAX_LIMIT) {
        swga_set_error("need 2 <= kmin <= kmax <= %d", SWGA_KMAX_LIMIT);
        return SWGA_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t words = swga_tables_words(kmin, kmax);
    TRY(swga_ensure(ctx, ctx->fwd, words * 4));
    TRY(swga_ensure(ctx, ctx->canon, words * 4));
    uint32_t *fwd = (uint32_t *)ctx->fwd.p, *canon = (uint32_t *)ctx->canon.p;
    CUDA_TRY(cudaMemsetAsync(fwd, 0, words * 4, ctx->s_comp));
    const bool trace = getenv("SWGA_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    bool midway = false;
    int rc = count_fasta_stream_impl(ctx, f, n_file, mode, kmin, kmax, fwd, n_bases, n_rec, mode_used, &midway);
    if (rc == SWGA_E_FORMAT && midway) {                       // mixed FASTA / FASTQ image: start over, byte-level reader
        CUDA_TRY(cudaStreamSynchronize(ctx->s_copy));
        CUDA_TRY(cudaMemsetAsync(fwd, 0, words * 4, ctx->s_comp));
        rc = count_fasta_sequential(ctx, f, n_file, mode, kmin, kmax, fwd, n_bases, n_rec, mode_used);
    }
    TRY(rc);
    TRY(swga_count_finalize(ctx, kmin, kmax, fwd, ctx->s_comp));
    TRY(swga_fold_canonical(ctx, kmin, kmax, fwd, canon, ctx->s_comp));
    const double t1 = now();
    if (trace) CUDA_TRY(cudaStreamSynchronize(ctx->s_comp));
    const double t2 = now();
    TRY(d2h_result(ctx, h_canon, canon, words * 4));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_copy));
    if (trace)
        fprintf(stderr, "[swga_count_fasta_host] stream %.1f ms, device tail %.1f ms, result copy %.1f ms\n", (t1 - t0) * 1e3,
                (t2 - t1) * 1e3, (now() - t2) * 1e3);
    return SWGA_OK;
}

// ---- DSK result file ------------------------------------------------------------------------
extern "C" int swga_write_dsk_file(const char *path, int k, const uint32_t *h_canon_k, uint32_t threshold,
                                   uint64_t *n_written)
{
    ARG_CHECK(path && h_canon_k && k >= 1 && k <= SWGA_KMAX_LIMIT);
    FILE *fp = fopen(path, "wb");
    if (!fp) {
        swga_set_error("cannot open %s for writing", path);
        return SWGA_E_ARG;
    }
    const int32_t hdr[2] = {64, k};
    fwrite(hdr, 4, 2, fp);
    const uint64_t n = 1ull << (2 * k);
    const uint32_t thr = threshold ? threshold : 1;
    std::vector<uint8_t> buf;
    buf.reserve(12 * 65536);
    uint64_t cnt = 0;
    for (uint64_t c = 0; c < n; ++c) {
        const uint32_t v = h_canon_k[c];
        if (v < thr || v > 2147483646u) continue;              // max_couv, ext/dsk/minia/Utils.cpp:6
        uint8_t rec[12];
        memcpy(rec, &c, 8);
        memcpy(rec + 8, &v, 4);
        buf.insert(buf.end(), rec, rec + 12);
        ++cnt;
        if (buf.size() >= 12 * 65536) {
            fwrite(buf.data(), 1, buf.size(), fp);
            buf.clear();
        }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), fp);
    if (fclose(fp) != 0) {
        swga_set_error("write to %s failed", path);
        return SWGA_E_ARG;
    }
    if (n_written) *n_written = cnt;
    return SWGA_OK;
}

// ---- set_finder output, a block at a time ------------------------------------------------------
// What score.read_set_finder_line (swga/score.py:16-24) does per line -- `ids, w = line.strip('\n').split(' ');
// [int(i) for i in ids.split(',')]; float(w)` -- for a whole block of stdout: the per-line Python of
// FindSets.process_lines (commands/find_sets.py:53-63) is what limits the reference's candidate rate.
// A line that Python would reject (not exactly two fields, an empty or non-integer id, a bad float) gets
// h_ok = 0 and, like there, is skipped by the caller; so does a set with more than max_m ids.
extern "C" int swga_parse_set_lines(const char *text, uint64_t n_bytes, uint32_t max_m, uint64_t cap,
                                    int32_t *h_ids, double *h_weight, uint8_t *h_ok, uint64_t *n_lines,
                                    uint64_t *consumed)
{
    ARG_CHECK((text || n_bytes == 0) && h_ids && h_weight && h_ok && n_lines && consumed && max_m > 0);
    uint64_t pos = 0, line = 0;
    while (pos < n_bytes && line < cap) {
        const char *nl = static_cast<const char *>(memchr(text + pos, '\n', n_bytes - pos));
        if (!nl) break;                                        // incomplete last line: left for the next block
        const uint64_t end = (uint64_t)(nl - text);
        int32_t *ids = h_ids + line * max_m;
        for (uint32_t j = 0; j < max_m; ++j) ids[j] = -1;
        bool ok = true;
        uint64_t p = pos;
        uint32_t m = 0;
        // ids: decimal integers separated by ','; Python's int() also takes a sign and surrounding blanks, which
        // set_finder never prints: anything but [+-]digits is a parse failure here as it is a ValueError there
        for (;;) {
            uint64_t q = p;
            bool neg = false;
            if (q < end && (text[q] == '+' || text[q] == '-')) { neg = text[q] == '-'; ++q; }
            if (q >= end || text[q] < '0' || text[q] > '9') { ok = false; break; }
            long long v = 0;
            while (q < end && text[q] >= '0' && text[q] <= '9') {
                v = v * 10 + (text[q] - '0');
                if (v > 2147483647ll) { ok = false; break; }
                ++q;
            }
            if (!ok) break;
            if (neg && v != 0) { ok = false; break; }              // no vertex has a negative id (-1 is the padding)
            if (m < max_m) ids[m] = (int32_t)v;
            ++m;
            p = q;
            if (p < end && text[p] == ',') { ++p; continue; }
            break;
        }
        if (ok && (p >= end || text[p] != ' ')) ok = false;    // exactly one ' ' between the two fields
        double w = 0.0;
        if (ok) {
            ++p;
            char tmp[64];
            const uint64_t len = end - p;
            if (len == 0 || len >= sizeof(tmp) || memchr(text + p, ' ', len)) {
                ok = false;
            } else {
                memcpy(tmp, text + p, len);
                tmp[len] = 0;
                char *e = nullptr;
                w = strtod(tmp, &e);
                if (e == tmp || *e != 0) ok = false;
            }
        }
        if (m > max_m) ok = false;
        h_weight[line] = w;
        h_ok[line] = ok ? 1 : 0;
        ++line;
        pos = end + 1;
    }
    *n_lines = line;
    *consumed = pos;
    return SWGA_OK;
}
