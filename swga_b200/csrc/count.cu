// count.cu -- kernels (a) 2-bit pack and (b) all-k counting, plus the finalize / fold passes.
//
// Results reproduced (reference = eclarke/swga, paths relative to its tree):
//   pack        : NT2int / code4NT / BinaryReads::write_read   ext/dsk/minia/Kmer.cpp:18-43,
//                 Bank.cpp:832-885; N-split SortingCount.cpp:243-261
//   accumulate  : compute_kmer_table_from_one_seq / KmersBuffer::readkmers + OAHash::increment
//                 Bank.cpp:889-902,1042-1070, OAHash.cpp:142-154
//   fold        : canonical = min(fwd, revcomp)                 Kmer.cpp:122,172-186, lut.h:8-13
//
// Data layout (see include/swga_b200.h): packed = 16 bases / uint32, first base in bits 31:30, so
// the 16 bases starting at any position are one funnel shift of two adjacent words and a k-mer code
// is its top 2k bits.  brk = 1 bit / base, first base in bit 31; bit i set <=> no window may hold
// both base i and base i+1.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// (a) pack: one thread = 32 bases = two 128-bit loads -> one uint2 packed store + one brk word.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ ascii, uint64_t n, int mask_mode,
       uint32_t *__restrict__ packed, uint32_t *__restrict__ brk, uint64_t n_groups)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_groups) return;
    const uint64_t base = t * 32;
    uint32_t w[8];
    uint32_t next = 0x41;                      // byte after this group ('A' when past the end)
    if (base + 32 <= n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ascii + base);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        if (base + 32 < n) next = ascii[base + 32];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t p = base + 4 * i + j;
                uint32_t c = (p < n) ? ascii[p] : 0x41u;      // pad with 'A' (code 0, never flagged)
                v |= c << (8 * j);
            }
            w[i] = v;
        }
    }
    uint32_t p0 = 0, p1 = 0, bad = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) p0 |= codes4(w[i]) << (24 - 8 * i);
#pragma unroll
    for (int i = 4; i < 8; ++i) p1 |= codes4(w[i]) << (24 - 8 * (i - 4));
    if (mask_mode != SWGA_MASK_NONE) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bad |= flags4(bad4(w[i], mask_mode)) << (28 - 4 * i);
        uint32_t nb = bad4(next, mask_mode) & 1u;             // flag of the first base of the next group
        // a bad base i forbids (i-1,i) and (i,i+1): brk[j] = bad[j] | bad[j+1]
        bad = bad | (bad << 1) | nb;
    }
    // end-of-buffer break: nothing follows base n-1
    if (n > 0 && (n - 1) / 32 == t) bad |= 1u << (31 - (uint32_t)((n - 1) & 31));
    reinterpret_cast<uint2 *>(packed)[t] = make_uint2(p0, p1);
    brk[t] = bad;
}

__global__ void k_mark_records(const uint64_t *__restrict__ rec_starts, uint64_t n_rec, uint64_t origin,
                               uint64_t n, uint32_t *__restrict__ brk)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t s = rec_starts[r];
    if (s <= origin) return;
    const uint64_t l = s - origin;                             // local start of record r
    if (l > n) return;
    const uint64_t b = l - 1;                                  // last base of the previous record
    atomicOr(&brk[b >> 5], 1u << (31 - (uint32_t)(b & 31)));
}

// ------------------------------------------------------------------------------------------------
// (b) accumulate: one thread = 32 start positions.  For start p with v = min(kmax, break-free run
// from p): one red.global.add on T_v[top 2v bits].  Almost every thread takes the mask-free fast
// path: 32 unconditional reds into T_kmax.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_count(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk, uint64_t n_starts,
        int kmin, int kmax, uint32_t *__restrict__ tables, TableOffs offs)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t p0 = t * 32;
    if (p0 >= n_starts) return;
    const uint2 w01 = __ldg(reinterpret_cast<const uint2 *>(packed) + t);
    const uint32_t w0 = w01.x, w1 = w01.y;
    const uint32_t w2 = __ldg(packed + 2 * t + 2);
    const uint32_t b0 = __ldg(brk + t), b1 = __ldg(brk + t + 1);
    const int need = kmax - 2;                                 // brk bits of the next word a full group needs
    const uint32_t tail = need > 0 ? (b1 >> (32 - need)) : 0u;
    if (p0 + 32 <= n_starts && (b0 | tail) == 0u) {
        uint32_t *__restrict__ T = tables + offs.off[kmax];
        const int sh = 32 - 2 * kmax;
#pragma unroll
        for (int o = 0; o < 16; ++o) atomicAdd(T + (__funnelshift_l(w1, w0, 2 * o) >> sh), 1u);
#pragma unroll
        for (int o = 0; o < 16; ++o) atomicAdd(T + (__funnelshift_l(w2, w1, 2 * o) >> sh), 1u);
        return;
    }
    const int lim = (int)min((uint64_t)32, n_starts - p0);
    for (int o = 0; o < lim; ++o) {
        const uint32_t W = o < 16 ? __funnelshift_l(w1, w0, 2 * o) : __funnelshift_l(w2, w1, 2 * (o - 16));
        const uint32_t B = __funnelshift_l(b1, b0, o);         // bit 31 = brk[p], bit 30 = brk[p+1], ...
        const int v = min(kmax, __clz(B) + 1);
        if (v >= kmin) atomicAdd(tables + offs.off[v] + (W >> (32 - 2 * v)), 1u);
    }
}

// finalize, one level: T_k[x] += T_{k+1}[4x .. 4x+3]
__global__ void __launch_bounds__(256)
k_marginalize(const uint4 *__restrict__ upper, uint32_t *__restrict__ lower, uint64_t n_lower)
{
    const uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (x >= n_lower) return;
    const uint4 v = upper[x];
    lower[x] += v.x + v.y + v.z + v.w;
}

// reverse complement of a k-mer code in DSK's encoding (complement = digit ^ 2)
__device__ __forceinline__ uint32_t rc_code(uint32_t x, int k)
{
    uint32_t r = __brev(x);                                    // reverses bits, and bits inside each digit
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);   // restore bit order inside each digit
    r >>= (32 - 2 * k);
    return r ^ (0xAAAAAAAAu >> (32 - 2 * k));
}

__device__ __forceinline__ void fold_small(const uint32_t *__restrict__ fwd, uint32_t *__restrict__ canon, int kmin, int kmax,
                                           const TableOffs &offs, uint64_t total, uint32_t block)
{
    const uint64_t i = block * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int k = kmin;
    while (k < kmax && i >= offs.off[k + 1]) ++k;
    const uint32_t x = (uint32_t)(i - offs.off[k]);
    const uint32_t r = rc_code(x, k);
    uint32_t c = 0;
    if (x < r) c = fwd[i] + fwd[offs.off[k] + r];
    else if (x == r) c = fwd[i];
    canon[i] = c;
}

// The same for ONE large table (k >= 7), as a tiled transpose: a code is [A: 3 digits][M: k-6 digits][B: 3 digits]
// and its reverse complement [rc(B)][rc(M)][rc(A)], so the block of middle value M (M <= rc(M)) loads the two
// 64 x 64 tiles (.., M, ..) and (.., rc(M), ..) with coalesced 256-byte rows and finds every partner in shared
// memory.  (fold_small reads the partner of every code with a stride of 4^(k-1) words: one sector per lane.)
__device__ __forceinline__ void fold_tile(const uint32_t *__restrict__ fwd, uint32_t *__restrict__ canon, int k, uint32_t M)
{
    __shared__ uint32_t t1[64][65], t2[64][65];
    __shared__ uint8_t rc3[64];                                // reverse complement of a 3-digit code
    const int md = k - 6, sh = 2 * (k - 3);
    const uint32_t rM = md ? rc_code(M, md) : 0u;
    if (M > rM) return;                                        // handled by the block of rc(M)
    const bool self = M == rM;
    if (threadIdx.x < 64) rc3[threadIdx.x] = (uint8_t)rc_code(threadIdx.x, 3);
    const uint32_t B = threadIdx.x & 63, A0 = threadIdx.x >> 6;   // this thread: column B, rows A0, A0 + 4, ...
    const uint32_t *src1 = fwd + ((M << 6) | B), *src2 = fwd + ((rM << 6) | B);
#pragma unroll 4
    for (uint32_t A = A0; A < 64; A += 4) {
        t1[A][B] = src1[(uint64_t)A << sh];
        if (!self) t2[A][B] = src2[(uint64_t)A << sh];
    }
    __syncthreads();
    const uint32_t rB = rc3[B];
    const uint32_t x1 = (M << 6) | B, x2 = (rM << 6) | B;      // codes without their A digits
    const uint32_t r1 = (rB << sh) | (rM << 6), r2 = (rB << sh) | (M << 6);   // partners without their last 3 digits
#pragma unroll 4
    for (uint32_t A = A0; A < 64; A += 4) {
        const uint32_t rA = rc3[A];
        {
            const uint32_t x = (A << sh) | x1, r = r1 | rA;
            const uint32_t v = t1[A][B], pv = self ? t1[rB][rA] : t2[rB][rA];
            canon[x] = x < r ? v + pv : (x == r ? v : 0u);
        }
        if (!self) {
            const uint32_t x = (A << sh) | x2, r = r2 | rA;
            const uint32_t v = t2[A][B], pv = t1[rB][rA];
            canon[x] = x < r ? v + pv : (x == r ? v : 0u);
        }
    }
}

// The whole fold in ONE launch (four before: k_fold + one k_fold_tiled per large table, ~50 us per genome of which the
// three small launches were mostly ramp and tail): the blocks of the largest table come first, then the tables below it
// down to k_split, then the one-thread-per-code blocks of the tables below k_split.
__global__ void __launch_bounds__(256)
k_fold_all(const uint32_t *__restrict__ fwd, uint32_t *__restrict__ canon, int kmin, int kmax, int k_split, TableOffs offs,
           uint64_t total_small)
{
    uint32_t b = blockIdx.x;
    for (int k = kmax; k >= k_split; --k) {
        const uint32_t nb = 1u << (2 * (k - 6));
        if (b < nb) {
            fold_tile(fwd + offs.off[k], canon + offs.off[k], k, b);
            return;
        }
        b -= nb;
    }
    fold_small(fwd, canon, kmin, min(kmax, k_split - 1), offs, total_small, b);
}

// ------------------------------------------------------------------------------------------------
// synthetic genome generator, bit-identical to swga_b200/synth.py
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256)
k_synth(uint64_t seed, uint64_t start, uint64_t n, uint32_t ta, uint32_t tc, uint32_t tg, uint8_t *__restrict__ out)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t base = t * 16;
    if (base >= n) return;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t r = (uint32_t)(splitmix64((seed << 40) + start + base + j) >> 40);
        const uint32_t c = r < ta ? 'A' : r < tc ? 'C' : r < tg ? 'G' : 'T';
        w[j >> 2] |= c << (8 * (j & 3));
    }
    if (base + 16 <= n) {
        reinterpret_cast<uint4 *>(out)[t] = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int j = 0; base + j < n; ++j) out[base + j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    }
}

// Candidate sets of workload C4, bit-identical to swga_b200/workloads.py c4_sets: set s has
// m = 6 + h(base) mod 5 members, base = (seed << 40) + 64 s: the first m distinct values of h(base + t) mod P,
// t = 1, 2, ... (t < 64); rows of 10 int32, -1 padded.
__global__ void __launch_bounds__(256)
k_synth_sets(uint64_t seed, uint64_t start, uint64_t n_sets, uint32_t n_primers, int32_t *__restrict__ out)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_sets) return;
    const uint64_t base = (seed << 40) + 64ull * (start + i);
    const int m = 6 + (int)(splitmix64(base) % 5ull);
    int32_t row[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) row[j] = -1;
    int filled = 0;
    for (uint64_t t = 1; t < 64 && filled < m; ++t) {
        const int32_t c = (int32_t)(splitmix64(base + t) % (uint64_t)n_primers);
        bool dup = false;
#pragma unroll
        for (int j = 0; j < 10; ++j) dup |= row[j] == c;
        if (!dup) {
#pragma unroll
            for (int j = 0; j < 10; ++j)
                if (j == filled) row[j] = c;
            ++filled;
        }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) out[i * 10 + j] = row[j];
}

// ------------------------------------------------------------------------------------------------
// atomic-throughput microbenchmarks (A_peak)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bench_red_global(uint32_t *__restrict__ table, uint32_t mask, uint64_t ops_per_thread)
{
    uint64_t x = splitmix64(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    for (uint64_t i = 0; i < ops_per_thread; i += 4) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t a = (uint32_t)(x >> 33);
        atomicAdd(table + (a & mask), 1u);
        atomicAdd(table + ((a * 2654435761u) & mask), 1u);
        atomicAdd(table + ((a * 2246822519u + 12345u) & mask), 1u);
        atomicAdd(table + ((a * 3266489917u + 777u) & mask), 1u);
    }
}
__global__ void __launch_bounds__(256)
k_bench_atom_shared(uint32_t mask, uint64_t ops_per_thread, uint32_t *__restrict__ sink)
{
    extern __shared__ uint32_t sh[];
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint64_t x = splitmix64(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    for (uint64_t i = 0; i < ops_per_thread; i += 4) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t a = (uint32_t)(x >> 33);
        atomicAdd(sh + (a & mask), 1u);
        atomicAdd(sh + ((a * 2654435761u) & mask), 1u);
        atomicAdd(sh + ((a * 2246822519u + 12345u) & mask), 1u);
        atomicAdd(sh + ((a * 3266489917u + 777u) & mask), 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(sink, sh[0]);
}



#define PT_AUTO_MIN_STARTS (14ull << 20)                    // auto mode: partitioned path from here on

int swga_mark_records_origin(swga_ctx *ctx, const uint64_t *d_rec_starts, uint64_t n_rec, uint64_t origin,
                             uint64_t n, uint32_t *d_brk, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// C ABI wrappers
// ------------------------------------------------------------------------------------------------
TableOffs swga_make_offs(int kmin, int kmax)
{
    TableOffs o;
    memset(&o, 0, sizeof(o));
    unsigned long long acc = 0;
    for (int k = kmin; k <= kmax + 1 && k <= SWGA_KMAX_LIMIT + 1; ++k) {
        o.off[k] = acc;
        acc += 1ull << (2 * k);
    }
    return o;
}

static int check_k(int kmin, int kmax)
{
    if (kmin < 2 || kmax < kmin || kmax > SWGA_KMAX_LIMIT) {
        swga_set_error("need 2 <= kmin <= kmax <= %d, got kmin=%d kmax=%d", SWGA_KMAX_LIMIT, kmin, kmax);
        return SWGA_E_ARG;
    }
    return SWGA_OK;
}

extern "C" uint64_t swga_tables_words(int kmin, int kmax)
{
    uint64_t acc = 0;
    for (int k = kmin; k <= kmax; ++k) acc += 1ull << (2 * k);
    return acc;
}
extern "C" uint64_t swga_table_offset(int kmin, int k)
{
    uint64_t acc = 0;
    for (int j = kmin; j < k; ++j) acc += 1ull << (2 * j);
    return acc;
}
extern "C" uint64_t swga_packed_words(uint64_t n) { return ((n + 31) / 32 + 1) * 2; }
extern "C" uint64_t swga_brk_words(uint64_t n) { return (n + 31) / 32 + 1; }

extern "C" int swga_pack(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, int mask_mode, uint32_t *d_packed,
                         uint32_t *d_brk, void *stream)
{
    ARG_CHECK(ctx && d_packed && d_brk && (d_ascii || n == 0));
    ARG_CHECK(mask_mode >= SWGA_MASK_NONE && mask_mode <= SWGA_MASK_STRICT);
    ARG_CHECK((reinterpret_cast<uintptr_t>(d_ascii) & 15) == 0);
    const uint64_t groups = (n + 31) / 32 + 1;                 // one zero pad group after the data
    SWGA_NOTE_LAUNCH(); k_pack<<<grid_for(groups, 256), 256, 0, as_stream(stream)>>>(d_ascii, n, mask_mode, d_packed, d_brk, groups);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

int swga_mark_records_origin(swga_ctx *ctx, const uint64_t *d_rec_starts, uint64_t n_rec, uint64_t origin,
                             uint64_t n, uint32_t *d_brk, cudaStream_t stream)
{
    if (n_rec == 0) return SWGA_OK;
    SWGA_NOTE_LAUNCH(); k_mark_records<<<grid_for(n_rec, 256), 256, 0, stream>>>(d_rec_starts, n_rec, origin, n, d_brk);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_mark_records(swga_ctx *ctx, const uint64_t *d_rec_starts, uint64_t n_rec, uint64_t n,
                                 uint32_t *d_brk, void *stream)
{
    ARG_CHECK(ctx && d_brk && (d_rec_starts || n_rec == 0));
    return swga_mark_records_origin(ctx, d_rec_starts, n_rec, 0, n, d_brk, as_stream(stream));
}

extern "C" int swga_count_accumulate(swga_ctx *ctx, const uint32_t *d_packed, const uint32_t *d_brk,
                                     uint64_t n_starts, int kmin, int kmax, uint32_t *d_tables, void *stream)
{
    ARG_CHECK(ctx && d_packed && d_brk && d_tables);
    TRY(check_k(kmin, kmax));
    if (n_starts == 0) return SWGA_OK;
    cudaStream_t st = as_stream(stream);
    const TableOffs offs = swga_make_offs(kmin, kmax);
    // Partitioned shared-memory path when the shape allows it (count_part.cu) and the input is large enough to
    // pay for its five launches (measured crossover with the direct kernel: 13 M starts), else direct L2 reds.
    // (Running the two concurrently on split ranges was measured in round 1 and does not overlap: both
    // are limited by the per-SM LSU issue path -- profiles/r01_summary.md.)
    if (ctx->count_impl == 2 || (ctx->count_impl == 0 && n_starts >= PT_AUTO_MIN_STARTS)) {
        int done = 0;
        GenomeSrc src;
        memset(&src, 0, sizeof(src));
        src.packed = d_packed;
        src.brk = d_brk;
        TRY(swga_count_accumulate_partitioned(ctx, src, SRC_PACKED, n_starts, kmin, kmax, d_tables, st, &done));
        if (done) return SWGA_OK;
        if (ctx->count_impl == 2) {
            swga_set_error("partitioned counting does not cover kmax=%d n_starts=%llu", kmax, (unsigned long long)n_starts);
            return SWGA_E_ARG;
        }
    }
    ctx->part_stats = nullptr;
    ctx->stage_valid = 0;
    SWGA_NOTE_LAUNCH(); k_count<<<grid_for((n_starts + 31) / 32, 256), 256, 0, st>>>(d_packed, d_brk, n_starts, kmin, kmax, d_tables, offs);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

// the end-of-buffer break of a mask that holds only record boundaries (the ASCII-fed counting path; k_pack sets it itself)
__global__ void k_mark_end(uint64_t n, uint32_t *__restrict__ brk)
{
    if (n > 0) atomicOr(&brk[(n - 1) >> 5], 1u << (31 - (uint32_t)((n - 1) & 31)));
}

int swga_count_ascii(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, const uint64_t *d_marks, uint64_t n_marks,
                     uint64_t origin, int mask_mode, uint64_t n_starts, int kmin, int kmax, uint32_t *d_packed,
                     uint32_t *d_brk, uint32_t *d_tables, cudaStream_t stream)
{
    ARG_CHECK(ctx && (d_ascii || n == 0) && d_packed && d_brk && d_tables && n_starts <= n);
    ARG_CHECK(mask_mode == SWGA_MASK_NONE || mask_mode == SWGA_MASK_N);
    ARG_CHECK((reinterpret_cast<uintptr_t>(d_ascii) & 15) == 0);
    TRY(check_k(kmin, kmax));
    if (n_starts == 0) return SWGA_OK;
    // The partitioned kernels read the ASCII bases themselves (packed in registers on the way in): no pack pass, no
    // packed copy of the genome.  Smaller inputs / other k ranges: pack, then the direct kernel.
    int part = 0;
    if (ctx->count_impl == 2 || (ctx->count_impl == 0 && n_starts >= PT_AUTO_MIN_STARTS)) TRY(swga_count_partition_supported(n_starts, kmax, &part));
    if (part && !ctx->count_via_pack) {
        CUDA_TRY(cudaMemsetAsync(d_brk, 0, swga_brk_words(n) * 4, stream));
        SWGA_NOTE_LAUNCH(); k_mark_end<<<1, 1, 0, stream>>>(n, d_brk);
        TRY(swga_mark_records_origin(ctx, d_marks, n_marks, origin, n, d_brk, stream));
        GenomeSrc src;
        memset(&src, 0, sizeof(src));
        src.brk = d_brk;
        src.ascii = d_ascii;
        src.n = n;
        int done = 0;
        TRY(swga_count_accumulate_partitioned(ctx, src, mask_mode == SWGA_MASK_N ? SRC_ASCII_N : SRC_ASCII, n_starts, kmin, kmax,
                                              d_tables, stream, &done));
        if (done) return SWGA_OK;
    }
    TRY(swga_pack(ctx, d_ascii, n, mask_mode, d_packed, d_brk, stream));
    TRY(swga_mark_records_origin(ctx, d_marks, n_marks, origin, n, d_brk, stream));
    return swga_count_accumulate(ctx, d_packed, d_brk, n_starts, kmin, kmax, d_tables, stream);
}

extern "C" int swga_count_accumulate_ascii(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, const uint64_t *d_rec_marks,
                                           uint64_t n_marks, int mask_mode, uint64_t n_starts, int kmin, int kmax,
                                           uint32_t *d_packed, uint32_t *d_brk, uint32_t *d_tables, void *stream)
{
    return swga_count_ascii(ctx, d_ascii, n, d_rec_marks, n_marks, 0, mask_mode, n_starts, kmin, kmax, d_packed, d_brk, d_tables,
                            as_stream(stream));
}

extern "C" int swga_merge_supported(int kmin, int kmax, int world, int *ok);
extern "C" int swga_merge_finalize(swga_ctx *ctx, const uint64_t *h_blocks, int world, int rank, int kmin, int kmax, void *stream);

extern "C" int swga_count_finalize(swga_ctx *ctx, int kmin, int kmax, uint32_t *d_tables, void *stream)
{
    ARG_CHECK(ctx && d_tables);
    TRY(check_k(kmin, kmax));
    // up to 7 levels below kmax (the default 5..12): ONE pass over the block (csrc/merge.cu with a single "rank"): every
    // CTA takes a span of T_kmax and walks its four-way sums down the levels in shared memory
    int fused = 0;
    TRY(swga_merge_supported(kmin, kmax, 1, &fused));
    if (fused && kmax > kmin && (reinterpret_cast<uintptr_t>(d_tables) & 15) == 0) {
        const uint64_t blk = reinterpret_cast<uint64_t>(d_tables);
        return swga_merge_finalize(ctx, &blk, 1, 0, kmin, kmax, stream);
    }
    const TableOffs o = swga_make_offs(kmin, kmax);
    for (int k = kmax - 1; k >= kmin; --k) {
        const uint64_t n_lower = 1ull << (2 * k);
        SWGA_NOTE_LAUNCH(); k_marginalize<<<grid_for(n_lower, 256), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4 *>(d_tables + o.off[k + 1]), d_tables + o.off[k], n_lower);
        CUDA_TRY(cudaGetLastError());
    }
    return SWGA_OK;
}

extern "C" int swga_fold_canonical(swga_ctx *ctx, int kmin, int kmax, const uint32_t *d_fwd, uint32_t *d_canon,
                                   void *stream)
{
    ARG_CHECK(ctx && d_fwd && d_canon && d_fwd != d_canon);
    TRY(check_k(kmin, kmax));
    const TableOffs offs = swga_make_offs(kmin, kmax);
    // small tables: one thread per code; tables of k >= 10 (94 % of the codes for 5..12): tiled transpose; one launch
    const int k_split = std::max(kmin, 10);
    uint64_t total_small = 0, blocks = 0;
    if (kmin < k_split) {
        total_small = swga_tables_words(kmin, std::min(kmax, k_split - 1));
        blocks += (total_small + 255) / 256;
    }
    for (int k = k_split; k <= kmax; ++k) blocks += 1ull << (2 * (k - 6));
    SWGA_NOTE_LAUNCH(); k_fold_all<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_fwd, d_canon, kmin, kmax, k_split, offs, total_small);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_synth_bases(swga_ctx *ctx, uint64_t seed, uint64_t start, uint64_t n, uint32_t t_a, uint32_t t_c,
                                uint32_t t_g, uint8_t *d_ascii, void *stream)
{
    ARG_CHECK(ctx && (d_ascii || n == 0));
    ARG_CHECK((reinterpret_cast<uintptr_t>(d_ascii) & 15) == 0);
    if (n == 0) return SWGA_OK;
    SWGA_NOTE_LAUNCH(); k_synth<<<grid_for((n + 15) / 16, 256), 256, 0, as_stream(stream)>>>(seed, start, n, t_a, t_c, t_g, d_ascii);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_synth_sets(swga_ctx *ctx, uint64_t seed, uint64_t start, uint64_t n_sets, uint32_t n_primers,
                               int32_t *d_sets, void *stream)
{
    ARG_CHECK(ctx && (d_sets || n_sets == 0) && n_primers >= 10);
    if (n_sets == 0) return SWGA_OK;
    SWGA_NOTE_LAUNCH(); k_synth_sets<<<grid_for(n_sets, 256), 256, 0, as_stream(stream)>>>(seed, start, n_sets, n_primers, d_sets);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_bench_red_global(swga_ctx *ctx, uint32_t *d_table, uint64_t table_words, uint64_t n_ops, float *ms)
{
    ARG_CHECK(ctx && d_table && ms && table_words && (table_words & (table_words - 1)) == 0);
    const unsigned blocks = ctx->sm_count * 8, threads = 256;
    uint64_t per = (n_ops / ((uint64_t)blocks * threads) + 3) & ~3ull;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    SWGA_NOTE_LAUNCH(); k_bench_red_global<<<blocks, threads>>>(d_table, (uint32_t)(table_words - 1), 4);   // warm-up
    CUDA_TRY(cudaEventRecord(e0));
    SWGA_NOTE_LAUNCH(); k_bench_red_global<<<blocks, threads>>>(d_table, (uint32_t)(table_words - 1), per);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SWGA_OK;
}

extern "C" int swga_bench_atom_shared(swga_ctx *ctx, uint32_t table_words, uint64_t n_ops, float *ms)
{
    ARG_CHECK(ctx && ms && table_words && (table_words & (table_words - 1)) == 0 && table_words <= 49152);
    const unsigned blocks = ctx->sm_count * 4, threads = 256;
    uint64_t per = (n_ops / ((uint64_t)blocks * threads) + 3) & ~3ull;
    const size_t smem = (size_t)table_words * 4;
    CUDA_TRY(cudaFuncSetAttribute(k_bench_atom_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TRY(swga_ensure(ctx, ctx->scratch, 256));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    SWGA_NOTE_LAUNCH(); k_bench_atom_shared<<<blocks, threads, smem>>>(table_words - 1, 4, (uint32_t *)ctx->scratch.p);
    CUDA_TRY(cudaEventRecord(e0));
    SWGA_NOTE_LAUNCH(); k_bench_atom_shared<<<blocks, threads, smem>>>(table_words - 1, per, (uint32_t *)ctx->scratch.p);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SWGA_OK;
}
