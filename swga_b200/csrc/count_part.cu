// count_part.cu -- kernel (b), partitioned form: the same additive "top-k + tails" counts as k_count
// (count.cu), but the updates of the top table T_kmax are taken out of the L2 atomic unit.
//
// Round-1 profile (profiles/r01_summary.md): k_count sits at the red.global ceiling (~190 G/s, LTS 87 %)
// while shared-memory atomics measure ~2600 G/s.  So the k-mers are first PARTITIONED by the top bits of
// their code into buckets, and each bucket is then histogrammed in a shared-memory table.
//
// Round 2: the ELEMENT that travels is the (kmax+1)-mer at every EVEN start position.  It holds two
// consecutive kmax-mers -- its first kmax bases and its last kmax bases -- so one shared-memory slot
// atomic, one 2-byte staging store and one histogram atomic now serve two table updates (round 1 spent them
// per k-mer and was bound by exactly these operations: 10.5 + 3.9 shared wavefronts per 32 k-mers).  The
// element code has 2 (kmax+1) bits: the top `pb` bits select one of nb = 2^pb buckets, the low PT_LB = 15
// bits are stored as uint16.  The histogram H of a bucket over those 15 bits is turned into table updates
// when it is flushed, by two linear maps:
//     first  kmax-mer of an element = code >> 2          : T[(b << 13) | (e >> 2)]          += H[e]
//     second kmax-mer               = low 2 kmax bits     : T[((b mod 2^(pb-2)) << 15) | e]  += H[e]
//
//   (1) k_bucket_hist   counts the prefixes of a strided SAMPLE of the genome (1 warp-tile in `stride`);
//   (2) k_bucket_layout turns the sample into (a) a global bucket layout -- estimated size + slack per
//       bucket -- and (b) per-tile slot ranges inside the shared-memory staging area of k_partition;
//   (3) k_partition     one tile = 64 Ki starts = 32 Ki elements: every element takes its staging slot with
//       ONE shared atomicAdd on a packed (guard | slot) counter and ONE 2-byte shared store; the
//       whole 16-byte chunks of every bucket's run are then copied to the bucket (one global atomicAdd per
//       bucket and tile reserves the range), the up to 7 elements left over wait for the next tile;
//   (4) k_bucket_count  shared-memory histogram of each bucket, flushed to T_kmax through the two maps.
//
// Nothing depends on the sample being right: an element that finds its staging range full goes through
// a small overflow list, an element that finds its global bucket full is added to T_kmax directly
// (red.global), groups of 32 starts that touch a break (record end, 'N', buffer end) keep the direct
// path of k_count, and single-base runs (poly-A, DSK mode B's N -> G) are one add of 32.  Results
// are therefore identical to k_count by construction (tests/test_count_gpu.py).
//
// Reference results reproduced: as count.cu (Bank.cpp:889-902,1042-1070; OAHash.cpp:142-154).
#include "common.cuh"

#define PT_THREADS 1024                             // one CTA per SM
#define PT_EPG 16                                   // elements per group of 32 starts (the even offsets)
#define PT_GP 2                                     // groups per thread and tile
#define PT_BATCH 4                                  // elements whose slot atomics are issued together
#define PT_TILE_STARTS (PT_THREADS * 32 * PT_GP)    // 65536 starts per tile
#define PT_TILE_ELEMS (PT_THREADS * PT_EPG * PT_GP) // 32768 elements per tile
#define PT_LB 15                                    // low bits kept per element (uint16), 32768-bin tables
#define PT_MIN_PB 7                                 // kmax = 10: 2 * 11 - 15
#define PT_MAX_PB 11                                // kmax = 12: 2 * 13 - 15
#define PT_STAGE 90112u                             // uint16 staging slots per tile (176 KB)
#define PT_CH 8                                     // bucket runs are stored in whole 16-byte chunks of 8 elements,
#define PT_PAD_WORD 0x80008000u                     // padded with values 0x8000 | r (r < PT_PAD_BINS), which land in
#define PT_PAD_BINS 64                              // dump bins behind k_bucket_count's 2^15-bin table
#define PT_OVF_CAP 512u                             // staging-overflow list entries per tile
#define PT_AGG 1024u                                // slots of the per-CTA table that aggregates overflow elements by code
#define PT_ADD ((1u << 17) + 1u)                    // one element: guard += 1, slot += 1
#define PT_CAP_MAX (0x4000u - 8u)                   // slots of one bucket's range (14-bit guard)
#define PT_ITEMS 2                                  // buckets per thread in the single-block layout kernels (nb <= 2048)
#define HIST_THREADS 512
#define BC_THREADS 1024
#define BC_ITEM (1u << 20)                          // elements per work item of k_bucket_count

// Where the bases of a group of 32 starts come from: the packed genome + break mask that k_pack wrote (SRC_PACKED),
// or the ASCII bases themselves (SRC_ASCII / SRC_ASCII_N: packed in registers on the way in, so that the pack kernel
// and its 1 + 1/4 + 1/8 bytes per base of DRAM traffic disappear from the counting path; `brk` then holds only the
// record boundaries and the end of the buffer, and SRC_ASCII_N adds DSK mode A's 'N' breaks on the fly).
struct Group {
    uint32_t w0, w1, w2;           // bases 0..15, 16..31 of the group and the 16 after it
    uint32_t b0, b1;               // break bits of the group and of the 32 bases after it
    bool fast;
};

// The raw words of a group, fetched ahead of their use.  Volatile asm loads: the compiler may neither sink them to
// the first use nor re-issue them there (a plain __ldg of read-only data is rematerialised, which turns a prefetch
// into a stall); nothing is computed from the values until finish_group().
template <int SRC> struct RawGroup;
template <> struct RawGroup<SRC_PACKED> {
    uint32_t w0, w1, w2, b0, b1;
};
struct RawAscii {
    uint4 a, b;                    // bases 0..15, 16..31; the 16 after them come from the next lane when the group is finished
    uint32_t b0, b1;
};
template <> struct RawGroup<SRC_ASCII> : RawAscii {};
template <> struct RawGroup<SRC_ASCII_N> : RawAscii {};

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const void *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// 16 bases from byte p on, 'A' past the end of the buffer (the break mask ends the last window before them)
__device__ __forceinline__ uint4 ascii16_edge(const uint8_t *__restrict__ ascii, uint64_t p, uint64_t n)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t q = p + 4 * i + j;
            v |= (q < n ? (uint32_t)ascii[q] : 0x41u) << (8 * j);
        }
        w[i] = v;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int SRC>
__device__ __forceinline__ RawGroup<SRC> prefetch_group(const GenomeSrc &S, uint64_t g, uint64_t n_starts)
{
    RawGroup<SRC> r;
    r.b0 = r.b1 = 0xffffffffu;
    if constexpr (SRC == SRC_PACKED) {
        r.w0 = r.w1 = r.w2 = 0;
        if (g * 32 >= n_starts) return r;
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.w0), "=r"(r.w1) : "l"(S.packed + 2 * g));
        r.w2 = ldg_nc_u32(S.packed + 2 * g + 2);
    } else {
        // loaded whenever the buffer has the bases (not only for owned starts): the lane before needs them as look-ahead
        r.a = r.b = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        const uint64_t p = g * 32;
        if (p >= S.n || n_starts == 0) return r;
        if (p + 32 <= S.n) {
            r.a = ldg_nc_v4(S.ascii + p);
            r.b = ldg_nc_v4(S.ascii + p + 16);
        } else {                                               // the last group of the buffer
            r.a = ascii16_edge(S.ascii, p, S.n);
            r.b = ascii16_edge(S.ascii, p + 16, S.n);
        }
    }
    r.b0 = ldg_nc_u32(S.brk + g);
    r.b1 = ldg_nc_u32(S.brk + g + 1);
    return r;
}

// Must be called by all 32 lanes of a warp, lane l holding group g0 + l (SRC_ASCII*: the 16 bases after a group are
// the first 16 of the next lane's; lane 31 reads them itself -- the next warp has them in L1 already).
template <int SRC>
__device__ __forceinline__ Group finish_group(const GenomeSrc &S, const RawGroup<SRC> &r, uint64_t g, uint64_t n_starts, int kmax)
{
    Group G;
    G.b0 = r.b0; G.b1 = r.b1;
    if constexpr (SRC == SRC_PACKED) {
        G.w0 = r.w0; G.w1 = r.w1; G.w2 = r.w2;
    } else {
        G.w0 = pack16(r.a); G.w1 = pack16(r.b);
        // 'N' flags (DSK mode A).  Bit 3 of a byte is clear in A, C, G, T (either case) and set in 'N': the exact
        // comparison (a third of this function's instructions) runs only when some lane of the warp holds such a byte.
        uint32_t fa = 0, fb = 0;
        if constexpr (SRC == SRC_ASCII_N) {
            const uint32_t sus = (r.a.x | r.a.y | r.a.z | r.a.w | r.b.x | r.b.y | r.b.z | r.b.w) & 0x08080808u;
            if (__any_sync(0xffffffffu, sus != 0u)) {
                fa = flags16(r.a, SWGA_MASK_N);
                fb = flags16(r.b, SWGA_MASK_N);
            }
        }
        uint32_t w2 = __shfl_down_sync(0xffffffffu, G.w0, 1), fc = __shfl_down_sync(0xffffffffu, fa, 1);
        if ((threadIdx.x & 31) == 31) {
            const uint64_t p = g * 32 + 32;
            uint4 c = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
            if (n_starts != 0 && p + 16 <= S.n) c = ldg_nc_v4(S.ascii + p);
            else if (n_starts != 0 && p < S.n) c = ascii16_edge(S.ascii, p, S.n);
            w2 = pack16(c);
            if constexpr (SRC == SRC_ASCII_N) {
                fc = 0;
                if ((c.x | c.y | c.z | c.w) & 0x08080808u) fc = flags16(c, SWGA_MASK_N);
            }
        }
        G.w2 = w2;
        if constexpr (SRC == SRC_ASCII_N) {
            // an 'N' at base i forbids the windows over (i-1, i) and (i, i+1): brk[j] = bad[j] | bad[j+1]  (k_pack)
            const uint32_t badA = (fa << 16) | fb;
            const uint32_t badC = fc << 16;
            G.b0 |= badA | (badA << 1) | (badC >> 31);
            G.b1 |= badC | (badC << 1);                        // only its leading kmax-2 bits are looked at
        }
    }
    const int need = kmax - 2;
    const uint32_t tail = need > 0 ? (G.b1 >> (32 - need)) : 0u;
    G.fast = (g * 32 + 32 <= n_starts) && ((G.b0 | tail) == 0u);
    return G;
}

template <int SRC>
__device__ __forceinline__ Group load_group(const GenomeSrc &S, uint64_t g, uint64_t n_starts, int kmax)
{
    return finish_group<SRC>(S, prefetch_group<SRC>(S, g, n_starts), g, n_starts, kmax);
}

// the 16 bases starting at offset o (0..31) of the group, first base in bits 31:30
__device__ __forceinline__ uint32_t window(const Group &g, int o)
{
    return o < 16 ? __funnelshift_l(g.w1, g.w0, 2 * o) : __funnelshift_l(g.w2, g.w1, 2 * (o - 16));
}

// the group (and the kmax-1 bases after it) is a repeat of period 1 or 2: all even-offset (kmax+1)-mers are equal
__device__ __forceinline__ bool uniform_group(const Group &g, int kmax)
{
    return g.w0 == g.w1 && g.w0 == __funnelshift_l(g.w0, g.w0, 4) && ((g.w2 ^ g.w0) >> (34 - 2 * kmax)) == 0u;
}

// direct path for a group that touches a break or the end of the owned range (same as k_count's slow path)
__device__ __forceinline__ void slow_group(const Group &G, uint64_t g, uint64_t n_starts, int kmin, int kmax,
                                           uint32_t *__restrict__ tables, const TableOffs &offs)
{
    const uint64_t p0 = g * 32;
    if (p0 >= n_starts) return;
    const int lim = (int)min((uint64_t)32, n_starts - p0);
    for (int o = 0; o < lim; ++o) {
        const uint32_t W = o < 16 ? __funnelshift_l(G.w1, G.w0, 2 * o) : __funnelshift_l(G.w2, G.w1, 2 * (o - 16));
        const uint32_t B = __funnelshift_l(G.b1, G.b0, o);
        const int v = min(kmax, __clz(B) + 1);
        if (v >= kmin) atomicAdd(tables + offs.off[v] + (W >> (32 - 2 * v)), 1u);
    }
}

// atom.shared.add.u32 on a 32-bit shared-memory address
__device__ __forceinline__ uint32_t atomicAdd_shared(uint32_t addr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}

// both kmax-mers of an element (code = 2 (kmax+1) bits), straight to the top table
__device__ __forceinline__ void add_element(uint32_t *__restrict__ top, uint32_t code, int kmax)
{
    atomicAdd(top + (code >> 2), 1u);
    atomicAdd(top + (code & ((1u << (2 * kmax)) - 1u)), 1u);
}

// ---- (1) per-prefix counts of the fast groups of a strided sample ------------------------------------
// Sampled group i is group (i / 32) * 32 * stride + (i % 32): one warp-tile of 1024 consecutive starts
// out of every `stride`.
template <int SRC>
__global__ void __launch_bounds__(HIST_THREADS)
k_bucket_hist(GenomeSrc S, uint64_t n_starts, int kmax, int pb, uint64_t n_sampled, uint32_t stride,
              uint32_t *__restrict__ bucket_cnt)
{
    extern __shared__ uint32_t h[];
    const uint32_t nb = 1u << pb;
    for (uint32_t i = threadIdx.x; i < nb; i += HIST_THREADS) h[i] = 0;
    __syncthreads();
    const int sh = 32 - pb;
    // warp-uniform trip count: load_group is a warp-wide call (the lanes hold 32 consecutive groups)
    for (uint64_t i0 = blockIdx.x * (uint64_t)HIST_THREADS + (threadIdx.x & ~31u); i0 < n_sampled; i0 += (uint64_t)gridDim.x * HIST_THREADS) {
        const uint64_t i = i0 + (threadIdx.x & 31u);
        const uint64_t g = (i >> 5) * 32 * stride + (i & 31);
        const Group G = load_group<SRC>(S, g, n_starts, kmax);
        if (i < n_sampled && G.fast && !uniform_group(G, kmax)) {   // uniform groups never reach the buckets
#pragma unroll
            for (int o = 0; o < 32; o += 2) atomicAdd(&h[window(G, o) >> sh], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += HIST_THREADS)
        if (h[i]) atomicAdd(bucket_cnt + i, h[i]);
}

// ---- (2) bucket layout from the sample -----------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_excl_scan_1024(unsigned long long v, unsigned long long *s_warp,
                                                                    unsigned long long *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    __syncthreads();                                           // s_warp may still be read by a previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = s_warp[lane];
        unsigned long long wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += o;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}

// Global slack of one bucket: 1/16 of the estimate + 8 sigma of the sampling noise + a constant.
__host__ __device__ static inline unsigned long long bucket_slack(unsigned long long est, float sqrt_s, uint32_t ratio)
{
    return est / 16 + (unsigned long long)(8.0f * (float)ratio * sqrt_s) + 1024ull;
}

// exclusive scan over PT_ITEMS consecutive values per thread (bucket b = PT_ITEMS * threadIdx.x + j)
__device__ __forceinline__ void block_excl_scan_items(const unsigned long long (&v)[PT_ITEMS], unsigned long long (&before)[PT_ITEMS],
                                                      unsigned long long *s_warp, unsigned long long *total)
{
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) sum += v[j];
    unsigned long long run = block_excl_scan_1024(sum, s_warp, total);
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        before[j] = run;
        run += v[j];
    }
}

// single block of 1024 threads; nb <= PT_ITEMS * 1024.  `pad` = expected padding per bucket (runs are stored in
// chunks of PT_CH elements).  n_groups / n_sampled: groups of 32 starts in the genome / in the sample.
__global__ void __launch_bounds__(1024)
k_bucket_layout(const uint32_t *__restrict__ sample_cnt, uint32_t nb, uint64_t n_groups, uint64_t n_sampled,
                uint32_t ratio, uint64_t pad, uint64_t capacity, uint32_t *__restrict__ base,
                uint32_t *__restrict__ cursor, uint32_t *__restrict__ init)
{
    __shared__ unsigned long long s_warp[33];
    unsigned long long s[PT_ITEMS], tmp[PT_ITEMS], cap_g[PT_ITEMS], off[PT_ITEMS], cap_t[PT_ITEMS], start[PT_ITEMS];
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        s[j] = b < nb ? sample_cnt[b] : 0u;
    }
    unsigned long long S;
    block_excl_scan_items(s, tmp, s_warp, &S);
    // (a) global layout
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        const unsigned long long est = n_sampled ? s[j] * n_groups / n_sampled : 0ull;
        cap_g[j] = b < nb ? (est + bucket_slack(est, sqrtf((float)s[j] + 1.0f), ratio) + pad + 7ull) & ~7ull : 0ull;
    }
    unsigned long long total;
    block_excl_scan_items(cap_g, off, s_warp, &total);
    const bool squeeze = total > capacity;                     // cannot happen with the host's bound; stay in range anyway
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        unsigned long long o = off[j];
        if (squeeze) o = (o * capacity / total) & ~7ull;
        if (b < nb) {
            base[b] = (uint32_t)o;
            cursor[b] = (uint32_t)o;
        }
    }
    if (threadIdx.x == 0) base[nb] = (uint32_t)(squeeze ? capacity & ~7ull : total);
    // (b) staging slots per tile: 5/4 of the expected share of the tile + beta, a multiple of 8
    const uint32_t beta = (PT_STAGE - PT_TILE_ELEMS * 5 / 4) / nb - 8;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        cap_t[j] = 0;
        if (b < nb) {
            const unsigned long long share = S ? s[j] * (PT_TILE_ELEMS * 5 / 4) / S : 0ull;
            cap_t[j] = min((unsigned long long)PT_CAP_MAX, (share + beta + 7ull) & ~7ull);
        }
    }
    unsigned long long tot_t;
    block_excl_scan_items(cap_t, start, s_warp, &tot_t);
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        if (b < nb) init[b] = ((0x4000u - (uint32_t)cap_t[j]) << 17) | (uint32_t)start[j];
    }
}

// ---- (3) scatter the low bits into prefix buckets --------------------------------------------------
// The last, partial chunk of a run: element of rank r sits at position r ^ x of the chunk (x = bucket & 7, the
// bank swizzle); positions whose rank is >= nv hold stale data and become padding.
__device__ __forceinline__ uint4 pad_partial_chunk(uint4 v, uint32_t nv, uint32_t x, const uint4 &padv)
{
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t pw[4] = {padv.x, padv.y, padv.z, padv.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (((uint32_t)q ^ x) >= nv) {
            const uint32_t m = 0xFFFFu << (16 * (q & 1));
            w[q >> 1] = (w[q >> 1] & ~m) | (pw[q >> 1] & m);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int KMAX, int SRC>
__global__ void __launch_bounds__(PT_THREADS, 1)
k_partition(GenomeSrc S, uint64_t n_starts, int kmin,
            uint32_t *__restrict__ tables, TableOffs offs, const uint32_t *__restrict__ base,
            const uint32_t *__restrict__ init_g, uint32_t *__restrict__ cursor, uint16_t *__restrict__ data,
            uint32_t n_tiles, unsigned long long *__restrict__ stats)
{
    constexpr int KE = KMAX + 1;                               // bases per element
    constexpr int PB = 2 * KE - PT_LB, SH = 32 - 2 * KMAX, SHE = 32 - 2 * KE;
    constexpr uint32_t NB = 1u << PB, LMASK = (1u << PT_LB) - 1u;
    constexpr int WARPS = PT_THREADS / 32, UNITS = NB / 32;      // a unit = 32 buckets, one per lane
    constexpr int ROUNDS = (UNITS + WARPS - 1) / WARPS;        // units per warp
    static_assert(NB <= PT_ITEMS * 1024 && (NB % 32) == 0, "bucket count");
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t *fill = sm;                                       // [NB] (guard << 17) | next slot of the bucket's range
    uint32_t *init = fill + NB;                                // [NB] value of fill at the start of a tile
    uint32_t *ovf = init + NB;                                 // [PT_OVF_CAP] codes that found their staging range full
    uint32_t *agg_key = ovf + PT_OVF_CAP;                      // [PT_AGG] code of an aggregated overflow element, ~0 = free
    uint32_t *agg_cnt = agg_key + PT_AGG;                      // [PT_AGG] its multiplicity in this tile
    uint16_t *stg = reinterpret_cast<uint16_t *>(agg_cnt + PT_AGG);   // [PT_STAGE]
    __shared__ uint32_t s_novf[2];                             // entries of `ovf`, by tile parity
    __shared__ uint32_t s_stats[4];                            // exits taken: staging full, bucket full, uniform, slow groups
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *__restrict__ top = tables + offs.off[KMAX];
    __shared__ uint32_t s_dummy[32];                           // target of the atomics of skipped elements, one word per lane
    const uint32_t fill_s = (uint32_t)__cvta_generic_to_shared(fill);
    const uint32_t dummy_s = (uint32_t)__cvta_generic_to_shared(s_dummy + lane);
    if (tid < 4) s_stats[tid] = 0;

    // Padding (only the last chunk a CTA writes into a bucket, and the chunks of the overflow list, carry any):
    // values 0x8000 | r, r < 64, which land in k_bucket_count's dump bins, spread so that they do not pile up
    // on one shared-memory word.
    uint4 padv;
    {
        const uint32_t r0 = lane * 8;
        padv.x = PT_PAD_WORD | ((r0 + 0) & 63) | (((r0 + 1) & 63) << 16);
        padv.y = PT_PAD_WORD | ((r0 + 2) & 63) | (((r0 + 3) & 63) << 16);
        padv.z = PT_PAD_WORD | ((r0 + 4) & 63) | (((r0 + 5) & 63) << 16);
        padv.w = PT_PAD_WORD | ((r0 + 6) & 63) | (((r0 + 7) & 63) << 16);
    }
    for (uint32_t i = tid; i < NB; i += PT_THREADS) {
        const uint32_t v = init_g[i];
        init[i] = v;
        fill[i] = v;
    }
    for (uint32_t i = tid; i < PT_AGG; i += PT_THREADS) {
        agg_key[i] = 0xffffffffu;
        agg_cnt[i] = 0;
    }
    if (tid < 2) s_novf[tid] = 0;
    __syncthreads();

    uint32_t par = 0;
    // group of (tile, gp, thread): consecutive threads read consecutive groups; the NEXT group's words are always in
    // flight while the current one is scattered (past the last group: past n_starts, nothing is loaded)
    const uint64_t n_groups_all = (uint64_t)n_tiles * (PT_THREADS * PT_GP);
    uint64_t g = (uint64_t)blockIdx.x * (PT_THREADS * PT_GP) + tid;
    RawGroup<SRC> R = prefetch_group<SRC>(S, g, n_starts);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int gp = 0; gp < PT_GP; ++gp) {
            const Group G = finish_group<SRC>(S, R, g, n_starts, KMAX);
            const uint64_t g_cur = g;
            g = gp + 1 < PT_GP ? g + PT_THREADS
                               : (uint64_t)(tile + gridDim.x) * (PT_THREADS * PT_GP) + tid;
            R = prefetch_group<SRC>(S, g, g < n_groups_all ? n_starts : 0);
            // ---- every element takes its staging slot ----
            if (G.fast) {
                if (uniform_group(G, KMAX)) {                  // poly-A, (AC)n, mode B's N -> G: 16 equal elements
                    atomicAdd(top + (G.w0 >> SH), 16u);
                    atomicAdd(top + ((G.w0 >> SHE) & ((1u << (2 * KMAX)) - 1u)), 16u);
                    atomicAdd(&s_stats[2], 1u);
                } else {
                    uint32_t full = 0;                         // bit 15-e: element e (offset 2e) found its range full
                    uint32_t prev = 0xffffffffu;
                    // batches of PT_BATCH elements: their slot atomics are in flight together, then their stores (an atomic
                    // may not pass an earlier shared store in program order, so one at a time exposes its whole latency)
#pragma unroll
                    for (int o0 = 0; o0 < 32; o0 += 2 * PT_BATCH) {
                        uint32_t W[PT_BATCH], sw[PT_BATCH], old[PT_BATCH];
#pragma unroll
                        for (int j = 0; j < PT_BATCH; ++j) {
                            W[j] = window(G, o0 + 2 * j);
                            const uint32_t boff = (W[j] >> (30 - PB)) & ((NB - 1u) << 2);
                            // An element that follows one of the same bucket takes the overflow path: a group then adds
                            // at most 8 to any counter, a tile at most 2^14, and the 14-bit guard cannot wrap (it would
                            // re-open a full range after 2^14 + range further elements of one tile).
                            // (branch-free: the skipped atomic goes to a per-lane dummy word instead; a conditional atomic
                            // compiles to a branch with a reconvergence barrier, five more instructions per element)
                            const bool dup = boff == prev;
                            const uint32_t got = atomicAdd_shared(dup ? dummy_s : fill_s + boff, PT_ADD);
                            old[j] = dup ? 0x80000000u : got;
                            prev = boff;
                            sw[j] = (boff >> 2) & 7u;
                        }
#pragma unroll
                        for (int j = 0; j < PT_BATCH; ++j) {
                            // slot inside its 8-element chunk is xor-ed with the bucket's low bits: all ranges start on a
                            // 16-byte boundary and fill in step, which would put every early store on 8 of the 32 banks
                            if ((int32_t)old[j] >= 0) stg[(old[j] & 0x1FFFFu) ^ sw[j]] = (uint16_t)((W[j] >> SHE) & LMASK);
                            full = __funnelshift_l(old[j], full, 1);
                        }
                    }
                    while (full) {                             // rare: the tile's composition is off the sample's
                        const int e = __clz(full) - 16;
                        full &= ~(0x8000u >> e);
                        const uint32_t code = window(G, 2 * e) >> SHE;
                        // Overflow elements are mostly REPEATS (a tandem repeat puts a whole tile into a few buckets):
                        // aggregate them by code in a small open-addressing table, flushed once per tile with two
                        // table updates per DISTINCT code -- per-element red.global on a handful of addresses
                        // serialises in L2 (measured: a 10x slower genome).  Two probes, then the list, then direct.
                        uint32_t h = (code * 2654435761u) >> 22;
                        bool done = false;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (!done) {
                                const uint32_t k = atomicCAS(&agg_key[h], 0xffffffffu, code);
                                if (k == 0xffffffffu || k == code) {
                                    atomicAdd(&agg_cnt[h], 1u);
                                    done = true;
                                }
                                h = (h + 1u) & (PT_AGG - 1u);
                            }
                        }
                        if (!done) {
                            const uint32_t slot = atomicAdd(&s_novf[par], 1u);
                            if (slot < PT_OVF_CAP) ovf[slot] = code;
                            else add_element(top, code, KMAX); // list full too: straight to the table
                        }
                        atomicAdd(&s_stats[0], 1u);
                    }
                }
            } else {
                slow_group(G, g_cur, n_starts, kmin, KMAX, tables, offs);
                if (g_cur * 32 < n_starts) atomicAdd(&s_stats[3], 1u);
            }
        }
        __syncthreads();
        // ---- per warp, lane = bucket: reserve the global range of the run, copy its whole chunks, re-arm ----
        // Only WHOLE chunks of 8 leave the staging area; the up to 7 elements left over move to the head of the
        // bucket's range for the next tile, so padding is written once per (CTA, bucket), with the last tile.
        const bool last = tile + gridDim.x >= n_tiles;
        // pass A: the reservations of ALL my buckets are in flight together (a global atomic with a result is a
        // ~2 us round trip here; the profile showed the write-out waiting for them one round at a time)
        uint32_t u_[ROUNDS], s0_[ROUNDS], at_[ROUNDS], lim_[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const uint32_t unit = r * WARPS + warp;
            u_[r] = s0_[r] = at_[r] = lim_[r] = 0;
            if (unit >= UNITS) continue;
            const uint32_t b = unit * 32 + lane;
            const uint32_t f = fill[b], i0 = init[b];
            const uint32_t cap = 0x4000u - (i0 >> 17);
            const uint32_t u = (int32_t)f < 0 ? cap : (f - i0) & 0x1FFFFu;      // elements staged
            const uint32_t ru = last ? (u + 7u) & ~7u : u & ~7u;               // elements that leave now
            if (ru) {
                at_[r] = atomicAdd(cursor + b, ru);
                lim_[r] = __ldg(base + b + 1);                 // end of the bucket in `data`
            }
            fill[b] = i0 + (last ? 0u : (u & 7u) * PT_ADD);
            u_[r] = u;
            s0_[r] = i0 & 0x1FFFFu;                            // first slot of the range
        }
        // (while the reservations are in flight) staging-overflow list: one chunk per element
        {
            const uint32_t n_all = s_novf[par], n_ovf = min(n_all, PT_OVF_CAP);
            if (tid == 0) {
                s_novf[par ^ 1] = 0;
                if (n_all > n_ovf) s_stats[1] += n_all - n_ovf;
            }
            // aggregated overflow elements: both k-mers of each distinct code, with its multiplicity
            for (uint32_t i = tid; i < PT_AGG; i += PT_THREADS) {
                const uint32_t code = agg_key[i];
                if (code != 0xffffffffu) {
                    const uint32_t c = agg_cnt[i];
                    atomicAdd(top + (code >> 2), c);
                    atomicAdd(top + (code & ((1u << (2 * KMAX)) - 1u)), c);
                    agg_key[i] = 0xffffffffu;
                    agg_cnt[i] = 0;
                }
            }
            for (uint32_t i = tid; i < n_ovf; i += PT_THREADS) {
                const uint32_t code = ovf[i], b = code >> PT_LB;
                const uint32_t at = atomicAdd(cursor + b, 8u);
                if (at + 8u <= __ldg(base + b + 1)) {
                    uint4 v = padv;
                    v.x = (v.x & 0xFFFF0000u) | (code & LMASK);
                    *reinterpret_cast<uint4 *>(data + at) = v;
                } else {
                    add_element(top, code, KMAX);
                    atomicAdd(&s_stats[1], 1u);
                }
            }
        }
        // pass B: copy the whole chunks -- a PAIR of lanes per bucket, so that the two chunks a bucket typically sends
        // per tile are adjacent 16-byte stores of one instruction (one L1 wavefront instead of two: the profile showed
        // the write-out bound by the wavefronts of uncoalesced stores) -- then move the left-over chunk to the head
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const uint32_t unit = r * WARPS + warp;
            if (unit >= UNITS) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int src = 16 * h + (int)(lane >> 1);
                const uint32_t b = unit * 32 + src;
                const uint32_t u = __shfl_sync(0xffffffffu, u_[r], src), s0 = __shfl_sync(0xffffffffu, s0_[r], src);
                const uint32_t at = __shfl_sync(0xffffffffu, at_[r], src), lim = __shfl_sync(0xffffffffu, lim_[r], src);
                const uint32_t ru = last ? (u + 7u) & ~7u : u & ~7u;
                for (uint32_t cc = (lane & 1u) * 8u; cc < ru; cc += 16u) {
                    uint4 v = *reinterpret_cast<const uint4 *>(stg + s0 + cc);
                    if (u - cc < 8u) v = pad_partial_chunk(v, u - cc, b & 7u, padv);   // last tile only
                    if (at + cc + 8u <= lim) {
                        *reinterpret_cast<uint4 *>(data + at + cc) = v;
                    } else {                                   // the run crosses the end of its bucket: straight to the table
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t e = (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;
                            if (e <= LMASK) {
                                add_element(top, (b << PT_LB) | e, KMAX);
                                atomicAdd(&s_stats[1], 1u);
                            }
                        }
                    }
                }
            }
            __syncwarp();                                      // the head chunk may still be read by the pair's other lane
            {
                const uint32_t u = u_[r];
                const uint32_t ru = last ? (u + 7u) & ~7u : u & ~7u;
                if (!last && (u & 7u) && ru) {
                    uint4 *hp = reinterpret_cast<uint4 *>(stg + s0_[r]);
                    *hp = hp[ru >> 3];
                }
            }
        }
        __syncthreads();
        par ^= 1;
    }
    __syncthreads();
    if (tid < 4 && s_stats[tid]) atomicAdd(stats + tid, (unsigned long long)s_stats[tid]);
}

// ---- work items of phase 4: ceil(size_b / BC_ITEM) per bucket ----------------------------------------
__global__ void __launch_bounds__(1024)
k_item_prefix(const uint32_t *__restrict__ bucket_base, const uint32_t *__restrict__ cursor, uint32_t nb,
              uint32_t *__restrict__ item_prefix)
{
    // single block of 1024 threads; nb <= PT_ITEMS * 1024
    __shared__ unsigned long long s_warp[33];
    unsigned long long items[PT_ITEMS], before[PT_ITEMS];
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        items[j] = 0;
        if (b < nb) {
            const uint32_t sz = min(cursor[b], bucket_base[b + 1]) - bucket_base[b];
            items[j] = (sz + BC_ITEM - 1) / BC_ITEM;
        }
    }
    unsigned long long total;
    block_excl_scan_items(items, before, s_warp, &total);
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        const uint32_t b = PT_ITEMS * threadIdx.x + j;
        if (b < nb) item_prefix[b] = (uint32_t)before[j];
    }
    if (threadIdx.x == 0) item_prefix[nb] = (uint32_t)total;
}

// ---- (4) shared-memory histogram of each bucket, added to the top table ------------------------------
__device__ __forceinline__ void count1(uint32_t *tab, uint32_t e)
{
    atomicAdd(&tab[e], 1u);                                    // padding (>= 2^14) goes to the dump bins
}
__device__ __forceinline__ void count8(uint32_t *tab, const uint4 x)
{
    count1(tab, x.x & 0xffffu); count1(tab, x.x >> 16);
    count1(tab, x.y & 0xffffu); count1(tab, x.y >> 16);
    count1(tab, x.z & 0xffffu); count1(tab, x.z >> 16);
    count1(tab, x.w & 0xffffu); count1(tab, x.w >> 16);
}

__global__ void __launch_bounds__(BC_THREADS)
k_bucket_count(const uint16_t *__restrict__ bucket_data, const uint32_t *__restrict__ bucket_base,
               const uint32_t *__restrict__ cursor, const uint32_t *__restrict__ item_prefix, uint32_t nb, int lb, int pb,
               uint32_t *__restrict__ top, uint32_t *__restrict__ work_counter)
{
    extern __shared__ uint32_t tab[];
    __shared__ uint32_t s_item;
    const uint32_t bins = 1u << lb;
    const uint32_t n_items = item_prefix[nb];
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1u);
        for (uint32_t i = threadIdx.x; i < bins + PT_PAD_BINS; i += BC_THREADS) tab[i] = 0;
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        // bucket = last b with item_prefix[b] <= item
        uint32_t lo = 0, hi = nb;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (item_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const uint32_t b = lo;
        const uint32_t part = item - item_prefix[b];
        const uint32_t bend = min(cursor[b], bucket_base[b + 1]);
        const uint32_t beg = bucket_base[b] + part * BC_ITEM;
        const uint32_t end = min(bend, beg + BC_ITEM);
        // head up to 16-byte alignment, vector body (two loads in flight per thread), tail
        const uint32_t abeg = min(end, (beg + 7u) & ~7u);
        const uint32_t aend = abeg + ((end - abeg) & ~7u);
        for (uint32_t i = beg + threadIdx.x; i < abeg; i += BC_THREADS) count1(tab, bucket_data[i]);
        const uint4 *v = reinterpret_cast<const uint4 *>(bucket_data + abeg);
        const uint32_t nv = (aend - abeg) >> 3;
        uint32_t i = threadIdx.x;
        for (; i + 3 * BC_THREADS < nv; i += 4 * BC_THREADS) {
            const uint4 x0 = __ldcs(v + i), x1 = __ldcs(v + i + BC_THREADS);
            const uint4 x2 = __ldcs(v + i + 2 * BC_THREADS), x3 = __ldcs(v + i + 3 * BC_THREADS);
            count8(tab, x0); count8(tab, x1); count8(tab, x2); count8(tab, x3);
        }
        for (; i < nv; i += BC_THREADS) count8(tab, __ldcs(v + i));
        for (uint32_t j = aend + threadIdx.x; j < end; j += BC_THREADS) count1(tab, bucket_data[j]);
        __syncthreads();
        // flush: H[e] counts elements of code (b << lb | e); each is two kmax-mers (see the head of this file)
        uint32_t *dst2 = top + ((uint64_t)(b & ((1u << (pb - 2)) - 1u)) << lb);    // second k-mer: low 2 kmax bits of the code
        for (uint32_t j = threadIdx.x; j < bins; j += BC_THREADS) {
            const uint32_t c = tab[j];
            if (c) atomicAdd(dst2 + j, c);
        }
        uint32_t *dst1 = top + ((uint64_t)b << (lb - 2));                          // first k-mer: code >> 2
        const uint4 *tab4 = reinterpret_cast<const uint4 *>(tab);
        for (uint32_t t = threadIdx.x; t < (bins >> 2); t += BC_THREADS) {
            const uint4 q = tab4[t];
            const uint32_t c = q.x + q.y + q.z + q.w;
            if (c) atomicAdd(dst1 + t, c);
        }
        __syncthreads();
    }
}

int swga_count_partition_supported(uint64_t n_starts, int kmax, int *ok)
{
    const int pb = 2 * (kmax + 1) - PT_LB;
    *ok = !(pb < PT_MIN_PB || pb > PT_MAX_PB || n_starts < (1u << 22) || n_starts > 0xE0000000ull);
    return SWGA_OK;
}

// Partitioned accumulate.  Returns SWGA_OK and sets *done = 1 when it handled the request, *done = 0
// when the shape is outside its range (the caller then uses the direct kernel).
int swga_count_accumulate_partitioned(swga_ctx *ctx, const GenomeSrc &src, int src_kind, uint64_t n_starts,
                                      int kmin, int kmax, uint32_t *d_tables, cudaStream_t stream, int *done)
{
    *done = 0;
    const int pb = 2 * (kmax + 1) - PT_LB;
    int ok = 0;
    swga_count_partition_supported(n_starts, kmax, &ok);
    if (!ok) return SWGA_OK;
    const uint32_t nb = 1u << pb;
    const TableOffs offs = swga_make_offs(kmin, kmax);
    // sample: one warp-tile (32 groups = 1024 starts) out of every `stride` (<= 64), at least ~2^15 groups
    const uint64_t n_groups = (n_starts + 31) / 32;
    const uint32_t stride = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, n_groups >> 15));
    const uint64_t n_sampled = (n_groups / (32ull * stride)) * 32 + std::min<uint64_t>(32, n_groups % (32ull * stride));
    const uint32_t ratio = (uint32_t)((n_groups + n_sampled - 1) / n_sampled);
    // upper bound of the layout (sum of estimate + slack over the buckets; Cauchy-Schwarz on the sqrt terms)
    const double s_tot = (double)n_sampled * PT_EPG;
    const uint32_t n_tiles = (uint32_t)((n_starts + PT_TILE_STARTS - 1) / PT_TILE_STARTS);
    const uint64_t pad = 2ull * ctx->sm_count * PT_CH;         // padding per bucket: one partial chunk per CTA, with its last tile
    const uint64_t capacity = (uint64_t)((double)n_groups * PT_EPG * (17.0 / 16.0) + 8.0 * ratio * sqrt((double)nb * (s_tot + nb)) +
                                         (1033.0 + (double)pad) * nb + 64.0 * ratio + 4096.0);
    if (capacity >= 0xFFFFFFF0ull) return SWGA_OK;            // offsets are 32-bit
    // workspace: bucket data (2 B per slot) + sample[nb] + base[nb+1] + cursor[nb] + init[nb] + item_prefix[nb+1] + work
    TRY(swga_ensure(ctx, ctx->bucket_data, (size_t)capacity * 2 + 64));
    TRY(swga_ensure(ctx, ctx->bucket_meta, (size_t)(5 * nb + 16) * 4));
    uint16_t *data = (uint16_t *)ctx->bucket_data.p;
    uint32_t *sample = (uint32_t *)ctx->bucket_meta.p;         // [nb]
    uint32_t *base = sample + nb;                              // [nb + 1]
    uint32_t *cursor = base + nb + 1;                          // [nb]
    uint32_t *init = cursor + nb;                              // [nb]
    uint32_t *item_prefix = init + nb;                         // [nb + 1]
    uint32_t *work = item_prefix + nb + 1;                     // [1]
    unsigned long long *stats = (unsigned long long *)(sample + 5 * nb + 4);   // [4] exits taken (swga_count_partition_stats)
    ctx->part_stats = stats;
    CUDA_TRY(cudaMemsetAsync(sample, 0, (size_t)(5 * nb + 16) * 4, stream));
    const bool timing = ctx->stage_timing != 0;
    if (timing) {
        for (int i = 0; i < 4; ++i)
            if (!ctx->ev_stage[i]) CUDA_TRY(cudaEventCreate(&ctx->ev_stage[i]));
        CUDA_TRY(cudaEventRecord(ctx->ev_stage[0], stream));
    }
    const int grid_h = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * 4, (n_sampled + HIST_THREADS - 1) / HIST_THREADS);
    SWGA_NOTE_LAUNCH();
    if (src_kind == SRC_PACKED) k_bucket_hist<SRC_PACKED><<<grid_h, HIST_THREADS, nb * 4, stream>>>(src, n_starts, kmax, pb, n_sampled, stride, sample);
    else if (src_kind == SRC_ASCII) k_bucket_hist<SRC_ASCII><<<grid_h, HIST_THREADS, nb * 4, stream>>>(src, n_starts, kmax, pb, n_sampled, stride, sample);
    else k_bucket_hist<SRC_ASCII_N><<<grid_h, HIST_THREADS, nb * 4, stream>>>(src, n_starts, kmax, pb, n_sampled, stride, sample);
    SWGA_NOTE_LAUNCH(); k_bucket_layout<<<1, 1024, 0, stream>>>(sample, nb, n_groups, n_sampled, ratio, pad, capacity, base, cursor, init);
    if (timing) CUDA_TRY(cudaEventRecord(ctx->ev_stage[1], stream));
    const size_t smem_p = (size_t)nb * 4 * 2 + PT_OVF_CAP * 4 + PT_AGG * 8 + PT_STAGE * 2;
    const uint32_t grid_p = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count);
#define LAUNCH_PARTITION(K, SRC)                                                                                         \
    do {                                                                                                                 \
        CUDA_TRY(cudaFuncSetAttribute(k_partition<K, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));   \
        SWGA_NOTE_LAUNCH();                                                                                              \
        k_partition<K, SRC><<<grid_p, PT_THREADS, smem_p, stream>>>(src, n_starts, kmin, d_tables, offs, base, init,     \
                                                                    cursor, data, n_tiles, stats);                       \
    } while (0)
#define LAUNCH_PARTITION_K(K)                                                                                            \
    do {                                                                                                                 \
        if (src_kind == SRC_PACKED) LAUNCH_PARTITION(K, SRC_PACKED);                                                     \
        else if (src_kind == SRC_ASCII) LAUNCH_PARTITION(K, SRC_ASCII);                                                  \
        else LAUNCH_PARTITION(K, SRC_ASCII_N);                                                                           \
    } while (0)
    if (kmax == 12) LAUNCH_PARTITION_K(12);
    else if (kmax == 11) LAUNCH_PARTITION_K(11);
    else LAUNCH_PARTITION_K(10);
#undef LAUNCH_PARTITION_K
#undef LAUNCH_PARTITION
    if (timing) CUDA_TRY(cudaEventRecord(ctx->ev_stage[2], stream));
    SWGA_NOTE_LAUNCH(); k_item_prefix<<<1, 1024, 0, stream>>>(base, cursor, nb, item_prefix);
    const size_t smem_c = ((size_t)4 << PT_LB) + 4 * PT_PAD_BINS;
    CUDA_TRY(cudaFuncSetAttribute(k_bucket_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    SWGA_NOTE_LAUNCH(); k_bucket_count<<<ctx->sm_count, BC_THREADS, smem_c, stream>>>(data, base, cursor, item_prefix, nb, PT_LB, pb,
                                                                 d_tables + offs.off[kmax], work);
    CUDA_TRY(cudaGetLastError());
    if (timing) {
        CUDA_TRY(cudaEventRecord(ctx->ev_stage[3], stream));
        ctx->stage_valid = 1;
    }
    *done = 1;
    return SWGA_OK;
}

extern "C" int swga_count_partition_stats(swga_ctx *ctx, uint64_t *h_stats)
{
    ARG_CHECK(ctx && h_stats);
    if (!ctx->part_stats) {
        memset(h_stats, 0, 4 * sizeof(uint64_t));
        return SWGA_OK;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h_stats, ctx->part_stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return SWGA_OK;
}

extern "C" int swga_ctx_set_stage_timing(swga_ctx *ctx, int on)
{
    ARG_CHECK(ctx);
    ctx->stage_timing = on ? 1 : 0;
    ctx->stage_valid = 0;
    return SWGA_OK;
}

extern "C" int swga_count_stage_ms(swga_ctx *ctx, float *h_ms)
{
    ARG_CHECK(ctx && h_ms);
    h_ms[0] = h_ms[1] = h_ms[2] = 0.f;
    if (!ctx->stage_timing || !ctx->stage_valid) return SWGA_OK;
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stage[3]));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(h_ms + i, ctx->ev_stage[i], ctx->ev_stage[i + 1]));
    return SWGA_OK;
}
