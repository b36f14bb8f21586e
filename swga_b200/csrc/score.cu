// score.cu -- kernel (d): batched scoring of candidate primer sets.
//
// Results reproduced (reference = eclarke/swga), one set per CTA-iteration:
//   swga/locate.py:22-36   linearize_binding_sites: union over the set's primers of their linearised
//                          sites, plus the first and last position of every record, de-duplicated
//   swga/stats.py:22-33    seq_diff: ascending sort, adjacent differences ("gaps")
//   swga/score.py:83-90    max_dist = max(gaps); reject when max_dist > max_fg_bind_dist
//   swga/stats.py:10-19    mean = float(sum)/n ; stdev = sqrt(sum((x-mu)^2)/(n-1))
//   swga/stats.py:36-54    gini with Python-2 floor division in fair_area = height*n/2
//   swga/commands/specfiles/find_sets.yaml:39-40  default score (fg_dist_mean*fg_dist_gini)/bg_dist_mean
//
// Method: the union is formed in a shared-memory bitmap, one coordinate tile of 2^19 positions at a
// time.  Every site sets one bit (shared-memory atomicOr; a second-level bitmap marks non-empty
// words), then each thread walks the set bits of its 1024-position slice in ascending order and
// emits gaps; the gap to the previous slice / tile comes from a block-wide running maximum.  Gaps
// feed exact integer accumulators (count, sum of squares, max) and a shared-memory value histogram
// from which the rank-weighted sum that Gini needs is obtained without sorting.  All integer
// statistics are exact; mean and Gini are bit-identical to the reference's floating point (one
// rounding each); stdev differs only by summation order (~1e-15 relative).
#include "common.cuh"

#define SC_THREADS 512
#define SC_WARPS (SC_THREADS / 32)
#define SC_TILE_SHIFT 19                                   // k_score: 2^19 positions per tile = 1024 per thread
#define DT_SHIFT 13                                        // tile offsets (toff) and k_score_dense: 2^13 positions per tile
#define SC_FINE_PER_TILE (1u << (SC_TILE_SHIFT - DT_SHIFT)) // fine tiles per k_score tile
#define SC_L0_WORDS ((1u << SC_TILE_SHIFT) / 32)           // 16384
#define SC_L0_PHYS (SC_L0_WORDS + SC_L0_WORDS / 32)        // padded: phys(w) = w + (w >> 5)
#define SC_L1_WORDS (SC_L0_WORDS / 32)                     // 512 == SC_THREADS
#define SC_MAX_LISTS 32                                    // set members + the record-bounds list

struct ScoreArgs {
    const uint32_t *sites;       // concatenated sorted-unique site lists (linear coordinates)
    const uint32_t *toff;        // [n_lists][n_tiles + 1] offsets into sites of the tiles of 2^DT_SHIFT positions
    uint32_t n_tiles;
    uint32_t bounds_list;        // id of the list holding every record's first and last position
    const int32_t *sets;         // [S][max_m] list ids, -1 padded
    uint32_t S, max_m;
    const double *bg_dist_mean;  // [S]
    uint32_t max_fg_bind_dist;
    uint32_t hist_bins;          // gaps < hist_bins are histogrammed (Gini); 0 disables
    int interactive;             // score.py:89: never reject
    // outputs, [S] each
    uint32_t *n_sites, *max_dist;
    double *mean, *stdev, *gini, *score;
    uint8_t *status;             // bit0: passed the max_dist filter; bit1: Gini needs the unbounded path
    // optional gap dump for the unbounded path (single set): gaps of set 0 are appended here
    uint32_t *gap_out;
    uint32_t *gap_count;
    uint32_t sort_cap;           // sets with at most this many raw sites are scored by k_score_sort, the rest by k_score
    uint32_t sort_route;         // k_score_sort: 1 = routed front end (counting sort by window), 0 = always merge the lists
    uint32_t *work;              // [2] dynamic work counters (k_score_sort, k_score)
    const uint32_t *ids;         // set ids this kernel scores (from k_score_classify); NULL = all of 0..S-1
    const uint32_t *n_ids;       // number of entries of ids
};

__device__ __forceinline__ uint32_t phys(uint32_t w) { return w + (w >> 5); }

__device__ __forceinline__ uint32_t warp_incl_max(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v = max(v, o);
    }
    return v;
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T ident, T *scratch /* SC_WARPS + 1 */)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        T w = threadIdx.x < SC_WARPS ? scratch[threadIdx.x] : ident;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w = op(w, __shfl_xor_sync(0xffffffffu, w, d));
        if (threadIdx.x == 0) scratch[SC_WARPS] = w;
    }
    __syncthreads();
    return scratch[SC_WARPS];
}

struct OpAddU32 { __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; } };
struct OpAddU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };
struct OpMaxU32 { __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return max(a, b); } };
struct OpMinU32 { __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return min(a, b); } };

__global__ void __launch_bounds__(SC_THREADS, 1)
k_score(ScoreArgs a)
{
    extern __shared__ uint32_t smem[];
    uint32_t *L0 = smem;                                   // SC_L0_PHYS
    uint32_t *L1 = L0 + SC_L0_PHYS;                        // SC_L1_WORDS
    uint32_t *hist = L1 + SC_L1_WORDS;                     // a.hist_bins
    __shared__ uint32_t s_lo[SC_MAX_LISTS], s_hi[SC_MAX_LISTS], s_list[SC_MAX_LISTS];
    __shared__ uint32_t s_warp[SC_WARPS + 1];
    __shared__ unsigned long long s_red64[SC_WARPS + 1];
    __shared__ uint32_t s_red32[SC_WARPS + 1];
    __shared__ uint32_t s_nlists;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < SC_L0_PHYS + SC_L1_WORDS + a.hist_bins; i += SC_THREADS) smem[i] = 0;
    __syncthreads();

    __shared__ uint32_t s_set;
    for (;;) {
        if (tid == 0) s_set = a.work ? atomicAdd(a.work + 1, 1u) : 0xffffffffu;
        __syncthreads();
        uint32_t s = s_set;
        if (!a.work) s = blockIdx.x;                      // single-set use (swga_gini_unbounded)
        if (a.ids) {
            if (s >= *a.n_ids) break;
            s = a.ids[s];
        }
        if (s >= a.S) break;
        // ---- members of this set (+ the record-bounds list) ----
        if (tid == 0) {
            uint32_t n = 0;
            for (uint32_t j = 0; j < a.max_m && n < SC_MAX_LISTS - 1; ++j) {
                const int32_t id = a.sets[(uint64_t)s * a.max_m + j];
                if (id >= 0) s_list[n++] = (uint32_t)id;
            }
            s_list[n++] = a.bounds_list;
            s_nlists = n;
        }
        __syncthreads();
        const uint32_t nl = s_nlists;

        uint32_t cnt = 0, maxgap = 0, first = 0xffffffffu;
        unsigned long long sumsq = 0;
        uint32_t carry = 0;                                // (last position so far) + 1; 0 = none

        const uint32_t n_coarse = (a.n_tiles + SC_FINE_PER_TILE - 1) / SC_FINE_PER_TILE;
        for (uint32_t t = 0; t < n_coarse; ++t) {
            if (tid < nl) {
                const uint32_t *row = a.toff + (uint64_t)s_list[tid] * (a.n_tiles + 1);
                s_lo[tid] = row[t * SC_FINE_PER_TILE];
                s_hi[tid] = row[min(a.n_tiles, (t + 1) * SC_FINE_PER_TILE)];
            }
            __syncthreads();
            uint32_t total = 0;
            for (uint32_t l = 0; l < nl; ++l) total += s_hi[l] - s_lo[l];
            if (total == 0) { __syncthreads(); continue; }   // uniform: nothing in this tile
            const uint32_t tile_base = t << SC_TILE_SHIFT;
            // ---- fill: one bit per site ----
            for (uint32_t l = 0; l < nl; ++l) {
                const uint32_t hi = s_hi[l];
                for (uint32_t i = s_lo[l] + tid; i < hi; i += SC_THREADS) {
                    const uint32_t p = a.sites[i] - tile_base;
                    const uint32_t w = p >> 5;
                    const uint32_t old = atomicOr(&L0[phys(w)], 1u << (p & 31));
                    if (old == 0) atomicOr(&L1[w >> 5], 1u << (w & 31));
                }
            }
            __syncthreads();
            // ---- scan: thread owns L1 word `tid` = L0 words [32 tid, 32 tid + 32) ----
            uint32_t m1 = L1[tid];
            L1[tid] = 0;
            uint32_t enc = 0;                              // (last position in my slice) + 1
            if (m1) {
                const uint32_t j = 31 - __clz(m1);
                const uint32_t w = L0[phys(32 * tid + j)];
                enc = tile_base + (((32 * tid + j) << 5) | (31 - __clz(w))) + 1;
            }
            // exclusive running max over threads, seeded with the carry from earlier tiles
            uint32_t inc = warp_incl_max(enc);
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = lane < SC_WARPS ? s_warp[lane] : 0;
                uint32_t vi = warp_incl_max(v);
                uint32_t ve = __shfl_up_sync(0xffffffffu, vi, 1);
                if (lane == 0) ve = 0;
                if (lane < SC_WARPS) s_warp[lane] = ve;
                if (lane == 31) s_warp[SC_WARPS] = vi;     // max over the whole tile
            }
            __syncthreads();
            uint32_t pred = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) pred = 0;
            pred = max(max(pred, s_warp[warp]), carry);
            carry = max(carry, s_warp[SC_WARPS]);
            // ---- enumerate my set bits in ascending order ----
            while (m1) {
                const uint32_t j = __ffs(m1) - 1;
                m1 &= m1 - 1;
                const uint32_t pw = phys(32 * tid + j);
                uint32_t w = L0[pw];
                L0[pw] = 0;
                const uint32_t wbase = tile_base + ((32 * tid + j) << 5);
                while (w) {
                    const uint32_t pos = wbase + (__ffs(w) - 1);
                    w &= w - 1;
                    if (pred) {
                        const uint32_t gap = pos - (pred - 1);
                        ++cnt;
                        sumsq += (unsigned long long)gap * gap;
                        maxgap = max(maxgap, gap);
                        if (gap < a.hist_bins) atomicAdd(&hist[gap], 1u);
                        if (a.gap_out && s == 0) a.gap_out[atomicAdd(a.gap_count, 1u)] = gap;
                    } else {
                        first = pos;
                    }
                    pred = pos + 1;
                }
            }
            __syncthreads();
        }

        // ---- reduce over the block ----
        const uint32_t n = block_reduce(cnt, OpAddU32(), 0u, s_red32);            // number of gaps
        const uint32_t mx = block_reduce(maxgap, OpMaxU32(), 0u, s_red32);
        const uint32_t fst = block_reduce(first, OpMinU32(), 0xffffffffu, s_red32);
        const unsigned long long ssq = block_reduce(sumsq, OpAddU64(), 0ull, s_red64);
        const uint32_t last = carry ? carry - 1 : 0;
        const unsigned long long H = (carry && fst != 0xffffffffu) ? (unsigned long long)(last - fst) : 0ull;

        const bool passed = a.interactive || mx <= a.max_fg_bind_dist;
        const bool hist_ok = mx < a.hist_bins;
        // ---- rank-weighted sum from the histogram (and clear it) ----
        unsigned long long sum_prefix = 0;
        if (a.hist_bins) {
            const uint32_t top = min(mx, a.hist_bins - 1);                    // highest bin that can be non-zero
            const uint32_t per = top / SC_THREADS + 1;
            const uint32_t b0 = tid * per, b1 = min(b0 + per, top + 1);
            uint32_t c_mine = 0;
            for (uint32_t v = b0; v < b1; ++v) c_mine += hist[v];
            // exclusive scan of c_mine over threads
            uint32_t inc = c_mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (uint32_t)d) inc += o;
            }
            __syncthreads();
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = lane < SC_WARPS ? s_warp[lane] : 0, vi = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t o = __shfl_up_sync(0xffffffffu, vi, d);
                    if (lane >= (uint32_t)d) vi += o;
                }
                if (lane < SC_WARPS) s_warp[lane] = vi - v;
            }
            __syncthreads();
            unsigned long long r = s_warp[warp] + inc - c_mine;               // gaps smaller than my first bin
            unsigned long long part = 0;
            for (uint32_t v = b0; v < b1; ++v) {
                const unsigned long long c = hist[v];
                hist[v] = 0;
                if (c) {
                    // ranks r+1 .. r+c (ascending): sum of (n - rank + 1) = c (n - r) - c (c - 1) / 2
                    part += (unsigned long long)v * (c * ((unsigned long long)n - r) - c * (c - 1) / 2);
                    r += c;
                }
            }
            sum_prefix = block_reduce(part, OpAddU64(), 0ull, s_red64);
        }

        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            double mean = nan, sd = nan, gini = nan, score = nan;
            if (n > 0) {
                mean = (double)H / (double)n;                                  // stats.py:12
                if (n > 1) {
                    // n * sum(x^2) - (sum x)^2 exactly in 128 bits (>= 0), then one conversion
                    const unsigned long long a_lo = (unsigned long long)n * ssq, a_hi = __umul64hi(n, ssq);
                    const unsigned long long b_lo = H * H, b_hi = __umul64hi(H, H);
                    const unsigned long long d_lo = a_lo - b_lo, d_hi = a_hi - b_hi - (a_lo < b_lo ? 1ull : 0ull);
                    const double num = (double)d_hi * 18446744073709551616.0 + (double)d_lo;
                    sd = sqrt(num / ((double)n * (double)(n - 1)));            // stats.py:17-19
                }
                if (a.hist_bins && hist_ok) {
                    const unsigned long long fair = (H * (unsigned long long)n) / 2;   // Python-2 floor, stats.py:53
                    if (fair > 0) {
                        const long long two_area = (long long)(2 * sum_prefix) - (long long)H;
                        gini = (double)((long long)(2 * fair) - two_area) / (double)(2 * fair);
                    }
                }
                score = (mean * gini) / a.bg_dist_mean[s];
            }
            a.n_sites[s] = carry ? n + 1 : 0;
            a.max_dist[s] = mx;
            a.mean[s] = mean;
            a.stdev[s] = sd;
            a.gini[s] = gini;
            a.score[s] = score;
            a.status[s] = (uint8_t)((passed ? 1 : 0) | ((!hist_ok) ? 2 : 0));
        }
        __syncthreads();
        if (!a.work) break;
    }
}


// ------------------------------------------------------------------------------------------------
// Dense sets: one WARP per fine tile of 2^13 positions, no CTA-wide barrier inside a set.  Every warp owns
// a contiguous range of tiles and a private 1 KB bitmap: (1) the lanes read their lists' segment bounds
// of the tile, (2) every site sets one bit, (3) each lane takes 8 bitmap words, the set bits are
// COMPACTED into a warp-private list of tile-relative positions (ascending), (4) the gaps are taken
// from the list with all lanes busy.  The last position of a tile is carried in a register to the next
// tile of the warp; the gaps between the ranges of two warps are added once per set.  Statistics as in
// k_score (exact integers, Gini from the shared gap histogram).
// ------------------------------------------------------------------------------------------------
#define DT_THREADS 1024
#define DT_WARPS (DT_THREADS / 32)
#define DT_WORDS ((1u << DT_SHIFT) / 32)                   // 256 bitmap words per warp
#define DT_LIST_CAP 512u                                   // compact list entries per warp (uint16)

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce_dt(T v, Op op, T ident, T *scratch /* DT_WARPS + 1 */)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        T w = threadIdx.x < DT_WARPS ? scratch[threadIdx.x] : ident;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w = op(w, __shfl_xor_sync(0xffffffffu, w, d));
        if (threadIdx.x == 0) scratch[DT_WARPS] = w;
    }
    __syncthreads();
    return scratch[DT_WARPS];
}

// A read-only load the compiler may neither sink to its first use nor re-issue there (prefetch into registers).
__device__ __forceinline__ uint32_t ld_pinned(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ uint4 ld_pinned4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 16-bit shared-memory accesses through a 32-bit shared address computed once (the generic-pointer form makes
// the compiler rebuild the shared window base inside every unrolled loop body)
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}

struct GapAcc {
    uint32_t cnt, maxgap;
    unsigned long long sumsq;
};
// `hist_s` is the 32-bit shared-memory address of the gap histogram: the predicated red keeps the loop bodies free
// of a branch and of the generic-to-shared address arithmetic the compiler otherwise repeats per iteration.
__device__ __forceinline__ void add_gap(GapAcc &g, uint32_t gap, uint32_t hist_s, uint32_t hist_bins)
{
    ++g.cnt;
    g.sumsq += (unsigned long long)gap * gap;
    g.maxgap = max(g.maxgap, gap);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], 1;\n\t}"
                 ::"r"(gap), "r"(hist_bins), "r"(hist_s + 4u * gap) : "memory");
}
// one bit of a shared-memory bitmap, when idx < len
__device__ __forceinline__ void set_bit_if(uint32_t idx, uint32_t len, uint32_t bm_s, uint32_t p)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p red.shared.or.b32 [%2], %3;\n\t}"
                 ::"r"(idx), "r"(len), "r"(bm_s + 4u * (p >> 5)), "r"(1u << (p & 31)) : "memory");
}


// ---- routed variant of the dense kernel (default) ------------------------------------------------------
// The bitmap variant spends 60 % of its instructions on divergent bit-extraction loops and the gap pass that follows
// them (profiles/r01_summary.md).  Here a tile is a per-warp COUNTING SORT instead: lane w owns the window of 256
// positions [256 w, 256 w + 256) of the tile; every site takes the next slot of its window with one shared atomicAdd
// on the window's counter and is stored there as it is (a 32-bit position); the owner reads its slots, leaves them
// empty (0xffffffff), sorts the values in registers (Batcher's network: 19 compare-exchanges for 8 values, 63 for 16;
// empty slots sort to the end) and takes the gaps from neighbouring registers -- straight-line predicated code, no
// loop whose trip count differs between lanes.  The predecessor of a window's first site is the running maximum of
// the windows before it (one warp scan).  The warp takes the form for the fullest window of the tile (4 .. 16 live
// slots; 89 % of the C4-dense tiles need at most 8) and dense_tile_slow() -- bitmap + per-lane bit walk, in the
// shared memory of the slots -- beyond 16.
// Shared memory of a warp (DT_RT_BYTES, 128-byte aligned): [32 counters][slot 0 of the 32 windows][slot 1] ... [slot 15],
// so that the address of slot s of a window is (address of its counter) + 128 (s + 1): one LEA after the atomic, and
// the stores of a slot index fall on 32 different banks.
#define DT_ROW_CAP 16u
#define DT_RT_BYTES (128u * (DT_ROW_CAP + 1u))
#define DT_EMPTY 0xffffffffu

// The next slot of a window: one shared atomicAdd on its counter.  Chunk entries that are not sites of the segment skip
// it by a branch (a predicated atomic compiles to a branch anyway); their "slot" is DT_ROW_CAP, which predicates the
// store off like that of a 17th site.
__device__ __forceinline__ uint32_t route_take(uint32_t addr)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}
__device__ __forceinline__ void route_put(uint32_t old, uint32_t cnt_addr, uint32_t v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, 16;\n\t@p st.shared.u32 [%1+128], %2;\n\t}"
                 ::"r"(old), "r"(cnt_addr + 128u * old), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// Statistics of one gap in the routed kernel.  The histogram has one more bin behind the last (gaps >= hist_bins) and 32
// dummy words behind that, one per lane, which take the increments of the empty register slots: the update is one
// unconditional shared-memory increment.  The NUMBER of gaps is not counted here: it is the sum of the bins.
__device__ __forceinline__ void add_gap_r(GapAcc &acc, uint32_t g, uint32_t bin, uint32_t hist_s)
{
    acc.sumsq += (unsigned long long)g * g;
    acc.maxgap = max(acc.maxgap, g);
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + 4u * bin) : "memory");
}
// The owner's half of a routed tile: v[0 .. M) = the slots of my window (DT_EMPTY = empty; v[M .. N) are not looked at: no
// window of the tile received more than M sites and slots fill in order).  Sorts them ascending -- the sites then
// come first -- and takes the gap of every site to its predecessor: the register before it, or, for the first, the
// last site before the window (running maximum over the lanes, `carry` from the tiles before, both as position + 1).
// FIRST: the tile may hold the first site of the warp's range (carry == 0): that site has no gap and is reported.
template <int N, int M, bool FIRST>
__device__ __forceinline__ void owner_gaps(uint32_t (&v)[N], uint32_t lane, uint32_t &carry, uint32_t &first, uint32_t (&g)[M])
{
    // (last position of my window) + 1, 0 = no site: the largest v + 1 (an empty slot wraps to 0), taken from the UNSORTED
    // values so that the warp scan below does not wait for the sorting network
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) run = max(run, v[i] + 1u);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) run = max(run, __shfl_up_sync(0xffffffffu, run, d));
    uint32_t pred = __shfl_up_sync(0xffffffffu, run, 1);
    if (lane == 0) pred = 0;
    pred = max(pred, carry);                                   // >= 1 unless FIRST
    // Batcher's odd-even merge sort of N = 8 or 16 inputs, fully unrolled (every index is a compile-time constant); the
    // inputs from M on count as +infinity, so a compare-exchange that touches one of them does nothing and is left out
#pragma unroll
    for (int p = 1; p < N; p <<= 1) {
#pragma unroll
        for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k) {
#pragma unroll
                for (int i = 0; i <= (k - 1 < N - j - k - 1 ? k - 1 : N - j - k - 1); ++i) {
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p) && i + j + k < M) {
                        const uint32_t lo_ = min(v[i + j], v[i + j + k]), hi_ = max(v[i + j], v[i + j + k]);
                        v[i + j] = lo_;
                        v[i + j + k] = hi_;
                    }
                }
            }
        }
    }
    // g[i] = gap of my i-th site to its predecessor; 0 = no gap here (empty slot, a site two lists share, the first site)
#pragma unroll
    for (int i = 0; i < M; ++i) {
        bool site = v[i] != DT_EMPTY;
        if (FIRST && i == 0 && site && pred == 0u) {
            first = v[0] + 1u;
            site = false;
        }
        g[i] = site ? (i ? v[i] - v[i ? i - 1 : 0] : v[0] + 1u - pred) : 0u;
    }
    carry = max(carry, __shfl_sync(0xffffffffu, run, 31));
}

// owner_gaps on the first M slots of my window (read and left empty here; `win_s` = address of my slot 0); every gap goes
// to the histogram, the slots without one to my dummy word behind it
template <int M, bool FIRST>
__device__ __forceinline__ void owner_form(uint32_t win_s, uint32_t lane, uint32_t &carry, uint32_t &first, GapAcc &acc,
                                           uint32_t hist_s, uint32_t bins, uint32_t dummy_bin)
{
    constexpr int N = M <= 8 ? 8 : 16;
    uint32_t v[N], g[M];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < M ? lds_u32(win_s + 128u * i) : DT_EMPTY;
#pragma unroll
    for (int i = 0; i < M; ++i) sts_u32(win_s + 128u * i, DT_EMPTY);
    __syncwarp();
    owner_gaps<N, M, FIRST>(v, lane, carry, first, g);
#pragma unroll
    for (int i = 0; i < M; ++i) add_gap_r(acc, g[i], g[i] ? min(g[i], bins) : dummy_bin, hist_s);
}

// One tile through the bitmap, every lane walking its own 8 words (any number of sites, with or without a predecessor).
// State in and out by value: a variable whose address is passed to a call lives in local memory for the whole kernel.
struct TileState {
    uint32_t carry, first, cnt, maxgap;
    unsigned long long sumsq;
};
__device__ __noinline__ TileState dense_tile_slow(const uint32_t *__restrict__ sites, const uint32_t *__restrict__ toff, uint32_t n_tiles,
                                                  uint32_t hist_bins, const uint32_t *s_list, uint32_t nl, uint32_t t, uint32_t *bm,
                                                  uint32_t hist_s, TileState st)
{
    const uint32_t lane = threadIdx.x & 31u, tile_base = t << DT_SHIFT;
    for (uint32_t l = 0; l < nl; ++l) {
        const uint32_t *r = toff + (uint64_t)s_list[l] * (n_tiles + 1) + t;
        const uint32_t lo_l = r[0], len_l = r[1] - lo_l;
        for (uint32_t i = lane; i < len_l; i += 32) {
            const uint32_t p = __ldg(sites + lo_l + i) - tile_base;
            atomicOr(&bm[p >> 5], 1u << (p & 31));
        }
    }
    __syncwarp();
    uint32_t w[8];
    {
        uint4 *bm4 = reinterpret_cast<uint4 *>(bm) + 2 * lane;
        const uint4 x = bm4[0], y = bm4[1];
        bm4[0] = make_uint4(0, 0, 0, 0);
        bm4[1] = make_uint4(0, 0, 0, 0);
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
    }
    GapAcc acc;
    acc.cnt = st.cnt; acc.maxgap = st.maxgap; acc.sumsq = st.sumsq;
    uint32_t enc = 0;                                          // (last position of my slice) + 1
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (w[q]) enc = tile_base + (lane * 8 + q) * 32 + (31 - __clz(w[q])) + 1;
    uint32_t run = enc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) run = max(run, __shfl_up_sync(0xffffffffu, run, d));
    uint32_t pred = __shfl_up_sync(0xffffffffu, run, 1);
    if (lane == 0) pred = 0;
    pred = max(pred, st.carry);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint32_t v = w[q];
        const uint32_t wb = tile_base + (lane * 8 + q) * 32;
        while (v) {
            const uint32_t cur = wb + __ffs(v) - 1;
            v &= v - 1;
            if (pred) add_gap_r(acc, cur - (pred - 1), min(cur - (pred - 1), hist_bins), hist_s);
            else st.first = cur + 1;
            pred = cur + 1;
        }
    }
    st.carry = max(st.carry, __shfl_sync(0xffffffffu, run, 31));
    st.cnt = acc.cnt; st.maxgap = acc.maxgap; st.sumsq = acc.sumsq;
    __syncwarp();
    return st;
}

template <bool ROUTED>
__global__ void __launch_bounds__(DT_THREADS, 1)
k_score_dense(ScoreArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_list[SC_MAX_LISTS];
    __shared__ uint32_t s_nlists, s_set, s_rawsites;
    __shared__ uint32_t s_first[DT_WARPS], s_last[DT_WARPS];              // (position + 1) of the first / last site of a warp's range
    __shared__ uint32_t s_red32[DT_WARPS + 1], s_scan[DT_WARPS + 1];
    __shared__ unsigned long long s_red64[DT_WARPS + 1];
    __shared__ GapAcc s_cross;
    __shared__ uint32_t s_tot_first, s_tot_last;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bitmap variant: [DT_WARPS][DT_WORDS] bitmaps, [DT_WARPS][DT_LIST_CAP] compact lists (uint16), histogram.
    // routed variant: [DT_WARPS][DT_RT_BYTES] counters + slots from the next 128-byte boundary, histogram; the first 1 KB
    // of a warp's slots is also the bitmap of dense_tile_slow
    uint32_t *bm, *hist;
    uint32_t cl_s = 0, cnt_s = 0;
    if constexpr (ROUTED) {
        const uint32_t pad = (128u - ((uint32_t)__cvta_generic_to_shared(smem) & 127u)) & 127u;
        uint8_t *rt0 = reinterpret_cast<uint8_t *>(smem) + pad, *rt = rt0 + warp * DT_RT_BYTES;
        cnt_s = (uint32_t)__cvta_generic_to_shared(rt);
        bm = reinterpret_cast<uint32_t *>(rt + 128);
        hist = reinterpret_cast<uint32_t *>(rt0 + DT_WARPS * DT_RT_BYTES);
        reinterpret_cast<uint32_t *>(rt)[lane] = 0;
        for (uint32_t i = 0; i < DT_ROW_CAP; ++i) bm[i * 32 + lane] = DT_EMPTY;
        if (tid == 0) hist[a.hist_bins] = 0;
    } else {
        uint32_t *bm_all = smem;
        uint16_t *cl_all = reinterpret_cast<uint16_t *>(bm_all + DT_WARPS * DT_WORDS);
        bm = bm_all + warp * DT_WORDS;
        hist = reinterpret_cast<uint32_t *>(cl_all + DT_WARPS * DT_LIST_CAP);
        cl_s = (uint32_t)__cvta_generic_to_shared(cl_all + warp * DT_LIST_CAP);
        for (uint32_t i = tid; i < DT_WARPS * DT_WORDS; i += DT_THREADS) bm_all[i] = 0;
    }
    const uint32_t bm_s = (uint32_t)__cvta_generic_to_shared(bm), hist_s = (uint32_t)__cvta_generic_to_shared(hist);
    const uint32_t dummy_bin = a.hist_bins + 1u + lane;                     // see add_gap_r
    for (uint32_t i = tid; i < a.hist_bins; i += DT_THREADS) hist[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_set = atomicAdd(a.work + 1, 1u);
        __syncthreads();
        if (s_set >= *a.n_ids) break;
        const uint32_t s = a.ids[s_set];
        if (warp == 0) {
            // the set's lists, lane = member (the record bounds last), and -- routed variant -- their total length: one round
            // trip for the whole warp instead of a serial walk by one thread
            int32_t id = -1;
            if (lane < a.max_m) id = a.sets[(uint64_t)s * a.max_m + lane];
            else if (lane == a.max_m) id = (int32_t)a.bounds_list;
            uint32_t raw = 0;
            if (ROUTED && id >= 0) {
                const uint32_t *r = a.toff + (uint64_t)id * (a.n_tiles + 1);
                raw = r[a.n_tiles] - r[0];
            }
            const uint32_t have = __ballot_sync(0xffffffffu, id >= 0);
            if (id >= 0) s_list[__popc(have & ((1u << lane) - 1u))] = (uint32_t)id;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, d);
            if (lane == 0) { s_nlists = __popc(have); s_rawsites = raw; }
        }
        __syncthreads();
        const uint32_t nl = s_nlists;
        // Routed variant: a thin set is walked in COARSE tiles of 8 or 64 fine ones (windows of 2^11 / 2^14 positions): per-tile
        // costs are fixed (~350 instructions), and a 7,000-site set spent them on 2,808 tiles of 2.5 sites (1.2 M sets/s; 50
        // times the time per site of a 2,000-site set in k_score_sort; 4.7 M sets/s with 351 tiles of 20).  The thresholds keep
        // a window below ~1 site on average: a window that overflows sends ALL fine tiles of its coarse tile through the
        // bitmap (measured with tiles of 160 sites: 2 % of the sets met one, and lost more than the coarser walk gained).
        uint32_t csh = 0;
        if constexpr (ROUTED) {
            const uint32_t per_tile = s_rawsites / a.n_tiles;
            csh = per_tile >= 8u ? 0u : (per_tile >= 1u ? 3u : 6u);
        }
        const uint32_t n_ct = (a.n_tiles + (1u << csh) - 1u) >> csh;          // (coarse) tiles of the set
        const uint32_t tpw = (n_ct + DT_WARPS - 1) / DT_WARPS;                // tiles per warp
        auto fine = [&](uint32_t t) { return min(a.n_tiles, t << csh); };      // first fine tile of (coarse) tile t

        GapAcc acc;
        acc.cnt = 0; acc.maxgap = 0; acc.sumsq = 0;
        uint32_t carry = 0, first = 0;                                     // (position + 1); 0 = none yet
        const uint32_t t0 = min(n_ct, warp * tpw), t1 = min(n_ct, t0 + tpw);
        // Software pipeline over the warp's tiles.  Lane = (list lane / L, part lane % L), L = 4, 3 or 2 lanes per list for up
        // to 8, 10 or 16 lists: it reads two aligned 16-byte chunks of its list's segment of a tile (part p: the chunks at
        // aligned element offsets 8 p and 8 p + 4), i.e. the first 8 L - 3 .. 8 L sites of every list in ONE load
        // instruction per lane group, one tile ahead (the loads of tile t + 1 fly while tile t is processed); the segment
        // bounds are read two tiles ahead.  Longer segments and lists beyond 16 take the scalar loop below.
        const uint32_t L = nl <= 8 ? 4u : (nl <= 10 ? 3u : 2u);
        const uint32_t mylist = lane / L;
        uint32_t half = lane - mylist * L, cov8 = 8u * L;
        asm volatile("" : "+r"(half), "+r"(cov8));                       // kept in registers: the compiler otherwise recomputes both per tile
        const uint32_t *row = a.toff + (uint64_t)s_list[mylist < nl ? mylist : 0] * (a.n_tiles + 1);
        uint32_t b0 = 0, b1 = 0, b2 = 0;                                   // row[t], row[t + 1], row[t + 2] of my list
        if (mylist < nl && t0 < t1) {
            b0 = row[fine(t0)];
            b1 = row[fine(t0 + 1)];
            b2 = row[fine(t0 + 2)];
        }
        uint4 pa = make_uint4(0, 0, 0, 0), pb = pa;                        // chunks of the current tile
        if (b1 > b0) {
            const uint4 *c = reinterpret_cast<const uint4 *>(a.sites + (b0 & ~3u)) + 2 * half;
            pa = ld_pinned4(c);
            pb = ld_pinned4(c + 1);
        }
        for (uint32_t t = t0; t < t1; ++t) {
            const uint32_t lo = b0, len = b1 - b0;
            // next tile: bounds are here already; its sites and the bounds after it start their trip now
            uint4 na = make_uint4(0, 0, 0, 0), nb4 = na;
            if (t + 1 < t1 && b2 > b1) {
                const uint4 *c = reinterpret_cast<const uint4 *>(a.sites + (b1 & ~3u)) + 2 * half;
                na = ld_pinned4(c);
                nb4 = ld_pinned4(c + 1);
            }
            const uint32_t b3 = (mylist < nl && t + 1 < t1) ? ld_pinned(row + fine(t + 3)) : b2;
            const uint32_t tile_base = t << DT_SHIFT;
            const uint32_t covered = cov8 - (lo & 3u);                   // sites of my list the chunks hold
            bool any = __any_sync(0xffffffffu, len != 0);
            uint32_t need = __ballot_sync(0xffffffffu, len > covered && half == 0u);   // lists (their first lane) with sites beyond the chunks
            if (nl > 16) {                                                 // lists no lane owns
                for (uint32_t l = 16; l < nl && !any; ++l) {
                    const uint32_t *r = a.toff + (uint64_t)s_list[l] * (a.n_tiles + 1);
                    any = r[fine(t + 1)] > r[fine(t)];
                }
            }
            if constexpr (ROUTED) {
            if (any) {
                const bool nopred = carry == 0u;                           // no site yet in the warp's range (warp-uniform)
                bool slow;
                uint32_t mx;
                {
                    // ---- (2) every site takes a slot of its window ----
                    {
                        const uint32_t v[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                        // entry q of my chunks is site number 8 half - (lo & 3) + q of the segment: a site for qlo <= q < qhi
                        const int qlo = half ? 0 : (int)(lo & 3u), qhi = (int)len - (int)(8u * half) + (int)(lo & 3u);
                        uint32_t ca[8], old[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            ca[q] = cnt_s | ((v[q] >> (6u + csh)) & 0x7Cu);   // the tile starts at a multiple of its size: 5 bits = window
                            const bool site = q < 3 ? (q >= qlo && q < qhi) : q < qhi;   // qlo <= 3
                            // (measured: sending the other lanes' atomics to per-lane dummy words instead keeps the code
                            // straight, 16 instructions per tile fewer, but adds a wavefront or two per atomic: same speed)
                            old[q] = DT_ROW_CAP;
                            if (site) old[q] = route_take(ca[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) route_put(old[q], ca[q], v[q]);
                    }
                    while (need) {                                         // what the chunks do not cover
                        const int l = __ffs(need) - 1;
                        need &= need - 1;
                        const uint32_t lo_l = __shfl_sync(0xffffffffu, lo, l), len_l = __shfl_sync(0xffffffffu, len, l);
                        for (uint32_t i = cov8 - (lo_l & 3u) + lane; i < len_l; i += 32) {
                            const uint32_t v = __ldg(a.sites + lo_l + i), ca = cnt_s | ((v >> (6u + csh)) & 0x7Cu);
                            route_put(route_take(ca), ca, v);
                        }
                    }
                    for (uint32_t l = 16; l < nl; ++l) {                   // lists no lane owns
                        const uint32_t *r = a.toff + (uint64_t)s_list[l] * (a.n_tiles + 1);
                        const uint32_t lo_l = r[fine(t)], len_l = r[fine(t + 1)] - lo_l;
                        for (uint32_t i = lane; i < len_l; i += 32) {
                            const uint32_t v = __ldg(a.sites + lo_l + i), ca = cnt_s | ((v >> (6u + csh)) & 0x7Cu);
                            route_put(route_take(ca), ca, v);
                        }
                    }
                    __syncwarp();
                    // ---- (3) my window's count (left at 0 for the next tile) ----
                    const uint32_t c = lds_u32(cnt_s + 4u * lane);
                    sts_u32(cnt_s + 4u * lane, 0u);
                    mx = __reduce_max_sync(0xffffffffu, c);                // most sites any window of the tile received
                    slow = mx > DT_ROW_CAP;
                }
                const uint32_t win_s = cnt_s + 128u + 4u * lane;
                if (slow) {
                    // the first 1 KB of the slots becomes the bitmap (all 16 slots of a window may hold sites); afterwards
                    // every slot is empty again
                    uint4 *rowp = reinterpret_cast<uint4 *>(bm) + 2 * lane;
                    rowp[0] = make_uint4(0, 0, 0, 0);
                    rowp[1] = make_uint4(0, 0, 0, 0);
                    __syncwarp();
                    TileState st;
                    st.carry = carry; st.first = first; st.cnt = 0; st.maxgap = acc.maxgap; st.sumsq = acc.sumsq;
                    for (uint32_t f = fine(t); f < fine(t + 1); ++f)       // the bitmap holds one fine tile
                        st = dense_tile_slow(a.sites, a.toff, a.n_tiles, a.hist_bins, s_list, nl, f, bm, hist_s, st);
                    carry = st.carry; first = st.first; acc.maxgap = st.maxgap; acc.sumsq = st.sumsq;
                    uint4 *all = reinterpret_cast<uint4 *>(bm) + 4 * lane;
#pragma unroll
                    for (int i = 0; i < 4; ++i) all[i] = make_uint4(DT_EMPTY, DT_EMPTY, DT_EMPTY, DT_EMPTY);
                    __syncwarp();
                } else if (nopred) {
                    owner_form<16, true>(win_s, lane, carry, first, acc, hist_s, a.hist_bins, dummy_bin);
                } else {
                    // ---- (4) my window's slots; sort, gaps: the form for the fullest window of the tile (warp-uniform) ----
#define DT_FORM(M) owner_form<M, false>(win_s, lane, carry, first, acc, hist_s, a.hist_bins, dummy_bin)
                    if (mx <= 6u) {
                        if (mx <= 4u) DT_FORM(4);
                        else if (mx == 5u) DT_FORM(5);
                        else DT_FORM(6);
                    } else if (mx <= 8u) {
                        if (mx == 7u) DT_FORM(7);
                        else DT_FORM(8);
                    } else if (mx <= 10u) DT_FORM(10);
                    else if (mx <= 12u) DT_FORM(12);
                    else DT_FORM(16);
#undef DT_FORM
                }
            }
            } else {
            if (any) {
            // ---- (2) one bit per site ----
            {
                const uint32_t v[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                const uint32_t first_idx = 8u * half - (lo & 3u);          // index in the segment of v[0] (may wrap below 0)
#pragma unroll
                for (int q = 0; q < 8; ++q)                                // unsigned compare: also false before the segment
                    set_bit_if(first_idx + q, len, bm_s, (v[q] - tile_base) & ((1u << DT_SHIFT) - 1u));
            }
            while (need) {                                                 // what the chunks do not cover
                const int l = __ffs(need) - 1;
                need &= need - 1;
                const uint32_t lo_l = __shfl_sync(0xffffffffu, lo, l), len_l = __shfl_sync(0xffffffffu, len, l);
                for (uint32_t i = cov8 - (lo_l & 3u) + lane; i < len_l; i += 32) {
                    const uint32_t p = __ldg(a.sites + lo_l + i) - tile_base;
                    atomicOr(&bm[p >> 5], 1u << (p & 31));
                }
            }
            for (uint32_t l = 16; l < nl; ++l) {
                const uint32_t *r = a.toff + (uint64_t)s_list[l] * (a.n_tiles + 1) + t;
                const uint32_t lo_l = r[0], len_l = r[1] - lo_l;
                for (uint32_t i = lane; i < len_l; i += 32) {
                    const uint32_t p = __ldg(a.sites + lo_l + i) - tile_base;
                    atomicOr(&bm[p >> 5], 1u << (p & 31));
                }
            }
            __syncwarp();
            // ---- (3) my 8 words; count, scan, compact ----
            uint32_t w[8];
            {
                uint4 *bm4 = reinterpret_cast<uint4 *>(bm) + 2 * lane;
                const uint4 x = bm4[0], y = bm4[1];
                bm4[0] = make_uint4(0, 0, 0, 0);
                bm4[1] = make_uint4(0, 0, 0, 0);
                w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
            }
            uint32_t c = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) c += __popc(w[q]);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += o;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total <= DT_LIST_CAP) {
                uint32_t ad = cl_s + 2u * (incl - c);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t v = w[q];
                    const uint32_t wb = (lane * 8 + q) * 32 - 1;
                    while (v) {
                        sts_u16(ad, wb + __ffs(v));
                        ad += 2;
                        v &= v - 1;
                    }
                }
                __syncwarp();
                // ---- (4) gaps, all lanes busy ----
                for (uint32_t j = lane; j < total; j += 32) {
                    const uint32_t cur = tile_base + lds_u16(cl_s + 2u * j);
                    const uint32_t prev1 = j ? tile_base + lds_u16(cl_s + 2u * j - 2u) + 1 : carry;
                    if (prev1) {
                        const uint32_t gap = cur - (prev1 - 1);
                        add_gap(acc, gap, hist_s, a.hist_bins);
                    } else {
                        first = cur + 1;
                    }
                }
                carry = tile_base + lds_u16(cl_s + 2u * total - 2u) + 1;
                __syncwarp();
            } else {
                // more distinct sites than the list holds: each lane walks its own bits; its predecessor is
                // the last bit of the lanes before it (exclusive running max), else the carry
                uint32_t enc = 0;                                          // (last position of my slice) + 1
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (w[q]) enc = tile_base + (lane * 8 + q) * 32 + (31 - __clz(w[q])) + 1;
                uint32_t run = enc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) run = max(run, __shfl_up_sync(0xffffffffu, run, d) * (lane >= (uint32_t)d));
                uint32_t pred = __shfl_up_sync(0xffffffffu, run, 1);
                if (lane == 0) pred = 0;
                pred = max(pred, carry);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t v = w[q];
                    const uint32_t wb = tile_base + (lane * 8 + q) * 32;
                    while (v) {
                        const uint32_t cur = wb + __ffs(v) - 1;
                        v &= v - 1;
                        if (pred) {
                            const uint32_t gap = cur - (pred - 1);
                            add_gap(acc, gap, hist_s, a.hist_bins);
                            } else {
                            first = cur + 1;
                        }
                        pred = cur + 1;
                    }
                }
                carry = max(carry, __shfl_sync(0xffffffffu, run, 31));
            }
            }                                                              // if (any)
            }                                                              // bitmap variant
            b0 = b1; b1 = b2; b2 = b3;
            pa = na; pb = nb4;
        }
        // ---- the set: every warp's partial statistics and first / last site meet in warp 0 (lane = warp): the gaps between
        // the ranges of consecutive non-empty warps, the totals -- two barriers for all of it ----
        first = __reduce_max_sync(0xffffffffu, first);                    // one lane set it
        {
            uint32_t wn = acc.cnt, wmx = acc.maxgap;
            unsigned long long wss = acc.sumsq;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                wn += __shfl_xor_sync(0xffffffffu, wn, d);
                wmx = max(wmx, __shfl_xor_sync(0xffffffffu, wmx, d));
                wss += __shfl_xor_sync(0xffffffffu, wss, d);
            }
            if (lane == 0) { s_first[warp] = first; s_last[warp] = carry; s_red32[warp] = wn; s_scan[warp] = wmx; s_red64[warp] = wss; }
        }
        __syncthreads();
        if (warp == 0) {
            static_assert(DT_WARPS == 32, "lane = warp below");
            const uint32_t f = s_first[lane], l = s_last[lane];
            uint32_t tn = s_red32[lane], tmx = s_scan[lane];
            unsigned long long tss = s_red64[lane];
            uint32_t pl = l;                                               // last site of the ranges up to mine (position + 1)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) pl = max(pl, __shfl_up_sync(0xffffffffu, pl, d));
            pl = __shfl_up_sync(0xffffffffu, pl, 1);
            if (lane == 0) pl = 0;
            if (l && pl) {                                                 // the gap between my range and the one before it
                const uint32_t gap = f - pl;                               // both carry the + 1
                GapAcc x;
                x.cnt = 0; x.maxgap = 0; x.sumsq = 0;
                if constexpr (ROUTED) add_gap_r(x, gap, min(gap, a.hist_bins), hist_s);
                else add_gap(x, gap, hist_s, a.hist_bins);
                tn += 1u; tmx = max(tmx, gap); tss += x.sumsq;
            }
            uint32_t fmin = l ? f : 0xffffffffu, lmax = l;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                tn += __shfl_xor_sync(0xffffffffu, tn, d);
                tmx = max(tmx, __shfl_xor_sync(0xffffffffu, tmx, d));
                tss += __shfl_xor_sync(0xffffffffu, tss, d);
                fmin = min(fmin, __shfl_xor_sync(0xffffffffu, fmin, d));
                lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
            }
            if (lane == 0) {
                s_cross.cnt = tn; s_cross.maxgap = tmx; s_cross.sumsq = tss;
                s_tot_first = fmin; s_tot_last = lmax;
                if constexpr (ROUTED) hist[0] = 0;                        // (dummy-free: never written; kept for safety)
            }
        }
        __syncthreads();
        // number of gaps: counted per gap (bitmap variant) or the sum of the histogram incl. its overflow bin (routed)
        uint32_t n = 0;
        if constexpr (!ROUTED) n = s_cross.cnt;
        const uint32_t mx = s_cross.maxgap;
        const unsigned long long ssq = s_cross.sumsq;
        const uint32_t fst1 = s_tot_first, lst1 = s_tot_last;
        const unsigned long long H = lst1 ? (unsigned long long)(lst1 - fst1) : 0ull;
        const bool passed = a.interactive || mx <= a.max_fg_bind_dist;
        const bool hist_ok = mx < a.hist_bins;
        // ---- rank-weighted sum from the histogram (and clear it) ----
        unsigned long long sum_prefix = 0;
        if (a.hist_bins) {
            const uint32_t top = min(mx, a.hist_bins - 1);                    // highest bin that can be non-zero
            const uint32_t per = top / DT_THREADS + 1;
            const uint32_t b0 = tid * per, b1 = min(b0 + per, top + 1);
            uint32_t c_mine = 0;
            for (uint32_t v = b0; v < b1; ++v) c_mine += hist[v];
            uint32_t inc = c_mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (uint32_t)d) inc += o;
            }
            __syncthreads();
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                const uint32_t v = s_scan[lane];
                uint32_t vi = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, vi, d);
                    if (lane >= (uint32_t)d) vi += o;
                }
                s_scan[lane] = vi - v;
                if (lane == 31) s_scan[DT_WARPS] = vi;
            }
            __syncthreads();
            if constexpr (ROUTED) n = s_scan[DT_WARPS] + hist[a.hist_bins];
            unsigned long long r = s_scan[warp] + inc - c_mine;               // gaps smaller than my first bin
            unsigned long long part = 0;
            for (uint32_t v = b0; v < b1; ++v) {
                const unsigned long long c = hist[v];
                hist[v] = 0;
                if (c) {
                    part += (unsigned long long)v * (c * ((unsigned long long)n - r) - c * (c - 1) / 2);
                    r += c;
                }
            }
            sum_prefix = block_reduce_dt(part, OpAddU64(), 0ull, s_red64);
        }
        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            double mean = nan, sd = nan, gini = nan, score = nan;
            if (n > 0) {
                mean = (double)H / (double)n;                                  // stats.py:12
                if (n > 1) {
                    const unsigned long long a_lo = (unsigned long long)n * ssq, a_hi = __umul64hi(n, ssq);
                    const unsigned long long b_lo = H * H, b_hi = __umul64hi(H, H);
                    const unsigned long long d_lo = a_lo - b_lo, d_hi = a_hi - b_hi - (a_lo < b_lo ? 1ull : 0ull);
                    const double num = (double)d_hi * 18446744073709551616.0 + (double)d_lo;
                    sd = sqrt(num / ((double)n * (double)(n - 1)));            // stats.py:17-19
                }
                if (a.hist_bins && hist_ok) {
                    const unsigned long long fair = (H * (unsigned long long)n) / 2;   // Python-2 floor, stats.py:53
                    if (fair > 0) {
                        const long long two_area = (long long)(2 * sum_prefix) - (long long)H;
                        gini = (double)((long long)(2 * fair) - two_area) / (double)(2 * fair);
                    }
                }
                score = (mean * gini) / a.bg_dist_mean[s];
            }
            a.n_sites[s] = lst1 ? n + 1 : 0;
            a.max_dist[s] = mx;
            a.mean[s] = mean;
            a.stdev[s] = sd;
            a.gini[s] = gini;
            a.score[s] = score;
            a.status[s] = (uint8_t)((passed ? 1 : 0) | ((!hist_ok) ? 2 : 0));
            if constexpr (ROUTED) hist[a.hist_bins] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Sparse sets: all raw sites of the set fit in shared memory -> gather, bitonic sort, adjacent
// differences (duplicates dropped), and for sets that pass the max_dist filter a second sort of the
// gaps gives Gini's rank-weighted sum directly.  128 threads per set, several CTAs per SM.
// ------------------------------------------------------------------------------------------------
#define SS_THREADS 128
#define SS_WARPS (SS_THREADS / 32)

#define MS_E 16u                                        // elements per thread-level segment of the merge sort
#define SS_RUN_PAD (MS_E * SC_MAX_LISTS)                // room for every list's tail up to a multiple of MS_E
#define SS_BINS 1024u                                   // coarse bins of the gap grouping (Gini)
#define SS_BIN_CROWD 256u                               // a bin with more gaps than this: use the merge sort
#define MS_PHYS(i) ((i) + ((i) >> 4))                   // one pad word per segment: lane-strided segment accesses hit 32 banks

// ascending sorting network (bitonic) on 16 registers
__device__ __forceinline__ void sort16(uint32_t (&v)[MS_E])
{
#pragma unroll
    for (int k = 2; k <= (int)MS_E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < (int)MS_E; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t lo = min(v[i], v[l]), hi = max(v[i], v[l]);
                    const bool up = (i & k) == 0;
                    v[i] = up ? lo : hi;
                    v[l] = up ? hi : lo;
                }
            }
        }
    }
}

// Ascending merge sort of a[0..n) in shared memory, b = second buffer of the same capacity; both must
// hold n rounded up to a multiple of MS_E (the tail is filled with 0xffffffff, which sorts last) in the
// padded layout MS_PHYS.
// Thread-level segments of 16 are sorted in registers, then ceil(log2(n / 16)) merge passes follow in
// which every thread produces 16 consecutive outputs of its pair of runs (merge path: one binary
// search for the split, then a sequential 2-way merge).  About 4 x fewer instructions than the
// bitonic network it replaces, and no padding to a power of two.  Returns the buffer with the result.
__device__ __forceinline__ uint32_t *block_merge_sort(uint32_t *a, uint32_t *b, uint32_t n)
{
    const uint32_t R = (n + MS_E - 1) / MS_E, np = R * MS_E;
    for (uint32_t seg = threadIdx.x; seg < R; seg += SS_THREADS) {
        uint32_t v[MS_E];
#pragma unroll
        for (int e = 0; e < (int)MS_E; ++e) {
            const uint32_t i = seg * MS_E + e;
            v[e] = i < n ? a[MS_PHYS(i)] : 0xffffffffu;
        }
        sort16(v);
#pragma unroll
        for (int e = 0; e < (int)MS_E; ++e) a[seg * (MS_E + 1) + e] = v[e];
    }
    __syncthreads();
    for (uint32_t run = MS_E; run < np; run <<= 1) {
        for (uint32_t seg = threadIdx.x; seg < R; seg += SS_THREADS) {
            const uint32_t o = seg * MS_E, pair = o & ~(2u * run - 1u), d = o - pair;
            const uint32_t la = min(run, np - pair), lb = min(run, np - pair - la);
            const uint32_t pb = pair + la;                     // run A = [pair, pair + la), run B = [pb, pb + lb)
            uint32_t lo = d > lb ? d - lb : 0u, hi = min(d, la);
            while (lo < hi) {                                  // first i with A[i] > B[d - 1 - i]
                const uint32_t mid = (lo + hi) >> 1;
                if (a[MS_PHYS(pair + mid)] <= a[MS_PHYS(pb + d - 1 - mid)]) lo = mid + 1; else hi = mid;
            }
            uint32_t i = lo, j = d - lo;
            uint32_t av = i < la ? a[MS_PHYS(pair + i)] : 0xffffffffu, bv = j < lb ? a[MS_PHYS(pb + j)] : 0xffffffffu;
            uint32_t out[MS_E];
#pragma unroll
            for (int e = 0; e < (int)MS_E; ++e) {
                const bool ta = j >= lb || (i < la && av <= bv);
                out[e] = ta ? av : bv;
                if (ta) { ++i; av = i < la ? a[MS_PHYS(pair + i)] : 0xffffffffu; }
                else { ++j; bv = j < lb ? a[MS_PHYS(pb + j)] : 0xffffffffu; }
            }
#pragma unroll
            for (int e = 0; e < (int)MS_E; ++e) b[seg * (MS_E + 1) + e] = out[e];
        }
        __syncthreads();
        uint32_t *t = a; a = b; b = t;
    }
    return a;
}

// Merge of nr ALREADY SORTED runs a[rb[r] .. rb[r + 1]) (boundaries multiples of MS_E, run tails filled with
// 0xffffffff) into one: ceil(log2(nr)) merge-path passes over pairs of neighbouring runs, no base sort.  The
// primers' site lists arrive sorted, so this replaces the first sort of a set (8 passes + the register
// network for ~2,000 sites) by 4 passes for 9-11 lists.  rb lives in shared memory and is overwritten.
__device__ __forceinline__ uint32_t *block_merge_runs(uint32_t *a, uint32_t *b, uint32_t *rb, uint32_t nr)
{
    while (nr > 1) {
        const uint32_t np = rb[nr], R = np / MS_E;
        for (uint32_t seg = threadIdx.x; seg < R; seg += SS_THREADS) {
            const uint32_t o = seg * MS_E;
            uint32_t p = 0;
            while (rb[min(2 * p + 2, nr)] <= o) ++p;               // pair of runs that holds output o
            const uint32_t pair = rb[2 * p], pb = rb[min(2 * p + 1, nr)];
            const uint32_t la = pb - pair, lb = rb[min(2 * p + 2, nr)] - pb, d = o - pair;
            uint32_t lo = d > lb ? d - lb : 0u, hi = min(d, la);
            while (lo < hi) {                                  // first i with A[i] > B[d - 1 - i]
                const uint32_t mid = (lo + hi) >> 1;
                if (a[MS_PHYS(pair + mid)] <= a[MS_PHYS(pb + d - 1 - mid)]) lo = mid + 1; else hi = mid;
            }
            uint32_t i = lo, j = d - lo;
            uint32_t av = i < la ? a[MS_PHYS(pair + i)] : 0xffffffffu, bv = j < lb ? a[MS_PHYS(pb + j)] : 0xffffffffu;
            uint32_t out[MS_E];
#pragma unroll
            for (int e = 0; e < (int)MS_E; ++e) {
                const bool ta = j >= lb || (i < la && av <= bv);
                out[e] = ta ? av : bv;
                if (ta) { ++i; av = i < la ? a[MS_PHYS(pair + i)] : 0xffffffffu; }
                else { ++j; bv = j < lb ? a[MS_PHYS(pb + j)] : 0xffffffffu; }
            }
#pragma unroll
            for (int e = 0; e < (int)MS_E; ++e) b[seg * (MS_E + 1) + e] = out[e];
        }
        __syncthreads();
        const uint32_t nr2 = (nr + 1) / 2;
        if (threadIdx.x == 0) {
            for (uint32_t q = 0; q < nr2; ++q) rb[q] = rb[2 * q];
            rb[nr2] = np;
        }
        nr = nr2;
        __syncthreads();
        uint32_t *t = a; a = b; b = t;
    }
    return a;
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce_ss(T v, Op op, T ident, T *scratch /* SS_WARPS + 1 */)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T w = ident;
#pragma unroll
    for (int i = 0; i < SS_WARPS; ++i) w = op(w, scratch[i]);
    return w;
}

// ---- routed front end of k_score_sort (round 2) -----------------------------------------------------------------
// The merge of the set's sorted lists (4 merge-path passes for 9-11 lists) is replaced by the counting sort of the dense
// kernel, over the WHOLE genome at once: windows of 2^rsh positions (at most RS_W of them; rsh <= 15 so that a position
// inside its window is a uint16 below 0x8000), every site takes the next slot of its window with one shared atomicAdd,
// then each warp walks its share of the windows 32 at a time -- lane = window: sort the slots in registers, gaps from
// neighbouring registers (owner_gaps) -- and appends the gaps to the list the Gini step reads.  A window with more than
// RS_CAP sites sends the set through the merge (its slots are cleared first); genomes above 2^15 RS_W = 33.5 Mbp too.
#define RS_W 1024u
#define RS_CAP 16u
#define RS_EMPTY16 0xffffu
#define RS_BYTES (RS_W * 4u + RS_W * RS_CAP * 2u)          // counters + slots (slot-major: slot s of window w at 2 (s RS_W + w))

__device__ __forceinline__ uint32_t lds_s16(uint32_t addr)    // sign-extended: an empty slot reads as DT_EMPTY
{
    int32_t v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return (uint32_t)v;
}

struct SparseAcc {
    uint32_t cnt, maxgap;
    unsigned long long sumsq;
};

// 32 windows (lane = window `w`, slot 0 at win_s) holding at most M sites each: gaps -> acc, and appended to gp
template <int M, bool FIRST>
__device__ __forceinline__ void sparse_form(uint32_t win_s, uint32_t wbase, uint32_t lane, uint32_t &carry, uint32_t &first,
                                            SparseAcc &acc, uint32_t *gp, uint32_t *s_ng)
{
    constexpr int N = M <= 8 ? 8 : 16;
    uint32_t v[N], g[M];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < M ? (lds_s16(win_s + 2u * RS_W * i) | wbase) : DT_EMPTY;   // empty | wbase = empty
#pragma unroll
    for (int i = 0; i < M; ++i) sts_u16(win_s + 2u * RS_W * i, RS_EMPTY16);
    owner_gaps<N, M, FIRST>(v, lane, carry, first, g);
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        k += g[i] != 0u;
        acc.sumsq += (unsigned long long)g[i] * g[i];
        acc.maxgap = max(acc.maxgap, g[i]);
    }
    acc.cnt += k;
    uint32_t incl = k;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(s_ng, incl);
    uint32_t idx = __shfl_sync(0xffffffffu, base, 31) + incl - k;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (g[i]) {
            gp[MS_PHYS(idx)] = g[i];
            ++idx;
        }
    }
}

__global__ void __launch_bounds__(SS_THREADS)
k_score_sort(ScoreArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t bufwords = MS_PHYS(a.sort_cap + SS_RUN_PAD);
    uint32_t *gaps = smem;                      // second buffer of the merge / the gap list; padded layout MS_PHYS
    uint32_t *buf = smem + bufwords;            // [cap2] positions; the routed front end's counters and slots lie here too
    __shared__ uint32_t s_src[SC_MAX_LISTS], s_len[SC_MAX_LISTS], s_rb[SC_MAX_LISTS + 1];   // s_rb: run starts, multiples of 16
    __shared__ uint32_t s_nl, s_raw, s_set, s_sid;
    __shared__ uint32_t r32[SS_WARPS + 1];
    __shared__ unsigned long long r64[SS_WARPS + 1];
    __shared__ uint32_t s_ng, s_ovf, s_first[SS_WARPS], s_last[SS_WARPS], s_rmx[SS_WARPS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    // routed front end: windows of 2^rsh positions, at most RS_W of them
    uint32_t rsh = DT_SHIFT;
    while (((a.n_tiles + (1u << (rsh - DT_SHIFT)) - 1u) >> (rsh - DT_SHIFT)) > RS_W) ++rsh;
    const uint32_t nwin = (a.n_tiles + (1u << (rsh - DT_SHIFT)) - 1u) >> (rsh - DT_SHIFT);
    const bool can_route = a.sort_route && rsh <= 15u;
    uint32_t *cnt32 = buf;
    const uint32_t slots_s = (uint32_t)__cvta_generic_to_shared(buf + RS_W);
    // words of `buf` (counters, then slots) to re-initialise before the next routed set: the merge and the Gini step write
    // over them (all of them at the start; the grouped Gini step touches its 1,024 bins and n words behind them)
    constexpr uint32_t RT_WORDS = RS_W + RS_W * RS_CAP / 2u;
    uint32_t rt_dirty = RT_WORDS;
    // Warp 0 carries the NEXT set's descriptor through the current set: the dependent global accesses of its chain (work
    // counter -> set id -> members -> extents of their lists, ~3 us end to end) are issued one phase of the current set
    // before their result is needed, instead of in a row at the top of the set with the other warps at the barrier.
    const uint32_t n_ids = *a.n_ids;
    uint32_t nx_set = 0xffffffffu, nx_s = 0, nx_lo = 0, nx_len = 0;
    int32_t nx_id = -1;
    auto pf_counter = [&]() {                   // lane 0: next work item
        if (tid == 0) nx_set = atomicAdd(a.work, 1u);
    };
    auto pf_setid = [&]() {                     // lane 0: its set id
        if (tid == 0) nx_s = nx_set < n_ids ? a.ids[nx_set] : 0u;
    };
    auto pf_members = [&]() {                   // lane = member (the record bounds last)
        if (wq == 0) {
            nx_set = __shfl_sync(0xffffffffu, nx_set, 0);
            nx_s = __shfl_sync(0xffffffffu, nx_s, 0);
            nx_id = -1;
            if (nx_set < n_ids) {
                if (lane < a.max_m) nx_id = a.sets[(uint64_t)nx_s * a.max_m + lane];
                else if (lane == a.max_m) nx_id = (int32_t)a.bounds_list;
            }
        }
    };
    auto pf_extents = [&]() {                   // the extent of every member's list
        if (wq == 0) {
            nx_lo = 0; nx_len = 0;
            if (nx_id >= 0) {
                const uint32_t *row = a.toff + (uint64_t)nx_id * (a.n_tiles + 1);
                nx_lo = row[0];
                nx_len = row[a.n_tiles] - nx_lo;
            }
        }
    };
    uint32_t pf_done = 0;                       // stages of the next descriptor issued so far (uniform)
    auto pf_to = [&](uint32_t k) {
        if (pf_done < 2u && k >= 2u) { pf_setid(); pf_done = 2u; }
        if (pf_done < 3u && k >= 3u) { pf_members(); pf_done = 3u; }
        if (pf_done < 4u && k >= 4u) { pf_extents(); pf_done = 4u; }
    };
    pf_counter(); pf_setid(); pf_members(); pf_extents();
    for (;;) {
        if (wq == 0) {
            const int32_t id = nx_id;
            const uint32_t lo = nx_lo, len = nx_len;
            const uint32_t have = __ballot_sync(0xffffffffu, id >= 0);
            const uint32_t rank = __popc(have & ((1u << lane) - 1u));
            const uint32_t padded = id >= 0 ? (len + MS_E - 1) & ~(MS_E - 1) : 0u;
            uint32_t at = padded, raw = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, at, d);
                if (lane >= (uint32_t)d) at += o;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, d);
            if (id >= 0) { s_src[rank] = lo; s_len[rank] = len; s_rb[rank] = at - padded; }
            if (lane == 31) s_rb[__popc(have)] = at;
            if (lane == 0) {
                s_set = nx_set; s_sid = nx_s;
                s_nl = __popc(have); s_raw = raw;
                s_ng = 0; s_ovf = 0;
            }
        }
        if (can_route && rt_dirty) {
            uint4 *b4 = reinterpret_cast<uint4 *>(buf);               // 16-byte aligned: bufwords is a multiple of 4
            for (uint32_t i = tid; i < RS_W / 4u; i += SS_THREADS) b4[i] = make_uint4(0, 0, 0, 0);
            for (uint32_t i = RS_W / 4u + tid; i < (rt_dirty + 3u) / 4u; i += SS_THREADS)
                b4[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            rt_dirty = 0;
        }
        __syncthreads();
        if (s_set >= n_ids) break;
        const uint32_t s = s_sid;
        pf_counter();
        pf_done = 1u;
        const uint32_t raw = s_raw, nl = s_nl;
        uint32_t cnt = 0, maxgap = 0, entries = 0;
        unsigned long long sumsq = 0, H = 0;
        uint32_t *pos = buf, *gp = gaps;
        bool routed = false;
        if (can_route) {
            // ---- every site takes a slot of its window ----
            const uint32_t rmask = (1u << rsh) - 1u;
            auto route = [&](uint32_t p) {
                const uint32_t w = p >> rsh;
                const uint32_t old = atomicAdd(&cnt32[w], 1u);
                if (old < RS_CAP) sts_u16(slots_s + 2u * (old * RS_W + w), p & rmask);
            };
            // the first two sites per thread of list l + 1 are in flight while list l is routed
            uint32_t v0 = 0, v1 = 0;
            if (nl) {
                const uint32_t *src = a.sites + s_src[0];
                if (tid < s_len[0]) v0 = ld_pinned(src + tid);
                if (tid + SS_THREADS < s_len[0]) v1 = ld_pinned(src + tid + SS_THREADS);
            }
            for (uint32_t l = 0; l < nl; ++l) {
                const uint32_t len = s_len[l];
                const uint32_t *src = a.sites + s_src[l];
                uint32_t n0 = 0, n1 = 0;
                if (l + 1 < nl) {
                    const uint32_t *nsrc = a.sites + s_src[l + 1];
                    if (tid < s_len[l + 1]) n0 = ld_pinned(nsrc + tid);
                    if (tid + SS_THREADS < s_len[l + 1]) n1 = ld_pinned(nsrc + tid + SS_THREADS);
                }
                if (tid < len) route(v0);
                if (tid + SS_THREADS < len) route(v1);
                for (uint32_t i = tid + 2u * SS_THREADS; i < len; i += SS_THREADS) route(__ldg(src + i));
                v0 = n0; v1 = n1;
            }
            __syncthreads();
            pf_to(2u);
            // ---- each warp its share of the windows, 32 at a time (lane = window) ----
            const uint32_t NR = (nwin + 31u) / 32u, RPW = (NR + SS_WARPS - 1u) / SS_WARPS;
            const uint32_t r1 = min(NR, (wq + 1u) * RPW);
            uint32_t carry = 0, first = 0;                             // (position + 1) of the last / first site of my range
            SparseAcc acc;
            acc.cnt = 0; acc.maxgap = 0; acc.sumsq = 0;
            for (uint32_t r = wq * RPW; r < r1; ++r) {
                const uint32_t w = r * 32u + lane;
                uint32_t c = 0;
                if (w < nwin) {
                    c = cnt32[w];
                    cnt32[w] = 0;
                }
                const uint32_t mx = __reduce_max_sync(0xffffffffu, c);
                if (mx == 0u) continue;
                const uint32_t win_s = slots_s + 2u * w, wbase = w << rsh;
                if (mx > RS_CAP) {                                     // the set goes through the merge; leave the slots empty
                    if (c) {
#pragma unroll
                        for (uint32_t q = 0; q < RS_CAP; ++q) sts_u16(win_s + 2u * RS_W * q, RS_EMPTY16);
                    }
                    if (lane == 0) s_ovf = 1;
                    continue;
                }
                // one instantiation per form: the "first site of the range" test of owner_gaps is a run-time one here (pred == 0
                // only while carry == 0); the kernel's code must stay inside the instruction cache (ncu: no_inst 23 % with
                // twelve instantiations)
#define RS_FORM(M) sparse_form<M, true>(win_s, wbase, lane, carry, first, acc, gaps, &s_ng)
                if (mx <= 4u) RS_FORM(4);
                else if (mx <= 8u) RS_FORM(8);
                else if (mx <= 12u) RS_FORM(12);
                else RS_FORM(16);
#undef RS_FORM
            }
            first = __reduce_max_sync(0xffffffffu, first);             // one lane set it
            // my warp's statistics travel with its first / last site through the barrier that ends the window walk: every
            // thread then adds up the four warps and the (at most three) gaps between their ranges itself -- no further
            // exchange for a set that is rejected
            {
                uint32_t wn = acc.cnt, wmx = acc.maxgap;
                unsigned long long wss = acc.sumsq;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    wn += __shfl_xor_sync(0xffffffffu, wn, d);
                    wmx = max(wmx, __shfl_xor_sync(0xffffffffu, wmx, d));
                    wss += __shfl_xor_sync(0xffffffffu, wss, d);
                }
                if (lane == 0) { s_first[wq] = first; s_last[wq] = carry; r32[wq] = wn; s_rmx[wq] = wmx; r64[wq] = wss; }
            }
            __syncthreads();
            pf_to(3u);
            routed = s_ovf == 0u;
            if (routed) {
                uint32_t prev = 0, fst = 0;
                entries = s_ng;
                for (uint32_t wv = 0; wv < SS_WARPS; ++wv) {
                    cnt += r32[wv]; maxgap = max(maxgap, s_rmx[wv]); sumsq += r64[wv];
                    if (!s_last[wv]) continue;
                    if (prev) {                                        // the gap between the ranges of two warps
                        const uint32_t d = s_first[wv] - prev;         // both carry the + 1
                        if (tid == 0) gaps[MS_PHYS(entries)] = d;      // read (Gini) behind a barrier
                        ++entries; ++cnt; maxgap = max(maxgap, d);
                        sumsq += (unsigned long long)d * d;
                    } else {
                        fst = s_first[wv];
                    }
                    prev = s_last[wv];
                }
                H = prev ? (unsigned long long)(prev - fst) : 0ull;
            }
        }
        if (!routed) {
        rt_dirty = can_route ? RT_WORDS : 0u;
        pf_to(2u);
        for (uint32_t l = 0; l < nl; ++l) {                                // every list is one sorted run, tail = 0xffffffff
            const uint32_t len = s_len[l], at = s_rb[l], end = s_rb[l + 1];
            const uint32_t *src = a.sites + s_src[l];
            for (uint32_t i = tid; i < end - at; i += SS_THREADS) buf[MS_PHYS(at + i)] = i < len ? src[i] : 0xffffffffu;
        }
        __syncthreads();
        pos = block_merge_runs(buf, gaps, s_rb, nl);                   // sorted positions (duplicates kept)
        pf_to(3u);
        gp = pos == buf ? gaps : buf;
        // gaps between distinct neighbours; duplicates become the sentinel, which sorts last
        cnt = 0; maxgap = 0; sumsq = 0;
        for (uint32_t i = tid; i < raw; i += SS_THREADS) {
            uint32_t g = 0xffffffffu;
            if (i + 1 < raw) {
                const uint32_t d = pos[MS_PHYS(i + 1)] - pos[MS_PHYS(i)];
                if (d) {
                    g = d; ++cnt; maxgap = max(maxgap, d);
                    sumsq += (unsigned long long)d * d;
                }
            }
            gp[MS_PHYS(i)] = g;
        }
        entries = raw;
        H = raw ? (unsigned long long)(pos[MS_PHYS(raw - 1)] - pos[0]) : 0ull;
        }
        pf_to(4u);
        // the three set-wide statistics: already totals after the routed front end, else one exchange (two barriers)
        uint32_t n = cnt, mx = maxgap;
        unsigned long long ssq = sumsq;
        if (!routed) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                n += __shfl_xor_sync(0xffffffffu, n, d);
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
                ssq += __shfl_xor_sync(0xffffffffu, ssq, d);
            }
            __syncthreads();
            if (lane == 0) { r32[wq] = n; s_rmx[wq] = mx; r64[wq] = ssq; }
            __syncthreads();
            n = 0; mx = 0; ssq = 0;
#pragma unroll
            for (int i = 0; i < SS_WARPS; ++i) { n += r32[i]; mx = max(mx, s_rmx[i]); ssq += r64[i]; }
        }
        const bool passed = a.interactive || mx <= a.max_fg_bind_dist;
        unsigned long long sum_prefix = 0;
        if (passed && n > 0) {
            if (can_route) rt_dirty = RT_WORDS;                       // the Gini step works in `pos` (narrowed below)
            // Gini's rank-weighted sum  sum_i g_(i) (n - i)  over the ascending gaps.  A counting sort by a coarse key
            // (gap >> shift, SS_BINS bins, in the buffer the positions no longer need) groups the gaps; inside a
            // bin (a handful of gaps) the rank is found by comparing with the bin's other members.  Exact for any
            // input; a bin too crowded for that (or a set too large for the buffer) takes the full merge sort.
            __syncthreads();
            uint32_t *cnt = pos, *grouped = pos + SS_BINS;            // [SS_BINS] counts -> bin ends; [n] gaps by bin
            const uint32_t words = MS_PHYS(a.sort_cap + SS_RUN_PAD);  // of one buffer
            const bool fits = words >= SS_BINS + n;                   // uniform
            uint32_t shift = 0, fullest = 0xffffffffu, before = 0;
            uint32_t mine[SS_BINS / SS_THREADS];
            if (fits) {
                while ((mx >> shift) >= SS_BINS) ++shift;
                for (uint32_t i = tid; i < SS_BINS; i += SS_THREADS) cnt[i] = 0;
                __syncthreads();
                for (uint32_t i = tid; i < entries; i += SS_THREADS) {
                    const uint32_t g = gp[MS_PHYS(i)];
                    if (g != 0xffffffffu) atomicAdd(&cnt[g >> shift], 1u);
                }
                __syncthreads();
                // exclusive scan of the counts (SS_BINS / SS_THREADS consecutive bins per thread) + the fullest bin
                uint32_t local = 0;
                fullest = 0;
#pragma unroll
                for (int q = 0; q < (int)(SS_BINS / SS_THREADS); ++q) {
                    mine[q] = cnt[tid * (SS_BINS / SS_THREADS) + q];
                    local += mine[q];
                    fullest = max(fullest, mine[q]);
                }
                uint32_t inc = local;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                    if ((tid & 31) >= (uint32_t)d) inc += o;
                }
                if ((tid & 31) == 31) r32[tid >> 5] = inc;
                __syncthreads();
                before = inc - local;
                for (uint32_t wq = 0; wq < (tid >> 5); ++wq) before += r32[wq];
                __syncthreads();
                fullest = block_reduce_ss(fullest, OpMaxU32(), 0u, r32);
            }
            if (fits && fullest <= SS_BIN_CROWD) {
                if (can_route && routed) rt_dirty = min(RT_WORDS, SS_BINS + n + 8u);   // bins + grouped gaps only
#pragma unroll
                for (int q = 0; q < (int)(SS_BINS / SS_THREADS); ++q) {
                    cnt[tid * (SS_BINS / SS_THREADS) + q] = before;      // start of the bin; the scatter turns it into its end
                    before += mine[q];
                }
                __syncthreads();
                for (uint32_t i = tid; i < entries; i += SS_THREADS) {
                    const uint32_t g = gp[MS_PHYS(i)];
                    if (g != 0xffffffffu) grouped[atomicAdd(&cnt[g >> shift], 1u)] = g;
                }
                __syncthreads();
                unsigned long long part = 0;
                for (uint32_t j = tid; j < n; j += SS_THREADS) {
                    const uint32_t g = grouped[j], b = g >> shift;
                    const uint32_t beg = b ? cnt[b - 1] : 0u, end = cnt[b];
                    uint32_t rank = beg;                                 // gaps that sort before this one
                    for (uint32_t k = beg; k < end; ++k) {
                        const uint32_t o = grouped[k];
                        rank += (o < g || (o == g && k < j)) ? 1u : 0u;
                    }
                    part += (unsigned long long)g * (n - rank);
                }
                sum_prefix = block_reduce_ss(part, OpAddU64(), 0ull, r64);
            } else {
                const uint32_t *sg = block_merge_sort(gp, pos, entries);   // ascending gaps
                unsigned long long part = 0;
                for (uint32_t i = tid; i < n; i += SS_THREADS) part += (unsigned long long)sg[MS_PHYS(i)] * (n - i);
                sum_prefix = block_reduce_ss(part, OpAddU64(), 0ull, r64);
            }
        }
        // The set's results are in registers of every thread: the last warp's first lane writes them out (double-precision
        // divisions and a square root) while warp 0 already publishes the next set -- no barrier at the end of a set: every
        // read of the set's shared-memory state lies before the last barrier above.
        if (tid == SS_THREADS - 32) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            double mean = nan, sd = nan, gini = nan, score = nan;
            if (n > 0) {
                mean = (double)H / (double)n;
                if (n > 1) {
                    const unsigned long long a_lo = (unsigned long long)n * ssq, a_hi = __umul64hi(n, ssq);
                    const unsigned long long b_lo = H * H, b_hi = __umul64hi(H, H);
                    const unsigned long long d_lo = a_lo - b_lo, d_hi = a_hi - b_hi - (a_lo < b_lo ? 1ull : 0ull);
                    const double num = (double)d_hi * 18446744073709551616.0 + (double)d_lo;
                    sd = sqrt(num / ((double)n * (double)(n - 1)));
                }
                if (passed) {
                    const unsigned long long fair = (H * (unsigned long long)n) / 2;
                    if (fair > 0) {
                        const long long two_area = (long long)(2 * sum_prefix) - (long long)H;
                        gini = (double)((long long)(2 * fair) - two_area) / (double)(2 * fair);
                    }
                }
                score = (mean * gini) / a.bg_dist_mean[s];
            }
            a.n_sites[s] = raw ? n + 1 : 0;
            a.max_dist[s] = mx;
            a.mean[s] = mean;
            a.stdev[s] = sd;
            a.gini[s] = gini;
            a.score[s] = score;
            a.status[s] = (uint8_t)(passed ? 1 : 0);
        }
    }
}

__global__ void k_score_all_ids(uint32_t *__restrict__ ids, uint32_t *__restrict__ n, uint32_t S)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) ids[s] = s;
    if (s == 0) *n = S;
}

// Splits the sets by raw site count: ids[0 .. n[0]) = sets for k_score_sort (raw <= cap), ids[S .. S + n[1]) = the rest.
__global__ void __launch_bounds__(256)
k_score_classify(ScoreArgs a, uint32_t *__restrict__ ids, uint32_t *__restrict__ n)
{
    const uint32_t s = blockIdx.x * 256u + threadIdx.x;
    if (s >= a.S) return;
    uint32_t raw = 0;
    for (uint32_t j = 0; j <= a.max_m; ++j) {
        const int32_t id = j < a.max_m ? a.sets[(uint64_t)s * a.max_m + j] : (int32_t)a.bounds_list;
        if (id < 0) continue;
        const uint32_t *row = a.toff + (uint64_t)id * (a.n_tiles + 1);
        raw += row[a.n_tiles] - row[0];
    }
    const uint32_t which = raw <= a.sort_cap ? 0u : 1u;
    // warp-aggregated append
    const uint32_t m = __match_any_sync(__activemask(), which);
    const uint32_t leader = __ffs(m) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(n + which, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    ids[(uint64_t)which * a.S + base + __popc(m & ((1u << lane) - 1u))] = s;
}

// Gini of an ascending-sorted gap array (the unbounded path, one set): sum_prefix by a block reduction
__global__ void __launch_bounds__(SC_THREADS)
k_gini_sorted(const uint32_t *__restrict__ gaps, uint32_t n, double *__restrict__ out)
{
    __shared__ unsigned long long s_red64[SC_WARPS + 1];
    unsigned long long part = 0, h = 0;
    for (uint32_t i = threadIdx.x; i < n; i += SC_THREADS) {
        part += (unsigned long long)gaps[i] * (n - i);                        // rank i+1 ascending: n - rank + 1
        h += gaps[i];
    }
    const unsigned long long sum_prefix = block_reduce(part, OpAddU64(), 0ull, s_red64);
    const unsigned long long H = block_reduce(h, OpAddU64(), 0ull, s_red64);
    if (threadIdx.x == 0) {
        double gini = __longlong_as_double(0x7ff8000000000000ll);
        const unsigned long long fair = (H * (unsigned long long)n) / 2;
        if (fair > 0) {
            const long long two_area = (long long)(2 * sum_prefix) - (long long)H;
            gini = (double)((long long)(2 * fair) - two_area) / (double)(2 * fair);
        }
        *out = gini;
    }
}

int swga_radix_sort_pairs(swga_ctx *ctx, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b,
                          uint64_t n, int bits, bool vals_are_index, int *result_in_a, cudaStream_t stream);

extern "C" uint32_t swga_score_tile_shift(void) { return DT_SHIFT; }
extern "C" uint32_t swga_score_max_hist_bins(void)
{
    return (227u * 1024u - 4u * (SC_L0_PHYS + SC_L1_WORDS) - 2048u) / 4u;
}

extern "C" int swga_score_sets(swga_ctx *ctx, const uint32_t *d_sites, const uint32_t *d_toff, uint32_t n_tiles,
                               uint32_t bounds_list, const int32_t *d_sets, uint32_t S, uint32_t max_m,
                               const double *d_bg_dist_mean, uint32_t max_fg_bind_dist, int interactive,
                               uint32_t *d_n_sites, uint32_t *d_max_dist, double *d_mean, double *d_stdev,
                               double *d_gini, double *d_score, uint8_t *d_status, void *stream)
{
    ARG_CHECK(ctx && d_toff && n_tiles > 0);
    if (S == 0) return SWGA_OK;
    ARG_CHECK(d_sets && d_bg_dist_mean && d_n_sites && d_max_dist && d_mean && d_stdev && d_gini && d_score && d_status);
    ARG_CHECK(max_m + 1 <= SC_MAX_LISTS);
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.sites = d_sites; a.toff = d_toff; a.n_tiles = n_tiles; a.bounds_list = bounds_list;
    a.sets = d_sets; a.S = S; a.max_m = max_m; a.bg_dist_mean = d_bg_dist_mean;
    a.max_fg_bind_dist = max_fg_bind_dist; a.interactive = interactive;
    const uint32_t cap = swga_score_max_hist_bins();
    // gaps up to max_fg_bind_dist can occur in a passing set; when that does not fit, sets whose largest
    // gap is >= cap are flagged (status bit1) for swga_gini_unbounded
    uint64_t want = interactive ? cap : (uint64_t)max_fg_bind_dist + 1;
    a.hist_bins = (uint32_t)(want < cap ? want : cap);
    a.n_sites = d_n_sites; a.max_dist = d_max_dist; a.mean = d_mean; a.stdev = d_stdev; a.gini = d_gini;
    a.score = d_score; a.status = d_status;
    // sparse sets first (k_score_sort: gather + sort in shared memory), then the rest (k_score: bitmap tiles)
    a.sort_cap = ctx->score_sort_cap;
    TRY(swga_ensure(ctx, ctx->score_tmp, 256 + 8ull * S));
    a.work = (uint32_t *)((uint8_t *)ctx->score_tmp.p + 128);             // [2] work counters, [2] list lengths
    uint32_t *n_ids = a.work + 2, *ids = (uint32_t *)((uint8_t *)ctx->score_tmp.p + 256);
    CUDA_TRY(cudaMemsetAsync(a.work, 0, 16, as_stream(stream)));
    if (a.sort_cap) {
        SWGA_NOTE_LAUNCH(); k_score_classify<<<grid_for(S, 256), 256, 0, as_stream(stream)>>>(a, ids, n_ids);
        CUDA_TRY(cudaGetLastError());
        a.ids = ids; a.n_ids = n_ids;
        a.sort_route = ctx->score_sort_impl == 0 ? 1u : 0u;
        const size_t bufb = 4ull * MS_PHYS(a.sort_cap + SS_RUN_PAD);
        const size_t smem_s = bufb + (a.sort_route ? std::max<size_t>(bufb, RS_BYTES) : bufb);
        CUDA_TRY(cudaFuncSetAttribute(k_score_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(16, (226 * 1024) / (smem_s + 1280)));
        const uint32_t grid_s = std::min<uint32_t>(S, (uint32_t)ctx->sm_count * per_sm);
        SWGA_NOTE_LAUNCH(); k_score_sort<<<grid_s, SS_THREADS, smem_s, as_stream(stream)>>>(a);
        CUDA_TRY(cudaGetLastError());
    }
    // the rest: one warp per fine tile (k_score_dense)
    const bool routed = ctx->score_dense_impl == 0;
    // routed: alignment slack + counters and slots + histogram with its overflow bin and 32 dummy words; the histogram of
    // the dense kernel is what fits beside them (sets with a larger gap are flagged for swga_gini_unbounded as always)
    if (routed) {
        const uint32_t fit = (uint32_t)((227u * 1024u - 2048u - 128u - DT_WARPS * DT_RT_BYTES) / 4u - 33u);   // 40,127
        a.hist_bins = std::min(a.hist_bins, fit);
    }
    const size_t smem = routed ? 128ull + (size_t)DT_WARPS * DT_RT_BYTES + 4ull * (a.hist_bins + 33u)
                               : 4ull * DT_WARPS * DT_WORDS + 2ull * DT_WARPS * DT_LIST_CAP + 4ull * a.hist_bins;
    CUDA_TRY(cudaFuncSetAttribute(routed ? k_score_dense<true> : k_score_dense<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t grid = S < (uint32_t)ctx->sm_count ? S : (uint32_t)ctx->sm_count;
    if (a.sort_cap) {
        a.ids = ids + S;
        a.n_ids = n_ids + 1;
    } else {                                                   // no sparse kernel: every set is "dense"
        SWGA_NOTE_LAUNCH(); k_score_all_ids<<<grid_for(S, 256), 256, 0, as_stream(stream)>>>(ids + S, n_ids + 1, S);
        a.ids = ids + S;
        a.n_ids = n_ids + 1;
    }
    SWGA_NOTE_LAUNCH();
    if (routed) k_score_dense<true><<<grid, DT_THREADS, smem, as_stream(stream)>>>(a);
    else k_score_dense<false><<<grid, DT_THREADS, smem, as_stream(stream)>>>(a);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_ctx_set_score_sort_impl(swga_ctx *ctx, int impl)
{
    ARG_CHECK(ctx && (impl == 0 || impl == 1));
    ctx->score_sort_impl = impl;
    return SWGA_OK;
}

extern "C" int swga_ctx_set_score_dense_impl(swga_ctx *ctx, int impl)
{
    ARG_CHECK(ctx && (impl == 0 || impl == 1));
    ctx->score_dense_impl = impl;
    return SWGA_OK;
}

extern "C" int swga_ctx_set_score_sort_cap(swga_ctx *ctx, uint32_t cap)
{
    ARG_CHECK(ctx && (cap == 0 || ((cap & (cap - 1)) == 0 && cap >= 64 && cap <= 16384)));
    ctx->score_sort_cap = cap;
    return SWGA_OK;
}

// Gini of ONE set without a bound on the gap size (interactive `swga score`, or max_fg_bind_dist beyond
// the histogram capacity): dump the gaps, radix-sort them, rank-weighted sum.  Synchronises.
extern "C" int swga_gini_unbounded(swga_ctx *ctx, const uint32_t *d_sites, const uint32_t *d_toff, uint32_t n_tiles,
                                   uint32_t bounds_list, const int32_t *d_set, uint32_t max_m, uint64_t max_gaps,
                                   double *h_gini, void *stream_)
{
    ARG_CHECK(ctx && d_toff && d_set && h_gini && n_tiles > 0 && max_m + 1 <= SC_MAX_LISTS);
    cudaStream_t stream = as_stream(stream_);
    const size_t n = max_gaps ? max_gaps : 1;
    TRY(swga_ensure(ctx, ctx->sort_keys[0], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_keys[1], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_vals[0], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_vals[1], n * 4));
    TRY(swga_ensure(ctx, ctx->score_tmp, 256));
    uint8_t *tmp = (uint8_t *)ctx->score_tmp.p;               // [0,8) bg  [8,16) gini  [16,20) gap_count, then outputs
    CUDA_TRY(cudaMemsetAsync(tmp, 0, 256, stream));
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.sites = d_sites; a.toff = d_toff; a.n_tiles = n_tiles; a.bounds_list = bounds_list;
    a.sets = d_set; a.S = 1; a.max_m = max_m; a.bg_dist_mean = (const double *)tmp;
    a.max_fg_bind_dist = 0xffffffffu; a.interactive = 1; a.hist_bins = 0;
    a.n_sites = (uint32_t *)(tmp + 32); a.max_dist = (uint32_t *)(tmp + 40); a.mean = (double *)(tmp + 48);
    a.stdev = (double *)(tmp + 56); a.gini = (double *)(tmp + 64); a.score = (double *)(tmp + 72);
    a.status = tmp + 80;
    a.gap_out = (uint32_t *)ctx->sort_keys[0].p;
    a.gap_count = (uint32_t *)(tmp + 16);
    const size_t smem = 4ull * (SC_L0_PHYS + SC_L1_WORDS);
    CUDA_TRY(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SWGA_NOTE_LAUNCH(); k_score<<<1, SC_THREADS, smem, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    uint32_t ng = 0;
    CUDA_TRY(cudaMemcpyAsync(&ng, tmp + 16, 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (ng > max_gaps) {
        swga_set_error("gap buffer too small: %u gaps > %llu", ng, (unsigned long long)max_gaps);
        return SWGA_E_OVERFLOW;
    }
    int in_a = 1;
    TRY(swga_radix_sort_pairs(ctx, (uint32_t *)ctx->sort_keys[0].p, (uint32_t *)ctx->sort_vals[0].p,
                              (uint32_t *)ctx->sort_keys[1].p, (uint32_t *)ctx->sort_vals[1].p, ng, 32, true, &in_a,
                              stream));
    SWGA_NOTE_LAUNCH(); k_gini_sorted<<<1, SC_THREADS, 0, stream>>>((const uint32_t *)ctx->sort_keys[in_a ? 0 : 1].p, ng,
                                                (double *)(tmp + 8));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_gini, tmp + 8, 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return SWGA_OK;
}
