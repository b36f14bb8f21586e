// index.cu -- kernel (c): CSR position index of a (foreground) genome by counting sort, and the
// O(1) locate built on it.
//
// Results reproduced (reference = eclarke/swga):
//   swga/locate.py:39-51   binding_sites: per record, 0-based starts of the primer (ascending) followed
//                          by those of its reverse complement (ascending); overlapping hits included
//   swga/locate.py:76-90   substr_indices: exact byte match of the upper-cased primer against the raw
//                          genome => only windows of upper-case A/C/G/T inside ONE record can match
//                          (strict mask, SWGA_MASK_STRICT + record breaks)
//
// Index for one k:  offsets[4^k + 2] (exclusive scan of the per-code histogram; slot 4^k is the
// sentinel bucket of invalid windows), positions[n_valid] = linearised start positions, ascending
// inside each code bucket (stable LSD radix sort of (code, position) pairs, 8-bit digits).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three launches: chunk sums, scan of sums, scan + offset)
// ------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_CHUNK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += o;
    }
    return v;
}

// exclusive scan across the block of one value per thread; returns the exclusive prefix, *total = sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const uint32_t incl = warp_incl_scan(v);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? wsum[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        wsum[lane] = wi - w;
        if (lane == 31) tot = wi;
    }
    __syncthreads();
    const uint32_t r = wsum[warp] + incl - v;
    if (total) *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ partial)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const uint64_t j = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
        if (j < n) s += in[j];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
k_scan_partials(uint32_t *__restrict__ partial, uint32_t n_part, uint32_t *__restrict__ grand_total)
{
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_part; base += 1024) {
        const uint32_t j = base + threadIdx.x;
        const uint32_t v = j < n_part ? partial[j] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, &tot);
        if (j < n_part) partial[j] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint32_t *__restrict__ partial,
             uint32_t *__restrict__ out)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t run = block_excl_scan(s, nullptr) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// out[0..n) = exclusive scan of in[0..n); out may alias in; *d_total (optional, device) = sum.
int swga_exclusive_scan(swga_ctx *ctx, const uint32_t *d_in, uint64_t n, uint32_t *d_out, uint32_t *d_total,
                        cudaStream_t stream)
{
    if (n == 0) {
        if (d_total) CUDA_TRY(cudaMemsetAsync(d_total, 0, 4, stream));
        return SWGA_OK;
    }
    const uint32_t n_part = (uint32_t)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    TRY(swga_ensure(ctx, ctx->scratch, (size_t)n_part * 4 + 256));
    uint32_t *partial = (uint32_t *)ctx->scratch.p;
    SWGA_NOTE_LAUNCH(); k_scan_reduce<<<n_part, SCAN_THREADS, 0, stream>>>(d_in, n, partial);
    SWGA_NOTE_LAUNCH(); k_scan_partials<<<1, 1024, 0, stream>>>(partial, n_part, d_total);
    SWGA_NOTE_LAUNCH(); k_scan_apply<<<n_part, SCAN_THREADS, 0, stream>>>(d_in, n, partial, d_out);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

// ------------------------------------------------------------------------------------------------
// keys: forward code of the k-window starting at every position, or the sentinel 4^k when the window
// is not wholly inside one record / contains a byte that cannot match (strict break mask)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_index_keys(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk, uint64_t n, int k,
             uint32_t *__restrict__ keys, uint32_t *__restrict__ hist)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t p0 = t * 32;
    if (p0 >= n) return;
    const uint32_t w0 = packed[2 * t], w1 = packed[2 * t + 1], w2 = packed[2 * t + 2];
    const uint32_t b0 = brk[t], b1 = brk[t + 1];
    const uint32_t sentinel = 1u << (2 * k);
    const int lim = (int)min((uint64_t)32, n - p0);
    for (int o = 0; o < lim; ++o) {
        const uint32_t W = o < 16 ? __funnelshift_l(w1, w0, 2 * o) : __funnelshift_l(w2, w1, 2 * (o - 16));
        const uint32_t B = __funnelshift_l(b1, b0, o);
        const bool ok = (__clz(B) + 1) >= k;                     // break-free run from p covers k bases
        const uint32_t key = ok ? (W >> (32 - 2 * k)) : sentinel;
        keys[p0 + o] = key;
        atomicAdd(hist + key, 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, value) pairs, 8-bit digits
// ------------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS)
k_rs_upsweep(const uint32_t *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ tile_hist,
             uint32_t n_tiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint64_t j = base + (uint64_t)i * RS_THREADS + threadIdx.x;
        if (j < n) atomicAdd(&h[(keys[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    tile_hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_downsweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
               uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n, int shift,
               const uint32_t *__restrict__ tile_base, uint32_t n_tiles)
{
    __shared__ uint32_t wcount[RS_WARPS][257];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    // each warp owns a contiguous run of 32*RS_ITEMS keys; round j takes keys [j*32, j*32+32) of it,
    // so (round, lane) order == memory order and ranks are stable
    const uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint64_t idx = wbase + j * 32 + lane;
        const bool valid = idx < n;
        key[j] = valid ? keys_in[idx] : 0u;
        const uint32_t d = valid ? ((key[j] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        const uint32_t base = wcount[warp][d];
        __syncwarp();
        if (before == 0) wcount[warp][d] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + before;
    }
    __syncthreads();
    {   // thread d turns per-warp counts of digit d into global output bases
        const int d = threadIdx.x;
        uint32_t run = tile_base[(uint64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint64_t idx = wbase + j * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[j] >> shift) & 255u;
            const uint32_t o = wcount[warp][d] + rank[j];
            keys_out[o] = key[j];
            vals_out[o] = vals_in ? vals_in[idx] : (uint32_t)idx;   // first pass: value = position
        }
    }
}

// Sorts n (key, value) pairs by the low `bits` bits of the key, stably.  vals_in == NULL means
// value = index.  Ping-pongs between (keys_a, vals_a) and (keys_b, vals_b); returns which pair holds
// the result in *result_in_a.
int swga_radix_sort_pairs(swga_ctx *ctx, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b,
                          uint64_t n, int bits, bool vals_are_index, int *result_in_a, cudaStream_t stream)
{
    *result_in_a = 1;
    if (n == 0) return SWGA_OK;
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    DevBuf &hb = ctx->sort_hist;
    TRY(swga_ensure(ctx, hb, (size_t)256 * n_tiles * 4));
    uint32_t *th = (uint32_t *)hb.p;
    bool in_a = true, first = true;
    for (int shift = 0; shift < bits; shift += 8) {
        uint32_t *ki = in_a ? keys_a : keys_b, *vi = in_a ? vals_a : vals_b;
        uint32_t *ko = in_a ? keys_b : keys_a, *vo = in_a ? vals_b : vals_a;
        SWGA_NOTE_LAUNCH(); k_rs_upsweep<<<n_tiles, RS_THREADS, 0, stream>>>(ki, n, shift, th, n_tiles);
        TRY(swga_exclusive_scan(ctx, th, (uint64_t)256 * n_tiles, th, nullptr, stream));
        SWGA_NOTE_LAUNCH(); k_rs_downsweep<<<n_tiles, RS_THREADS, 0, stream>>>(ki, (first && vals_are_index) ? nullptr : vi, ko, vo, n,
                                                           shift, th, n_tiles);
        CUDA_TRY(cudaGetLastError());
        in_a = !in_a;
        first = false;
    }
    *result_in_a = in_a ? 1 : 0;
    return SWGA_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: index build, locate, per-primer union, tile offsets
// ------------------------------------------------------------------------------------------------
extern "C" uint64_t swga_index_offsets_words(int k) { return (1ull << (2 * k)) + 2; }

extern "C" int swga_build_index(swga_ctx *ctx, const uint32_t *d_packed, const uint32_t *d_brk, uint64_t n, int k,
                                uint32_t *d_offsets, uint32_t *d_positions, uint64_t *n_valid, void *stream_)
{
    ARG_CHECK(ctx && d_packed && d_brk && d_offsets && (d_positions || n == 0) && n_valid);
    ARG_CHECK(k >= 2 && k <= SWGA_KMAX_LIMIT && n < (1ull << 32));
    cudaStream_t stream = as_stream(stream_);
    const uint64_t nbins = (1ull << (2 * k)) + 1;               // + sentinel bucket
    CUDA_TRY(cudaMemsetAsync(d_offsets, 0, (nbins + 1) * 4, stream));
    *n_valid = 0;
    if (n == 0) return SWGA_OK;
    TRY(swga_ensure(ctx, ctx->sort_keys[0], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_keys[1], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_vals[0], n * 4));
    TRY(swga_ensure(ctx, ctx->sort_vals[1], n * 4));
    uint32_t *ka = (uint32_t *)ctx->sort_keys[0].p, *kb = (uint32_t *)ctx->sort_keys[1].p;
    uint32_t *va = (uint32_t *)ctx->sort_vals[0].p, *vb = (uint32_t *)ctx->sort_vals[1].p;
    SWGA_NOTE_LAUNCH(); k_index_keys<<<grid_for((n + 31) / 32, 256), 256, 0, stream>>>(d_packed, d_brk, n, k, ka, d_offsets);
    CUDA_TRY(cudaGetLastError());
    // offsets = exclusive scan of the histogram (nbins entries) with the grand total in slot nbins
    TRY(swga_exclusive_scan(ctx, d_offsets, nbins, d_offsets, d_offsets + nbins, stream));
    int in_a = 1;
    TRY(swga_radix_sort_pairs(ctx, ka, va, kb, vb, n, 2 * k + 1, true, &in_a, stream));
    uint32_t nv = 0;
    CUDA_TRY(cudaMemcpyAsync(&nv, d_offsets + (nbins - 1), 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    *n_valid = nv;
    if (nv) CUDA_TRY(cudaMemcpyAsync(d_positions, in_a ? va : vb, (size_t)nv * 4, cudaMemcpyDeviceToDevice, stream));
    return SWGA_OK;
}

__global__ void k_locate_counts(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ codes, uint32_t q,
                                uint32_t *__restrict__ counts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < q) counts[i] = offsets[codes[i] + 1] - offsets[codes[i]];
}

__global__ void __launch_bounds__(256)
k_locate_gather(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ positions,
                const uint32_t *__restrict__ codes, const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out)
{
    const uint32_t q = blockIdx.x;
    const uint32_t lo = offsets[codes[q]], hi = offsets[codes[q] + 1];
    uint32_t *dst = out + out_off[q];
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i - lo] = positions[i];
}

extern "C" int swga_locate_counts(swga_ctx *ctx, int k, const uint32_t *d_offsets, const uint32_t *d_codes, uint32_t q,
                                  uint32_t *d_counts, void *stream)
{
    ARG_CHECK(ctx && d_offsets && (d_codes || q == 0) && (d_counts || q == 0) && k >= 2 && k <= SWGA_KMAX_LIMIT);
    if (q == 0) return SWGA_OK;
    SWGA_NOTE_LAUNCH(); k_locate_counts<<<grid_for(q, 256), 256, 0, as_stream(stream)>>>(d_offsets, d_codes, q, d_counts);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

extern "C" int swga_locate_gather(swga_ctx *ctx, int k, const uint32_t *d_offsets, const uint32_t *d_positions,
                                  const uint32_t *d_codes, uint32_t q, const uint64_t *d_out_off, uint32_t *d_out,
                                  void *stream)
{
    ARG_CHECK(ctx && d_offsets && (d_codes || q == 0) && k >= 2 && k <= SWGA_KMAX_LIMIT);
    if (q == 0) return SWGA_OK;
    ARG_CHECK(d_out_off && d_out);
    SWGA_NOTE_LAUNCH(); k_locate_gather<<<q, 256, 0, as_stream(stream)>>>(d_offsets, d_positions, d_codes, d_out_off, d_out);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

// union of the forward and reverse-complement site lists of each primer (sorted, unique, linear
// coordinates) -- what swga/locate.py:22-36 feeds to set() for one primer, minus the record bounds.
// The two lists are disjoint unless the primer is its own reverse complement (then they are equal).
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t v)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_union2(const uint32_t *__restrict__ hits, const uint64_t *__restrict__ hit_off, const uint8_t *__restrict__ is_pal,
         const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out)
{
    const uint32_t p = blockIdx.x;
    const uint32_t *A = hits + hit_off[2 * p], *B = hits + hit_off[2 * p + 1];
    const uint32_t na = (uint32_t)(hit_off[2 * p + 1] - hit_off[2 * p]);
    const uint32_t nb = is_pal[p] ? 0u : (uint32_t)(hit_off[2 * p + 2] - hit_off[2 * p + 1]);
    uint32_t *dst = out + out_off[p];
    for (uint32_t i = threadIdx.x; i < na; i += blockDim.x) dst[i + lower_bound_u32(B, nb, A[i])] = A[i];
    for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) dst[j + lower_bound_u32(A, na, B[j])] = B[j];
}

extern "C" int swga_union_strands(swga_ctx *ctx, const uint32_t *d_hits, const uint64_t *d_hit_off,
                                  const uint8_t *d_is_pal, uint32_t n_primers, const uint64_t *d_out_off,
                                  uint32_t *d_out, void *stream)
{
    ARG_CHECK(ctx);
    if (n_primers == 0) return SWGA_OK;
    ARG_CHECK(d_hits && d_hit_off && d_is_pal && d_out_off && d_out);
    SWGA_NOTE_LAUNCH(); k_union2<<<n_primers, 256, 0, as_stream(stream)>>>(d_hits, d_hit_off, d_is_pal, d_out_off, d_out);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

// toff[l * (n_tiles + 1) + t] = index (into the concatenated site array) of the first site of list l
// that is >= t << tile_shift
__global__ void k_tile_offsets(const uint32_t *__restrict__ sites, const uint64_t *__restrict__ site_off, uint32_t n_lists,
                               int tile_shift, uint32_t n_tiles, uint32_t *__restrict__ toff)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_lists * (n_tiles + 1)) return;
    const uint32_t l = (uint32_t)(i / (n_tiles + 1)), t = (uint32_t)(i % (n_tiles + 1));
    const uint64_t lo = site_off[l];
    const uint32_t len = (uint32_t)(site_off[l + 1] - lo);
    const uint64_t bound = (uint64_t)t << tile_shift;
    const uint32_t lb = bound > 0xffffffffull ? len : lower_bound_u32(sites + lo, len, (uint32_t)bound);
    toff[i] = (uint32_t)(lo + lb);
}

extern "C" int swga_tile_offsets(swga_ctx *ctx, const uint32_t *d_sites, const uint64_t *d_site_off, uint32_t n_lists,
                                 int tile_shift, uint32_t n_tiles, uint32_t *d_toff, void *stream)
{
    ARG_CHECK(ctx && d_site_off && d_toff && tile_shift >= 5 && tile_shift < 32);
    if (n_lists == 0) return SWGA_OK;
    SWGA_NOTE_LAUNCH(); k_tile_offsets<<<grid_for((uint64_t)n_lists * (n_tiles + 1), 256), 256, 0, as_stream(stream)>>>(
        d_sites, d_site_off, n_lists, tile_shift, n_tiles, d_toff);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}
