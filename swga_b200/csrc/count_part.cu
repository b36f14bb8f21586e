// count_part.cu -- kernel (b), partitioned form: the same additive "top-k + tails" counts as k_count
// (count.cu), but the updates of the top table T_kmax are taken out of the L2 atomic unit.
//
// Round-1 profile (profiles/r01_summary.md): k_count sits at the red.global ceiling (~190-210 G/s, LTS
// 87 %, 8.3 B/base of DRAM traffic from evicted table sectors) while shared-memory atomics measure
// ~2600 G/s.  So: (1) k_bucket_hist counts the kmax-mers per PREFIX (top `pb` bits of the code),
// (2) k_partition scatters the low `lb = 2 kmax - pb` bits of every kmax-mer (uint16) into its
// prefix bucket through a shared-memory counting sort of a 32 Ki-start tile (coalesced bucket runs),
// (3) k_bucket_count histograms each bucket in a 2^lb-word shared-memory table and adds the non-zero
// bins to T_kmax.  Groups of 32 starts that touch a break (record end, 'N', buffer end) keep the
// direct path of k_count, so results are identical by construction.
//
// Reference results reproduced: as count.cu (Bank.cpp:889-902,1042-1070; OAHash.cpp:142-154).
#include "common.cuh"

#define PT_THREADS 512
#define PT_ROUNDS 1
#define PT_GROUPS (PT_THREADS * PT_ROUNDS)          // groups of 32 starts per tile
#define PT_TILE (PT_GROUPS * 32)                    // 16384 starts per tile
#define PT_LB 14                                    // low bits kept per element (uint16), 16384-bin tables
#define PT_MAX_PB 12                                // at most 4096 buckets
#define BC_THREADS 512
#define BC_ITEM (1u << 20)                          // elements per work item of k_bucket_count

struct Group {
    uint32_t w0, w1, w2;
    bool fast;
};

__device__ __forceinline__ Group load_group(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk,
                                            uint64_t g, uint64_t n_starts, int kmax)
{
    Group r;
    r.fast = false;
    r.w0 = r.w1 = r.w2 = 0;
    const uint64_t p0 = g * 32;
    if (p0 >= n_starts) return r;
    const uint2 w01 = __ldg(reinterpret_cast<const uint2 *>(packed) + g);
    r.w0 = w01.x; r.w1 = w01.y;
    r.w2 = __ldg(packed + 2 * g + 2);
    const uint32_t b0 = __ldg(brk + g), b1 = __ldg(brk + g + 1);
    const int need = kmax - 2;
    const uint32_t tail = need > 0 ? (b1 >> (32 - need)) : 0u;
    r.fast = (p0 + 32 <= n_starts) && ((b0 | tail) == 0u);
    return r;
}

__device__ __forceinline__ uint32_t code_at(const Group &g, int o, int sh)
{
    return (o < 16 ? __funnelshift_l(g.w1, g.w0, 2 * o) : __funnelshift_l(g.w2, g.w1, 2 * (o - 16))) >> sh;
}

// direct path for a group that touches a break or the end of the owned range (same as k_count's slow path)
__device__ __forceinline__ void slow_group(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk,
                                           uint64_t g, uint64_t n_starts, int kmin, int kmax,
                                           uint32_t *__restrict__ tables, const TableOffs &offs)
{
    const uint64_t p0 = g * 32;
    if (p0 >= n_starts) return;
    const uint32_t w0 = packed[2 * g], w1 = packed[2 * g + 1], w2 = packed[2 * g + 2];
    const uint32_t b0 = brk[g], b1 = brk[g + 1];
    const int lim = (int)min((uint64_t)32, n_starts - p0);
    for (int o = 0; o < lim; ++o) {
        const uint32_t W = o < 16 ? __funnelshift_l(w1, w0, 2 * o) : __funnelshift_l(w2, w1, 2 * (o - 16));
        const uint32_t B = __funnelshift_l(b1, b0, o);
        const int v = min(kmax, __clz(B) + 1);
        if (v >= kmin) atomicAdd(tables + offs.off[v] + (W >> (32 - 2 * v)), 1u);
    }
}

// ---- (1) per-prefix counts of the fast groups ------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
k_bucket_hist(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk, uint64_t n_starts, int kmax,
              int pb, uint32_t *__restrict__ bucket_cnt)
{
    extern __shared__ uint32_t h[];
    const uint32_t nb = 1u << pb;
    for (uint32_t i = threadIdx.x; i < nb; i += PT_THREADS) h[i] = 0;
    __syncthreads();
    const uint64_t n_groups = (n_starts + 31) / 32;
    const int sh = 32 - pb;
    for (uint64_t g = blockIdx.x * (uint64_t)PT_THREADS + threadIdx.x; g < n_groups; g += (uint64_t)gridDim.x * PT_THREADS) {
        const Group G = load_group(packed, brk, g, n_starts, kmax);
        if (G.fast) {
#pragma unroll
            for (int o = 0; o < 32; ++o) atomicAdd(&h[code_at(G, o, sh)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += PT_THREADS)
        if (h[i]) atomicAdd(bucket_cnt + i, h[i]);
}

// ---- (2) scatter the low bits into prefix buckets --------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
k_partition(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ brk, uint64_t n_starts, int kmin,
            int kmax, int pb, uint32_t *__restrict__ tables, TableOffs offs, uint32_t *__restrict__ cursor,
            uint16_t *__restrict__ bucket_data, uint32_t n_tiles)
{
    extern __shared__ uint32_t sm[];
    const uint32_t nb = 1u << pb;
    uint32_t *cnt = sm;                      // [nb] per-tile bucket counts (then reused as running ranks)
    uint32_t *start = cnt + nb;              // [nb] exclusive scan of cnt
    uint32_t *goff = start + nb;             // [nb] (reserved global offset) - (start in tile)
    uint32_t *staging = goff + nb;           // [PT_TILE] (bucket << 16) | low bits, bucket-sorted
    __shared__ uint32_t s_wsum[PT_THREADS / 32];
    const int lb = 2 * kmax - pb;
    const int sh_code = 32 - 2 * kmax;
    const uint32_t lmask = (1u << lb) - 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = tid; i < nb; i += PT_THREADS) cnt[i] = 0;
        __syncthreads();
        Group G[PT_ROUNDS];
        uint32_t rk[PT_ROUNDS][16];
        // ---- pass A: rank of every element inside its bucket (for this tile) ----
#pragma unroll
        for (int r = 0; r < PT_ROUNDS; ++r) {
            const uint64_t g = (uint64_t)tile * PT_GROUPS + (uint64_t)r * PT_THREADS + tid;
            G[r] = load_group(packed, brk, g, n_starts, kmax);
            if (G[r].fast) {
#pragma unroll
                for (int o = 0; o < 32; o += 2) {
                    const uint32_t c0 = code_at(G[r], o, sh_code), c1 = code_at(G[r], o + 1, sh_code);
                    const uint32_t r0 = atomicAdd(&cnt[c0 >> lb], 1u);
                    const uint32_t r1 = atomicAdd(&cnt[c1 >> lb], 1u);
                    rk[r][o >> 1] = r0 | (r1 << 16);
                }
            } else {
                slow_group(packed, brk, g, n_starts, kmin, kmax, tables, offs);
            }
        }
        __syncthreads();
        // ---- exclusive scan of cnt -> start; reserve global space per bucket ----
        {
            const uint32_t per = nb / PT_THREADS ? nb / PT_THREADS : 1;      // nb is a power of two
            uint32_t local = 0;
            const uint32_t b0 = tid * per;
            if (b0 < nb)
                for (uint32_t i = 0; i < per; ++i) local += cnt[b0 + i];
            uint32_t inc = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (uint32_t)d) inc += o;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = lane < PT_THREADS / 32 ? s_wsum[lane] : 0, vi = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, vi, d);
                    if (lane >= (uint32_t)d) vi += o;
                }
                if (lane < PT_THREADS / 32) s_wsum[lane] = vi - v;
            }
            __syncthreads();
            uint32_t run = s_wsum[warp] + inc - local;
            if (b0 < nb)
                for (uint32_t i = 0; i < per; ++i) {
                    const uint32_t c = cnt[b0 + i];
                    start[b0 + i] = run;
                    goff[b0 + i] = (c ? atomicAdd(cursor + b0 + i, c) : 0u) - run;
                    run += c;
                }
        }
        __syncthreads();
        // ---- pass B: place the low bits at start[bucket] + rank ----
#pragma unroll
        for (int r = 0; r < PT_ROUNDS; ++r) {
            if (G[r].fast) {
#pragma unroll
                for (int o = 0; o < 32; ++o) {
                    const uint32_t c = code_at(G[r], o, sh_code);
                    const uint32_t rank = (rk[r][o >> 1] >> ((o & 1) * 16)) & 0xffffu;
                    staging[start[c >> lb] + rank] = ((c >> lb) << 16) | (c & lmask);
                }
            }
        }
        __syncthreads();
        // ---- write out: slot j of the bucket-sorted tile goes to goff[bucket] + j (runs are contiguous) ----
        {
            const uint32_t total = start[nb - 1] + cnt[nb - 1];
            for (uint32_t j = tid; j < total; j += PT_THREADS) {
                const uint32_t e = staging[j];
                bucket_data[goff[e >> 16] + j] = (uint16_t)e;
            }
        }
        __syncthreads();
    }
}

// ---- work items of phase 3: ceil(size_b / BC_ITEM) per bucket ----------------------------------------
__global__ void k_item_prefix(const uint32_t *__restrict__ bucket_base, uint32_t nb, uint32_t *__restrict__ item_prefix)
{
    // single block; nb <= 4096
    __shared__ uint32_t s[4097];
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        const uint32_t sz = bucket_base[b + 1] - bucket_base[b];
        s[b] = (sz + BC_ITEM - 1) / BC_ITEM;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t c = s[b];
            item_prefix[b] = run;
            run += c;
        }
        item_prefix[nb] = run;
    }
}

// ---- (3) shared-memory histogram of each bucket, added to the top table ------------------------------
__global__ void __launch_bounds__(BC_THREADS)
k_bucket_count(const uint16_t *__restrict__ bucket_data, const uint32_t *__restrict__ bucket_base,
               const uint32_t *__restrict__ item_prefix, uint32_t nb, int lb, uint32_t *__restrict__ top,
               uint32_t *__restrict__ work_counter)
{
    extern __shared__ uint32_t tab[];
    __shared__ uint32_t s_item;
    const uint32_t bins = 1u << lb;
    const uint32_t n_items = item_prefix[nb];
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1u);
        for (uint32_t i = threadIdx.x; i < bins; i += BC_THREADS) tab[i] = 0;
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
        const uint32_t beg = bucket_base[b] + part * BC_ITEM;
        const uint32_t end = min(bucket_base[b + 1], beg + BC_ITEM);
        // head up to 16-byte alignment, vector body, tail
        const uint32_t abeg = min(end, (beg + 7u) & ~7u);
        const uint32_t aend = abeg + ((end - abeg) & ~7u);
        for (uint32_t i = beg + threadIdx.x; i < abeg; i += BC_THREADS) atomicAdd(&tab[bucket_data[i]], 1u);
        const uint4 *v = reinterpret_cast<const uint4 *>(bucket_data + abeg);
        const uint32_t nv = (aend - abeg) >> 3;
        for (uint32_t i = threadIdx.x; i < nv; i += BC_THREADS) {
            const uint4 x = __ldg(v + i);
            atomicAdd(&tab[x.x & 0xffffu], 1u); atomicAdd(&tab[x.x >> 16], 1u);
            atomicAdd(&tab[x.y & 0xffffu], 1u); atomicAdd(&tab[x.y >> 16], 1u);
            atomicAdd(&tab[x.z & 0xffffu], 1u); atomicAdd(&tab[x.z >> 16], 1u);
            atomicAdd(&tab[x.w & 0xffffu], 1u); atomicAdd(&tab[x.w >> 16], 1u);
        }
        for (uint32_t i = aend + threadIdx.x; i < end; i += BC_THREADS) atomicAdd(&tab[bucket_data[i]], 1u);
        __syncthreads();
        uint32_t *dst = top + ((uint64_t)b << lb);
        for (uint32_t i = threadIdx.x; i < bins; i += BC_THREADS) {
            const uint32_t c = tab[i];
            if (c) atomicAdd(dst + i, c);
        }
        __syncthreads();
    }
}

int swga_exclusive_scan(swga_ctx *ctx, const uint32_t *d_in, uint64_t n, uint32_t *d_out, uint32_t *d_total,
                        cudaStream_t stream);

int swga_count_partition_supported(uint64_t n_starts, int kmax, int *ok)
{
    const int pb = 2 * kmax - PT_LB;
    *ok = !(pb < 1 || pb > PT_MAX_PB || n_starts < (1u << 22) || n_starts >= 0xfff00000ull);
    return SWGA_OK;
}

// Partitioned accumulate.  Returns SWGA_OK and sets *done = 1 when it handled the request, *done = 0
// when the shape is outside its range (the caller then uses the direct kernel).
int swga_count_accumulate_partitioned(swga_ctx *ctx, const uint32_t *d_packed, const uint32_t *d_brk, uint64_t n_starts,
                                      int kmin, int kmax, uint32_t *d_tables, cudaStream_t stream, int *done)
{
    *done = 0;
    const int pb = 2 * kmax - PT_LB;
    int ok = 0;
    swga_count_partition_supported(n_starts, kmax, &ok);
    if (!ok) return SWGA_OK;
    const uint32_t nb = 1u << pb;
    const TableOffs offs = swga_make_offs(kmin, kmax);
    // workspace: bucket data (2 B per start) + bucket_cnt[nb+2] + cursor[nb] + item_prefix[nb+1] + work counter
    TRY(swga_ensure(ctx, ctx->bucket_data, (size_t)n_starts * 2 + 64));
    TRY(swga_ensure(ctx, ctx->bucket_meta, (size_t)(4 * nb + 16) * 4));
    uint16_t *data = (uint16_t *)ctx->bucket_data.p;
    uint32_t *base = (uint32_t *)ctx->bucket_meta.p;          // [nb + 1] counts -> exclusive bases (+ total)
    uint32_t *cursor = base + nb + 2;                          // [nb]
    uint32_t *item_prefix = cursor + nb;                       // [nb + 1]
    uint32_t *work = item_prefix + nb + 1;                     // [1]
    CUDA_TRY(cudaMemsetAsync(base, 0, (size_t)(4 * nb + 16) * 4, stream));
    const int grid_h = ctx->sm_count * 4;
    k_bucket_hist<<<grid_h, PT_THREADS, nb * 4, stream>>>(d_packed, d_brk, n_starts, kmax, pb, base);
    TRY(swga_exclusive_scan(ctx, base, nb, base, base + nb, stream));
    CUDA_TRY(cudaMemcpyAsync(cursor, base, nb * 4, cudaMemcpyDeviceToDevice, stream));
    const uint32_t n_tiles = (uint32_t)((n_starts + PT_TILE - 1) / PT_TILE);
    const size_t smem_p = (size_t)3 * nb * 4 + (size_t)PT_TILE * 4;
    CUDA_TRY(cudaFuncSetAttribute(k_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    const int per_sm_p = smem_p <= 100 * 1024 ? 2 : 1;
    const uint32_t grid_p = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * per_sm_p);
    k_partition<<<grid_p, PT_THREADS, smem_p, stream>>>(d_packed, d_brk, n_starts, kmin, kmax, pb, d_tables, offs, cursor,
                                                       data, n_tiles);
    k_item_prefix<<<1, 1024, 0, stream>>>(base, nb, item_prefix);
    const size_t smem_c = (size_t)4 << PT_LB;
    CUDA_TRY(cudaFuncSetAttribute(k_bucket_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    k_bucket_count<<<ctx->sm_count * 3, BC_THREADS, smem_c, stream>>>(data, base, item_prefix, nb, PT_LB,
                                                                     d_tables + offs.off[kmax], work);
    CUDA_TRY(cudaGetLastError());
    *done = 1;
    return SWGA_OK;
}
