// merge.cu -- the cross-GPU merge of the additive count-table blocks, fused with the finalize pass.
//
// Every rank has accumulated its shard of the genome into its own block (additive "top-k + tails" form,
// include/swga_b200.h).  The reference of this step is a sum over ranks followed by the prefix marginalisation
// T_k[x] += sum_b T_{k+1}[4x + b] (swga_count_finalize).  Round 1 did it as an NCCL all_reduce of the 89.5 MB
// block plus finalize on every rank: every rank summed and marginalised EVERYTHING.
//
// Here the code space is split by its leading bits: rank r owns, at every level k, the codes whose position
// in the table is in [r, r + 1) / world -- and the children of an owned code are owned too.  One kernel per
// rank, over the peer mappings of all blocks (NVLink / NVSwitch loads and stores, no staging copy, no NCCL):
//
//     for its span of 4^d codes of T_kmax a CTA loads the 16-byte words of ALL ranks' blocks, sums them,
//     stores the sum into ALL ranks' blocks (in place: only the owner ever touches an owned word), and keeps
//     the four-way sums in shared memory; these are the marginal part of the level below, to which it adds the
//     ranks' tail counts of that level (same load / sum / store), and so on down to kmin.
//
// With world = 1 the same kernel is the single-GPU finalize: one pass over the block instead of one launch per level.
// Afterwards every block holds the FINALIZED forward tables of the whole genome: reduce-scatter, finalize of
// the owned slice and all-gather in one pass; each rank sums and marginalises 1 / world of the tables.
// The caller brackets the kernel with two device-side barriers across the ranks (swga_b200/dist.py).
#include "common.cuh"

#define MG_THREADS 256                              // 1,024 threads measured: merge 0.28 ms either way, step 1.13 vs 1.10 ms at N = 8
#define MG_MAX_WORLD 8
#define MG_UNROLL 8                                 // multicast kernel: reductions in flight per thread
#define MG_MAX_DEPTH 7                              // levels below kmax finished inside one CTA: span = 4^7 codes of T_kmax

struct PeerBlocks {
    uint32_t *p[MG_MAX_WORLD];
};

__device__ __forceinline__ uint4 ld_peer(const uint32_t *a)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a));
    return v;
}
__device__ __forceinline__ void st_peer(uint32_t *a, const uint4 v)
{
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One CTA: `span` consecutive codes of level kmax starting at code c0 (span = 4^depth), then the span / 4 codes of
// level kmax - 1 starting at c0 / 4, ... down to level kmax - depth.
template <int WORLD>
__global__ void __launch_bounds__(MG_THREADS)
k_merge_finalize(PeerBlocks blocks, TableOffs offs, int kmax, int depth, uint64_t slice_start, uint32_t span)
{
    // marg[i] = sum of the four children of code i of the level being processed (the marginal part of its value)
    __shared__ __align__(16) uint32_t bufA[4096], bufB[1024];
    uint32_t *marg = bufA, *next = bufB;
    uint64_t c0 = slice_start + (uint64_t)blockIdx.x * span;
    uint32_t n = span;                                         // codes of this level in the CTA's span
    for (int lv = 0; lv <= depth; ++lv) {
        const int k = kmax - lv;
        const uint64_t off = offs.off[k] + c0;
        uint32_t *out = lv == 0 ? marg : next;                 // where this level's four-way sums go
        if (n >= 4) {
            for (uint32_t i = threadIdx.x; i < n / 4; i += MG_THREADS) {
                uint4 v[WORLD];
#pragma unroll
                for (int p = 0; p < WORLD; ++p) v[p] = ld_peer(blocks.p[p] + off + 4ull * i);
                uint4 s = v[0];
#pragma unroll
                for (int p = 1; p < WORLD; ++p) {
                    s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w;
                }
                if (lv > 0) {
                    const uint4 m = reinterpret_cast<const uint4 *>(marg)[i];
                    s.x += m.x; s.y += m.y; s.z += m.z; s.w += m.w;
                }
                if (WORLD > 1 || lv > 0) {                     // one rank, top level: the word already holds its value
#pragma unroll
                    for (int p = 0; p < WORLD; ++p) st_peer(blocks.p[p] + off + 4ull * i, s);
                }
                out[i] = s.x + s.y + s.z + s.w;
            }
        } else if (threadIdx.x == 0) {                         // n == 1: the last level of the span
            uint32_t s = lv > 0 ? marg[0] : 0u;
#pragma unroll
            for (int p = 0; p < WORLD; ++p) s += *reinterpret_cast<volatile const uint32_t *>(blocks.p[p] + off);
#pragma unroll
            for (int p = 0; p < WORLD; ++p) *reinterpret_cast<volatile uint32_t *>(blocks.p[p] + off) = s;
        }
        __syncthreads();
        if (lv > 0) {                                          // `next` becomes the marginals of the level below
            uint32_t *t = marg;
            marg = next;
            next = t;
        }
        c0 >>= 2;
        n >>= 2;
    }
}

// ------------------------------------------------------------------------------------------------
// The same pass through the NVSwitch: `mc` is the MULTICAST mapping of the ranks' blocks (one address = the same
// offset in every block).  multimem.ld_reduce returns the sum over all ranks, added up INSIDE the switch (each GPU
// sends its words once, the owner receives one reduced word instead of one per rank), multimem.st writes the result
// into every block.  Pairs of 32-bit counts travel as one 64-bit integer: exact while every summed count stays below
// 2^32 (the table's own range; beyond it the peer-pointer kernel wraps each count modulo 2^32 and this one carries
// into the neighbour -- the reference drops such counts, ext/dsk/minia/Utils.cpp:6).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mc_ld_reduce_u64(const uint32_t *a)
{
    unsigned long long v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_u64(uint32_t *a, unsigned long long v)
{
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t mc_ld_reduce_u32(const uint32_t *a)
{
    uint32_t v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_u32(uint32_t *a, uint32_t v)
{
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}

// One CTA: span codes of level kmax from c0 on, then the levels below (as k_merge_finalize).  A lane handles one PAIR of
// adjacent codes per step (8 contiguous bytes: a warp reads 256 contiguous bytes); the four children of a code are two
// pairs = two adjacent lanes.
__global__ void __launch_bounds__(MG_THREADS)
k_merge_finalize_mc(uint32_t *__restrict__ mc, TableOffs offs, int kmax, int depth, uint64_t slice_start, uint32_t span)
{
    __shared__ __align__(16) uint32_t bufA[4096], bufB[1024];
    uint32_t *marg = bufA, *next = bufB;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t c0 = slice_start + (uint64_t)blockIdx.x * span;
    uint32_t n = span;
    for (int lv = 0; lv <= depth; ++lv) {
        const int k = kmax - lv;
        uint32_t *tab = mc + offs.off[k] + c0;
        uint32_t *out = lv == 0 ? marg : next;
        if (n >= 4) {
            const uint32_t pairs = n / 2;
            // MG_UNROLL steps of the CTA at a time: all their reductions are in flight before the first is used (a round
            // trip through the switch is microseconds; one 8-byte load per thread at a time ran at 320 GB/s)
            for (uint32_t p0 = threadIdx.x & ~31u; p0 < pairs; p0 += MG_UNROLL * MG_THREADS) {   // warp-uniform trip count
                unsigned long long v[MG_UNROLL];
#pragma unroll
                for (int u = 0; u < MG_UNROLL; ++u) {
                    const uint32_t p = p0 + u * MG_THREADS + lane;
                    v[u] = p < pairs ? mc_ld_reduce_u64(tab + 2ull * p) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < MG_UNROLL; ++u) {
                    const uint32_t p = p0 + u * MG_THREADS + lane;
                    const bool on = p < pairs;
                    uint32_t lo = (uint32_t)v[u], hi = (uint32_t)(v[u] >> 32);
                    if (on) {
                        if (lv > 0) {
                            const uint2 m = reinterpret_cast<const uint2 *>(marg)[p];
                            lo += m.x;
                            hi += m.y;
                        }
                        mc_st_u64(tab + 2ull * p, (unsigned long long)lo | ((unsigned long long)hi << 32));
                    }
                    uint32_t s = lo + hi;
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    if (on && !(p & 1u)) out[p >> 1] = s;
                }
            }
        } else if (threadIdx.x == 0) {
            uint32_t s = lv > 0 ? marg[0] : 0u;
            s += mc_ld_reduce_u32(tab);
            mc_st_u32(tab, s);
        }
        __syncthreads();
        if (lv > 0) {
            uint32_t *t = marg;
            marg = next;
            next = t;
        }
        c0 >>= 2;
        n >>= 2;
    }
}

extern "C" int swga_merge_supported(int kmin, int kmax, int world, int *ok)
{
    ARG_CHECK(ok);
    const int depth = kmax - kmin;
    // a rank's slice of T_kmax must be whole spans of 4^depth codes
    *ok = kmin >= 2 && kmax <= SWGA_KMAX_LIMIT && depth >= 0 && depth <= MG_MAX_DEPTH &&
          (world == 1 || world == 2 || world == 4 || world == 8) &&
          ((1ull << (2 * kmax)) / (unsigned)world) % (1ull << (2 * depth)) == 0;
    return SWGA_OK;
}

// d_multicast: the multicast mapping of the ranks' blocks (torch symmetric memory: multicast_ptr)
extern "C" int swga_merge_finalize_multicast(swga_ctx *ctx, uint32_t *d_multicast, int world, int rank, int kmin, int kmax,
                                             void *stream)
{
    ARG_CHECK(ctx && d_multicast && rank >= 0 && rank < world && world >= 2);
    int ok = 0;
    TRY(swga_merge_supported(kmin, kmax, world, &ok));
    if (!ok) {
        swga_set_error("swga_merge_finalize_multicast: unsupported shape kmin=%d kmax=%d world=%d", kmin, kmax, world);
        return SWGA_E_ARG;
    }
    const TableOffs offs = swga_make_offs(kmin, kmax);
    const int depth = kmax - kmin;
    const uint64_t slice = (1ull << (2 * kmax)) / (unsigned)world;
    const uint32_t span = 1u << (2 * depth);
    SWGA_NOTE_LAUNCH();
    k_merge_finalize_mc<<<(unsigned)(slice / span), MG_THREADS, 0, as_stream(stream)>>>(d_multicast, offs, kmax, depth, slice * rank, span);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}

// h_blocks: `world` device pointers (this rank's view: peer mappings) of the ranks' table blocks, in rank order.
extern "C" int swga_merge_finalize(swga_ctx *ctx, const uint64_t *h_blocks, int world, int rank, int kmin, int kmax, void *stream)
{
    ARG_CHECK(ctx && h_blocks && rank >= 0 && rank < world);
    int ok = 0;
    TRY(swga_merge_supported(kmin, kmax, world, &ok));
    if (!ok) {
        swga_set_error("swga_merge_finalize: unsupported shape kmin=%d kmax=%d world=%d", kmin, kmax, world);
        return SWGA_E_ARG;
    }
    PeerBlocks pb;
    for (int i = 0; i < MG_MAX_WORLD; ++i) pb.p[i] = reinterpret_cast<uint32_t *>(i < world ? h_blocks[i] : 0);
    const TableOffs offs = swga_make_offs(kmin, kmax);
    const int depth = kmax - kmin;
    const uint64_t slice = (1ull << (2 * kmax)) / (unsigned)world;
    const uint32_t span = 1u << (2 * depth);
    const unsigned grid = (unsigned)(slice / span);
    cudaStream_t st = as_stream(stream);
    SWGA_NOTE_LAUNCH();
    if (world == 1) k_merge_finalize<1><<<grid, MG_THREADS, 0, st>>>(pb, offs, kmax, depth, slice * rank, span);
    else if (world == 2) k_merge_finalize<2><<<grid, MG_THREADS, 0, st>>>(pb, offs, kmax, depth, slice * rank, span);
    else if (world == 4) k_merge_finalize<4><<<grid, MG_THREADS, 0, st>>>(pb, offs, kmax, depth, slice * rank, span);
    else k_merge_finalize<8><<<grid, MG_THREADS, 0, st>>>(pb, offs, kmax, depth, slice * rank, span);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}
