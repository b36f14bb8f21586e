// graph.cu -- the heterodimer compatibility graph of `swga find_sets`.
//
// Result reproduced: graph.build_edges (swga/graph.py:7-17) with kmers.max_sequential_nt (swga/kmers.py:59-91):
// primers i < j (in the order given) are joined when neither sequence contains the other and the longest run of
// consecutive complementary bases between them, over all offsets of the longer one against the reversed shorter
// one, is at most max_binding.  The reference does this with a Python double loop per pair, O(P^2 k^2).
//
// One thread per ordered pair (i, j > i); a warp packs its 32 answers into one word of row i of the adjacency
// bit matrix (bit b of word w of row i = pair (i, 32 w + b)), so that reading the set bits row by row gives the
// edges in itertools.combinations order.
#include "common.cuh"

#define GE_THREADS 128
#define GE_MAX_LEN 256

__device__ __forceinline__ uint8_t complement(uint8_t c)
{
    // A<->T, C<->G; anything else (the '_' padding of the reference) pairs with nothing
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 0;
}

__device__ bool contains(const uint8_t *hay, uint32_t nh, const uint8_t *needle, uint32_t nn)
{
    if (nn > nh) return false;
    for (uint32_t s = 0; s + nn <= nh; ++s) {
        uint32_t x = 0;
        while (x < nn && hay[s + x] == needle[x]) ++x;
        if (x == nn) return true;
    }
    return false;
}

__global__ void __launch_bounds__(GE_THREADS)
k_build_edges(const uint8_t *__restrict__ seqs, uint32_t n, uint32_t stride, const uint32_t *__restrict__ len,
              int max_binding, uint32_t *__restrict__ adj, uint32_t words_per_row)
{
    __shared__ uint8_t s_i[GE_MAX_LEN];
    const uint32_t i = blockIdx.y;
    const uint32_t j = blockIdx.x * GE_THREADS + threadIdx.x;
    const uint32_t li = len[i];
    for (uint32_t x = threadIdx.x; x < li; x += GE_THREADS) s_i[x] = seqs[(size_t)i * stride + x];
    __syncthreads();
    bool edge = false;
    if (j > i && j < n) {
        const uint8_t *pj = seqs + (size_t)j * stride;
        const uint32_t lj = len[j];
        // mer1 = the longer one (the first one on a tie), mer2 = the other one, read backwards
        const bool swap = lj > li;
        const uint8_t *m1 = swap ? pj : s_i, *m2 = swap ? s_i : pj;
        const uint32_t n1 = swap ? lj : li, n2 = swap ? li : lj;
        if (!contains(m1, n1, m2, n2)) {                       // (equal lengths: containment either way = equality)
            int best = 0;
            for (uint32_t off = 0; off < n1; ++off) {
                int run = 0;
                for (uint32_t x = 0; x < n2; ++x) {
                    const uint8_t a = off + x < n1 ? complement(m1[off + x]) : 0;
                    if (a != 0 && a == m2[n2 - 1 - x]) {
                        ++run;
                        best = max(best, run);
                    } else {
                        run = 0;
                    }
                }
            }
            edge = best <= max_binding;
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, edge);
    if ((threadIdx.x & 31) == 0 && (j >> 5) < words_per_row) adj[(size_t)i * words_per_row + (j >> 5)] = word;
}

extern "C" int swga_build_edges(swga_ctx *ctx, const uint8_t *d_seqs, uint32_t n, uint32_t stride, const uint32_t *d_len,
                                uint32_t max_len, int max_binding, uint32_t *d_adj, uint32_t words_per_row, void *stream)
{
    ARG_CHECK(ctx && (n == 0 || (d_seqs && d_len && d_adj)));
    ARG_CHECK(max_len <= stride && words_per_row >= (n + 31) / 32);
    if (max_len > GE_MAX_LEN) {
        swga_set_error("primers longer than %d bases are not supported by swga_build_edges (got %u)", GE_MAX_LEN, max_len);
        return SWGA_E_ARG;
    }
    if (n > 65535) {
        swga_set_error("swga_build_edges takes at most 65535 primers (got %u)", n);
        return SWGA_E_ARG;
    }
    if (n == 0) return SWGA_OK;
    const dim3 grid((n + GE_THREADS - 1) / GE_THREADS, n);
    SWGA_NOTE_LAUNCH(); k_build_edges<<<grid, GE_THREADS, 0, as_stream(stream)>>>(d_seqs, n, stride, d_len, max_binding, d_adj, words_per_row);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}
