// filter.cu -- the count -> primer-row merge of `swga count` on the device.
//
// Results reproduced (reference = eclarke/swga):
//   swga/commands/count.py:98-104   keep k-mers found in the foreground that are not in the exclusion set
//   swga/commands/count.py:117-135  primer_dict: fg_freq >= min_fg_bind, bg_freq <= max_bg_bind (-1 = no limit),
//                                   max_sequential_nt(seq, seq) <= max_dimer_bp  (homodimer test)
//   swga/kmers.py:59-91             max_sequential_nt
// One thread per table slot of the concatenated canonical tables; survivors are appended (order is
// unspecified in the reference too: DSK hash order / Python-2 dict order).
#include "common.cuh"

// longest run of consecutive complementary bases between seq (offset by `o`) and reversed seq
__device__ __forceinline__ int max_sequential_self(uint32_t code, int k)
{
    int best = 0;
    for (int o = 0; o < k; ++o) {
        int run = 0;
        for (int x = 0; x + o < k; ++x) {                       // positions past the end are '_' and never bind
            const uint32_t a = (code >> (2 * (k - 1 - (o + x)))) & 3u;    // mer1[o + x]
            const uint32_t b = (code >> (2 * x)) & 3u;                     // reversed mer2[x] = seq[k-1-x]
            if ((a ^ 2u) == b) { ++run; best = max(best, run); }
            else run = 0;
        }
    }
    return best;
}

__global__ void __launch_bounds__(256)
k_filter_primers(const uint32_t *__restrict__ fg, const uint32_t *__restrict__ bg, const uint32_t *__restrict__ ex,
                 int kmin, int kmax, TableOffs offs, uint64_t total, uint32_t min_fg, uint32_t max_bg, int max_dimer,
                 uint32_t ex_threshold, uint4 *__restrict__ rows, uint32_t *__restrict__ n_rows, uint32_t cap)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t f = fg[i];
    if (f == 0 || f < min_fg) return;                           // `for seq in fg`: only k-mers seen in the foreground
    const uint32_t b = bg[i];
    if (b > max_bg) return;
    if (ex && ex[i] >= ex_threshold) return;                    // `seq not in ex`
    int k = kmin;
    while (k < kmax && i >= offs.off[k + 1]) ++k;
    const uint32_t code = (uint32_t)(i - offs.off[k]);
    if (max_sequential_self(code, k) > max_dimer) return;
    const uint32_t slot = atomicAdd(n_rows, 1u);
    if (slot < cap) rows[slot] = make_uint4((uint32_t)k, code, f, b);
}

extern "C" int swga_filter_primers(swga_ctx *ctx, const uint32_t *d_canon_fg, const uint32_t *d_canon_bg,
                                   const uint32_t *d_canon_ex, int kmin, int kmax, uint32_t min_fg_bind,
                                   int64_t max_bg_bind, int max_dimer_bp, uint32_t exclude_threshold,
                                   uint32_t *d_rows, uint32_t row_capacity, uint32_t *d_n_rows, void *stream)
{
    ARG_CHECK(ctx && d_canon_fg && d_canon_bg && d_n_rows && (d_rows || row_capacity == 0));
    ARG_CHECK(kmin >= 2 && kmax >= kmin && kmax <= SWGA_KMAX_LIMIT);
    const uint64_t total = swga_tables_words(kmin, kmax);
    CUDA_TRY(cudaMemsetAsync(d_n_rows, 0, 4, as_stream(stream)));
    const uint32_t max_bg = max_bg_bind < 0 ? 0xffffffffu : (max_bg_bind > 0xffffffffll ? 0xffffffffu : (uint32_t)max_bg_bind);
    SWGA_NOTE_LAUNCH(); k_filter_primers<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        d_canon_fg, d_canon_bg, d_canon_ex, kmin, kmax, swga_make_offs(kmin, kmax), total, min_fg_bind, max_bg, max_dimer_bp,
        exclude_threshold ? exclude_threshold : 1u, reinterpret_cast<uint4 *>(d_rows), d_n_rows, row_capacity);
    CUDA_TRY(cudaGetLastError());
    return SWGA_OK;
}
