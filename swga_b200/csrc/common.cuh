// common.cuh -- shared declarations for libswga_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>

#include "../../include/swga_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libswga_b200 is written for sm_100a (B200) only"
#endif

// ---- error plumbing -------------------------------------------------------------------------
void swga_set_error(const char *fmt, ...);
#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            swga_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                           __FILE__, __LINE__);                                          \
            return SWGA_E_CUDA;                                                          \
        }                                                                                \
    } while (0)
#define ARG_CHECK(cond)                                                                  \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            swga_set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__);   \
            return SWGA_E_ARG;                                                           \
        }                                                                                \
    } while (0)
#define TRY(expr)                                                                        \
    do {                                                                                 \
        int _r = (expr);                                                                 \
        if (_r != SWGA_OK) return _r;                                                    \
    } while (0)

// ---- launch counter (bench.py's gpu_launches): every kernel launch of the library goes through this ----
void swga_note_launch();
#define SWGA_NOTE_LAUNCH() swga_note_launch()

// ---- context --------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct swga_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;   // used by the *_host pipelines
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    DevBuf packed, brk, ascii[2], packed2[2], brk2[2], fwd, canon, recs, scratch;
    void *h_stage[2] = {nullptr, nullptr};             // pinned staging for pageable host genomes (64 MiB each)
    size_t h_stage_cap = 0;
    void *h_recs[2] = {nullptr, nullptr};              // pinned: record starts of the chunk in flight (file-image path)
    size_t h_recs_cap = 0;
    DevBuf sort_hist, sort_keys[2], sort_vals[2], score_tmp, bucket_data, bucket_meta;
    uint32_t score_sort_cap = 4096;                    // raw sites per set up to which k_score_sort is used (0 = never)
    int score_sort_impl = 0;                           // k_score_sort: 0 = routed front end (window counting sort), 1 = always merge the lists
    int score_dense_impl = 0;                          // k_score_dense: 0 = routed (window rows + register sort), 1 = bitmap + compaction
    unsigned long long *part_stats = nullptr;          // inside bucket_meta: exits taken by the last partitioned count
    int stage_timing = 0;                              // record events between the kernels of the partitioned count
    cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    int stage_valid = 0;
    int count_impl = 0;                                // 0 = auto, 1 = direct red.global kernel, 2 = partitioned
    int count_via_pack = 0;                            // 1 = the partitioned kernels read k_pack's output instead of the ASCII bases
};

int swga_ensure(swga_ctx *ctx, DevBuf &b, size_t bytes);   // grow-only workspace

struct TableOffs {
    unsigned long long off[SWGA_KMAX_LIMIT + 2];   // off[k] = word offset of table k (k >= kmin)
};
TableOffs swga_make_offs(int kmin, int kmax);

// ---- base encoding shared by the pack kernel and the kernels that read ASCII bases directly ----------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t codes4(uint32_t w)
{
    // 4 ASCII bytes (first base in the low byte) -> 8 bits, first base in bits 7:6.
    // (c>>1)&3 per byte, then one multiply gathers the four 2-bit fields (no carries: the partial
    // products land on disjoint bit positions).
    uint32_t x = (w >> 1) & 0x03030303u;
    return (x * ((1u << 30) | (1u << 20) | (1u << 10) | 1u)) >> 24;
}
__device__ __forceinline__ uint32_t flags4(uint32_t m)
{
    // m has 0x01 in each flagged byte (first base in the low byte) -> 4 bits, first base in bit 3.
    return (m * ((1u << 31) | (1u << 22) | (1u << 13) | (1u << 4))) >> 28;
}
__device__ __forceinline__ uint32_t bad4(uint32_t w, int mask_mode)
{
    if (mask_mode == SWGA_MASK_N) return __vcmpeq4(w, 0x4E4E4E4Eu) & 0x01010101u;   // 'N'
    uint32_t ok = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) |            // A C
                  __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);             // G T
    return ~ok & 0x01010101u;
}

// 16 ASCII bytes -> 16 bases (32 bits, first base in bits 31:30) / their 16 flag bits (first base in bit 15)
__device__ __forceinline__ uint32_t pack16(const uint4 v)
{
    return (codes4(v.x) << 24) | (codes4(v.y) << 16) | (codes4(v.z) << 8) | codes4(v.w);
}
__device__ __forceinline__ uint32_t flags16(const uint4 v, int mask_mode)
{
    return (flags4(bad4(v.x, mask_mode)) << 12) | (flags4(bad4(v.y, mask_mode)) << 8) | (flags4(bad4(v.z, mask_mode)) << 4) |
           flags4(bad4(v.w, mask_mode));
}
#endif

// where the partitioned counting kernels read a group's bases from (count_part.cu)
enum { SRC_PACKED = 0, SRC_ASCII = 1, SRC_ASCII_N = 2 };
struct GenomeSrc {
    const uint32_t *packed;        // SRC_PACKED
    const uint32_t *brk;           // break mask (SRC_PACKED: complete; SRC_ASCII*: record boundaries + end of buffer)
    const uint8_t *ascii;          // SRC_ASCII*: 16-byte aligned
    uint64_t n;                    // SRC_ASCII*: bases in `ascii`
};

int swga_count_accumulate_partitioned(swga_ctx *ctx, const GenomeSrc &src, int src_kind, uint64_t n_starts, int kmin, int kmax,
                                      uint32_t *d_tables, cudaStream_t stream, int *done);
int swga_count_partition_supported(uint64_t n_starts, int kmax, int *ok);
// pack (if needed) + record marks + accumulate of a device-resident ASCII buffer; `d_marks` = the starts of the records
// that begin inside (origin, origin + n] in genome coordinates (device array, may be NULL when n_marks == 0)
int swga_count_ascii(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, const uint64_t *d_marks, uint64_t n_marks,
                     uint64_t origin, int mask_mode, uint64_t n_starts, int kmin, int kmax, uint32_t *d_packed,
                     uint32_t *d_brk, uint32_t *d_tables, cudaStream_t stream);

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline unsigned grid_for(uint64_t n_threads, unsigned block)
{
    uint64_t g = (n_threads + block - 1) / block;
    return (unsigned)(g ? g : 1);
}
