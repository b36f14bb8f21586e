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
    unsigned long long *part_stats = nullptr;          // inside bucket_meta: exits taken by the last partitioned count
    int stage_timing = 0;                              // record events between the kernels of the partitioned count
    cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    int stage_valid = 0;
    int count_impl = 0;                                // 0 = auto, 1 = direct red.global kernel, 2 = partitioned
};

int swga_ensure(swga_ctx *ctx, DevBuf &b, size_t bytes);   // grow-only workspace

struct TableOffs {
    unsigned long long off[SWGA_KMAX_LIMIT + 2];   // off[k] = word offset of table k (k >= kmin)
};
TableOffs swga_make_offs(int kmin, int kmax);

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline unsigned grid_for(uint64_t n_threads, unsigned block)
{
    uint64_t g = (n_threads + block - 1) / block;
    return (unsigned)(g ? g : 1);
}
