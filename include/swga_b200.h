/*
 * swga_b200.h -- C ABI of libswga_b200.so: the B200 (sm_100a) implementation of swga's
 * data-parallel hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (eclarke/swga, citations relative to its tree) has no FFI: its seams are Python
 * functions and two subprocess protocols.  Each entry point below names the reference interface
 * whose RESULTS it reproduces; INTEGRATION.md shows the ctypes stubs a swga maintainer would add
 * to swga/kmers.py, swga/locate.py and swga/score.py.
 *
 * Conventions
 *   - every call returns 0 on success or a negative SWGA_E* code; swga_last_error() returns a
 *     thread-local message for the last failing call on this thread
 *   - pointers named d_* are device pointers (cudaMalloc / swga_malloc / torch data_ptr), h_* are
 *     host pointers; all buffers are caller-owned unless stated
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - nucleotide code is DSK's: (byte >> 1) & 3  =>  A=0 C=1 T=2 G=3 (ext/dsk/minia/Kmer.cpp:18-24);
 *     a k-mer code is the base-4 number whose most significant digit is the first base
 *     (Kmer.cpp:65-76); reverse complement is digit reversal with digit ^ 2 (lut.h:8-10)
 *   - "packed genome": uint32 words, 16 bases per word, first base in bits 31:30; plus a "break
 *     mask": 1 bit per base (uint32 per 32 bases, first base in bit 31); bit i set means NO k-mer
 *     window may contain both base i and base i+1 (record ends, mode-A 'N', genome end)
 *   - count tables: uint32, direct-indexed by k-mer code, the tables for k = kmin..kmax
 *     concatenated (swga_tables_words(kmin,kmax) words in total)
 */
#ifndef SWGA_B200_H
#define SWGA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWGA_OK            0
#define SWGA_E_CUDA       -1   /* a CUDA runtime call failed (message has the CUDA error string) */
#define SWGA_E_ARG        -2   /* invalid argument */
#define SWGA_E_NOMEM      -3   /* allocation failed */
#define SWGA_E_FORMAT     -4   /* malformed input (e.g. not a FASTA image) */
#define SWGA_E_OVERFLOW   -5   /* an output buffer supplied by the caller is too small */

#define SWGA_KMAX_LIMIT   15   /* dense direct-indexed tables: 4^15 uint32 = 4 GiB */

/* mask_mode of swga_pack */
#define SWGA_MASK_NONE    0    /* DSK mode B: every byte is a base ('N' -> G), SortingCount.cpp:392-421 */
#define SWGA_MASK_N       1    /* DSK mode A: upper-case 'N' splits reads, SortingCount.cpp:243-261   */
#define SWGA_MASK_STRICT  2    /* locate: only upper-case A,C,G,T can match, swga/locate.py:76-90      */

typedef struct swga_ctx swga_ctx;

/* ---- library / context -------------------------------------------------------------------- */
int         swga_abi_version(void);
const char *swga_last_error(void);
/* One context per (host thread, device).  Owns workspaces, two streams and events used by the
 * *_host pipelines.  Fails with SWGA_E_CUDA when no CUDA device is usable: there is NO CPU path. */
int  swga_ctx_create(int device, swga_ctx **out);
void swga_ctx_destroy(swga_ctx *ctx);
int  swga_ctx_sm_count(const swga_ctx *ctx);
/* Counting kernel choice: 0 = automatic (partitioned shared-memory histograms when the shape allows,
 * else direct L2 reds), 1 = always the direct red.global kernel, 2 = partitioned or fail. */
int  swga_ctx_set_count_impl(swga_ctx *ctx, int impl);

/* ---- memory helpers for hosts that do not bring their own allocator ------------------------- */
int swga_malloc(swga_ctx *ctx, void **d_ptr, size_t bytes);
int swga_free(swga_ctx *ctx, void *d_ptr);
int swga_host_alloc(swga_ctx *ctx, void **h_ptr, size_t bytes);      /* pinned */
int swga_host_free(swga_ctx *ctx, void *h_ptr);
int swga_memset(swga_ctx *ctx, void *d_ptr, int value, size_t bytes, void *stream);
int swga_h2d(swga_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int swga_d2h(swga_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);
int swga_stream_sync(swga_ctx *ctx, void *stream);

/* ---- sizes ----------------------------------------------------------------------------------- */
uint64_t swga_tables_words(int kmin, int kmax);            /* sum_{k=kmin..kmax} 4^k              */
uint64_t swga_table_offset(int kmin, int k);               /* word offset of table k in the block */
uint64_t swga_packed_words(uint64_t n_bases);              /* uint32 words swga_pack writes        */
uint64_t swga_brk_words(uint64_t n_bases);                 /* uint32 words of the break mask       */

/* ---- host-side FASTA reader (native host code; no GPU) ---------------------------------------
 * Replaces DSK's Bank::get_next_seq_from_file (ext/dsk/minia/Bank.cpp:150-209): records start at a
 * '>' or '@' that begins a line, lines are joined, one trailing '\r' per line is dropped under
 * Bank.cpp:124-125's rule.  FASTQ records ('+' at a line start, Bank.cpp:192-201) are read with the
 * reference's byte-level rules by one thread (any image that begins with '@' or holds such a line).  Pass 1 (h_seq == NULL) returns sizes; pass 2 fills the buffers (it reuses the
 * line ranges pass 1 found when it follows it on the same thread with the same, unmodified image).  Both
 * passes scan the image with up to 16 host threads.
 *   h_seq        [*n_bases]   joined sequence bytes of all records, no separators
 *   h_rec_starts [*n_rec + 1] start of each record in h_seq, last entry = *n_bases
 *   h_name_off   [*n_rec * 2] (offset, length) into the file image of each record's first header
 *                             token (pyfaidx's record key, relied on by swga/locate.py:42)
 */
int swga_fasta_scan(const uint8_t *h_file, uint64_t n_file, uint8_t *h_seq, uint64_t *h_rec_starts,
                    uint64_t *h_name_off, uint64_t *n_bases, uint64_t *n_rec);
/* DSK's read-mode choice: 1 (mode B, raw) iff one of the first 10,001 records is longer than
 * 1,000,000 bases (Bank.cpp:470-494, SortingCount.cpp:81-82); else 0 (mode A, split at 'N'). */
int swga_dsk_mode(const uint64_t *h_rec_starts, uint64_t n_rec);

/* ---- kernel (a): 2-bit pack -------------------------------------------------------------------
 * Replaces NT2int/code4NT/BinaryReads::write_read (Kmer.cpp:18-43, Bank.cpp:832-885) and the N-split
 * (SortingCount.cpp:243-261).  d_ascii must be 16-byte aligned.  Writes swga_packed_words(n) words
 * of d_packed and swga_brk_words(n) words of d_brk (mask per `mask_mode`, plus the end-of-buffer
 * break on base n-1). */
int swga_pack(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, int mask_mode,
              uint32_t *d_packed, uint32_t *d_brk, void *stream);
/* ORs record boundaries into the break mask: for every r in 1..n_rec-1 sets the bit of base
 * d_rec_starts[r]-1 (k-mers never span records: one DSK "read" per record, Bank.cpp:150-209;
 * swga/locate.py:43-47 searches per record).  Entries outside (0, n] are ignored. */
int swga_mark_records(swga_ctx *ctx, const uint64_t *d_rec_starts, uint64_t n_rec, uint64_t n,
                      uint32_t *d_brk, void *stream);

/* ---- kernel (b): all-k counting ----------------------------------------------------------------
 * Replaces compute_kmer_table_from_one_seq / KmersBuffer::readkmers + OAHash::increment
 * (Bank.cpp:889-902,1042-1070, OAHash.cpp:142-154) for every k in [kmin,kmax] at once.
 *
 * swga_count_accumulate ADDS, for each start position p < n_starts, one count to the forward-code
 * table of v = min(kmax, run of break-free bases from p), if v >= kmin ("top-k + tails" form).
 * The caller zeroes d_tables first and may accumulate several chunks / GPUs (the form is additive:
 * an allreduce-sum across ranks may be applied to it).  The buffers must hold the kmax-1 bases
 * after the last owned start (right halo) or end with a break.
 * swga_count_finalize turns the accumulated block into full forward tables in place:
 *   T_k[x] += sum_b T_{k+1}[4x+b]   for k = kmax-1 .. kmin.
 * swga_fold_canonical writes DSK's canonical counts (Kmer.cpp:122: min(fwd, revcomp)):
 *   C_k[x] = T_k[x] + T_k[rc(x)] if x < rc(x);  T_k[x] if x == rc(x);  0 otherwise.
 */
/* The same accumulate fed with the ASCII bases (device-resident, 16-byte aligned): for inputs large enough for
 * the partitioned kernels they read the bases themselves and pack them in registers -- no pack pass over the
 * genome (its 1 + 1/4 + 1/8 bytes per base of DRAM traffic and 0.6 ms per 3.1 Gbp are gone); smaller inputs are
 * packed into d_packed / d_brk (workspaces of swga_packed_words(n) / swga_brk_words(n) words, always overwritten)
 * and take the direct kernel.  d_rec_marks: device array of the n_marks record starts s with 0 < s <= n (the
 * boundaries inside the buffer), may be NULL.  mask_mode: SWGA_MASK_NONE (DSK mode B) or SWGA_MASK_N (mode A).
 * swga_ctx_set_count_via_pack(ctx, 1) makes the partitioned kernels read k_pack's output instead (A/B timing). */
int swga_count_accumulate_ascii(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n, const uint64_t *d_rec_marks,
                                uint64_t n_marks, int mask_mode, uint64_t n_starts, int kmin, int kmax,
                                uint32_t *d_packed, uint32_t *d_brk, uint32_t *d_tables, void *stream);
int swga_ctx_set_count_via_pack(swga_ctx *ctx, int on);
int swga_count_accumulate(swga_ctx *ctx, const uint32_t *d_packed, const uint32_t *d_brk,
                          uint64_t n_starts, int kmin, int kmax, uint32_t *d_tables, void *stream);
int swga_count_finalize(swga_ctx *ctx, int kmin, int kmax, uint32_t *d_tables, void *stream);
/* Diagnostics of the last partitioned swga_count_accumulate on this context (synchronises the device):
 * h_stats[0] = elements that found their shared-memory staging range full, [1] = elements that found
 * their global bucket full (added to the table directly), [2] = single-base-run groups of 32 starts,
 * [3] = groups that took the direct path because they touch a break.  All zero when the direct kernel ran. */
int swga_count_partition_stats(swga_ctx *ctx, uint64_t *h_stats);
/* Per-kernel timing of the partitioned swga_count_accumulate (bench.py's roofline of the dominant kernel):
 * with timing on, CUDA events are recorded on the call's stream between its kernels; swga_count_stage_ms
 * waits for the last call and returns milliseconds of {sample histogram + layout, k_partition,
 * k_item_prefix + k_bucket_count} (zeros when the direct kernel ran). */
int swga_ctx_set_stage_timing(swga_ctx *ctx, int on);
int swga_count_stage_ms(swga_ctx *ctx, float *h_ms);
int swga_fold_canonical(swga_ctx *ctx, int kmin, int kmax, const uint32_t *d_fwd,
                        uint32_t *d_canon, void *stream);

/* The cross-GPU merge of the ranks' additive blocks fused with swga_count_finalize (csrc/merge.cu): h_blocks holds
 * `world` (1, 2, 4 or 8) device pointers -- this rank's peer mappings of every rank's block, in rank order, e.g. the
 * buffer_ptrs of a torch symmetric-memory allocation.  Rank r sums, finalizes and writes back, into ALL blocks,
 * the slice [r, r + 1) / world of every table; when all ranks have run it (the caller puts a barrier across the
 * ranks before and after), every block holds the finalized forward tables of the whole genome.  Needs
 * kmax - kmin <= 7 (swga_merge_supported).  With all pointers on one device and the ranks run one after another
 * it is also a plain multi-block sum + finalize (tests). */
int swga_merge_supported(int kmin, int kmax, int world, int *ok);
/* The same pass through the NVSwitch: d_multicast is the MULTICAST mapping of the ranks' blocks (one address = the
 * same offset in every block; torch symmetric memory's multicast_ptr).  multimem.ld_reduce sums the ranks' words inside
 * the switch, multimem.st writes the finalized slice into every block. */
int swga_merge_finalize_multicast(swga_ctx *ctx, uint32_t *d_multicast, int world, int rank, int kmin, int kmax,
                                  void *stream);
int swga_merge_finalize(swga_ctx *ctx, const uint64_t *h_blocks, int world, int rank, int kmin, int kmax,
                        void *stream);

/* Whole counting pass on a device-resident genome: pack + mark + accumulate + finalize + fold,
 * using context-owned workspaces.  d_canon (and d_fwd if non-NULL) receive swga_tables_words()
 * words.  `finalize` = 0 leaves d_fwd in the additive form and skips the fold (multi-GPU: allreduce
 * d_fwd, then call swga_count_finalize + swga_fold_canonical). */
int swga_count_genome_device(swga_ctx *ctx, const uint8_t *d_ascii, uint64_t n,
                             const uint64_t *d_rec_starts, uint64_t n_rec, int mask_mode,
                             int kmin, int kmax, uint64_t n_starts, int finalize,
                             uint32_t *d_fwd, uint32_t *d_canon, void *stream);

/* The call swga/kmers.py:42-46 would make instead of spawning `dsk`, for all k at once, with HOST
 * buffers: streams h_seq to the device in chunks (H2D overlapped with pack/count), returns the
 * canonical tables in h_canon.  mode_b as returned by swga_dsk_mode. */
int swga_count_genome_host(swga_ctx *ctx, const uint8_t *h_seq, uint64_t n,
                           const uint64_t *h_rec_starts, uint64_t n_rec, int mode_b,
                           int kmin, int kmax, uint32_t *h_canon);

/* The streaming half of swga_count_genome_host for multi-GPU hosts: accumulates the additive form of
 * the owned starts [0, n_starts) of this rank's slice (h_seq holds n >= n_starts bases: the slice
 * plus its right halo) into the caller-zeroed device block d_fwd, asynchronously on the context's
 * compute stream (swga_ctx_compute_stream); the caller then allreduces, finalizes and folds. */
int   swga_count_stream_host(swga_ctx *ctx, const uint8_t *h_seq, uint64_t n,
                             const uint64_t *h_rec_starts, uint64_t n_rec, int mode_b,
                             int kmin, int kmax, uint64_t n_starts, uint32_t *d_fwd);
void *swga_ctx_compute_stream(swga_ctx *ctx);

/* The same two calls fed with the FASTA FILE IMAGE (the bytes of `genome_fp`, e.g. a read-only mapping) --
 * the closest equivalent of `dsk <genome_fp> <k>` (swga/kmers.py:44-47; Bank.cpp:150-209 record rules as in
 * swga_fasta_scan).  No host copy of the joined bases is made: host threads parse line-aligned ranges of the
 * image straight into pinned staging blocks, chunk by chunk, while the previous chunk is copied and counted.
 * mode: 0 = A (split at 'N'), 1 = B (raw), -1 = DSK's own choice (swga_dsk_mode); the mode used, the number
 * of bases and of records are returned through the last three pointers (each may be NULL).
 * swga_count_fasta_stream accumulates the additive form into the caller-zeroed d_fwd on the context's compute
 * stream and returns without synchronising it; swga_count_fasta_host returns the canonical tables in h_canon. */
int swga_count_fasta_stream(swga_ctx *ctx, const uint8_t *h_file, uint64_t n_file, int mode, int kmin, int kmax,
                            uint32_t *d_fwd, uint64_t *n_bases, uint64_t *n_rec, int *mode_used);
int swga_count_fasta_host(swga_ctx *ctx, const uint8_t *h_file, uint64_t n_file, int mode, int kmin, int kmax,
                          uint32_t *h_canon, uint64_t *n_bases, uint64_t *n_rec, int *mode_used);

/* The heterodimer compatibility graph of `swga find_sets`: graph.build_edges (swga/graph.py:7-17) with
 * kmers.max_sequential_nt (swga/kmers.py:59-91) for all pairs at once.  d_seqs holds n primers as upper-case
 * ACGT ASCII rows of `stride` bytes, d_len their lengths (<= max_len <= 256).  Row i of the bit matrix d_adj
 * (words_per_row >= ceil(n/32) words per row, every word of the first ceil(n/32) is written) gets bit (j & 31) of
 * word (j >> 5) set when i < j and the pair is compatible: neither sequence contains the other and their
 * longest complementary run is <= max_binding.  Set bits read row by row = itertools.combinations order. */
int swga_build_edges(swga_ctx *ctx, const uint8_t *d_seqs, uint32_t n, uint32_t stride, const uint32_t *d_len,
                     uint32_t max_len, int max_binding, uint32_t *d_adj, uint32_t words_per_row, void *stream);

/* The count -> primer-row merge of `swga count` (swga/commands/count.py:98-135) over the canonical
 * tables of the foreground, the background and (optionally, may be NULL) the exclusion sequences:
 * a k-mer survives when fg >= max(1, min_fg_bind), bg <= max_bg_bind (negative = no limit), its
 * exclusion count is below exclude_threshold, and max_sequential_nt(seq, seq) <= max_dimer_bp
 * (swga/kmers.py:59-91).  Survivors are appended to d_rows as (k, code, fg_freq, bg_freq) uint32
 * quadruples (16-byte aligned) in unspecified order; *d_n_rows receives the number of survivors, which
 * may exceed row_capacity (only the first row_capacity are stored: call again with a larger buffer). */
int swga_filter_primers(swga_ctx *ctx, const uint32_t *d_canon_fg, const uint32_t *d_canon_bg,
                        const uint32_t *d_canon_ex, int kmin, int kmax, uint32_t min_fg_bind,
                        int64_t max_bg_bind, int max_dimer_bp, uint32_t exclude_threshold,
                        uint32_t *d_rows, uint32_t row_capacity, uint32_t *d_n_rows, void *stream);

/* set_finder's stdout, a block at a time: what score.read_set_finder_line (swga/score.py:16-24) does per line
 * ("id,id,... weight\n", set_finder.c:427-444) for every complete line of `text`, at most `cap` of them.
 * h_ids [cap][max_m] (-1 padded), h_weight [cap], h_ok [cap] (0 = a line Python would reject, or more than
 * max_m ids); *consumed = bytes of the complete lines parsed (an unfinished last line stays for the next block). */
int swga_parse_set_lines(const char *text, uint64_t n_bytes, uint32_t max_m, uint64_t cap,
                         int32_t *h_ids, double *h_weight, uint8_t *h_ok, uint64_t *n_lines,
                         uint64_t *consumed);

/* DSK result-file writer (ext/dsk/minia/SortingCount.cpp:149-156,649-653; read back by
 * swga/kmers.py:94-115): [i32 64][i32 k] then (u64 code, u32 count) for every count >= threshold. */
int swga_write_dsk_file(const char *path, int k, const uint32_t *h_canon_k, uint32_t threshold,
                        uint64_t *n_written);

/* ---- kernel (c): CSR position index + locate ------------------------------------------------
 * Replaces the per-primer str.find scans of swga/locate.py:39-51,76-90 with one index per k:
 *   d_offsets   [swga_index_offsets_words(k)] exclusive scan of the per-code histogram; slot 4^k holds
 *               *n_valid (start of the sentinel bucket of windows that can never match)
 *   d_positions [n] (the first *n_valid are used) linearised start positions, ascending inside a code
 * d_packed/d_brk come from swga_pack(..., SWGA_MASK_STRICT, ...) + swga_mark_records, so only windows
 * of upper-case A/C/G/T inside one record are indexed (locate.py:83-86 matches raw bytes per record).
 * Synchronises the stream (returns *n_valid). */
uint64_t swga_index_offsets_words(int k);
int swga_build_index(swga_ctx *ctx, const uint32_t *d_packed, const uint32_t *d_brk, uint64_t n, int k,
                     uint32_t *d_offsets, uint32_t *d_positions, uint64_t *n_valid, void *stream);
/* O(1) locate: d_counts[i] = number of sites of code d_codes[i]; then gather each code's sites
 * (ascending) to d_out + d_out_off[i].  binding_sites(kmer) = sites of code(kmer) followed by sites of
 * code(revcomp(kmer)) (locate.py:45-47), split per record by the host. */
int swga_locate_counts(swga_ctx *ctx, int k, const uint32_t *d_offsets, const uint32_t *d_codes, uint32_t q,
                       uint32_t *d_counts, void *stream);
int swga_locate_gather(swga_ctx *ctx, int k, const uint32_t *d_offsets, const uint32_t *d_positions,
                       const uint32_t *d_codes, uint32_t q, const uint64_t *d_out_off, uint32_t *d_out,
                       void *stream);
/* Per primer p: sorted union of its forward hits d_hits[d_hit_off[2p] .. d_hit_off[2p+1]) and its
 * reverse-complement hits [d_hit_off[2p+1] .. d_hit_off[2p+2]) into d_out + d_out_off[p].  The two are
 * disjoint unless d_is_pal[p] (primer == its reverse complement: the lists are equal and the union is
 * the forward list; locate.py:33's set() removes the duplicates the same way). */
int swga_union_strands(swga_ctx *ctx, const uint32_t *d_hits, const uint64_t *d_hit_off,
                       const uint8_t *d_is_pal, uint32_t n_primers, const uint64_t *d_out_off,
                       uint32_t *d_out, void *stream);
/* d_toff[l * (n_tiles+1) + t] = index into d_sites of the first site of list l that is >= t << tile_shift
 * (d_site_off has n_lists + 1 entries).  The scoring kernel walks the genome in such tiles. */
int swga_tile_offsets(swga_ctx *ctx, const uint32_t *d_sites, const uint64_t *d_site_off, uint32_t n_lists,
                      int tile_shift, uint32_t n_tiles, uint32_t *d_toff, void *stream);

/* ---- kernel (d): batched scoring of candidate primer sets ---------------------------------------
 * Replaces score.score_set / default_score_set (swga/score.py:27-53,73-98) with
 * locate.linearize_binding_sites (swga/locate.py:22-36) and stats.seq_diff/mean/stdev/gini
 * (swga/stats.py:10-54) for S sets per launch.
 *   d_sites / d_toff  site lists (sorted, unique, linearised) of every primer plus one extra list,
 *                     id `bounds_list`, holding the first and last position of every record
 *                     (locate.py:30-32 adds them for every primer); tile offsets from
 *                     swga_tile_offsets with tile_shift = swga_score_tile_shift().  d_sites must be
 *                     16-byte aligned and readable for 16 words past the last site (the dense kernel
 *                     fetches the lists with aligned 16-byte loads)
 *   d_sets            [S][max_m] list ids, -1 padded
 *   d_bg_dist_mean    [S] (set_finder's weight or score.calculate_bg_dist_mean)
 * Outputs per set: n_sites (unique positions), max_dist, mean / stdev / gini of the gaps, the default
 * score (mean*gini)/bg_dist_mean, status bit0 = passed (interactive or max_dist <= max_fg_bind_dist),
 * bit1 = largest gap beyond the in-kernel histogram (gini is NaN: use swga_gini_unbounded). */
/* Sets with at most `cap` raw sites (a power of two, 64..16384; 0 disables) are scored by the shared-memory
 * sort kernel, the rest by the bitmap-tile kernel; results are identical.  Default 4096. */
int swga_ctx_set_score_sort_cap(swga_ctx *ctx, uint32_t cap);
/* Which form of the dense-set kernel scores the sets above `cap`: 0 (default) = routed -- within a tile of 2^13
 * positions every site takes a slot of its 256-position window (one shared atomicAdd), the window's owner lane sorts
 * its slots in registers and takes the gaps from neighbouring registers; 1 = bitmap + compaction (round 1).
 * Results are identical (swga/score.py:73-98, swga/stats.py:10-54 for both).  The routed form histograms gaps up to
 * 40,126 (bitmap form: swga_score_max_hist_bins() - 1); a passing set with a larger gap is flagged (status bit1) for
 * swga_gini_unbounded by either. */
int swga_ctx_set_score_dense_impl(swga_ctx *ctx, int impl);
/* Front end of the sparse-set kernel: 0 (default) = routed -- the set's sites are counting-sorted by genome window
 * (one shared atomicAdd per site, register sorting network per window); genomes above 33.5 Mbp and sets with more
 * than 16 sites in one window take the merge; 1 = always merge the set's sorted lists (round 1).  Same results. */
int swga_ctx_set_score_sort_impl(swga_ctx *ctx, int impl);
uint32_t swga_score_tile_shift(void);
uint32_t swga_score_max_hist_bins(void);
int swga_score_sets(swga_ctx *ctx, const uint32_t *d_sites, const uint32_t *d_toff, uint32_t n_tiles,
                    uint32_t bounds_list, const int32_t *d_sets, uint32_t S, uint32_t max_m,
                    const double *d_bg_dist_mean, uint32_t max_fg_bind_dist, int interactive,
                    uint32_t *d_n_sites, uint32_t *d_max_dist, double *d_mean, double *d_stdev,
                    double *d_gini, double *d_score, uint8_t *d_status, void *stream);
int swga_gini_unbounded(swga_ctx *ctx, const uint32_t *d_sites, const uint32_t *d_toff, uint32_t n_tiles,
                        uint32_t bounds_list, const int32_t *d_set, uint32_t max_m, uint64_t max_gaps,
                        double *h_gini, void *stream);

/* ---- synthetic genomes (SURVEY.md 8d generator, bit-identical to swga_b200/synth.py) ----------- */
int swga_synth_bases(swga_ctx *ctx, uint64_t seed, uint64_t start, uint64_t n,
                     uint32_t t_a, uint32_t t_c, uint32_t t_g, uint8_t *d_ascii, void *stream);

/* Candidate sets of workload C4 (bit-identical to swga_b200/workloads.py): n_sets rows of 10
 * int32 primer indices in [0, n_primers), 6..10 per set, -1 padded, for sets start .. start + n_sets - 1. */
int swga_synth_sets(swga_ctx *ctx, uint64_t seed, uint64_t start, uint64_t n_sets, uint32_t n_primers,
                    int32_t *d_sets, void *stream);

/* Kernel launches issued by this library since it was loaded (all contexts; bench.py's gpu_launches). */
uint64_t swga_launch_count(void);

/* ---- atomic-throughput microbenchmarks (the A_peak of SURVEY.md 8d) ----------------------------
 * Issues n_ops uniformly random red.global.add.u32 on a table of `table_words` words (global) or
 * shared-memory atomics on a per-CTA table of `table_words` words; returns milliseconds. */
int swga_bench_red_global(swga_ctx *ctx, uint32_t *d_table, uint64_t table_words, uint64_t n_ops,
                          float *ms);
int swga_bench_atom_shared(swga_ctx *ctx, uint32_t table_words, uint64_t n_ops, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* SWGA_B200_H */
