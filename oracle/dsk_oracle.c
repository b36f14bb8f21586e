/*
 * oracle/dsk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the k-mer counting semantics of the DSK
 * program that swga runs behind swga/kmers.py:23-56.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this restatement against
 *   (1) golden vectors produced by the real DSK binary compiled from /root/reference/ext/dsk
 *       (recipe: oracle/Makefile -> oracle/_ref/dsk; generator: tests/golden/make_golden.py),
 *   (2) the known answers of SURVEY.md Appendix B1/B2.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ext/dsk/minia/Kmer.cpp:18-24  NT2int: (c>>1)&3  =>  A=0 C=1 T=2 G=3, every byte aliases. */
static inline unsigned nt2int(uint8_t c) { return (c >> 1) & 3u; }

/* ext/dsk/minia/lut.h:8-10  comp_NT = {2,3,0,1}  (complement == code ^ 2). */
static const unsigned comp_nt[4] = {2, 3, 0, 1};

/*
 * ext/dsk/minia/Bank.cpp:889-902 (compute_kmer_table_from_one_seq, mode B) and
 * ext/dsk/minia/Bank.cpp:1049-1070 (KmersBuffer::readkmers, mode A): rolling forward and
 * reverse-complement codes, canonical = min(fwd, rc) (Kmer.cpp:122), one increment per
 * position (OAHash::increment, OAHash.cpp:142-154).  A read shorter than k contributes
 * nothing (SortingCount.cpp:405 nbkmers <= 0; Bank.cpp:1038 binSeq_toread <= 0).
 */
static void count_read(const uint8_t *s, uint64_t len, int k, uint32_t *table)
{
    if (len < (uint64_t)k) return;
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fwd = 0, rc = 0;
    for (int i = 0; i < k; ++i) {                       /* codeSeed, Kmer.cpp:65-76 */
        unsigned b = nt2int(s[i]);
        fwd = fwd * 4 + b;
        rc |= (uint64_t)comp_nt[b] << (2 * i);          /* revcomp(), Kmer.cpp:172-186 */
    }
    table[fwd < rc ? fwd : rc]++;
    for (uint64_t i = 1; i + k <= len; ++i) {
        unsigned b = nt2int(s[i + k - 1]);
        fwd = (fwd * 4 + b) & mask;
        rc = ((rc >> 2) + ((uint64_t)comp_nt[b] << (2 * (k - 1)))) & mask;
        table[fwd < rc ? fwd : rc]++;
    }
}

/*
 * Count canonical k-mers of a set of records into a dense 4^k uint32 table (indexed by the
 * canonical code; non-canonical slots stay 0).
 *   mode_b == 0: DSK "compressed reads" path -- each record is split at runs of upper-case
 *                'N' first (SortingCount.cpp:243-261), each N-free piece is one read.
 *   mode_b == 1: DSK raw path (SortingCount.cpp:392-421) -- the record is one read and 'N'
 *                is encoded like any other byte ('N' -> 3 == G).
 * seq = records concatenated without separators; rec_starts has n_rec+1 entries.
 */
void orc_count_k(const uint8_t *seq, const uint64_t *rec_starts, uint32_t n_rec,
                 int mode_b, int k, uint32_t *table)
{
    for (uint32_t r = 0; r < n_rec; ++r) {
        const uint8_t *rs = seq + rec_starts[r];
        uint64_t rl = rec_starts[r + 1] - rec_starts[r];
        if (mode_b) { count_read(rs, rl, k, table); continue; }
        uint64_t p = 0;
        while (p < rl) {
            while (p < rl && rs[p] == 'N') p++;         /* skips NN */
            uint64_t q = p;
            while (q < rl && rs[q] != 'N') q++;         /* goes to next N or end of seq */
            count_read(rs + p, q - p, k, table);
            p = q;
        }
    }
}

/* all k in [kmin,kmax]; tables = concatenated dense canonical tables 4^kmin .. 4^kmax */
void orc_count_allk(const uint8_t *seq, const uint64_t *rec_starts, uint32_t n_rec,
                    int mode_b, int kmin, int kmax, uint32_t *tables)
{
    uint64_t off = 0;
    for (int k = kmin; k <= kmax; ++k) {
        orc_count_k(seq, rec_starts, n_rec, mode_b, k, tables + off);
        off += 1ull << (2 * k);
    }
}

/*
 * FASTA / FASTQ reader restating Bank::get_next_seq_from_file (ext/dsk/minia/Bank.cpp:150-209) with
 * buffered_getc / buffered_gets (Bank.cpp:66-72,77-128) on an in-memory file image.  A '+' at a line
 * start ends the sequence and switches to the FASTQ branch (Bank.cpp:192-201): the rest of the '+' line
 * is skipped, then quality lines are read into `dummy` -- which still holds the remainder of the header
 * line (Bank.cpp:167, append = true) -- until dummy is at least as long as the read; the next header is
 * then searched byte by byte (last_char = 0).  Bytes that read as (signed char)-1 (0xFF) end the file or
 * the record exactly as in the reference.
 *
 * Writes the joined sequence bytes to seq_out (capacity >= n) and record boundaries to
 * rec_starts (capacity max_rec+1).  Returns the number of records, or <0 on error.
 */
typedef struct { const uint8_t *b; uint64_t n, pos; } rd_t;
typedef struct { uint8_t *s; uint64_t len, max; } vs_t;     /* variable_string_t */
static int rd_getc(rd_t *r) {
    if (r->pos >= r->n) return -1;
    int c = (signed char)r->b[r->pos++];
    return c;            /* may itself be -1 for byte 0xFF, exactly like buffered_getc */
}
/* buffered_gets(bf, s, dret, append, allow_spaces): appends up to the delimiter ('\n', or any isspace()
 * byte when !allow_spaces); *dret = the delimiter (0 if the end of the file came first); returns the
 * length of s, or -1 when already at the end of the file.  `fixed`: s->s is the caller's large buffer. */
static int64_t rd_gets(rd_t *r, vs_t *s, int *dret, int append, int allow_spaces, int fixed)
{
    if (dret) *dret = 0;
    if (!append) s->len = 0;
    if (r->pos >= r->n) return -1;
    uint64_t i = r->pos;
    if (allow_spaces) { while (i < r->n && r->b[i] != '\n') i++; }
    else              { while (i < r->n && !isspace(r->b[i])) i++; }
    if (!fixed && s->max < s->len + (i - r->pos) + 1) {
        s->max = 2 * (s->len + (i - r->pos) + 1);
        s->s = (uint8_t *)realloc(s->s, s->max);
    }
    memcpy(s->s + s->len, r->b + r->pos, i - r->pos);
    s->len += i - r->pos;
    r->pos = i;
    if (i < r->n) { if (dret) *dret = r->b[i]; r->pos = i + 1; }
    if (allow_spaces && s->len > 1 && s->s[s->len - 1] == '\r') s->len--;      /* Bank.cpp:124-125 */
    return (int64_t)s->len;
}

int64_t orc_fasta_parse(const uint8_t *buf, uint64_t n, uint8_t *seq_out,
                        uint64_t *rec_starts, uint64_t max_rec)
{
    rd_t r = {buf, n, 0};
    vs_t header = {NULL, 0, 0}, dummy = {NULL, 0, 0};
    int last_char = 0, c;
    uint64_t total = 0, nrec = 0;
    int64_t ret = 0;
    for (;;) {
        if (last_char == 0) {                                       /* Bank.cpp:154-160 */
            while ((c = rd_getc(&r)) != -1 && c != '>' && c != '@') {}
            if (c == -1) break;
            last_char = c;
        }
        dummy.len = 0;                                              /* Bank.cpp:161 */
        if (rd_gets(&r, &header, &c, 0, 0, 0) < 0) break;           /* header token */
        if (c != '\n') rd_gets(&r, &dummy, NULL, 1, 1, 0);          /* rest of the header line, into dummy */
        if (nrec >= max_rec) { ret = -3; break; }
        rec_starts[nrec] = total;
        vs_t read = {seq_out + total, 0, 0};
        while ((c = rd_getc(&r)) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;                                /* empty line */
            read.s[read.len++] = (uint8_t)c;
            rd_gets(&r, &read, NULL, 1, 1, 1);
        }
        if (c == '>' || c == '@') last_char = c;
        if (c == '+') {                                             /* fastq, Bank.cpp:192-201 */
            while ((c = rd_getc(&r)) != -1 && c != '\n') {}         /* rest of the quality comment */
            while (rd_gets(&r, &dummy, NULL, 1, 1, 0) >= 0 && dummy.len < read.len) {}
            last_char = 0;
        }
        total += read.len;
        nrec++;
    }
    free(header.s);
    free(dummy.s);
    if (ret < 0) return ret;
    rec_starts[nrec] = total;
    return (int64_t)nrec;
}

/*
 * DSK picks its read mode from the longest of the first 10,001 records
 * (Bank::estimate_max_readlen, Bank.cpp:470-494: NbRead++ == 10000 breaks after the
 * 10,001st; SortingCount.cpp:81-82: > 1000000 => raw mode).  Returns 1 for mode B.
 */
int orc_dsk_mode(const uint64_t *rec_starts, uint64_t n_rec)
{
    uint64_t lim = n_rec < 10001 ? n_rec : 10001, mx = 0;
    for (uint64_t r = 0; r < lim; ++r) {
        uint64_t l = rec_starts[r + 1] - rec_starts[r];
        if (l > mx) mx = l;
    }
    return mx > 1000000 ? 1 : 0;
}
