/*
 * deflate.c — oracle restatement of the reference compressor (TEST ONLY).
 *
 * Restates, function by function (citations into /root/reference):
 *   Bitstream                         src/compress/bitstream.rs:1-223
 *   make_huffman_code & helpers       src/compress/huffman_comp.rs:8-155
 *   HtMatchFinder (level 1)           src/compress/matchfinder.rs:1109-1232
 *   MatchFinder, hash chains (2..9)   src/compress/matchfinder.rs:721-1107
 *   BtMatchFinder (10..12)            src/compress/matchfinder.rs:1308-1776   [bt.inc.c]
 *   level table / init_params         src/compress/mod.rs:543-602
 *   BlockSplitStats                   src/compress/mod.rs:271-416
 *   compress_loop / compress          src/compress/mod.rs:604-790
 *   decide_greedy_sequences           src/compress/mod.rs:1261-1373
 *   compress_uncompressed (level 0)   src/compress/mod.rs:1400-1464
 *   compress_greedy_block             src/compress/mod.rs:1466-1584
 *   compress_near_optimal_block       src/compress/mod.rs:1586-1773
 *   write_dynamic_huffman_header_impl src/compress/mod.rs:1775-1883
 *   write_sequences_to_bitstream      src/compress/mod.rs:1952-2155
 *   bounds, zlib and gzip framing     src/compress/mod.rs:2236-2357
 *
 * Compressed bytes are "parity unpinned" (see oracle.h): this file IS the
 * definition of byte-identity for levels 0..9 in this repository.
 *
 * Simplifications that cannot change the output:
 *  - One matchfinder state per call starting from cleared tables.  The
 *    reference reuses a Compressor per rayon worker and advances base_offset;
 *    every stale entry then fails `cur_pos < base_offset`
 *    (matchfinder.rs:791,1041,1086,1165), which is the same as cleared tables.
 *  - Bitstream space checks: every checked write in the reference fails
 *    exactly when a completed byte no longer fits (bitstream.rs:143-189,
 *    194-222; the fast paths in mod.rs:1968,2071 only run with >= 16 spare
 *    bytes), so "ceil(total_bits / 8) <= capacity" is the success condition.
 *  - match_len_*: all variants return the common-prefix length capped at
 *    max_len (matchfinder.rs:245-694).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define MIN_MATCH 3
#define MAX_MATCH 258
#define MAX_OFFSET 32768
#define HASH_ORDER 15
#define HASH_SIZE (1u << HASH_ORDER)
#define WINDOW_MASK 32767u
#define NUM_LITLEN 288
#define NUM_OFFSET 32
#define MIN_BLOCK_LENGTH 5000u
#define SOFT_MAX_BLOCK_LENGTH 300000u
#define MAX_LITLEN_CODEWORD_LEN 14
#define MAX_OFFSET_CODEWORD_LEN 15
#define MAX_PRE_CODEWORD_LEN 7

/* RFC 1951 length / offset code tables.  The reference packs the same data
 * into LENGTH_WRITE_TABLE / OFFSET_*_TABLE / OFFSET_SLOT_TABLE_512
 * (src/compress/mod.rs:16-105); SURVEY appendix A records that those were
 * machine-checked to be the standard tables. */
static const uint16_t len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,
                                      15, 17, 19, 23, 27, 31, 35, 43, 51,  59,
                                      67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                      2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t off_base[30] = {
    1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
    33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
    1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t off_extra[30] = {0, 0, 0,  0,  1,  1,  2,  2,  3,  3,
                                      4, 4, 5,  5,  6,  6,  7,  7,  8,  8,
                                      9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static uint8_t len_slot_tab[MAX_MATCH + 1];
static uint8_t off_slot_small[257]; /* offsets 1..256 */
static int tabs_ready;

static void init_tabs(void)
{
    for (unsigned s = 0; s < 29; s++) {
        unsigned hi = s == 28 ? 258 : len_base[s] + (1u << len_extra[s]) - 1;
        for (unsigned l = len_base[s]; l <= hi && l <= MAX_MATCH; l++)
            len_slot_tab[l] = (uint8_t)s; /* 258 ends in slot 28 */
    }
    for (unsigned o = 1; o <= 256; o++) {
        unsigned s = 0;
        while (s + 1 < 30 && off_base[s + 1] <= o)
            s++;
        off_slot_small[o] = (uint8_t)s;
    }
    __atomic_store_n(&tabs_ready, 1, __ATOMIC_RELEASE);
}

/* get_offset_slot, src/compress/mod.rs:2197-2207 */
static inline unsigned offset_slot(unsigned off)
{
    if (off <= 256)
        return off_slot_small[off];
    unsigned v = off - 1;
    unsigned l = 31u - (unsigned)__builtin_clz(v);
    return 2 * l + ((v >> (l - 1)) & 1);
}
static inline unsigned length_slot(unsigned len) { return len_slot_tab[len]; }

/* ------------------------------------------------------------------ bits */

typedef struct {
    uint8_t *out;
    size_t cap, pos;
    uint64_t buf;
    unsigned cnt;
    int overflow;
} bitw;

static inline void bw_put(bitw *b, uint32_t bits, unsigned n)
{
    /* LSB-first concatenation, src/compress/bitstream.rs:123-192 */
    b->buf |= (uint64_t)bits << b->cnt;
    b->cnt += n;
    while (b->cnt >= 8) {
        if (b->pos < b->cap)
            b->out[b->pos] = (uint8_t)b->buf;
        else
            b->overflow = 1;
        b->pos++;
        b->buf >>= 8;
        b->cnt -= 8;
    }
}
/* Bitstream::flush, src/compress/bitstream.rs:194-222: zero-pad to a byte */
static inline void bw_flush(bitw *b)
{
    if (b->cnt)
        bw_put(b, 0, 8 - b->cnt);
}

/* --------------------------------------------------------------- huffman */

#define SYM_BITS 10
#define SYM_MASK ((1u << SYM_BITS) - 1)
#define FREQ_MASK (~SYM_MASK)

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

/* make_huffman_code, src/compress/huffman_comp.rs:125-155, with
 * sort_symbols :8-36, build_tree :38-62, compute_length_counts :64-89,
 * gen_codewords :95-123 inlined in that order. */
static void make_huffman_code(unsigned num_syms, unsigned max_len,
                              const uint32_t *freqs, uint8_t *lens,
                              uint32_t *a /* codewords, doubles as scratch */)
{
    uint32_t counters[NUM_LITLEN];
    memset(counters, 0, sizeof(counters));
    /* counting sort on min(freq, num_syms-1); the last bucket holds every
     * larger frequency and is ordered by the packed (freq<<10 | sym) key */
    for (unsigned s = 0; s < num_syms; s++) {
        uint32_t f = freqs[s];
        counters[f < num_syms - 1 ? f : num_syms - 1]++;
    }
    uint32_t run = 0;
    for (unsigned i = 1; i < num_syms; i++) {
        uint32_t c = counters[i];
        counters[i] = run;
        run += c;
    }
    unsigned used = run;
    for (unsigned s = 0; s < num_syms; s++) {
        uint32_t f = freqs[s];
        if (f) {
            uint32_t *slot = &counters[f < num_syms - 1 ? f : num_syms - 1];
            a[(*slot)++] = s | (f << SYM_BITS);
        } else {
            lens[s] = 0;
        }
    }
    {
        uint32_t lo = counters[num_syms - 2], hi = counters[num_syms - 1];
        if (hi > lo)
            qsort(a + lo, hi - lo, sizeof(uint32_t), cmp_u32);
    }
    if (used < 2) {
        /* degenerate alphabets, huffman_comp.rs:133-144 */
        unsigned sym = used ? (a[0] & SYM_MASK) : 0;
        unsigned nz = sym ? sym : 1;
        a[0] = 0;
        lens[0] = 1;
        a[nz] = 1;
        lens[nz] = 1;
        return;
    }
    /* in-place two-queue tree build: leaves a[i..], internal nodes a[b..e) */
    {
        unsigned last = used - 1, i = 0, b = 0, e = 0;
        while (e < last) {
            uint32_t nf;
            if (i < last && (b == e || (a[i + 1] & FREQ_MASK) <= (a[b] & FREQ_MASK))) {
                nf = (a[i] & FREQ_MASK) + (a[i + 1] & FREQ_MASK);
                i += 2;
            } else if (b + 2 <= e && (i > last || (a[b + 1] & FREQ_MASK) < (a[i] & FREQ_MASK))) {
                nf = (a[b] & FREQ_MASK) + (a[b + 1] & FREQ_MASK);
                a[b] = (e << SYM_BITS) | (a[b] & SYM_MASK);
                a[b + 1] = (e << SYM_BITS) | (a[b + 1] & SYM_MASK);
                b += 2;
            } else {
                nf = (a[i] & FREQ_MASK) + (a[b] & FREQ_MASK);
                a[b] = (e << SYM_BITS) | (a[b] & SYM_MASK);
                i += 1;
                b += 1;
            }
            a[e] = nf | (a[e] & SYM_MASK);
            e++;
        }
    }
    /* depth counts with the overflow fix-up to max_len */
    uint32_t len_counts[16];
    memset(len_counts, 0, sizeof(len_counts));
    {
        unsigned root = used - 2;
        len_counts[1] = 2;
        a[root] &= SYM_MASK;
        for (int node = (int)root - 1; node >= 0; node--) {
            unsigned parent = a[node] >> SYM_BITS;
            unsigned depth = (a[parent] >> SYM_BITS) + 1;
            a[node] = (a[node] & SYM_MASK) | (depth << SYM_BITS);
            if (depth >= max_len) {
                depth = max_len - 1;
                while (len_counts[depth] == 0)
                    depth--;
            }
            len_counts[depth]--;
            len_counts[depth + 1] += 2;
        }
    }
    /* assign lengths (longest to the rarest), then canonical codewords,
     * stored bit-reversed for the LSB-first writer */
    {
        unsigned i = 0;
        for (unsigned len = max_len; len >= 1; len--)
            for (uint32_t c = len_counts[len]; c > 0; c--)
                lens[a[i++] & SYM_MASK] = (uint8_t)len;
        uint32_t next[16];
        next[0] = 0;
        next[1] = 0;
        for (unsigned len = 2; len <= max_len; len++)
            next[len] = (next[len - 1] + len_counts[len - 1]) << 1;
        for (unsigned s = 0; s < num_syms; s++) {
            unsigned l = lens[s];
            if (l) {
                uint32_t c = next[l]++, r = 0;
                for (unsigned k = 0; k < l; k++)
                    r |= ((c >> k) & 1) << (l - 1 - k);
                a[s] = r;
            }
        }
    }
}

/* ------------------------------------------------------------ compressor */

typedef struct {
    uint32_t litrunlen;
    uint16_t length; /* 0 terminates the block */
    uint16_t offset;
} sequence;

/* BlockSplitStats, src/compress/mod.rs:271-416 */
typedef struct {
    uint32_t new_obs[14], obs[14];
    uint32_t num_new, num;
} split_stats;

typedef struct {
    int level;
    unsigned max_depth, nice_len;
    /* hash chains / hash table; positions are stream-relative, -1 = empty */
    int32_t *head;  /* HASH_SIZE */
    uint16_t *prev; /* 32768, indexed pos & 32767 (matchfinder.rs:794) */
    uint32_t litlen_freqs[NUM_LITLEN], offset_freqs[NUM_OFFSET];
    uint32_t litlen_codes[NUM_LITLEN], offset_codes[NUM_OFFSET];
    uint8_t litlen_lens[NUM_LITLEN], offset_lens[NUM_OFFSET];
    sequence *seqs;
    size_t nseq, seq_cap;
    split_stats st;
    struct bt_state *bt; /* levels >= 10 */
    uint32_t *dp_cost, *dp_path;
    size_t dp_cap;
} compressor;

static inline uint32_t ld24(const uint8_t *p)
{
    return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16;
}
static inline uint32_t ld32(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
static inline uint32_t hash3(uint32_t v24) { return (v24 * 0x1E35A7BDu) >> (32 - HASH_ORDER); }

static inline unsigned prefix_len(const uint8_t *a, const uint8_t *b, unsigned max)
{
    unsigned n = 0;
    while (n + 8 <= max) {
        uint64_t x, y;
        memcpy(&x, a + n, 8);
        memcpy(&y, b + n, 8);
        if (x != y)
            return n + ((unsigned)__builtin_ctzll(x ^ y) >> 3);
        n += 8;
    }
    while (n < max && a[n] == b[n])
        n++;
    return n;
}

/* init_params, src/compress/mod.rs:543-602 */
static void init_params(compressor *c)
{
    static const uint16_t tab[13][2] = {
        {0, 0},   {2, 32},   {6, 10},   {12, 14},  {16, 30},  {16, 30}, {35, 65},
        {100, 130}, {300, 258}, {600, 258}, {35, 75}, {100, 150}, {300, 258}};
    int l = c->level > 12 ? 12 : c->level;
    c->max_depth = tab[l][0];
    c->nice_len = tab[l][1];
}

static void push_seq(compressor *c, uint32_t litrun, unsigned len, unsigned off)
{
    if (c->nseq == c->seq_cap) {
        c->seq_cap = c->seq_cap ? c->seq_cap * 2 : 4096;
        c->seqs = (sequence *)realloc(c->seqs, c->seq_cap * sizeof(sequence));
    }
    c->seqs[c->nseq].litrunlen = litrun;
    c->seqs[c->nseq].length = (uint16_t)len;
    c->seqs[c->nseq].offset = (uint16_t)off;
    c->nseq++;
}

/* --- level 1: HtMatchFinder, src/compress/matchfinder.rs:1139-1231 */
static unsigned ht_find_match(compressor *c, const uint8_t *d, size_t n, size_t pos,
                              unsigned *off_out)
{
    if (pos + 3 > n)
        return 0;
    uint32_t v = ld24(d + pos);
    uint32_t h = hash3(v);
    int32_t cur = c->head[h];
    c->head[h] = (int32_t)pos; /* bucket overwritten before any check (:1162-1163) */
    if (cur < 0)
        return 0;
    size_t off = pos - (size_t)cur;
    if (off > MAX_OFFSET)
        return 0;
    if (ld24(d + cur) != v)
        return 0;
    size_t room = n - pos;
    *off_out = (unsigned)off;
    return prefix_len(d + cur, d + pos, room < MAX_MATCH ? (unsigned)room : MAX_MATCH);
}

/* --- levels 2..9: MatchFinder, src/compress/matchfinder.rs:754-891 */
static inline void hc_insert(compressor *c, const uint8_t *d, size_t n, size_t pos)
{
    /* skip_match :1020-1057 == skip_positions :1059-1106 per position */
    if (pos + 3 > n)
        return;
    uint32_t h = hash3(ld24(d + pos));
    int32_t cur = c->head[h];
    c->head[h] = (int32_t)pos;
    size_t link = cur >= 0 ? pos - (size_t)cur : 0;
    c->prev[pos & WINDOW_MASK] = link > 0xFFFF ? 0 : (uint16_t)link;
}
static void hc_skip(compressor *c, const uint8_t *d, size_t n, size_t pos, size_t count)
{
    for (size_t i = 0; i < count; i++)
        hc_insert(c, d, n, pos + i);
}

static unsigned hc_find_match(compressor *c, const uint8_t *d, size_t n, size_t pos,
                              unsigned *off_out)
{
    if (pos + 3 > n)
        return 0;
    const int can4 = pos + 4 <= n;
    const uint8_t *src = d + pos;
    const uint32_t v4 = can4 ? ld32(src) : 0;
    const uint32_t v3 = can4 ? (v4 & 0xFFFFFF) : ld24(src);
    const uint32_t h = hash3(v3);
    int64_t cur = c->head[h];
    c->head[h] = (int32_t)pos;
    if (cur < 0) {
        c->prev[pos & WINDOW_MASK] = 0;
        return 0;
    }
    {
        size_t link = pos - (size_t)cur;
        c->prev[pos & WINDOW_MASK] = link > 0xFFFF ? 0 : (uint16_t)link;
    }
    unsigned best = 0, best_off = 0, depth = 0;
    const unsigned room = n - pos < MAX_MATCH ? (unsigned)(n - pos) : MAX_MATCH;
    while (cur >= 0 && depth < c->max_depth) {
        /* cur < 0 other than the initial -1 can only come from the aliased
         * link of a candidate at offset exactly 32768 (:880-887); the
         * reference then fails the offset test on the wrapped value. */
        size_t off = pos - (size_t)cur;
        if (off > MAX_OFFSET)
            break;
        if (pos + best >= n)
            break;
        const uint8_t *m = d + cur;
        if (!(best >= 3 && m[best] != src[best])) {
            if (can4) {
                uint32_t m4 = ld32(m);
                if (m4 == v4) {
                    unsigned len = prefix_len(m, src, room);
                    if (len > best) {
                        best = len;
                        best_off = (unsigned)off;
                        if (len >= c->nice_len || len == MAX_MATCH)
                            break;
                    }
                } else if (best < 3 && (m4 & 0xFFFFFF) == v3) {
                    best = 3;
                    best_off = (unsigned)off;
                }
            } else if (ld24(m) == v3) {
                unsigned len = prefix_len(m, src, room);
                if (len > best) {
                    best = len;
                    best_off = (unsigned)off;
                    if (len >= c->nice_len || len == MAX_MATCH)
                        break;
                }
            }
        }
        unsigned link = c->prev[(size_t)cur & WINDOW_MASK];
        if (!link)
            break;
        cur -= link;
        depth++;
    }
    *off_out = best_off;
    return best;
}

#include "bt.inc.c"

/* matchfinder dispatch by level, src/compress/mod.rs:476-482 */
static inline unsigned mf_find(compressor *c, const uint8_t *d, size_t n, size_t pos,
                               unsigned *off)
{
    if (c->level == 1)
        return ht_find_match(c, d, n, pos, off);
    if (c->level >= 10)
        return bt_find_match(c->bt, d, n, pos, c->max_depth, c->nice_len, off);
    return hc_find_match(c, d, n, pos, off);
}
static inline void mf_skip(compressor *c, const uint8_t *d, size_t n, size_t pos, size_t count)
{
    if (c->level == 1)
        return; /* HtMatchFinder::skip_positions is a no-op (:1231) */
    if (c->level >= 10)
        bt_skip_positions(c->bt, d, n, pos, count, c->max_depth, c->nice_len);
    else
        hc_skip(c, d, n, pos, count);
}

/* --- block split statistics */
static inline unsigned bsr32(uint32_t v) { return 31u - (unsigned)__builtin_clz(v); }
static inline void st_reset(split_stats *s) { memset(s, 0, sizeof(*s)); }
static inline void st_literal(split_stats *s, uint8_t lit)
{
    s->new_obs[lit >> 5]++;
    s->num_new++;
}
/* observe_match_with_slot + SLOT_TO_OBS_IDX, src/compress/mod.rs:120-125,332-346 */
static inline void st_match(split_stats *s, unsigned len, unsigned slot)
{
    s->new_obs[8 + (len >= 8)]++;
    s->new_obs[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)]++;
    s->num_new += 2;
}
static int st_should_end(split_stats *s, size_t block_len, size_t remaining)
{
    /* should_end_block :387-415, do_end_block_check :359-384 */
    if (s->num_new < 2048 && block_len < SOFT_MAX_BLOCK_LENGTH)
        return 0;
    if (remaining <= MIN_BLOCK_LENGTH)
        return 0;
    if (block_len >= SOFT_MAX_BLOCK_LENGTH)
        return 1;
    if (block_len >= MIN_BLOCK_LENGTH) {
        if (s->num != 0) {
            uint32_t old_bits = 0, new_bits = 0;
            uint32_t lg_all = bsr32(s->num), lg_new = bsr32(s->num_new);
            for (int i = 0; i < 14; i++) {
                uint32_t k = s->new_obs[i];
                if (k) {
                    uint32_t lo = bsr32(s->obs[i] + 1), ln = bsr32(k + 1);
                    old_bits += k * (lg_all > lo ? lg_all - lo : 0);
                    new_bits += k * (lg_new > ln ? lg_new - ln : 0);
                }
            }
            if ((int32_t)old_bits - (int32_t)new_bits > (int32_t)block_len / 16)
                return 1;
        }
        for (int i = 0; i < 14; i++) {
            s->obs[i] += s->new_obs[i];
            s->new_obs[i] = 0;
        }
        s->num += s->num_new;
        s->num_new = 0;
    }
    return 0;
}

/* --- emission */

/* write_dynamic_huffman_header_impl, src/compress/mod.rs:1775-1883 */
static void write_dynamic_header(compressor *c, bitw *bs)
{
    static const uint8_t perm[19] = {16, 17, 18, 0, 8,  7, 9,  6, 10, 5,
                                     11, 4,  12, 3, 13, 2, 14, 1, 15};
    unsigned nlit = NUM_LITLEN, noff = NUM_OFFSET;
    while (nlit > 257 && c->litlen_lens[nlit - 1] == 0)
        nlit--;
    while (noff > 1 && c->offset_lens[noff - 1] == 0)
        noff--;
    bw_put(bs, nlit - 257, 5);
    bw_put(bs, noff - 1, 5);
    uint8_t lens[NUM_LITLEN + NUM_OFFSET];
    unsigned total = nlit + noff;
    memcpy(lens, c->litlen_lens, nlit);
    memcpy(lens + nlit, c->offset_lens, noff);

    uint32_t pre_freq[19] = {0};
    uint16_t items[NUM_LITLEN + NUM_OFFSET];
    unsigned nitems = 0;
    for (unsigned i = 0; i < total;) {
        unsigned len = lens[i], run = 1;
        while (i + run < total && lens[i + run] == len)
            run++;
        i += run;
        if (len == 0) {
            while (run >= 11) {
                unsigned k = run < 138 ? run : 138;
                items[nitems++] = (uint16_t)(18 << 8 | (k - 11));
                pre_freq[18]++;
                run -= k;
            }
            if (run >= 3) {
                unsigned k = run < 10 ? run : 10;
                items[nitems++] = (uint16_t)(17 << 8 | (k - 3));
                pre_freq[17]++;
                run -= k;
            }
        } else if (run >= 4) {
            items[nitems++] = (uint16_t)(len << 8);
            pre_freq[len]++;
            run--;
            while (run >= 3) {
                unsigned k = run < 6 ? run : 6;
                items[nitems++] = (uint16_t)(16 << 8 | (k - 3));
                pre_freq[16]++;
                run -= k;
            }
        }
        while (run--) {
            items[nitems++] = (uint16_t)(len << 8);
            pre_freq[len]++;
        }
    }
    uint8_t pre_lens[19] = {0};
    uint32_t pre_codes[NUM_LITLEN]; /* scratch sized for make_huffman_code */
    make_huffman_code(19, MAX_PRE_CODEWORD_LEN, pre_freq, pre_lens, pre_codes);
    unsigned npre = 19;
    while (npre > 4 && pre_lens[perm[npre - 1]] == 0)
        npre--;
    bw_put(bs, npre - 4, 4);
    for (unsigned j = 0; j < npre; j++)
        bw_put(bs, pre_lens[perm[j]], 3);
    for (unsigned k = 0; k < nitems; k++) {
        unsigned sym = items[k] >> 8, extra = items[k] & 0xFF;
        bw_put(bs, pre_codes[sym], pre_lens[sym]);
        if (sym == 16)
            bw_put(bs, extra, 2);
        else if (sym == 17)
            bw_put(bs, extra, 3);
        else if (sym == 18)
            bw_put(bs, extra, 7);
    }
}

/* write_sequences_to_bitstream + write_sym(256), src/compress/mod.rs:1952-2165.
 * The 4/2/1-literal packing and the match_len_table / offset_table layouts
 * (:509-541) are speed devices; the emitted bits are the plain concatenation
 * code, length extra bits, offset code, offset extra bits. */
static void write_sequences(compressor *c, bitw *bs, const uint8_t *in, size_t start)
{
    size_t p = start;
    for (size_t k = 0; k < c->nseq; k++) {
        const sequence *s = &c->seqs[k];
        for (uint32_t i = 0; i < s->litrunlen; i++, p++)
            bw_put(bs, c->litlen_codes[in[p]], c->litlen_lens[in[p]]);
        unsigned len = s->length;
        if (len >= 3) {
            unsigned ls = length_slot(len), os = offset_slot(s->offset);
            bw_put(bs, c->litlen_codes[257 + ls], c->litlen_lens[257 + ls]);
            if (len_extra[ls])
                bw_put(bs, len - len_base[ls], len_extra[ls]);
            bw_put(bs, c->offset_codes[os], c->offset_lens[os]);
            if (off_extra[os])
                bw_put(bs, s->offset - off_base[os], off_extra[os]);
            p += len;
        }
    }
    bw_put(bs, c->litlen_codes[256], c->litlen_lens[256]);
}

static void record_literal(compressor *c, uint8_t b)
{
    st_literal(&c->st, b);
    c->litlen_freqs[b]++;
}

/* decide_greedy_sequences, src/compress/mod.rs:1261-1373 */
static size_t decide_greedy_sequences(compressor *c, const uint8_t *in, size_t n,
                                      size_t start, int lazy)
{
    c->nseq = 0;
    st_reset(&c->st);
    memset(c->litlen_freqs, 0, sizeof(c->litlen_freqs));
    memset(c->offset_freqs, 0, sizeof(c->offset_freqs));
    uint32_t litrun = 0;
    size_t p = start;
    while (p < n) {
        if (st_should_end(&c->st, p - start, n - p))
            break;
        unsigned off = 0, len = mf_find(c, in, n, p, &off);
        if (len < 3) {
            record_literal(c, in[p]);
            litrun++;
            p++;
            continue;
        }
        unsigned skipped = 0; /* positions after p already inserted by a probe */
        if (lazy >= 1 && p + 1 < n && len < c->nice_len) {
            unsigned off1 = 0, len1 = mf_find(c, in, n, p + 1, &off1);
            if (len1 > len) {
                if (lazy >= 2 && p + 2 < n) {
                    unsigned off2 = 0, len2 = mf_find(c, in, n, p + 2, &off2);
                    if (len2 > len1) {
                        record_literal(c, in[p++]);
                        record_literal(c, in[p++]);
                        litrun += 2;
                        len = len2;
                        off = off2;
                    } else {
                        record_literal(c, in[p++]);
                        litrun += 1;
                        len = len1;
                        off = off1;
                        skipped = 1;
                    }
                } else {
                    record_literal(c, in[p++]);
                    litrun += 1;
                    len = len1;
                    off = off1;
                }
            } else {
                skipped = 1;
            }
        }
        unsigned slot = offset_slot(off);
        push_seq(c, litrun, len, off);
        st_match(&c->st, len, slot);
        c->litlen_freqs[257 + length_slot(len)]++;
        c->offset_freqs[slot]++;
        litrun = 0;
        if (len - 1 > skipped)
            mf_skip(c, in, n, p + 1 + skipped, len - 1 - skipped);
        p += len;
    }
    push_seq(c, litrun, 0, 0);
    c->litlen_freqs[256]++;
    return p - start;
}

static void make_block_codes(compressor *c)
{
    make_huffman_code(NUM_LITLEN, MAX_LITLEN_CODEWORD_LEN, c->litlen_freqs,
                      c->litlen_lens, c->litlen_codes);
    make_huffman_code(NUM_OFFSET, MAX_OFFSET_CODEWORD_LEN, c->offset_freqs,
                      c->offset_lens, c->offset_codes);
}

/* write_dynamic_block_with_sequences, src/compress/mod.rs:1375-1398 */
static void write_dynamic_block(compressor *c, bitw *bs, const uint8_t *in, size_t start,
                                int is_final)
{
    bw_put(bs, is_final ? 1 : 0, 1);
    bw_put(bs, 2, 2);
    write_dynamic_header(c, bs);
    write_sequences(c, bs, in, start);
}

/* compute_static_tables / load_static_huffman_codes, src/compress/mod.rs:161-234,1885-1895 */
static void load_static_codes(compressor *c)
{
    unsigned i = 0;
    for (; i < 144; i++) c->litlen_lens[i] = 8;
    for (; i < 256; i++) c->litlen_lens[i] = 9;
    for (; i < 280; i++) c->litlen_lens[i] = 7;
    for (; i < 288; i++) c->litlen_lens[i] = 8;
    for (i = 0; i < 32; i++) c->offset_lens[i] = 5;
    /* gen_codewords_from_lens :131-151 */
    for (int which = 0; which < 2; which++) {
        uint8_t *lens = which ? c->offset_lens : c->litlen_lens;
        uint32_t *codes = which ? c->offset_codes : c->litlen_codes;
        unsigned nsym = which ? 32 : 288, maxl = which ? 5 : 9;
        uint32_t cnt[16] = {0}, next[16] = {0}, code = 0;
        for (unsigned s = 0; s < nsym; s++) cnt[lens[s]]++;
        cnt[0] = 0;
        for (unsigned l = 1; l <= maxl; l++) {
            code = (code + cnt[l - 1]) << 1;
            next[l] = code;
        }
        for (unsigned s = 0; s < nsym; s++) {
            unsigned l = lens[s];
            uint32_t v = next[l]++, r = 0;
            for (unsigned k = 0; k < l; k++)
                r |= ((v >> k) & 1) << (l - 1 - k);
            codes[s] = r;
        }
    }
}

/* compress_greedy_block, src/compress/mod.rs:1466-1584 */
static size_t compress_greedy_block(compressor *c, const uint8_t *in, size_t n, size_t start,
                                    bitw *bs, int lazy, int final_block)
{
    if (c->level >= 2) {
        size_t done = decide_greedy_sequences(c, in, n, start, lazy);
        make_block_codes(c);
        write_dynamic_block(c, bs, in, start, start + done >= n && final_block);
        return done;
    }
    /* level 1: static codes, single-probe hash table */
    load_static_codes(c);
    c->nseq = 0;
    st_reset(&c->st);
    uint32_t litrun = 0;
    size_t p = start;
    const int split = n > 65536; /* <= 64 KiB: one block, no split statistics (:1505) */
    while (p < n) {
        if (split && st_should_end(&c->st, p - start, n - p))
            break;
        unsigned off = 0, len = mf_find(c, in, n, p, &off);
        if (len >= 3) {
            if (split)
                st_match(&c->st, len, offset_slot(off));
            push_seq(c, litrun, len, off);
            litrun = 0;
            p += len;
        } else {
            if (split)
                st_literal(&c->st, in[p]);
            litrun++;
            p++;
        }
    }
    push_seq(c, litrun, 0, 0);
    size_t done = p - start;
    bw_put(bs, (start + done >= n && final_block) ? 1 : 0, 1);
    bw_put(bs, 1, 2);
    write_sequences(c, bs, in, start);
    return done;
}

#include "nearopt.inc.c"

/* compress_uncompressed, src/compress/mod.rs:1400-1464 */
static int compress_stored(const uint8_t *in, size_t n, uint8_t *out, size_t cap, int finish,
                           int sync, size_t *out_size)
{
    size_t ip = 0, op = 0;
    while (ip < n) {
        size_t blk = n - ip < 65535 ? n - ip : 65535;
        int bfinal = (ip + 65535 >= n) && finish;
        if (op + 1 > cap || op + 1 + 4 + blk > cap)
            return ORC_INSUFFICIENT_SPACE;
        out[op++] = (uint8_t)bfinal; /* 3 header bits, then flush to a byte */
        out[op++] = (uint8_t)blk;
        out[op++] = (uint8_t)(blk >> 8);
        out[op++] = (uint8_t)~blk;
        out[op++] = (uint8_t)(~blk >> 8);
        memcpy(out + op, in + ip, blk);
        op += blk;
        ip += blk;
    }
    if (sync) {
        if (op + 5 > cap)
            return ORC_INSUFFICIENT_SPACE;
        out[op++] = 0;
        out[op++] = 0;
        out[op++] = 0;
        out[op++] = 0xFF;
        out[op++] = 0xFF;
    }
    *out_size = op;
    return ORC_OK;
}

static void compressor_free(compressor *c);

/* compress_loop for one buffer of at most 256 KiB, src/compress/mod.rs:604-691,774-789 */
static int compress_one(int level, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                        int finish, int sync, size_t *out_size)
{
    if (level == 0)
        return compress_stored(in, n, out, cap, finish, sync, out_size);
    compressor c;
    memset(&c, 0, sizeof(c));
    c.level = level;
    init_params(&c);
    c.head = (int32_t *)malloc(HASH_SIZE * sizeof(int32_t));
    memset(c.head, 0xFF, HASH_SIZE * sizeof(int32_t));
    c.prev = (uint16_t *)calloc(32768, sizeof(uint16_t));
    if (level >= 10)
        c.bt = bt_new();
    bitw bs = {out, cap, 0, 0, 0, 0};
    size_t p = 0;
    const int lazy = level >= 8 ? 2 : level >= 5 ? 1 : 0;
    while (p < n) {
        if (level >= 10)
            p += compress_near_optimal_block(&c, in, n, p, &bs, finish);
        else
            p += compress_greedy_block(&c, in, n, p, &bs, lazy, finish);
    }
    if (n == 0 && finish) {
        /* empty input: one final block holding only end-of-block (:648-660) */
        if (level >= 10)
            compress_near_optimal_block(&c, in, 0, 0, &bs, 1);
        else
            compress_greedy_block(&c, in, 0, 0, &bs, 0, 1);
    }
    if (sync) {
        /* FlushMode::Sync, :662-681: empty stored block 00 00 FF FF */
        bw_put(&bs, 0, 3);
        bw_flush(&bs);
        bw_put(&bs, 0x0000, 16);
        bw_put(&bs, 0xFFFF, 16);
    }
    bw_flush(&bs);
    compressor_free(&c);
    if (bs.overflow || bs.pos > cap)
        return ORC_INSUFFICIENT_SPACE;
    *out_size = bs.pos;
    return ORC_OK;
}

static void compressor_free(compressor *c)
{
    free(c->head);
    free(c->prev);
    free(c->seqs);
    free(c->dp_cost);
    free(c->dp_path);
    if (c->bt)
        bt_free(c->bt);
}

size_t orc_compress_bound(int format, size_t len)
{
    size_t b = len + (len / 65535 + 1) * 5 + 10;
    return b + (format == ORC_FMT_ZLIB ? 6 : format == ORC_FMT_GZIP ? 18 : 0);
}

/* Compressor::compress, src/compress/mod.rs:693-790: inputs above 256 KiB are
 * cut into 256 KiB chunks, each through a fresh compressor, all but the last
 * ending in a sync flush (:699-772). */
static int compress_raw(int level, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                        size_t *out_size)
{
    const size_t chunk = 256 * 1024;
    if (n <= chunk)
        return compress_one(level, in, n, out, cap, 1, 0, out_size);
    size_t op = 0;
    for (size_t ip = 0; ip < n; ip += chunk) {
        size_t len = n - ip < chunk ? n - ip : chunk;
        int last = ip + len >= n;
        size_t bound = orc_compress_bound(ORC_FMT_RAW, len), sz = 0;
        uint8_t *tmp = (uint8_t *)malloc(bound);
        int st = compress_one(level, in + ip, len, tmp, bound, last, !last, &sz);
        if (st == ORC_OK && op + sz > cap)
            st = ORC_INSUFFICIENT_SPACE;
        if (st == ORC_OK) {
            memcpy(out + op, tmp, sz);
            op += sz;
        }
        free(tmp);
        if (st != ORC_OK)
            return st;
    }
    *out_size = op;
    return ORC_OK;
}

int orc_compress_unit(int level, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                      int finish, int sync, size_t *out_size)
{
    if (!__atomic_load_n(&tabs_ready, __ATOMIC_ACQUIRE))
        init_tabs();
    *out_size = 0;
    if (level < 0)
        level = 0;
    if (in_len > 256 * 1024)
        return ORC_BAD_DATA;
    return compress_one(level, in, in_len, out, out_cap, finish, sync, out_size);
}

int orc_compress(int level, int format, const uint8_t *in, size_t in_len, uint8_t *out,
                 size_t out_cap, size_t *out_size)
{
    if (!__atomic_load_n(&tabs_ready, __ATOMIC_ACQUIRE))
        init_tabs();
    *out_size = 0;
    if (level < 0)
        level = 0;
    size_t sz = 0;
    int st;
    if (format == ORC_FMT_RAW)
        return compress_raw(level, in, in_len, out, out_cap, out_size);
    if (format == ORC_FMT_ZLIB) {
        /* compress_zlib, src/compress/mod.rs:2248-2298 */
        if (out_cap < 6)
            return ORC_INSUFFICIENT_SPACE;
        unsigned hint = level < 2 ? 0 : level < 6 ? 1 : level < 8 ? 2 : 3;
        unsigned hdr = (8u << 8) | (7u << 12) | (hint << 6);
        hdr |= 31 - (hdr % 31);
        out[0] = (uint8_t)(hdr >> 8);
        out[1] = (uint8_t)hdr;
        st = compress_raw(level, in, in_len, out + 2, out_cap - 6, &sz);
        if (st != ORC_OK)
            return st;
        uint32_t a = orc_adler32(1, in, in_len);
        uint8_t *f = out + 2 + sz;
        f[0] = (uint8_t)(a >> 24);
        f[1] = (uint8_t)(a >> 16);
        f[2] = (uint8_t)(a >> 8);
        f[3] = (uint8_t)a;
        *out_size = sz + 6;
        return ORC_OK;
    }
    /* compress_gzip, src/compress/mod.rs:2300-2357 */
    if (out_cap < 18)
        return ORC_INSUFFICIENT_SPACE;
    static const uint8_t ghdr[8] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0};
    memcpy(out, ghdr, 8);
    out[8] = level < 2 ? 4 : level >= 8 ? 2 : 0;
    out[9] = 255;
    st = compress_raw(level, in, in_len, out + 10, out_cap - 18, &sz);
    if (st != ORC_OK)
        return st;
    uint32_t crc = orc_crc32(0, in, in_len), isz = (uint32_t)in_len;
    uint8_t *f = out + 10 + sz;
    for (int k = 0; k < 4; k++) {
        f[k] = (uint8_t)(crc >> (8 * k));
        f[4 + k] = (uint8_t)(isz >> (8 * k));
    }
    *out_size = sz + 18;
    return ORC_OK;
}

#include "size.inc.c"
