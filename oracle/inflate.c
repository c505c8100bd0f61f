/*
 * inflate.c — oracle restatement of the reference decompressor (TEST ONLY).
 *
 * Follows the reference's portable path, which is what every build runs for
 * the tail of a stream and what non-BMI2 targets run throughout:
 *   Decompressor::decompress_streaming_ptr   src/decompress/mod.rs:213-353
 *   read_dynamic_huffman_header              src/decompress/mod.rs:403-507
 *   decompress_huffman_block_ptr             src/decompress/mod.rs:509-1072
 *   build_decode_table                       src/decompress/mod.rs:1324-1495
 *   decode result tables / entry flags       src/decompress/tables.rs
 *   zlib / gzip unwrap + footer check        src/decompress/mod.rs:1074-1240
 * The BMI2 driver (src/decompress/x86.rs:2193-2424) differs for stored blocks
 * (all-or-nothing copy, BadData instead of ShortInput); since the batch API
 * only distinguishes success from failure (src/batch.rs:93-97) the oracle
 * keeps the portable statuses.
 *
 * Bit reader: refill_bits! (src/decompress/mod.rs:16-35) is restated exactly
 * (same refill points, same byte accounting) because the reference reports
 * `in_consumed` as the READER position, prefetched bytes included, and the
 * zlib/gzip wrappers look for the footer at that position.
 *
 * REFERENCE DEFECT (documented, deliberately not mirrored): when a block's
 * end-of-block code is longer than litlen_tablebits (= min(11, longest code))
 * the reference finds it through a sub-table and then mishandles it — the
 * portable loop falls through and treats the entry as a length symbol
 * (src/decompress/mod.rs:948-976: no END_OF_BLOCK test after the sub-table
 * load), and the BMI2 loop consumes it without ending the block
 * (src/decompress/x86.rs:2304-2309: `break` without `eob_found = true`).
 * C libdeflate and RFC 1951 end the block there, and so does this oracle;
 * orc_inflate_last_ref_defect() reports whether the last stream decoded on
 * this thread would have hit that defect, so tests can say for which inputs
 * the reference itself is well defined.  A second defect is flagged the same
 * way: build_decode_table leaves its short-codeword loop as soon as
 * `len > table_bits` even when no codeword has that length
 * (src/decompress/mod.rs:1430-1442; C libdeflate keeps advancing while the
 * count is zero), so a code with codewords longer than table_bits but none of
 * length table_bits+1 gets a corrupt sub-table (and `len_counts[len] -= 1`
 * underflows at :1487).  The oracle skips the empty lengths, as the format
 * requires.  Two smaller points are also resolved
 * towards the format: an end-of-block code that needs more bits than remain
 * is ShortInput (the reference subtracts without a check, :943-946), and a
 * match whose offset reaches before the start of the output is BadData on
 * every path (the reference's pointer fast loop `break`s with the length bits
 * already consumed, :711-713).
 */
#include "oracle.h"
#include <string.h>

#define PRECODE_TABLEBITS 7
#define LITLEN_TABLEBITS 11
#define OFFSET_TABLEBITS 8
#define PRECODE_ENOUGH 128
#define LITLEN_ENOUGH 2342
#define OFFSET_ENOUGH 402
#define MAX_CODEWORD_LEN 15

#define F_LITERAL 0x80000000u
#define F_EXCEPTIONAL 0x00008000u
#define F_SUBTABLE 0x00004000u
#define F_EOB 0x00002000u

static const uint16_t k_len_base[29] = {
    3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
    31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t k_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
                                        1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                        4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t k_off_base[30] = {
    1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
    33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
    1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t k_off_extra[30] = {0, 0, 0,  0,  1,  1,  2,  2,  3,  3,
                                        4, 4, 5,  5,  6,  6,  7,  7,  8,  8,
                                        9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* Per-symbol decode results, src/decompress/tables.rs:10-30,55-379.  Litlen
 * symbols 286/287 decode as length 258 and offset symbols 30/31 as base 24577
 * with 13 extra bits (tables.rs:342-343,377-378) — kept. */
static uint32_t litlen_result(unsigned sym)
{
    if (sym < 256)
        return F_LITERAL | (sym << 16);
    if (sym == 256)
        return F_EXCEPTIONAL | F_EOB;
    unsigned s = sym - 257;
    if (s > 28)
        s = 28;
    return ((uint32_t)k_len_base[s] << 16) | k_len_extra[s];
}
static uint32_t offset_result(unsigned sym)
{
    if (sym > 29)
        sym = 29;
    return ((uint32_t)k_off_base[sym] << 16) | k_off_extra[sym];
}
static uint32_t precode_result(unsigned sym) { return sym << 16; }

typedef uint32_t (*result_fn)(unsigned);

typedef struct {
    uint32_t precode_tab[PRECODE_ENOUGH];
    uint32_t litlen_tab[LITLEN_ENOUGH];
    uint32_t offset_tab[OFFSET_ENOUGH];
    uint8_t precode_lens[19];
    uint8_t lens[288 + 32 + 137];
    uint16_t sorted[288];
    unsigned litlen_tablebits;
    int static_loaded;
    /* bit reader */
    const uint8_t *in;
    size_t in_len, in_idx;
    uint64_t bitbuf;
    uint32_t bitsleft;
    int ref_defect;
} inflater;

static __thread int tls_last_ref_defect;
int orc_inflate_last_ref_defect(void) { return tls_last_ref_defect; }

/* refill_bits!, src/decompress/mod.rs:16-35 */
static inline void refill(inflater *d)
{
    if (d->bitsleft >= 32)
        return;
    if (d->in_len - d->in_idx >= 8) {
        uint64_t w;
        memcpy(&w, d->in + d->in_idx, 8);
        d->bitbuf |= w << d->bitsleft;
        d->in_idx += (63 - d->bitsleft) >> 3;
        d->bitsleft |= 56;
    } else {
        while (d->bitsleft < 32 && d->in_idx < d->in_len) {
            d->bitbuf |= (uint64_t)d->in[d->in_idx++] << d->bitsleft;
            d->bitsleft += 8;
        }
    }
}

static inline unsigned bsr32(uint32_t v) { return 31u - (unsigned)__builtin_clz(v); }

/* build_decode_table, src/decompress/mod.rs:1324-1495.  Returns 0 when the
 * code is over-subscribed or incomplete in a way the reference rejects. */
static int build_table(uint32_t *tab, const uint8_t *lens, unsigned num_syms,
                       result_fn res, unsigned table_bits, unsigned max_len,
                       uint16_t *sorted, unsigned *table_bits_ret,
                       int *ref_defect)
{
    uint32_t cnt[MAX_CODEWORD_LEN + 1] = {0};
    uint32_t offs[MAX_CODEWORD_LEN + 1];
    for (unsigned s = 0; s < num_syms; s++)
        cnt[lens[s]]++;
    unsigned top = max_len;
    while (top > 1 && cnt[top] == 0)
        top--;
    if (table_bits_ret) {
        if (table_bits > top)
            table_bits = top;
        *table_bits_ret = table_bits;
    }
    offs[0] = 0;
    offs[1] = cnt[0];
    uint32_t used = 0;
    for (unsigned l = 1; l < top; l++) {
        offs[l + 1] = offs[l] + cnt[l];
        used = (used << 1) + cnt[l];
    }
    used = (used << 1) + cnt[top];
    for (unsigned s = 0; s < num_syms; s++)
        sorted[offs[lens[s]]++] = (uint16_t)s;
    /* offs[0] now indexes the first symbol with a non-zero length */
    if (used > (1u << top))
        return 0;
    if (used < (1u << top)) {
        unsigned sym;
        if (used == 0) {
            sym = 0;
        } else {
            if (used != (1u << (top - 1)) || cnt[1] != 1)
                return 0;
            sym = sorted[offs[0]];
        }
        uint32_t e = res(sym) + (1u << 8) + 1u;
        for (unsigned i = 0; i < (1u << table_bits); i++)
            tab[i] = e;
        return 1;
    }

    /* Complete code.  Walk symbols in canonical (length, symbol) order; the
     * codeword is kept bit-reversed so it indexes the table directly. */
    const uint16_t *sp = sorted + offs[0];
    uint32_t codeword = 0;
    unsigned len = 1;
    while (cnt[len] == 0)
        len++;
    unsigned cur_end = 1u << len;
    while (len <= table_bits) {
        for (uint32_t c = cnt[len]; c > 0; c--) {
            tab[codeword] = res(*sp++) + (len << 8) + len;
            if (codeword == cur_end - 1) {
                /* last codeword: replicate up to the full table and stop */
                for (; len < table_bits; len++) {
                    memcpy(tab + cur_end, tab, cur_end * sizeof(uint32_t));
                    cur_end <<= 1;
                }
                return 1;
            }
            uint32_t bit = 1u << bsr32(codeword ^ (cur_end - 1));
            codeword = (codeword & (bit - 1)) | bit;
        }
        do {
            len++;
            if (len <= table_bits) {
                memcpy(tab + cur_end, tab, cur_end * sizeof(uint32_t));
                cur_end <<= 1;
            }
        } while (len <= table_bits && cnt[len] == 0);
        while (len > table_bits && cnt[len] == 0) {
            /* the reference stops at table_bits+1 regardless (defect #2) */
            if (ref_defect)
                *ref_defect = 1;
            len++;
        }
    }
    /* Codewords longer than table_bits go to sub-tables. */
    cur_end = 1u << table_bits;
    uint32_t prefix = 0xFFFFFFFFu;
    unsigned sub_start = 0;
    for (;;) {
        if ((codeword & ((1u << table_bits) - 1)) != prefix) {
            prefix = codeword & ((1u << table_bits) - 1);
            sub_start = cur_end;
            unsigned sub_bits = len - table_bits;
            uint32_t sub_used = cnt[len];
            while (sub_used < (1u << sub_bits)) {
                sub_bits++;
                sub_used = (sub_used << 1) +
                           (table_bits + sub_bits <= MAX_CODEWORD_LEN
                                ? cnt[table_bits + sub_bits]
                                : 0);
            }
            cur_end = sub_start + (1u << sub_bits);
            tab[prefix] = ((uint32_t)sub_start << 16) | F_EXCEPTIONAL |
                          F_SUBTABLE | (sub_bits << 8) | table_bits;
        }
        uint32_t e = res(*sp++) + ((len - table_bits) << 8) + (len - table_bits);
        unsigned stride = 1u << (len - table_bits);
        for (unsigned i = sub_start + (codeword >> table_bits); i < cur_end;
             i += stride)
            tab[i] = e;
        if (codeword == (1u << len) - 1)
            return 1;
        uint32_t bit = 1u << bsr32(codeword ^ ((1u << len) - 1));
        codeword = (codeword & (bit - 1)) | bit;
        cnt[len]--;
        while (cnt[len] == 0) {
            len++;
            if (len > MAX_CODEWORD_LEN)
                return 1;
        }
    }
}

/* load_static_huffman_codes, src/decompress/mod.rs:355-401 */
static void load_static(inflater *d)
{
    if (d->static_loaded)
        return;
    unsigned i = 0;
    for (; i < 144; i++) d->lens[i] = 8;
    for (; i < 256; i++) d->lens[i] = 9;
    for (; i < 280; i++) d->lens[i] = 7;
    for (; i < 288; i++) d->lens[i] = 8;
    for (; i < 320; i++) d->lens[i] = 5;
    build_table(d->offset_tab, d->lens + 288, 32, offset_result,
                OFFSET_TABLEBITS, MAX_CODEWORD_LEN, d->sorted, 0, 0);
    build_table(d->litlen_tab, d->lens, 288, litlen_result, LITLEN_TABLEBITS,
                MAX_CODEWORD_LEN, d->sorted, &d->litlen_tablebits, 0);
    d->static_loaded = 1;
}

/* read_dynamic_huffman_header, src/decompress/mod.rs:403-507 */
static int read_dynamic_header(inflater *d)
{
    static const uint8_t perm[19] = {16, 17, 18, 0, 8,  7, 9,  6, 10, 5,
                                     11, 4,  12, 3, 13, 2, 14, 1, 15};
    refill(d);
    if (d->bitsleft < 14)
        return ORC_SHORT_INPUT;
    unsigned nlit = 257 + (unsigned)(d->bitbuf & 0x1F);
    unsigned noff = 1 + (unsigned)((d->bitbuf >> 5) & 0x1F);
    unsigned npre = 4 + (unsigned)((d->bitbuf >> 10) & 0xF);
    d->bitbuf >>= 14;
    d->bitsleft -= 14;
    for (unsigned i = 0; i < npre; i++) {
        refill(d);
        if (d->bitsleft < 3)
            return ORC_SHORT_INPUT;
        d->precode_lens[perm[i]] = (uint8_t)(d->bitbuf & 7);
        d->bitbuf >>= 3;
        d->bitsleft -= 3;
    }
    for (unsigned i = npre; i < 19; i++)
        d->precode_lens[perm[i]] = 0;
    if (!build_table(d->precode_tab, d->precode_lens, 19, precode_result,
                     PRECODE_TABLEBITS, 7, d->sorted, 0, 0))
        return ORC_BAD_DATA;
    unsigned total = nlit + noff, i = 0;
    while (i < total) {
        refill(d);
        uint32_t e = d->precode_tab[d->bitbuf & ((1u << PRECODE_TABLEBITS) - 1)];
        uint32_t nb = e & 0xFF;
        if (d->bitsleft < nb)
            return ORC_SHORT_INPUT;
        d->bitbuf >>= nb;
        d->bitsleft -= nb;
        unsigned presym = e >> 16;
        unsigned rep, need;
        uint8_t val;
        if (presym < 16) {
            d->lens[i++] = (uint8_t)presym;
            continue;
        } else if (presym == 16) {
            if (i == 0)
                return ORC_BAD_DATA;
            val = d->lens[i - 1];
            need = 2;
            rep = 3;
        } else if (presym == 17) {
            val = 0;
            need = 3;
            rep = 3;
        } else {
            val = 0;
            need = 7;
            rep = 11;
        }
        if (d->bitsleft < need)
            return ORC_SHORT_INPUT;
        rep += (unsigned)(d->bitbuf & ((1u << need) - 1));
        d->bitbuf >>= need;
        d->bitsleft -= need;
        /* overruns are clamped, not rejected (:462-467,475-480,488-493) */
        while (rep-- && i < total)
            d->lens[i++] = val;
    }
    if (!build_table(d->offset_tab, d->lens + nlit, noff, offset_result,
                     OFFSET_TABLEBITS, MAX_CODEWORD_LEN, d->sorted, 0,
                     &d->ref_defect))
        return ORC_BAD_DATA;
    if (!build_table(d->litlen_tab, d->lens, nlit, litlen_result,
                     LITLEN_TABLEBITS, MAX_CODEWORD_LEN, d->sorted,
                     &d->litlen_tablebits, &d->ref_defect))
        return ORC_BAD_DATA;
    d->static_loaded = 0;
    if (d->lens[256] > d->litlen_tablebits)
        d->ref_defect = 1;
    return ORC_OK;
}

/* decompress_huffman_block_ptr, src/decompress/mod.rs:509-1072.  The three
 * loops of the reference share refill points (top of symbol, before the
 * offset code) and differ only in bounds-check style, so one loop restates
 * them. */
static int huffman_block(inflater *d, uint8_t *out, size_t out_len,
                         size_t *out_idx)
{
    const uint32_t lmask = (1u << d->litlen_tablebits) - 1;
    size_t op = *out_idx;
    for (;;) {
        refill(d);
        uint32_t e = d->litlen_tab[d->bitbuf & lmask];
        if (e & F_EXCEPTIONAL) {
            if (!(e & F_EOB)) { /* sub-table pointer */
                uint32_t mb = e & 0xFF;
                if (d->bitsleft < mb)
                    return ORC_SHORT_INPUT;
                d->bitbuf >>= mb;
                d->bitsleft -= mb;
                e = d->litlen_tab[(e >> 16) +
                                  (d->bitbuf & ((1u << ((e >> 8) & 0x3F)) - 1))];
            }
            if (e & F_EOB) {
                uint32_t nb = e & 0xFF;
                if (d->bitsleft < nb)
                    return ORC_SHORT_INPUT;
                d->bitbuf >>= nb;
                d->bitsleft -= nb;
                *out_idx = op;
                return ORC_OK;
            }
        }
        uint64_t saved = d->bitbuf;
        uint32_t tb = e & 0xFF;
        if (d->bitsleft < tb)
            return ORC_SHORT_INPUT;
        d->bitbuf >>= tb;
        d->bitsleft -= tb;
        if (e & F_LITERAL) {
            if (op >= out_len)
                return ORC_INSUFFICIENT_SPACE;
            out[op++] = (uint8_t)(e >> 16);
            continue;
        }
        size_t length = e >> 16;
        uint32_t cl = (e >> 8) & 0xFF;
        if (tb > cl)
            length += (size_t)((saved >> cl) & ((1u << (tb - cl)) - 1));

        refill(d);
        e = d->offset_tab[d->bitbuf & ((1u << OFFSET_TABLEBITS) - 1)];
        if (e & F_SUBTABLE) {
            uint32_t mb = e & 0xFF;
            if (d->bitsleft < mb)
                return ORC_SHORT_INPUT;
            d->bitbuf >>= mb;
            d->bitsleft -= mb;
            e = d->offset_tab[(e >> 16) +
                              (d->bitbuf & ((1u << ((e >> 8) & 0x3F)) - 1))];
        }
        saved = d->bitbuf;
        tb = e & 0xFF;
        if (d->bitsleft < tb)
            return ORC_SHORT_INPUT;
        d->bitbuf >>= tb;
        d->bitsleft -= tb;
        size_t offset = e >> 16;
        cl = (e >> 8) & 0xFF;
        if (tb > cl)
            offset += (size_t)((saved >> cl) & ((1u << (tb - cl)) - 1));
        if (offset > op)
            return ORC_BAD_DATA;
        if (op + length > out_len)
            return ORC_INSUFFICIENT_SPACE;
        /* LZ77 copy out[i] = out[i - offset]; every specialisation in the
         * reference (prepare_pattern :1259-1317, copy_match_bmi2) is this. */
        const uint8_t *src = out + op - offset;
        uint8_t *dst = out + op;
        if (offset >= length) {
            memcpy(dst, src, length);
        } else {
            for (size_t i = 0; i < length; i++)
                dst[i] = src[i];
        }
        op += length;
    }
}

/* decompress_streaming_ptr, src/decompress/mod.rs:213-353 (one-shot use:
 * state machine collapsed to a loop over blocks). */
static int inflate_raw(inflater *d, const uint8_t *in, size_t in_len,
                       uint8_t *out, size_t out_len, size_t *in_consumed,
                       size_t *out_size)
{
    d->in = in;
    d->in_len = in_len;
    d->in_idx = 0;
    d->bitbuf = 0;
    d->bitsleft = 0;
    d->ref_defect = 0;
    size_t op = 0;
    int st;
    for (;;) {
        refill(d);
        if (d->bitsleft < 3) {
            st = ORC_SHORT_INPUT;
            break;
        }
        int final = (int)(d->bitbuf & 1);
        unsigned type = (unsigned)((d->bitbuf >> 1) & 3);
        d->bitbuf >>= 3;
        d->bitsleft -= 3;
        if (type == 0) {
            /* UncompressedHeader/Body, :282-346 */
            d->bitsleft -= d->bitsleft & 7;
            size_t unused = d->bitsleft / 8;
            d->in_idx = d->in_idx > unused ? d->in_idx - unused : 0;
            d->bitbuf = 0;
            d->bitsleft = 0;
            if (d->in_idx + 4 > in_len) {
                st = ORC_SHORT_INPUT;
                break;
            }
            unsigned len = in[d->in_idx] | (unsigned)in[d->in_idx + 1] << 8;
            unsigned nlen = in[d->in_idx + 2] | (unsigned)in[d->in_idx + 3] << 8;
            d->in_idx += 4;
            if (len != (~nlen & 0xFFFF)) {
                st = ORC_BAD_DATA;
                break;
            }
            size_t avail_in = in_len - d->in_idx, avail_out = out_len - op;
            size_t n = len;
            if (n > avail_in) n = avail_in;
            if (n > avail_out) n = avail_out;
            memcpy(out + op, in + d->in_idx, n);
            d->in_idx += n;
            op += n;
            if (n != len) {
                st = avail_out < len && avail_out <= avail_in
                         ? ORC_INSUFFICIENT_SPACE
                         : ORC_SHORT_INPUT;
                break;
            }
        } else if (type == 3) {
            st = ORC_BAD_DATA;
            break;
        } else {
            if (type == 1) {
                load_static(d);
            } else {
                st = read_dynamic_header(d);
                if (st != ORC_OK)
                    break;
            }
            st = huffman_block(d, out, out_len, &op);
            if (st != ORC_OK)
                break;
        }
        if (final) {
            st = ORC_OK;
            break;
        }
    }
    *in_consumed = d->in_idx;
    *out_size = op;
    tls_last_ref_defect = d->ref_defect;
    return st;
}

static uint32_t be32(const uint8_t *p)
{
    return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
}
static uint32_t le32(const uint8_t *p)
{
    return (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0];
}

int orc_decompress(int format, const uint8_t *in, size_t in_len, uint8_t *out,
                   size_t out_cap, size_t *in_consumed, size_t *out_size)
{
    static __thread inflater d; /* one codec state per worker thread */
    size_t ic = 0, os = 0;
    int st;
    d.static_loaded = 0;
    if (format == ORC_FMT_RAW) {
        st = inflate_raw(&d, in, in_len, out, out_cap, &ic, &os);
    } else if (format == ORC_FMT_ZLIB) {
        /* decompress_zlib_uninit, src/decompress/mod.rs:1074-1127 */
        if (in_len < 6) {
            st = ORC_SHORT_INPUT;
        } else {
            unsigned hdr = (unsigned)in[0] << 8 | in[1];
            if (hdr % 31 != 0 || ((hdr >> 8) & 0xF) != 8 ||
                ((hdr >> 12) & 0xF) > 7 || ((hdr >> 5) & 1)) {
                st = ORC_BAD_DATA;
            } else {
                st = inflate_raw(&d, in + 2, in_len - 6, out, out_cap, &ic, &os);
                ic += 2;
                if (st == ORC_OK) {
                    if (orc_adler32(1, out, os) != be32(in + ic))
                        st = ORC_BAD_DATA;
                    ic += 4;
                }
            }
        }
    } else {
        /* decompress_gzip_uninit, src/decompress/mod.rs:1144-1240 */
        if (in_len < 18) {
            st = ORC_SHORT_INPUT;
        } else if (in[0] != 0x1F || in[1] != 0x8B || in[2] != 8 || (in[3] & 0xE0)) {
            st = ORC_BAD_DATA;
        } else {
            unsigned flg = in[3];
            size_t p = 10;
            st = ORC_OK;
            if (flg & 0x04) {
                if (p + 2 > in_len)
                    st = ORC_SHORT_INPUT;
                else
                    p += 2 + (in[p] | (size_t)in[p + 1] << 8);
            }
            if (st == ORC_OK && (flg & 0x08)) {
                while (p < in_len && in[p]) p++;
                p++;
            }
            if (st == ORC_OK && (flg & 0x10)) {
                while (p < in_len && in[p]) p++;
                p++;
            }
            if (st == ORC_OK && (flg & 0x02))
                p += 2;
            if (st == ORC_OK && p + 8 > in_len)
                st = ORC_SHORT_INPUT;
            if (st == ORC_OK) {
                st = inflate_raw(&d, in + p, in_len - 8 - p, out, out_cap, &ic, &os);
                ic += p;
                if (st == ORC_OK) {
                    if (orc_crc32(0, out, os) != le32(in + ic) ||
                        (uint32_t)os != le32(in + ic + 4))
                        st = ORC_BAD_DATA;
                    ic += 8;
                }
            }
        }
    }
    if (in_consumed) *in_consumed = ic;
    if (out_size) *out_size = os;
    return st;
}
