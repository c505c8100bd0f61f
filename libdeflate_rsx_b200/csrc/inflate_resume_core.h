// inflate_resume_core.h — one resumable DEFLATE decoder step, scalar code for ONE decoder state.
//
// Replaces Decompressor::decompress_streaming (reference src/decompress/mod.rs:204-372): decode as
// much of `in` into the window as fits, remember where the stream stands, continue with the next
// call.  The reference keeps its bit buffer and decode tables inside the Decompressor between
// calls; here the state is a small plain struct (bdf_inflate_state, include/bdeflate.h) so that
// thousands of decoder states can be advanced by one kernel launch, one LANE per state
// (inflate_resume.cuh): position inside the stream (bits of the first input byte already used),
// the kind of block it stands in, and for a Huffman block the code lengths — the tables are rebuilt
// from them when a call resumes inside a block.
//
// The file is plain C++ without CUDA headers and compiles for the device (through
// inflate_resume.cuh) and for the host (tests/host_harness/resume_host.py compiles exactly this
// file with g++), so the CPU tests run the code the kernel runs.
//
// Stop rules.  A step stops only BETWEEN symbols:
//   * BDF_INSUFFICIENT_SPACE — fewer than 258 bytes of room left in the window (a match may need
//     them): the caller drains / shifts the window and calls again;
//   * BDF_SHORT_INPUT — not enough input for the next unit: 3 bits in front of a block, 48 bits in
//     front of a symbol (15 + 5 + 15 + 13), the whole header in front of a dynamic block (at most
//     570 bytes).  With in_final set the decoder reads zero bits past the end instead and reports
//     BDF_SHORT_INPUT only if the stream really ends inside a unit (the truncated-stream error);
//   * BDF_BAD_DATA — sticky.
// BDF_OK means the final block has ended.
#pragma once
#include <stdint.h>
#include "../../include/bdeflate.h"

#ifndef BDF_HD
#ifdef __CUDACC__
#define BDF_HD __host__ __device__ __forceinline__
#define BDF_HDM __host__ __device__ __forceinline__
#else
#define BDF_HD static inline
#define BDF_HDM inline
#endif
#endif

namespace bdf_rs {

constexpr int LTB = 10, OTB = 8;             // direct-table bits (litlen / offset)
constexpr uint32_t HDR_MAX_BITS = 3 + 14 + 19 * 3 + 320 * 14;
constexpr uint32_t SYM_MAX_BITS = 48;
constexpr uint32_t ROOM_MIN = 258;

enum { PH_START = 0, PH_STORED = 1, PH_HUFF = 2, PH_DONE = 3, PH_FAILED = 4 };

// entry: [3:0] codeword bits (0: longer than the table), [15:4] symbol
struct Code {
    uint16_t first[16], count[16], offs[16];
};
struct Tables {                              // one per decoder state while a step runs
    uint16_t lit_tab[1 << LTB];
    uint16_t off_tab[1 << OTB];
    uint16_t lit_sorted[288];
    uint16_t off_sorted[32];
    Code lit, off;
};

BDF_HD uint32_t rev_bits(uint32_t v, uint32_t n)     // the low n bits of v, reversed
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// Accept / reject rules of build_decode_table (src/decompress/mod.rs:1365-1383): over-subscribed
// codes are rejected, incomplete ones too unless there is no codeword at all or a single 1-bit one.
BDF_HD bool build_table(const uint8_t *lens, uint32_t nsyms, uint16_t *tab, int tbits, uint16_t *sorted, Code &c)
{
    for (int l = 0; l < 16; l++) c.count[l] = 0;
    for (uint32_t s = 0; s < nsyms; s++) c.count[lens[s]]++;
    c.count[0] = 0;
    uint32_t used = 0, total = 0, code = 0, off = 0;
    for (int l = 1; l < 16; l++) {
        used += (uint32_t)c.count[l] << (15 - l);
        total += c.count[l];
        code = (code + (l > 1 ? c.count[l - 1] : 0)) << 1;
        c.first[l] = (uint16_t)code;
        c.offs[l] = (uint16_t)off;
        off += c.count[l];
    }
    c.first[0] = 0; c.offs[0] = 0;
    if (used > (1u << 15)) return false;
    if (used < (1u << 15)) {
        if (!(total == 0 || (total == 1 && c.count[1] == 1))) return false;
        uint32_t sym = 0;
        if (total == 1)
            for (uint32_t s = 0; s < nsyms; s++)
                if (lens[s] == 1) { sym = s; break; }
        for (uint32_t i = 0; i < (1u << tbits); i++) tab[i] = (uint16_t)(sym << 4 | 1u);
        for (int l = 0; l < 16; l++) c.count[l] = 0;          // no long codewords
        return true;
    }
    uint16_t next[16];
    for (int l = 0; l < 16; l++) next[l] = 0;
    for (uint32_t s = 0; s < nsyms; s++) {
        const uint32_t l = lens[s];
        if (l == 0) continue;
        const uint32_t rank = next[l]++;
        sorted[c.offs[l] + rank] = (uint16_t)s;
        const uint32_t rev = rev_bits(c.first[l] + rank, l);
        if (l <= (uint32_t)tbits) {
            for (uint32_t i = rev; i < (1u << tbits); i += 1u << l) tab[i] = (uint16_t)(s << 4 | l);
        } else {
            tab[rev & ((1u << tbits) - 1u)] = 0;
        }
    }
    return true;
}

// codeword longer than the direct table: canonical search on the next 15 bits; 0 = not a codeword
BDF_HD uint32_t decode_long(uint32_t bits15, int tbits, const uint16_t *sorted, const Code &c)
{
    const uint32_t x = rev_bits(bits15, 15);
    for (uint32_t l = (uint32_t)tbits + 1; l <= 15; l++) {
        const uint32_t d = (x >> (15 - l)) - c.first[l];
        if (d < c.count[l]) return (uint32_t)sorted[c.offs[l] + d] << 4 | l;
    }
    return 0;
}

BDF_HD void length_slot(uint32_t slot, uint32_t &base, uint32_t &extra)
{
    if (slot < 8) { base = 3 + slot; extra = 0; }
    else if (slot >= 28) { base = 258; extra = 0; }          // 286 / 287 decode as 258 (tables.rs:342-343)
    else { extra = (slot >> 2) - 1; base = ((4 + (slot & 3)) << extra) + 3; }
}
BDF_HD void offset_slot(uint32_t slot, uint32_t &base, uint32_t &extra)
{
    if (slot > 29) slot = 29;                                // 30 / 31 alias 29 (tables.rs:377-378)
    if (slot < 4) { base = slot + 1; extra = 0; }
    else { extra = (slot >> 1) - 1; base = ((2 + (slot & 1)) << extra) + 1; }
}

// LSB-first reader over in[0, len) that starts `bit_off` bits into in[0]; zero bits past the end
struct Reader {
    const uint8_t *in;
    uint64_t len, ip;
    uint64_t buf;
    uint32_t cnt;
    BDF_HDM void start(const uint8_t *p, uint64_t n, uint32_t bit_off)
    {
        in = p; len = n; ip = 0; buf = 0; cnt = 0;
        fill();
        buf >>= bit_off; cnt -= bit_off;
    }
    BDF_HDM void fill()
    {
        while (cnt <= 56) {
            const uint64_t b = ip < len ? in[ip] : 0u;
            buf |= b << cnt;
            cnt += 8;
            ip++;
        }
    }
    BDF_HDM uint32_t peek(uint32_t n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    BDF_HDM void drop(uint32_t n) { buf >>= n; cnt -= n; }
    BDF_HDM uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    BDF_HDM uint64_t used_bits() const { return ip * 8 - cnt; }            // bits consumed since in[0] bit 0
    BDF_HDM uint64_t left_bits() const { const uint64_t u = used_bits(); return len * 8 > u ? len * 8 - u : 0; }
    BDF_HDM bool overrun() const { return used_bits() > len * 8; }
    BDF_HDM void seek_byte(uint64_t at) { ip = at; buf = 0; cnt = 0; }
};

// code lengths of a dynamic block (read_dynamic_huffman_header, src/decompress/mod.rs:403-507) into
// S.lens; the reader stands behind the 3 block bits
BDF_HD int read_dynamic_lengths(Reader &r, bdf_inflate_state &S, Tables &T)
{
    r.fill();
    S.nlit = 257 + r.take(5);
    S.noff = 1 + r.take(5);
    const uint32_t npre = 4 + r.take(4);
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t pl[19];
    for (int i = 0; i < 19; i++) pl[i] = 0;
    for (uint32_t i = 0; i < npre; i++) {
        r.fill();
        pl[order[i]] = (uint8_t)r.take(3);
    }
    if (r.overrun()) return BDF_SHORT_INPUT;
    // the precode table borrows the offset table (7 bits); its canonical description borrows T.off
    if (!build_table(pl, 19, T.off_tab, 7, T.off_sorted, T.off)) return BDF_BAD_DATA;
    const uint32_t total = S.nlit + S.noff;
    uint32_t i = 0, prev = 0;
    while (i < total) {
        r.fill();
        const uint32_t e = T.off_tab[r.peek(7)];
        r.drop(e & 15u);
        const uint32_t sym = e >> 4;
        if (sym < 16) { S.lens[i++] = (uint8_t)sym; prev = sym; continue; }
        uint32_t rep, val;
        if (sym == 16) {
            if (i == 0) return BDF_BAD_DATA;
            rep = 3 + r.take(2); val = prev;
        } else if (sym == 17) { rep = 3 + r.take(3); val = 0; }
        else { rep = 11 + r.take(7); val = 0; }
        if (rep > total - i) rep = total - i;                // overruns are clamped (:462-493)
        for (uint32_t q = 0; q < rep; q++) S.lens[i + q] = (uint8_t)val;
        prev = val;
        i += rep;
    }
    if (r.overrun()) return BDF_SHORT_INPUT;
    return BDF_OK;
}

BDF_HD bool build_block_tables(const bdf_inflate_state &S, Tables &T)
{
    return build_table(S.lens + S.nlit, S.noff, T.off_tab, OTB, T.off_sorted, T.off) &&
           build_table(S.lens, S.nlit, T.lit_tab, LTB, T.lit_sorted, T.lit);
}

// One step of one decoder.  win[0, *pos) is the output so far (history for matches, at least the
// last 32 KiB of it), new bytes go to win[*pos, cap).  Returns the status; *consumed = whole input
// bytes used up (the bits used of the next byte are kept in S.bit_off).
BDF_HD int resume_step(bdf_inflate_state &S, const uint8_t *in, uint64_t in_len, bool in_final, uint8_t *win, uint64_t cap,
                       uint64_t *pos_io, uint64_t *consumed, Tables &T)
{
    uint64_t pos = *pos_io;
    *consumed = 0;
    if (S.phase == PH_DONE) return BDF_OK;
    if (S.phase == PH_FAILED || S.phase > PH_FAILED || S.bit_off > 7 || pos > cap) { S.phase = PH_FAILED; return BDF_BAD_DATA; }
    Reader r;
    r.start(in, in_len, S.bit_off);
    bool tables_ready = false;
    int st;
    // every exit commits the reader position (whole bytes + bits of the next byte)
#define RS_RETURN(code)                                                      \
    do {                                                                     \
        st = (code);                                                         \
        goto out;                                                            \
    } while (0)
    for (;;) {
        if (S.phase == PH_START) {
            r.fill();
            if (r.left_bits() < 3) RS_RETURN(BDF_SHORT_INPUT);
            const uint32_t hdr = r.peek(3);
            const uint32_t type = hdr >> 1;
            if (type == 3) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
            if (type == 0) {
                // stored block (src/decompress/mod.rs:282-346): LEN / NLEN follow at the next byte boundary
                const uint64_t at = (r.used_bits() + 3 + 7) >> 3;
                if (at + 4 > in_len) RS_RETURN(BDF_SHORT_INPUT);
                const uint32_t blen = in[at] | (uint32_t)in[at + 1] << 8, nlen = in[at + 2] | (uint32_t)in[at + 3] << 8;
                r.seek_byte(at + 4);
                if (blen != (~nlen & 0xFFFFu)) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
                S.final_block = hdr & 1u;
                S.stored_rem = blen;
                S.phase = PH_STORED;
                continue;
            }
            if (type == 1) {
                r.drop(3);
                S.final_block = hdr & 1u;
                for (uint32_t s = 0; s < 320; s++) S.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5);
                S.nlit = 288; S.noff = 32;
            } else {
                if (!in_final && r.left_bits() < HDR_MAX_BITS) RS_RETURN(BDF_SHORT_INPUT);
                Reader save = r;
                r.drop(3);
                const int hs = read_dynamic_lengths(r, S, T);
                if (hs == BDF_SHORT_INPUT) { r = save; RS_RETURN(BDF_SHORT_INPUT); }
                if (hs != BDF_OK) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
                S.final_block = hdr & 1u;
            }
            if (!build_block_tables(S, T)) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
            tables_ready = true;
            S.phase = PH_HUFF;
            continue;
        }
        if (S.phase == PH_STORED) {
            // the reader is byte aligned and empty here
            uint64_t n = S.stored_rem;
            const uint64_t at = r.used_bits() >> 3;
            if (n > in_len - at) n = in_len - at;
            if (n > cap - pos) n = cap - pos;
            for (uint64_t i = 0; i < n; i++) win[pos + i] = in[at + i];
            pos += n;
            r.seek_byte(at + n);
            S.stored_rem -= (uint32_t)n;
            if (S.stored_rem == 0) {
                S.phase = S.final_block ? PH_DONE : PH_START;
                if (S.phase == PH_DONE) RS_RETURN(BDF_OK);
                continue;
            }
            RS_RETURN(pos == cap ? BDF_INSUFFICIENT_SPACE : BDF_SHORT_INPUT);
        }
        // ---- PH_HUFF
        if (!tables_ready) {
            // the state is the caller's memory: nothing in it is trusted
            bool sane = S.nlit >= 257 && S.nlit <= 288 && S.noff >= 1 && S.noff <= 32;
            for (uint32_t s = 0; sane && s < S.nlit + S.noff; s++) sane = S.lens[s] <= 15;
            if (!sane || !build_block_tables(S, T)) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
            tables_ready = true;
        }
        for (;;) {
            if (cap - pos < ROOM_MIN) RS_RETURN(BDF_INSUFFICIENT_SPACE);
            r.fill();
            if (!in_final && r.left_bits() < SYM_MAX_BITS) RS_RETURN(BDF_SHORT_INPUT);
            const Reader save = r;
            uint32_t e = T.lit_tab[r.peek(LTB)];
            if ((e & 15u) == 0) e = decode_long(r.peek(15), LTB, T.lit_sorted, T.lit);
            if (e == 0) {                                    // no codeword in the next 15 bits — or no 15 bits
                if (r.left_bits() < 15) RS_RETURN(BDF_SHORT_INPUT);
                S.phase = PH_FAILED;
                RS_RETURN(BDF_BAD_DATA);
            }
            r.drop(e & 15u);
            const uint32_t sym = e >> 4;
            if (sym < 256) {
                if (r.overrun()) { r = save; RS_RETURN(BDF_SHORT_INPUT); }
                win[pos++] = (uint8_t)sym;
                continue;
            }
            if (sym == 256) {
                if (r.overrun()) { r = save; RS_RETURN(BDF_SHORT_INPUT); }
                S.phase = S.final_block ? PH_DONE : PH_START;
                break;
            }
            uint32_t lb, le, ob, oe;
            length_slot(sym - 257, lb, le);
            const uint32_t length = lb + r.take(le);
            r.fill();
            uint32_t f = T.off_tab[r.peek(OTB)];
            if ((f & 15u) == 0) f = decode_long(r.peek(15), OTB, T.off_sorted, T.off);
            if (f == 0) {
                if (r.left_bits() < 15) { r = save; RS_RETURN(BDF_SHORT_INPUT); }
                S.phase = PH_FAILED;
                RS_RETURN(BDF_BAD_DATA);
            }
            r.drop(f & 15u);
            offset_slot(f >> 4, ob, oe);
            const uint32_t offset = ob + r.take(oe);
            if (r.overrun()) { r = save; RS_RETURN(BDF_SHORT_INPUT); }
            if (offset > pos) { S.phase = PH_FAILED; RS_RETURN(BDF_BAD_DATA); }
            for (uint32_t i = 0; i < length; i++) win[pos + i] = win[pos - offset + i];
            pos += length;
        }
        if (S.phase == PH_DONE) RS_RETURN(BDF_OK);
    }
out:
#undef RS_RETURN
    {
        const uint64_t used = r.used_bits();
        *consumed = used >> 3;
        S.bit_off = (uint32_t)(used & 7u);
        S.total_in += used >> 3;
        S.total_out += pos - *pos_io;
        *pos_io = pos;
    }
    return st;
}

}  // namespace bdf_rs
