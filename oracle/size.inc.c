/*
 * size.inc.c — oracle restatement of Compressor::compress_to_size
 * (src/compress/mod.rs:792-1259), included by deflate.c.  TEST ONLY.
 *
 * The estimator runs the same matchfinders as the compressor but keeps only
 * histograms: size = ceil(sum over blocks of (3 + header bits + data bits) / 8).
 * Differences from the emitting path that the restatement keeps:
 *  - level 1 always runs the block-split statistics (the emitting path skips
 *    them for inputs of at most 64 KiB, :1505) -> more, shorter static blocks;
 *  - levels 10..12 classify pass-1 offsets with OFF_IDX_TABLE (observe_match,
 *    :107-112,311-330) instead of the slot table, and seed the DP costs from a
 *    second greedy parse of the block slice with reset tables (:902-934);
 *  - an empty input costs 0 bytes (the loop never runs), final_block only
 *    matters at level 0 (:1073-1082).
 */

/* calculate_block_data_size, src/compress/mod.rs:1142-1173 */
static size_t block_data_bits(const compressor *c)
{
    size_t bits = 0;
    for (unsigned s = 0; s < NUM_LITLEN; s++)
        bits += (size_t)c->litlen_freqs[s] * c->litlen_lens[s];
    for (unsigned s = 0; s < NUM_OFFSET; s++)
        bits += (size_t)c->offset_freqs[s] * c->offset_lens[s];
    for (unsigned s = 0; s < 29; s++)
        bits += (size_t)c->litlen_freqs[257 + s] * len_extra[s];
    for (unsigned s = 0; s < 30; s++)
        bits += (size_t)c->offset_freqs[s] * off_extra[s];
    return bits;
}

/* calculate_dynamic_header_size, src/compress/mod.rs:1175-1259 */
static size_t dynamic_header_bits(const compressor *c)
{
    static const uint8_t perm[19] = {16, 17, 18, 0, 8,  7, 9,  6, 10, 5,
                                     11, 4,  12, 3, 13, 2, 14, 1, 15};
    size_t bits = 5 + 5 + 4;
    unsigned nlit = NUM_LITLEN, noff = NUM_OFFSET;
    while (nlit > 257 && c->litlen_lens[nlit - 1] == 0)
        nlit--;
    while (noff > 1 && c->offset_lens[noff - 1] == 0)
        noff--;
    uint8_t lens[NUM_LITLEN + NUM_OFFSET];
    const unsigned total = nlit + noff;
    memcpy(lens, c->litlen_lens, nlit);
    memcpy(lens + nlit, c->offset_lens, noff);
    uint32_t pre_freq[19] = {0};
    for (unsigned i = 0; i < total;) {
        const unsigned len = lens[i];
        unsigned span = 1;
        while (i + span < total && lens[i + span] == len)
            span++;
        unsigned run = span;
        if (len == 0) {
            while (run >= 11) {
                pre_freq[18]++;
                run -= run < 138 ? run : 138;
            }
            if (run >= 3) {
                pre_freq[17]++;
                run -= run < 10 ? run : 10;
            }
        } else if (run >= 4) {
            pre_freq[len]++;
            run--;
            while (run >= 3) {
                pre_freq[16]++;
                run -= run < 6 ? run : 6;
            }
        }
        pre_freq[len] += run;
        i += span;
    }
    uint8_t pre_lens[19] = {0};
    uint32_t pre_codes[NUM_LITLEN];
    make_huffman_code(19, MAX_PRE_CODEWORD_LEN, pre_freq, pre_lens, pre_codes);
    unsigned npre = 19;
    while (npre > 4 && pre_lens[perm[npre - 1]] == 0)
        npre--;
    bits += npre * 3;
    for (unsigned s = 0; s < 19; s++)
        if (pre_freq[s])
            bits += (size_t)pre_freq[s] * (pre_lens[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0));
    return bits;
}

/* accumulate_greedy_frequencies, src/compress/mod.rs:1096-1140 */
static size_t accumulate_greedy_frequencies(compressor *c, const uint8_t *in, size_t n, size_t start)
{
    st_reset(&c->st);
    memset(c->litlen_freqs, 0, sizeof(c->litlen_freqs));
    memset(c->offset_freqs, 0, sizeof(c->offset_freqs));
    size_t p = start;
    while (p < n) {
        if (st_should_end(&c->st, p - start, n - p))
            break;
        unsigned off = 0, len = mf_find(c, in, n, p, &off);
        if (len >= 3) {
            unsigned slot = offset_slot(off);
            st_match(&c->st, len, slot);
            c->litlen_freqs[257 + length_slot(len)]++;
            c->offset_freqs[slot]++;
            mf_skip(c, in, n, p + 1, len - 1);
            p += len;
        } else {
            record_literal(c, in[p]);
            p++;
        }
    }
    c->litlen_freqs[256]++;
    return p - start;
}

/* observe_match + OFF_IDX_TABLE, src/compress/mod.rs:107-112,311-330 */
static inline void st_match_by_offset(split_stats *s, unsigned len, unsigned off)
{
    const unsigned lg = bsr32(off);
    s->new_obs[8 + (len >= 8)]++;
    s->new_obs[10 + (lg < 8 ? 0 : lg < 12 ? 1 : lg < 15 ? 2 : 3)]++;
    s->num_new += 2;
}

/* calculate_block_size_near_optimal, src/compress/mod.rs:873-1071 */
static size_t size_block_near_optimal(compressor *c, const uint8_t *in, size_t n, size_t start,
                                      size_t *bits_out)
{
    st_reset(&c->st);
    size_t p = start;
    while (p < n) {
        if (st_should_end(&c->st, p - start, n - p))
            break;
        unsigned off = 0, len = mf_find(c, in, n, p, &off);
        if (len >= 3) {
            st_match_by_offset(&c->st, len, off);
            mf_skip(c, in, n, p + 1, len - 1); /* skip_match per position == skip_positions for BT */
            p += len;
        } else {
            st_literal(&c->st, in[p]);
            p++;
        }
    }
    const size_t done = p - start;
    const uint8_t *blk = in + start;

    /* greedy parse of the block slice with reset tables -> first costs (:902-934) */
    memset(c->litlen_freqs, 0, sizeof(c->litlen_freqs));
    memset(c->offset_freqs, 0, sizeof(c->offset_freqs));
    bt_reset(c->bt);
    for (size_t q = 0; q < done;) {
        unsigned off = 0, len = bt_find_match(c->bt, blk, done, q, c->max_depth, c->nice_len, &off);
        if (len >= 3) {
            c->litlen_freqs[257 + length_slot(len)]++;
            c->offset_freqs[offset_slot(off)]++;
            bt_skip_positions(c->bt, blk, done, q + 1, len - 1, c->max_depth, c->nice_len);
            q += len;
        } else {
            c->litlen_freqs[blk[q]]++;
            q++;
        }
    }
    c->litlen_freqs[256]++;
    make_block_codes(c);

    uint32_t length_cost[MAX_MATCH + 1], slot_cost[30];
    for (unsigned len = 3; len <= MAX_MATCH; len++) {
        unsigned s = length_slot(len);
        length_cost[len] = c->litlen_lens[257 + s] + len_extra[s];
    }
    for (unsigned s = 0; s < 30; s++)
        slot_cost[s] = c->offset_lens[s] + off_extra[s];
    if (c->dp_cap < done + 1) {
        c->dp_cap = done + 1;
        c->dp_cost = (uint32_t *)realloc(c->dp_cost, c->dp_cap * sizeof(uint32_t));
        c->dp_path = (uint32_t *)realloc(c->dp_path, c->dp_cap * sizeof(uint32_t));
    }
    uint32_t *cost = c->dp_cost, *path = c->dp_path;
    for (size_t i = 0; i <= done; i++)
        cost[i] = 0x3FFFFFFF;
    cost[0] = 0;

    bt_reset(c->bt);
    uint16_t list[260][2];
    size_t q = 0;
    while (q < done) {
        uint32_t here = cost[q];
        if (here >= 0x3FFFFFFF) {
            q++;
            continue;
        }
        uint32_t lit = c->litlen_lens[blk[q]];
        if (here + lit < cost[q + 1]) {
            cost[q + 1] = here + lit;
            path[q + 1] = 1;
        }
        unsigned nm = bt_find_matches(c->bt, blk, done, q, c->max_depth, c->nice_len, list);
        unsigned best = 0;
        for (unsigned k = 0; k < nm; k++) {
            unsigned len = list[k][0], off = list[k][1];
            if (len > best)
                best = len;
            uint32_t mc = length_cost[len] + slot_cost[offset_slot(off)];
            if (here + mc < cost[q + len]) {
                cost[q + len] = here + mc;
                path[q + len] = len | ((uint32_t)off << 16);
            }
        }
        if (best >= c->nice_len) {
            bt_skip_positions(c->bt, blk, done, q + 1, best - 1, c->max_depth, c->nice_len);
            q += best;
        } else {
            q++;
        }
    }

    memset(c->litlen_freqs, 0, sizeof(c->litlen_freqs));
    memset(c->offset_freqs, 0, sizeof(c->offset_freqs));
    c->litlen_freqs[256] = 1;
    for (size_t r = done; r > 0;) {
        const unsigned len = path[r] & 0xFFFF, off = path[r] >> 16;
        r -= len;
        if (len == 1) {
            c->litlen_freqs[blk[r]]++;
        } else {
            c->litlen_freqs[257 + length_slot(len)]++;
            c->offset_freqs[offset_slot(off)]++;
        }
    }
    make_block_codes(c);
    *bits_out = 3 + dynamic_header_bits(c) + block_data_bits(c);
    return done;
}

/* compress_to_size + compress_to_size_loop, src/compress/mod.rs:792-871,1073-1094 */
size_t orc_compress_to_size(int level, const uint8_t *in, size_t n, int final_block)
{
    if (!__atomic_load_n(&tabs_ready, __ATOMIC_ACQUIRE))
        init_tabs();
    if (level < 0)
        level = 0;
    if (level == 0) {
        size_t blocks = n / 65535 + ((n % 65535 != 0 || (n == 0 && final_block)) ? 1 : 0);
        return n + blocks * 5;
    }
    compressor c;
    memset(&c, 0, sizeof(c));
    c.level = level;
    init_params(&c);
    c.head = (int32_t *)malloc(HASH_SIZE * sizeof(int32_t));
    memset(c.head, 0xFF, HASH_SIZE * sizeof(int32_t));
    c.prev = (uint16_t *)calloc(32768, sizeof(uint16_t));
    if (level >= 10)
        c.bt = bt_new();
    const int lazy = level >= 8 ? 2 : level >= 5 ? 1 : 0;
    size_t p = 0, total = 0;
    while (p < n) {
        size_t bits = 0;
        if (level < 2) {
            p += accumulate_greedy_frequencies(&c, in, n, p);
            load_static_codes(&c);
            bits = 3 + block_data_bits(&c);
        } else if (level >= 10) {
            p += size_block_near_optimal(&c, in, n, p, &bits);
        } else {
            p += decide_greedy_sequences(&c, in, n, p, lazy);
            make_block_codes(&c);
            bits = 3 + dynamic_header_bits(&c) + block_data_bits(&c);
        }
        total += bits;
    }
    compressor_free(&c);
    return (total + 7) / 8;
}
