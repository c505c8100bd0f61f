/*
 * nearopt.inc.c — oracle restatement of compress_near_optimal_block
 * (src/compress/mod.rs:1586-1773, costs :2209-2234), included by deflate.c.
 * TEST ONLY.  Levels 10..12 are the ratio-tolerance tier: the GPU engine may
 * use another parser; this restatement provides the reference's size.
 *
 * Kept quirks: the DP pass runs on the block slice with
 * freshly reset hash tables, so optimal-parse matches never reach back before
 * the block; afterwards the tables hold block-relative positions which the
 * next block's greedy pass reads as input-relative candidates; symbols absent
 * from the greedy pass have cost 0 in the DP (their code length is 0).
 */
static size_t compress_near_optimal_block(compressor *c, const uint8_t *in, size_t n,
                                          size_t start, bitw *bs, int final_block)
{
    /* pass 1: greedy parse with the binary-tree matchfinder -> split point + first costs */
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
    const size_t done = p - start;
    const uint8_t *blk = in + start;
    const int is_final = start + done >= n && final_block;
    c->litlen_freqs[256]++;
    make_block_codes(c);

    /* update_costs, :2209-2224 */
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

    /* pass 2: forward DP over the block slice */
    bt_reset(c->bt);
    uint16_t list[260][2]; /* hash3 hit + strictly increasing lengths 4..258 */
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

    /* backtrack, then rebuild sequences / histograms in forward order (:1719-1751) */
    c->nseq = 0;
    memset(c->litlen_freqs, 0, sizeof(c->litlen_freqs));
    memset(c->offset_freqs, 0, sizeof(c->offset_freqs));
    c->litlen_freqs[256] = 1;
    size_t nsteps = 0;
    for (size_t r = done; r > 0; r -= path[r] & 0xFFFF)
        nsteps++;
    uint32_t *steps = (uint32_t *)malloc((nsteps ? nsteps : 1) * sizeof(uint32_t));
    {
        size_t k = nsteps;
        for (size_t r = done; r > 0; r -= path[r] & 0xFFFF)
            steps[--k] = path[r];
    }
    uint32_t litrun = 0;
    size_t at = 0;
    for (size_t k = 0; k < nsteps; k++) {
        unsigned len = steps[k] & 0xFFFF, off = steps[k] >> 16;
        if (len == 1) {
            c->litlen_freqs[blk[at]]++;
            litrun++;
            at++;
        } else {
            push_seq(c, litrun, len, off);
            c->litlen_freqs[257 + length_slot(len)]++;
            c->offset_freqs[offset_slot(off)]++;
            litrun = 0;
            at += len;
        }
    }
    free(steps);
    push_seq(c, litrun, 0, 0);
    make_block_codes(c);
    write_dynamic_block(c, bs, in, start, is_final);
    return done;
}
