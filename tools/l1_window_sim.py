"""Lane-level model of the level-1 "whole window" round of deflate_l1_kernel (csrc/deflate_l1.cuh).

The reference's level-1 parse (HtMatchFinder, src/compress/matchfinder.rs:1139-1231) inserts only
the positions it probes, so the parse is serial.  The kernel's round resolves a window of 32
consecutive positions at once: every lane reads its bucket BEFORE the round, same-hash lower lanes
stand in for the bucket writes of the round, and the warp walks from match to match inside the
window without touching memory again except for the match lengths.  This file states that round
with plain lists (one entry per lane) so that it can be checked against the serial parse on the
CPU (tests/test_l1_window.py); the CUDA code follows it step for step.
"""

EMPTY = -1
MAX_OFFSET = 32768
MAX_MATCH = 258


def ld24(d, p):
    return d[p] | d[p + 1] << 8 | d[p + 2] << 16


def hash3(v):
    return ((v * 0x1E35A7BD) & 0xFFFFFFFF) >> 17


def prefix_len(d, a, b, room):
    n = 0
    while n < room and d[a + n] == d[b + n]:
        n += 1
    return n


def serial_tokens(d):
    """The reference's parse: a list of ints (literals) and (length, offset) tuples."""
    n = len(d)
    head = {}
    out = []
    pos = 0
    while pos < n:
        mlen = 0
        if pos + 3 <= n:
            v = ld24(d, pos)
            h = hash3(v)
            cur = head.get(h, EMPTY)
            head[h] = pos
            if cur != EMPTY and pos - cur <= MAX_OFFSET and ld24(d, cur) == v:
                mlen = prefix_len(d, cur, pos, min(n - pos, MAX_MATCH))
                off = pos - cur
        if mlen:
            out.append((mlen, off))
            pos += mlen
        else:
            out.append(d[pos])
            pos += 1
    return out


def window_tokens(d, spec_cap=None, per_round=None):
    """The kernel's round.  spec_cap: the per-lane speculative compare stops there (the kernel
    finishes longer matches cooperatively); it must not change the result.  per_round: a list that
    receives (first position, per-lane symbols) of every round."""
    n = len(d)
    table = {}
    out = []
    pos = 0
    rounds = 0
    static_steps = dynamic_steps = 0
    while pos < n:
        rounds += 1
        W = 32
        p = [pos + l for l in range(W)]
        hashable = [q + 3 <= n for q in p]
        v = [ld24(d, q) if hb else None for q, hb in zip(p, hashable)]
        h = [hash3(x) if hb else 0x10000 + l for l, (x, hb) in enumerate(zip(v, hashable))]
        tcand = [table.get(h[l], EMPTY) if hashable[l] else EMPTY for l in range(W)]
        tfound = [hashable[l] and tcand[l] != EMPTY and p[l] - tcand[l] <= MAX_OFFSET and ld24(d, tcand[l]) == v[l]
                  for l in range(W)]
        lower = [sum(1 << j for j in range(l) if h[j] == h[l]) for l in range(W)]
        # speculative candidate / length: as if every lower lane had been inserted
        spec_cand = [EMPTY] * W
        spec_len = [0] * W
        for l in range(W):
            if not hashable[l]:
                continue
            if lower[l]:
                j = lower[l].bit_length() - 1
                if v[j] == v[l]:
                    spec_cand[l] = pos + j
            elif tfound[l]:
                spec_cand[l] = tcand[l]
            if spec_cand[l] != EMPTY:
                room = min(n - p[l], MAX_MATCH)
                if spec_cap is not None:
                    room = min(room, spec_cap)
                spec_len[l] = prefix_len(d, spec_cand[l], p[l], room)
        inserted = 0
        cur = 0
        kind = [None] * W           # None = inside a match / past the end, int = literal, tuple = match
        # lanes whose candidate depends on what the round itself inserts, and the hits that do not
        depmask = sum(1 << l for l in range(W) if lower[l])
        smask = sum(1 << l for l in range(W) if tfound[l] and not lower[l])
        while True:
            below_cur = (1 << cur) - 1
            from_cur = ~below_cur & 0xFFFFFFFF
            sh = smask & from_cur
            ks = (sh & -sh).bit_length() - 1 if sh else W
            first = None
            if depmask & from_cur & ((1 << ks) - 1) == 0:
                # STATIC step: nothing between cur and the first plain hit looks at the round's own
                # insertions, so the hit is that lane with its bucket candidate (bit operations only)
                if ks < W:
                    first = (ks, tcand[ks])
                    assert spec_cand[ks] == tcand[ks]
                static_steps += 1
            else:
                for l in range(cur, W):
                    if not hashable[l]:
                        continue
                    # lower lanes that count: inserted in earlier steps, or assumed literal from cur on
                    eff = lower[l] & (inserted | ~below_cur)
                    if eff:
                        j = eff.bit_length() - 1
                        ok, c = v[j] == v[l], pos + j
                    else:
                        ok, c = tfound[l], tcand[l]
                    if ok:
                        first = (l, c)
                        break
                dynamic_steps += 1
            last_lit = W if first is None else first[0]
            for l in range(cur, last_lit):
                if p[l] < n:
                    kind[l] = d[p[l]]
            upto = W - 1 if first is None else first[0]
            for l in range(cur, upto + 1):
                if hashable[l]:
                    inserted |= 1 << l
            if first is None:
                nxt = pos + W
                break
            k, c = first
            room = min(n - p[k], MAX_MATCH)
            if c == spec_cand[k] and (spec_cap is None or spec_len[k] < spec_cap or spec_len[k] >= room):
                mlen = spec_len[k]
            else:
                mlen = prefix_len(d, c, p[k], room)
            kind[k] = (mlen, p[k] - c)
            cur = k + mlen
            if cur >= W:
                nxt = pos + cur
                break
        # bucket writes: the highest inserted lane of a hash wins
        for l in range(W):
            if inserted >> l & 1:
                table[h[l]] = p[l]
        out.extend(x for x in kind if x is not None)
        if per_round is not None:
            per_round.append((pos, kind))
        pos = nxt
    window_tokens.last_steps = (static_steps, dynamic_steps)
    return out, rounds


# --- block splitting on top of the window rounds (units above 64 KiB, size estimation):
# BlockSplitStats::should_end_block, src/compress/mod.rs:387-415; compress_greedy_block's level-1
# loop asks it in front of every symbol (:1531-1564)
MIN_BLOCK_LENGTH = 5000
SOFT_MAX_BLOCK_LENGTH = 300000


class SplitStats:
    def __init__(self):
        self.reset()

    def reset(self):
        self.new_obs = [0] * 14
        self.obs = [0] * 14
        self.num_new = 0
        self.num = 0

    def literal(self, b):
        self.new_obs[b >> 5] += 1
        self.num_new += 1

    def match(self, length, offset):
        slot = max(i for i in range(30) if _OFF_BASE[i] <= offset)
        self.new_obs[8 + (length >= 8)] += 1
        self.new_obs[10 + (0 if slot < 16 else 1 if slot < 24 else 2 if slot < 30 else 0)] += 1
        self.num_new += 2

    def should_end(self, block_len, remaining):
        if self.num_new < 2048 and block_len < SOFT_MAX_BLOCK_LENGTH:
            return False
        if remaining <= MIN_BLOCK_LENGTH:
            return False
        if block_len >= SOFT_MAX_BLOCK_LENGTH:
            return True
        if block_len >= MIN_BLOCK_LENGTH:
            if self.num:
                lg_all, lg_new = self.num.bit_length() - 1, self.num_new.bit_length() - 1
                old_bits = new_bits = 0
                for o_, k in zip(self.obs, self.new_obs):
                    if k:
                        lo, ln = (o_ + 1).bit_length() - 1, (k + 1).bit_length() - 1
                        old_bits += k * max(lg_all - lo, 0)
                        new_bits += k * max(lg_new - ln, 0)
                if old_bits - new_bits > block_len // 16:
                    return True
            self.obs = [a + b for a, b in zip(self.obs, self.new_obs)]
            self.new_obs = [0] * 14
            self.num += self.num_new
            self.num_new = 0
        return False


def serial_blocks(d):
    """The reference's level-1 block loop on an input above 64 KiB: a list of token lists."""
    n = len(d)
    tokens = serial_tokens(d)
    blocks, cur, st, p, start = [], [], SplitStats(), 0, 0
    for t in tokens:
        if st.should_end(p - start, n - p):
            blocks.append(cur)
            cur, start = [], p
            st.reset()
        if isinstance(t, tuple):
            st.match(*t)
            p += t[0]
        else:
            st.literal(t)
            p += 1
        cur.append(t)
    blocks.append(cur)
    return blocks


def window_blocks(d, spec_cap=16):
    """The kernel's way: the symbols of a window go to the statistics in lane order — in bulk while
    the round cannot bring 2048 observations together, one by one otherwise — and a cut lands in
    front of the symbol of one lane."""
    n = len(d)
    rounds = []
    window_tokens(d, spec_cap, rounds)
    blocks, cur, st, start = [], [], SplitStats(), 0
    bulk = slow = 0
    for pos, kind in rounds:
        syms = [(l, t) for l, t in enumerate(kind) if t is not None]
        nobs = sum(2 if isinstance(t, tuple) else 1 for _, t in syms)
        if st.num_new + nobs < 2048:
            bulk += 1
            for _, t in syms:                       # order does not matter here: plain counters
                st.match(*t) if isinstance(t, tuple) else st.literal(t)
            cur.extend(t for _, t in syms)
        else:
            slow += 1
            for l, t in syms:
                if st.should_end(pos + l - start, n - (pos + l)):
                    blocks.append(cur)
                    cur, start = [], pos + l
                    st.reset()
                st.match(*t) if isinstance(t, tuple) else st.literal(t)
                cur.append(t)
    blocks.append(cur)
    window_blocks.last_rounds = (bulk, slow)
    return blocks


# --- static-Huffman bit stream (RFC 1951 3.2.6), for the comparison with the oracle's bytes
_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_OFF_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
             4097, 6145, 8193, 12289, 16385, 24577]
_OFF_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def _rev(x, n):
    return int(format(x, "0%db" % n)[::-1], 2)


def encode_static(tokens, blocks=None):
    """One final static block of `tokens`, or the given list of blocks (the last one final)."""
    acc, nbits = 0, 0

    def put(bits, n):
        nonlocal acc, nbits
        acc |= bits << nbits
        nbits += n

    def litlen(sym):
        if sym < 144:
            put(_rev(0x30 + sym, 8), 8)
        elif sym < 256:
            put(_rev(0x190 + sym - 144, 9), 9)
        elif sym < 280:
            put(_rev(sym - 256, 7), 7)
        else:
            put(_rev(0xC0 + sym - 280, 8), 8)

    for bi, toks in enumerate([tokens] if blocks is None else blocks):
        last = blocks is None or bi == len(blocks) - 1
        put((1 if last else 0) | 2, 3)              # BFINAL, BTYPE = 01
        for t in toks:
            if isinstance(t, tuple):
                ln, off = t
                s = max(i for i in range(29) if _LEN_BASE[i] <= ln)
                litlen(257 + s)
                put(ln - _LEN_BASE[s], _LEN_EXTRA[s])
                s = max(i for i in range(30) if _OFF_BASE[i] <= off)
                put(_rev(s, 5), 5)
                put(off - _OFF_BASE[s], _OFF_EXTRA[s])
            else:
                litlen(t)
        litlen(256)
    return acc.to_bytes((nbits + 7) // 8, "little")
