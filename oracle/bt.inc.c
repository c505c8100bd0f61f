/*
 * bt.inc.c — oracle restatement of BtMatchFinder (levels 10..12), included by
 * deflate.c.  TEST ONLY.
 *
 *   BtMatchFinder::{new,reset}          src/compress/matchfinder.rs:1316-1331
 *   advance_one_byte_generic            src/compress/matchfinder.rs:1344-1463
 *   visitors (best / all / no-op)       src/compress/matchfinder.rs:1234-1306
 *   find_match / skip_positions / find_matches   :1465-1775
 *
 * Positions are relative to the `data` pointer of each call and base_offset
 * is 0: within one stream the reference only changes base_offset through
 * reset() (-> 0) and the end-of-stream advance(), after which every stored
 * position is below the new base and is ignored (:1374,1404), i.e. the next
 * stream starts as if the tables were clear.  child_tab is deliberately NOT
 * cleared by reset(), as in the reference (:1327-1331).
 */
struct bt_state {
    int32_t (*hash3)[2]; /* 1 << 16 buckets, 2-way */
    int32_t *hash4;      /* 1 << 16 */
    int32_t (*child)[2]; /* 32768, indexed pos & 32767 */
};

static void bt_reset(struct bt_state *b)
{
    memset(b->hash3, 0xFF, sizeof(int32_t[2]) << 16);
    memset(b->hash4, 0xFF, sizeof(int32_t) << 16);
}
static struct bt_state *bt_new(void)
{
    struct bt_state *b = (struct bt_state *)malloc(sizeof(*b));
    b->hash3 = (int32_t(*)[2])malloc(sizeof(int32_t[2]) << 16);
    b->hash4 = (int32_t *)malloc(sizeof(int32_t) << 16);
    b->child = (int32_t(*)[2])calloc(32768, sizeof(int32_t[2]));
    bt_reset(b);
    return b;
}
static void bt_free(struct bt_state *b)
{
    free(b->hash3);
    free(b->hash4);
    free(b->child);
    free(b);
}

typedef struct {
    int mode; /* 0 no-op, 1 best, 2 all */
    unsigned best_len, best_off;
    uint16_t (*list)[2];
    unsigned nlist;
} bt_visitor;

static inline void bt_on_hash3(bt_visitor *v, unsigned len, unsigned off)
{
    if (v->mode == 1) {
        v->best_len = len;
        v->best_off = off;
    } else if (v->mode == 2) {
        v->list[v->nlist][0] = (uint16_t)len;
        v->list[v->nlist][1] = (uint16_t)off;
        v->nlist++;
    }
}
static inline void bt_on_match(bt_visitor *v, unsigned len, unsigned off)
{
    if (v->mode == 0 || len <= v->best_len)
        return;
    v->best_len = len;
    if (v->mode == 1) {
        v->best_off = off;
    } else {
        v->list[v->nlist][0] = (uint16_t)len;
        v->list[v->nlist][1] = (uint16_t)off;
        v->nlist++;
    }
}

static void bt_advance_one_byte(struct bt_state *b, const uint8_t *d, size_t n, size_t pos,
                                unsigned max_depth, unsigned nice_len, bt_visitor *v)
{
    if (pos + 4 > n)
        return;
    const uint8_t *src = d + pos;
    uint32_t val = ld32(src);
    uint32_t h3 = ((val & 0xFFFFFF) * 0x1E35A7BDu) >> 16;
    uint32_t h4 = (val * 0x1E35A7BDu) >> 16;
    int32_t self = (int32_t)pos;
    int32_t c3 = b->hash3[h3][0];
    b->hash3[h3][0] = self;
    int32_t c3b = b->hash3[h3][1];
    b->hash3[h3][1] = c3;
    int32_t cutoff = self - 32768;

    if (c3 != -1 && c3 > cutoff) {
        if (ld24(d + c3) == (val & 0xFFFFFF))
            bt_on_hash3(v, 3, (unsigned)(self - c3));
        else if (c3b != -1 && c3b > cutoff && ld24(d + c3b) == (val & 0xFFFFFF))
            bt_on_hash3(v, 3, (unsigned)(self - c3b));
    }

    int32_t cur = b->hash4[h4];
    b->hash4[h4] = self;
    unsigned me = (unsigned)pos & 32767u;
    if (cur == -1 || cur <= cutoff) {
        b->child[me][0] = -1;
        b->child[me][1] = -1;
        return;
    }
    unsigned depth_left = max_depth;
    unsigned lt_node = me, lt_side = 0, gt_node = me, gt_side = 1;
    unsigned room = n - pos < MAX_MATCH ? (unsigned)(n - pos) : MAX_MATCH;
    for (;;) {
        unsigned ci = (unsigned)cur & 32767u;
        const uint8_t *m = d + cur;
        unsigned len = prefix_len(m, src, room);
        bt_on_match(v, len, (unsigned)(self - cur));
        if (len >= nice_len || len == room) {
            b->child[lt_node][lt_side] = b->child[ci][0];
            b->child[gt_node][gt_side] = b->child[ci][1];
            return;
        }
        if (m[len] < src[len]) {
            b->child[lt_node][lt_side] = cur;
            lt_node = ci;
            lt_side = 1;
            cur = b->child[ci][1];
        } else {
            b->child[gt_node][gt_side] = cur;
            gt_node = ci;
            gt_side = 0;
            cur = b->child[ci][0];
        }
        if (cur == -1 || cur <= cutoff || --depth_left == 0) {
            b->child[lt_node][lt_side] = -1;
            b->child[gt_node][gt_side] = -1;
            return;
        }
    }
}

static unsigned bt_find_match(struct bt_state *b, const uint8_t *d, size_t n, size_t pos,
                              unsigned max_depth, unsigned nice_len, unsigned *off)
{
    bt_visitor v = {1, 0, 0, 0, 0};
    bt_advance_one_byte(b, d, n, pos, max_depth, nice_len, &v);
    *off = v.best_off;
    return v.best_len;
}
static void bt_skip_positions(struct bt_state *b, const uint8_t *d, size_t n, size_t pos,
                              size_t count, unsigned max_depth, unsigned nice_len)
{
    bt_visitor v = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < count; i++)
        bt_advance_one_byte(b, d, n, pos + i, max_depth, nice_len, &v);
}
/* find_matches: strictly increasing lengths; a length-3 match can only come
 * from the hash3 probe because the tree visitor starts at best_len = 3
 * (:1274-1295). */
static unsigned bt_find_matches(struct bt_state *b, const uint8_t *d, size_t n, size_t pos,
                                unsigned max_depth, unsigned nice_len, uint16_t (*list)[2])
{
    bt_visitor v = {2, 3, 0, list, 0};
    bt_advance_one_byte(b, d, n, pos, max_depth, nice_len, &v);
    return v.nlist;
}
