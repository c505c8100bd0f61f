/*
 * checksum.c — oracle restatement of the reference checksums (TEST ONLY).
 *
 *   Adler-32: src/adler32/mod.rs:8-104  (adler32_chunk + adler32_generic:
 *             sums accumulated in u32 over chunks of at most 4096 bytes, one
 *             modulo per chunk, modulus 65521).  The SIMD kernels in
 *             src/adler32/{x86,arm}.rs compute the same function.
 *   CRC-32:   src/crc32/mod.rs:4-9 (slice-by-1), :12-322 (slice-by-8) and the
 *             !f(!crc, data) wrapper at :364; tables src/crc32_tables.rs:1-35
 *             are the standard reflected 0xEDB88320 tables, generated here
 *             instead of being transcribed.  The PCLMULQDQ folds in
 *             src/crc32/x86.rs compute the same function.
 */
#include "oracle.h"

#define ADLER_DIVISOR 65521u
#define ADLER_MAX_CHUNK 4096u

uint32_t orc_adler32(uint32_t adler, const uint8_t *p, size_t n)
{
    uint32_t s1 = adler & 0xFFFF;
    uint32_t s2 = adler >> 16;
    while (n > 0) {
        size_t chunk = n < ADLER_MAX_CHUNK ? n : ADLER_MAX_CHUNK;
        /* 4096 * 255 * 4097 / 2 + 65520 * 4097 < 2^32: no overflow per chunk */
        for (size_t i = 0; i < chunk; i++) {
            s1 += p[i];
            s2 += s1;
        }
        s1 %= ADLER_DIVISOR;
        s2 %= ADLER_DIVISOR;
        p += chunk;
        n -= chunk;
    }
    return (s2 % ADLER_DIVISOR) << 16 | (s1 % ADLER_DIVISOR);
}

static uint32_t crc_tab[8][256];
static int crc_tab_ready;

static void crc_init(void)
{
    for (uint32_t b = 0; b < 256; b++) {
        uint32_t c = b;
        for (int k = 0; k < 8; k++)
            c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1)));
        crc_tab[0][b] = c;
    }
    for (uint32_t b = 0; b < 256; b++)
        for (int s = 1; s < 8; s++)
            crc_tab[s][b] =
                (crc_tab[s - 1][b] >> 8) ^ crc_tab[0][crc_tab[s - 1][b] & 0xFF];
    __atomic_store_n(&crc_tab_ready, 1, __ATOMIC_RELEASE);
}

uint32_t orc_crc32(uint32_t crc, const uint8_t *p, size_t n)
{
    if (!__atomic_load_n(&crc_tab_ready, __ATOMIC_ACQUIRE))
        crc_init();
    uint32_t c = ~crc;
    while (n >= 8) {
        uint32_t lo = c ^ ((uint32_t)p[0] | (uint32_t)p[1] << 8 |
                           (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
        uint32_t hi = (uint32_t)p[4] | (uint32_t)p[5] << 8 |
                      (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        c = crc_tab[7][lo & 0xFF] ^ crc_tab[6][(lo >> 8) & 0xFF] ^
            crc_tab[5][(lo >> 16) & 0xFF] ^ crc_tab[4][lo >> 24] ^
            crc_tab[3][hi & 0xFF] ^ crc_tab[2][(hi >> 8) & 0xFF] ^
            crc_tab[1][(hi >> 16) & 0xFF] ^ crc_tab[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--)
        c = (c >> 8) ^ crc_tab[0][(c ^ *p++) & 0xFF];
    return ~c;
}
