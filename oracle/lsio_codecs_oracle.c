/*
 * lsio_codecs_oracle.c -- CPU ORACLE for the codec layer.  TEST INFRASTRUCTURE ONLY: loaded by tests/,
 * __graft_entry__.smoke() and bench legs that time a CPU baseline; nothing under light-speed-io_b200/
 * links or loads it.
 *
 * The reference only PLANS this layer ("lsio_codecs: Compression / decompression",
 * planned_design.md:83-90; "Wrap a few decompression algorithms", README.md:68-70) -- it holds no codec
 * source, tests or vectors, so there is nothing of the reference's to restate.  What is restated here are
 * the PUBLISHED formats the Zarr ecosystem stores chunks in (the reference's stated target format,
 * planned_design.md:48-49, :100-104):
 *   - LZ4 block format (lz4 "lz4_Block_format.md", format as of liblz4 1.9.x): sequences of
 *     token | literal-length bytes | literals | 2-byte LE offset | match-length bytes; last sequence has
 *     literals only.  numcodecs.LZ4 (Zarr v2) = 4-byte LE uncompressed size + one LZ4 block.
 *   - byte shuffle (numcodecs.Shuffle / Blosc SHUFFLE): byte j of element i moves to plane j, position i;
 *     the len % elementsize trailing bytes are copied unchanged.
 *   - Zarr v3 `crc32c` codec: chunk bytes followed by their CRC-32C as 4 little-endian bytes.
 * PARITY PINNING: liblz4 itself (through pyarrow's "lz4_raw" codec, pyarrow 24 in this image) -- this
 * decoder must invert pyarrow's compressor and pyarrow must invert this compressor, on the committed
 * vectors of tests/golden/codecs.json and on random inputs (tests/test_oracle_codecs.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- LZ4 block decoder: returns the decoded length, or a negative code -------------------------------
 * -1 input overrun / truncated, -2 output overflow, -3 offset 0 or reaching before the output start */
int64_t lsio_oracle_lz4_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t ip = 0, op = 0;
    if (n == 0) return -1;
    for (;;) {
        if (ip >= n) return -1;
        const unsigned token = src[ip++];
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= n) return -1;
                b = src[ip++];
                lit += b;
            } while (b == 255);
        }
        if (lit > n - ip) return -1;
        if (lit > cap - op) return -2;
        memcpy(dst + op, src + ip, lit);
        ip += lit;
        op += lit;
        if (ip == n) return (int64_t)op; /* last sequence: literals only */
        if (n - ip < 2) return -1;
        const size_t offset = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
        ip += 2;
        if (offset == 0 || offset > op) return -3;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (ip >= n) return -1;
                b = src[ip++];
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (ml > cap - op) return -2;
        for (size_t i = 0; i < ml; i++) dst[op + i] = dst[op + i - offset]; /* may overlap: byte order matters */
        op += ml;
    }
}

/* ---- LZ4 block compressor (greedy, one hash probe): any valid LZ4 block will do for a test-data generator.
 * Honours the format's end-of-block rules: the last 5 bytes are literals and the last match starts at
 * least 12 bytes before the end.  Returns the compressed length or -2 if `cap` is too small. */
static size_t put_len(uint8_t* dst, size_t op, size_t v) {
    while (v >= 255) {
        dst[op++] = 255;
        v -= 255;
    }
    dst[op++] = (uint8_t)v;
    return op;
}
size_t lsio_oracle_lz4_compress_bound(size_t n) { return n + n / 255 + 16; }
int64_t lsio_oracle_lz4_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    enum { HB = 14 };
    if (cap < lsio_oracle_lz4_compress_bound(n)) return -2;
    uint32_t* table = (uint32_t*)calloc((size_t)1 << HB, sizeof(uint32_t)); /* position + 1 */
    if (!table) return -2;
    size_t ip = 0, anchor = 0, op = 0;
    const size_t mflimit = n >= 12 ? n - 12 : 0; /* a match may not start beyond this */
    const size_t matchlimit = n >= 5 ? n - 5 : 0; /* nor extend beyond this */
    while (n >= 13 && ip <= mflimit) {
        uint32_t seq;
        memcpy(&seq, src + ip, 4);
        const uint32_t h = (seq * 2654435761u) >> (32 - HB);
        const size_t cand = table[h];
        table[h] = (uint32_t)(ip + 1);
        uint32_t cseq = 0;
        if (cand) memcpy(&cseq, src + cand - 1, 4);
        if (!cand || cseq != seq || ip - (cand - 1) > 65535) {
            ip++;
            continue;
        }
        const size_t ref = cand - 1;
        size_t ml = 4;
        while (ip + ml < matchlimit && src[ref + ml] == src[ip + ml]) ml++;
        const size_t lit = ip - anchor;
        const size_t tok = op++;
        dst[tok] = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
        if (lit >= 15) op = put_len(dst, op, lit - 15);
        memcpy(dst + op, src + anchor, lit);
        op += lit;
        dst[op++] = (uint8_t)((ip - ref) & 255);
        dst[op++] = (uint8_t)((ip - ref) >> 8);
        dst[tok] |= (uint8_t)(ml - 4 >= 15 ? 15 : ml - 4);
        if (ml - 4 >= 15) op = put_len(dst, op, ml - 4 - 15);
        ip += ml;
        anchor = ip;
    }
    const size_t lit = n - anchor;
    const size_t tok = op++;
    dst[tok] = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
    if (lit >= 15) op = put_len(dst, op, lit - 15);
    memcpy(dst + op, src + anchor, lit);
    op += lit;
    free(table);
    return (int64_t)op;
}

/* ---- byte shuffle / unshuffle (numcodecs.Shuffle, Blosc SHUFFLE) --------------------------------------- */
void lsio_oracle_shuffle(const uint8_t* src, uint8_t* dst, size_t n, size_t elem) {
    const size_t cnt = elem ? n / elem : 0;
    for (size_t i = 0; i < cnt; i++)
        for (size_t j = 0; j < elem; j++) dst[j * cnt + i] = src[i * elem + j];
    memcpy(dst + cnt * elem, src + cnt * elem, n - cnt * elem);
}
void lsio_oracle_unshuffle(const uint8_t* src, uint8_t* dst, size_t n, size_t elem) {
    const size_t cnt = elem ? n / elem : 0;
    for (size_t i = 0; i < cnt; i++)
        for (size_t j = 0; j < elem; j++) dst[i * elem + j] = src[j * cnt + i];
    memcpy(dst + cnt * elem, src + cnt * elem, n - cnt * elem);
}
