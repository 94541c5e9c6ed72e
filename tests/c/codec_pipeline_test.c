/* The layers above the I/O path from plain C (include/lsio_codecs.h, lsio_uring_cuda.h, lsio_aligned_bytes.h): ordered
 * groups, a batch of completion items handed to the native codec pipeline [LZ4, crc32c], the decoded rows copied back
 * and compared.  What a `lsio_codecs` crate (planned_design.md:83-90) would do with the C ABI; no Python, no C++.
 * Build:  gcc -std=c11 -O2 -Iinclude tests/c/codec_pipeline_test.c -Llight-speed-io_b200 -llsio_b200 -Wl,-rpath,$PWD/light-speed-io_b200
 * Usage:  codec_pipeline_test <directory>      exit code 0 = pass */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "lsio_codecs.h"
#include "lsio_uring_cuda.h"

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)
enum { N_CHUNKS = 200, DECODED = 4096, SEED_LITS = 16, TAIL_LITS = 5, MATCH = DECODED - SEED_LITS - TAIL_LITS };

static uint32_t crc32c(const unsigned char* p, size_t n) { /* CRC-32C, bitwise (poly 0x82F63B78 reflected) */
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
    }
    return ~c;
}

/* One LZ4 block by hand: 16 literals, a match of MATCH bytes at offset 16 (periodic extension), 5 final literals.
 * Writes the stored chunk (block + crc32c trailer, Zarr v3 crc32c codec) and its decoded form. Returns the stored size. */
static size_t make_chunk(int k, unsigned char* stored, unsigned char* decoded) {
    size_t o = 0;
    unsigned char seed[SEED_LITS], tail[TAIL_LITS];
    for (int i = 0; i < SEED_LITS; i++) seed[i] = (unsigned char)(k * 31 + i * 7 + 1);
    for (int i = 0; i < TAIL_LITS; i++) tail[i] = (unsigned char)(k + 200 + i);
    stored[o++] = 0xFF;                   /* token: 15 literals + extension, match length 15 + extension */
    stored[o++] = SEED_LITS - 15;         /* literal length extension */
    memcpy(stored + o, seed, SEED_LITS);
    o += SEED_LITS;
    stored[o++] = SEED_LITS;              /* offset 16, little endian */
    stored[o++] = 0;
    unsigned rest = MATCH - 4 - 15;       /* match length extension */
    while (rest >= 255) { stored[o++] = 255; rest -= 255; }
    stored[o++] = (unsigned char)rest;
    stored[o++] = TAIL_LITS << 4;         /* last sequence: literals only */
    memcpy(stored + o, tail, TAIL_LITS);
    o += TAIL_LITS;
    const uint32_t c = crc32c(stored, o);
    for (int i = 0; i < 4; i++) stored[o++] = (unsigned char)(c >> (8 * i));
    for (int i = 0; i < SEED_LITS + MATCH; i++) decoded[i] = seed[i % SEED_LITS];
    memcpy(decoded + SEED_LITS + MATCH, tail, TAIL_LITS);
    return o;
}

int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : "/tmp";
    static unsigned char file[N_CHUNKS * 128 + 64], expect[N_CHUNKS][DECODED], got[N_CHUNKS * DECODED];
    int64_t starts[N_CHUNKS], ends[N_CHUNKS];
    uint64_t user_data[N_CHUNKS];
    size_t o = 3; /* unaligned start */
    for (int k = 0; k < N_CHUNKS; k++) {
        starts[k] = (int64_t)o;
        o += make_chunk(k, file + o, expect[k]);
        ends[k] = (int64_t)o;
        user_data[k] = (uint64_t)k;
    }
    char filename[512];
    snprintf(filename, sizeof filename, "%s/lsio_codec_tempfile_%d", dir, (int)getpid());
    int fd = open(filename, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0 || write(fd, file, o + 7) != (ssize_t)(o + 7) || fsync(fd) || close(fd)) FAIL("cannot write %s", filename);

    lsio_uring_cuda* uring = NULL;
    int rc = lsio_uring_cuda_new(2, &uring);
    if (rc) FAIL("lsio_uring_cuda_new: %d (%s)", rc, lsio_uring_cuda_last_error());
    /* two ordered groups (lsio_io lib.rs:6-8): the first half of the chunks, then the second */
    if (lsio_uring_cuda_get_ranges(uring, filename, starts, ends, N_CHUNKS / 2, user_data, N_CHUNKS / 2)) FAIL("get_ranges");
    if (lsio_uring_cuda_next_group(uring) != 1 || lsio_uring_cuda_current_group(uring) != 1) FAIL("next_group");
    if (lsio_uring_cuda_get_ranges(uring, filename, starts + N_CHUNKS / 2, ends + N_CHUNKS / 2, N_CHUNKS / 2, user_data + N_CHUNKS / 2, N_CHUNKS / 2)) FAIL("get_ranges 2");

    lsio_completion* rx = lsio_uring_cuda_completion(uring);
    static lsio_output items[N_CHUNKS];
    size_t n = 0;
    while (n < N_CHUNKS) {
        size_t m = 0;
        if (lsio_completion_recv_batch(rx, items + n, N_CHUNKS - n, 10000, &m) != LSIO_RECV_OK) FAIL("recv_batch after %zu items", n);
        n += m;
    }
    for (size_t i = 0; i < n; i++) {
        if (items[i].status != 0) FAIL("item %zu: status %d (%s)", i, items[i].status, items[i].context);
        if (items[i].group != (items[i].chunk.user_data < N_CHUNKS / 2 ? 0u : 1u)) FAIL("item %zu: group %llu", i, (unsigned long long)items[i].group);
        if (i && items[i].group < items[i - 1].group) FAIL("groups out of order at item %zu", i);
        if (!lsio_codecs_crc32c_trailer_ok(items[i].crc32c)) FAIL("item %zu: crc32c trailer", i);
    }

    /* decode the whole batch: stages in encode order (compressor, then checksum) */
    const lsio_codec_stage stages[2] = {{LSIO_CODEC_STAGE_LZ4, 0}, {LSIO_CODEC_STAGE_CRC32C, 0}};
    lsio_codec_pipeline* pipe = NULL;
    rc = lsio_codec_pipeline_create(0, stages, 2, DECODED, N_CHUNKS, &pipe);
    if (rc) FAIL("lsio_codec_pipeline_create: %d", rc);
    lsio_aligned_bytes_mut* rows = lsio_aligned_bytes_mut_new_device((size_t)N_CHUNKS * DECODED, 512, 0);
    if (!rows) FAIL("device allocation");
    rc = lsio_codec_pipeline_decode(pipe, items, n, lsio_aligned_bytes_mut_as_mut_ptr(rows), NULL);
    if (rc) FAIL("lsio_codec_pipeline_decode: %d", rc);
    static int32_t status[N_CHUNKS];
    const int64_t bad = lsio_codec_pipeline_finish(pipe, status);
    if (bad != 0) FAIL("lsio_codec_pipeline_finish: %lld bad, first status %d", (long long)bad, status[0]);
    lsio_aligned_bytes* frozen = NULL;
    if (lsio_aligned_bytes_mut_freeze(rows, &frozen)) FAIL("freeze");
    if (lsio_aligned_bytes_copy_to_host(frozen, got, sizeof got) != (int64_t)sizeof got) FAIL("copy_to_host");
    for (size_t i = 0; i < n; i++)
        if (memcmp(got + i * DECODED, expect[items[i].chunk.user_data], DECODED)) FAIL("row %zu (chunk %llu) differs", i, (unsigned long long)items[i].chunk.user_data);

    /* a corrupted trailer is reported per item, the others still decode */
    items[7].crc32c ^= 1u;
    if (lsio_codec_pipeline_decode(pipe, items, n, (void*)lsio_aligned_bytes_as_ptr(frozen), NULL)) FAIL("decode 2");
    if (lsio_codec_pipeline_finish(pipe, status) != 1 || status[7] != LSIO_CODEC_ERR_CHECKSUM || status[8] != LSIO_CODEC_OK) FAIL("checksum error not reported");

    for (size_t i = 0; i < n; i++) lsio_aligned_bytes_drop(items[i].chunk.buffer);
    lsio_aligned_bytes_drop(frozen);
    lsio_codec_pipeline_destroy(pipe);
    lsio_completion_drop(rx);
    lsio_uring_cuda_free(uring);
    unlink(filename);
    printf("%d chunks in 2 ordered groups decoded by the native pipeline and match\n", N_CHUNKS);
    return 0;
}
