/* crates/lsio_uring/tests/integration_test.rs:13-120 restated in plain C against the C ABI of
 * liblsio_b200.so (include/lsio_uring_cuda.h, lsio_io.h, lsio_aligned_bytes.h): no Python, no C++.
 * Build:  gcc -std=c11 -O2 -Iinclude tests/c/integration_test.c -Llight-speed-io_b200 -llsio_b200 -Wl,-rpath,$PWD/light-speed-io_b200
 * Usage:  integration_test <directory>      exit code 0 = pass */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "lsio_uring_cuda.h"

#define KIBIBYTE 1024
#define MEBIBYTE (KIBIBYTE * 1024)
#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

int main(int argc, char** argv) {
    enum { N_WORKER_THREADS = 4, FILE_SIZE = MEBIBYTE, CHUNK_SIZE = KIBIBYTE * 4, N_CHUNKS = FILE_SIZE / CHUNK_SIZE };
    const char* dir = argc > 1 ? argv[1] : "/tmp";
    /* random printable ASCII with period 1.5 * chunk so adjacent chunks differ (integration_test.rs:22-31) */
    static unsigned char file_contents[FILE_SIZE], assembled[FILE_SIZE];
    const int period = CHUNK_SIZE * 3 / 2;
    srand(12345);
    for (int i = 0; i < period; i++) file_contents[i] = (unsigned char)(32 + rand() % 95);
    for (int i = period; i < FILE_SIZE; i++) file_contents[i] = file_contents[i % period];
    char filename[512];
    snprintf(filename, sizeof filename, "%s/lsio_uring_tempfile_%d", dir, (int)getpid());
    int fd = open(filename, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0 || write(fd, file_contents, FILE_SIZE) != FILE_SIZE || fsync(fd) || close(fd)) FAIL("cannot write %s", filename);

    int64_t starts[N_CHUNKS], ends[N_CHUNKS];
    uint64_t user_data[N_CHUNKS];
    for (int i = 0; i < N_CHUNKS; i++) { starts[i] = (int64_t)i * CHUNK_SIZE; ends[i] = starts[i] + CHUNK_SIZE; user_data[i] = (uint64_t)i; }

    lsio_uring_cuda* uring = NULL;
    int rc = lsio_uring_cuda_new(N_WORKER_THREADS, &uring); /* IoUring::new(N_WORKER_THREADS) */
    if (rc) FAIL("lsio_uring_cuda_new: %d (%s)", rc, lsio_uring_cuda_last_error());
    /* a mismatch of the two lengths is refused (the reference asserts, get_ranges.rs:28) */
    if (lsio_uring_cuda_get_ranges(uring, filename, starts, ends, N_CHUNKS, user_data, N_CHUNKS - 1) == 0) FAIL("length mismatch accepted");
    rc = lsio_uring_cuda_get_ranges(uring, filename, starts, ends, N_CHUNKS, user_data, N_CHUNKS);
    if (rc) FAIL("get_ranges: %d", rc);

    lsio_completion* completion = lsio_uring_cuda_completion(uring);
    lsio_aligned_bytes* vec[N_CHUNKS] = {0};
    for (int i = 0; i < N_CHUNKS; i++) {
        lsio_output out;
        rc = lsio_completion_recv(completion, &out, 5000); /* recv_timeout */
        if (rc == LSIO_RECV_TIMEOUT) FAIL("Timed out waiting for chunk %d!", i);
        if (rc == LSIO_RECV_DISCONNECTED) FAIL("Disconnected whilst waiting for chunk %d!", i);
        if (out.status != 0) FAIL("Error reading chunk %d! %s", i, out.context);
        if (out.kind != LSIO_OUTPUT_CHUNK || out.chunk.user_data >= N_CHUNKS || vec[out.chunk.user_data]) FAIL("bad or duplicate chunk");
        if (lsio_aligned_bytes_memory_kind(out.chunk.buffer) != LSIO_MEM_DEVICE || lsio_aligned_bytes_len(out.chunk.buffer) != CHUNK_SIZE) FAIL("not a device view of the right size");
        vec[out.chunk.user_data] = out.chunk.buffer;
    }
    /* the receiver survives the backend and then reports the disconnect (integration_test.rs:92-96) */
    lsio_completion* clone = lsio_completion_clone(completion);
    lsio_completion_drop(completion);
    lsio_uring_cuda_free(uring); /* drop(uring) */
    lsio_output out;
    if (lsio_completion_recv(clone, &out, -1) != LSIO_RECV_DISCONNECTED) FAIL("expected a disconnected channel after drop");
    lsio_completion_drop(clone);
    /* chunks outlive the backend; reassemble by user_data and compare with the file */
    for (int i = 0; i < N_CHUNKS; i++) {
        if (lsio_aligned_bytes_copy_to_host(vec[i], assembled + (size_t)i * CHUNK_SIZE, CHUNK_SIZE) != CHUNK_SIZE) FAIL("copy_to_host failed for chunk %d", i);
        lsio_aligned_bytes_drop(vec[i]);
    }
    if (memcmp(assembled, file_contents, FILE_SIZE) != 0) FAIL("assembled buffer differs from the file");
    unlink(filename);
    printf("ok: %d chunks of %d bytes landed in HBM and match the file\n", N_CHUNKS, CHUNK_SIZE);
    return 0;
}
