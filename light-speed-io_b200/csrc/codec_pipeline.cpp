// codec_pipeline.cpp -- the decode side of a Zarr codec chain over a batch of completion-channel items
// (include/lsio_codecs.h).  The reference plans this layer as `lsio_codecs` consuming the I/O layer's output
// channel (planned_design.md:83-90) and holds no code for it.  Host side only: descriptor building and launches;
// the kernels are in codecs.cu.  No oracle, no CPU fallback.
#include <vector>

#include "aligned_bytes.hpp"
#include "common.hpp"
#include "lsio_codecs.h"
#include "lsio_cuda.h"

struct lsio_codec_pipeline {
    int device = 0;
    std::vector<lsio_codec_stage> stages;  // encode order
    uint64_t nbytes = 0;
    uint32_t max_batch = 0;
    bool has_crc = false, has_lz4 = false, has_shuffle = false;
    uint32_t lz4_header = 0, elem = 0;
    // pinned + mapped: the kernels read descriptors from and write results to host memory directly
    lsio_codec_item* items_lz4 = nullptr;
    lsio_codec_item* items_shuf = nullptr;
    lsio_codec_result* results = nullptr;
    lsio_codec_item *d_items_lz4 = nullptr, *d_items_shuf = nullptr;
    lsio_codec_result* d_results = nullptr;
    uint32_t* d_tile_map = nullptr;  // rows are uniform: the {row, tile} table of max_batch rows serves any n
    uint32_t tiles_per_row = 0;
    uint8_t* scratch = nullptr;  // lz4 output when a shuffle stage follows
    cudaEvent_t done = nullptr, t0 = nullptr, t1 = nullptr;  // t0..t1: the lz4 launch, t1..done: the rest (timing aid)
    std::vector<int32_t> host_status;  // per item, decided on the host (crc, bad item)
    size_t n_inflight = 0;
    bool inflight = false;
};

using namespace lsio;

static void free_pipeline(lsio_codec_pipeline* p) {
    DeviceGuard dg(p->device);
    if (p->items_lz4) cudaFreeHost(p->items_lz4);
    if (p->items_shuf) cudaFreeHost(p->items_shuf);
    if (p->results) cudaFreeHost(p->results);
    if (p->d_tile_map) cudaFree(p->d_tile_map);
    if (p->scratch) cudaFree(p->scratch);
    if (p->done) cudaEventDestroy(p->done);
    if (p->t0) cudaEventDestroy(p->t0);
    if (p->t1) cudaEventDestroy(p->t1);
    delete p;
}

static int init_pipeline(lsio_codec_pipeline* p) {
    DeviceGuard dg(p->device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    const unsigned flags = cudaHostAllocPortable | cudaHostAllocMapped;
    LSIO_CUDA_TRY(cudaHostAlloc((void**)&p->items_lz4, (size_t)p->max_batch * sizeof(lsio_codec_item), flags));
    LSIO_CUDA_TRY(cudaHostAlloc((void**)&p->items_shuf, (size_t)p->max_batch * sizeof(lsio_codec_item), flags));
    LSIO_CUDA_TRY(cudaHostAlloc((void**)&p->results, (size_t)p->max_batch * sizeof(lsio_codec_result), flags));
    LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&p->d_items_lz4, p->items_lz4, 0));
    LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&p->d_items_shuf, p->items_shuf, 0));
    LSIO_CUDA_TRY(cudaHostGetDevicePointer((void**)&p->d_results, p->results, 0));
    LSIO_CUDA_TRY(cudaEventCreate(&p->done));
    LSIO_CUDA_TRY(cudaEventCreate(&p->t0));
    LSIO_CUDA_TRY(cudaEventCreate(&p->t1));
    if (p->has_shuffle) {
        p->tiles_per_row = lsio_codecs_shuffle_tiles((uint32_t)p->nbytes);
        std::vector<uint32_t> lens(p->max_batch, (uint32_t)p->nbytes), map(2 * (size_t)p->max_batch * p->tiles_per_row);
        if (lsio_codecs_shuffle_tile_map(lens.data(), p->max_batch, map.data(), (size_t)p->max_batch * p->tiles_per_row) < 0) return -E2BIG;
        LSIO_CUDA_TRY(cudaMalloc((void**)&p->d_tile_map, map.size() * sizeof(uint32_t)));
        LSIO_CUDA_TRY(cudaMemcpy(p->d_tile_map, map.data(), map.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (p->has_lz4 && p->has_shuffle) LSIO_CUDA_TRY(cudaMalloc((void**)&p->scratch, (size_t)p->max_batch * p->nbytes));
    p->host_status.assign(p->max_batch, 0);
    return 0;
}

extern "C" {

int lsio_codec_pipeline_create(int device, const lsio_codec_stage* stages, size_t n_stages, uint64_t decoded_nbytes, uint32_t max_batch,
                               lsio_codec_pipeline** out) {
    if (!out || (!stages && n_stages) || decoded_nbytes == 0 || decoded_nbytes % 16 || decoded_nbytes > 0xFFFFFFF0ull || max_batch == 0) return -EINVAL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount (the codec layer has no CPU fallback)", __FILE__, __LINE__);
    if (device < 0 || device >= ndev) return -ENODEV;
    auto* p = new lsio_codec_pipeline();
    p->device = device;
    p->nbytes = decoded_nbytes;
    p->max_batch = max_batch;
    // encode order must be [shuffle] [lz4] [crc32c]: filters, then the compressor, then the checksum
    uint32_t last = 0;
    for (size_t i = 0; i < n_stages; i++) {
        const lsio_codec_stage& st = stages[i];
        if (st.kind <= last || st.kind > LSIO_CODEC_STAGE_CRC32C) {
            delete p;
            return -EINVAL;
        }
        last = st.kind;
        if (st.kind == LSIO_CODEC_STAGE_SHUFFLE) p->has_shuffle = true, p->elem = st.param;
        if (st.kind == LSIO_CODEC_STAGE_LZ4) p->has_lz4 = true, p->lz4_header = st.param;
        if (st.kind == LSIO_CODEC_STAGE_CRC32C) p->has_crc = true;
        p->stages.push_back(st);
    }
    if (p->has_shuffle && p->elem == 0) {
        delete p;
        return -EINVAL;
    }
    int rc = init_pipeline(p);
    if (rc) {
        free_pipeline(p);
        return rc;
    }
    *out = p;
    return 0;
}

int lsio_codec_pipeline_decode(lsio_codec_pipeline* p, const lsio_output* outputs, size_t n, void* out_dev, void* stream) {
    if (!p || (!outputs && n) || (!out_dev && n)) return -EINVAL;
    if (n > p->max_batch || ((uintptr_t)out_dev & 15)) return -EINVAL;
    if (p->inflight) return -EBUSY;
    DeviceGuard dg(p->device);
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* out = (uint8_t*)out_dev;
    uint8_t* lz_out = p->has_shuffle ? p->scratch : out;
    const uint32_t trailer = p->has_crc ? 4 : 0;
    std::vector<lsio_cuda_segment> copies;  // no data-moving stage: rows are gathered with the pack kernel
    for (size_t i = 0; i < n; i++) {
        const lsio_output& o = outputs[i];
        int32_t hs = LSIO_CODEC_OK;
        const lsio_aligned_bytes* v = o.status == 0 ? o.chunk.buffer : nullptr;
        uint64_t src = 0, len = 0;
        if (!v || v->buf->kind != MemKind::Device || v->buf->device != p->device) {
            hs = LSIO_CODEC_ERR_ITEM;
        } else {
            src = (uint64_t)(uintptr_t)(v->buf->buf + v->start);
            len = v->end - v->start;
            if (p->has_crc && (!(o.checksums & LSIO_CHECKSUM_CRC32C) || len < 4 || !lsio_codecs_crc32c_trailer_ok(o.crc32c))) hs = LSIO_CODEC_ERR_CHECKSUM;
            else if (len < trailer + (p->has_lz4 ? p->lz4_header : 0) || len - trailer > 0xFFFFFFF0ull) hs = LSIO_CODEC_ERR_INPUT;
            else if (!p->has_lz4 && len - trailer != p->nbytes) hs = LSIO_CODEC_ERR_OUTPUT;
        }
        p->host_status[i] = hs;
        const bool ok = hs == LSIO_CODEC_OK;
        if (p->has_lz4) {
            lsio_codec_item& it = p->items_lz4[i];
            it.src = ok ? src + p->lz4_header : 0;
            it.dst = (uint64_t)(uintptr_t)(lz_out + i * p->nbytes);
            it.src_len = ok ? (uint32_t)(len - trailer - p->lz4_header) : 0;  // an empty block decodes to an error status and writes nothing
            it.dst_cap = (uint32_t)p->nbytes;
        }
        if (p->has_shuffle) {
            lsio_codec_item& it = p->items_shuf[i];
            it.src = p->has_lz4 ? (uint64_t)(uintptr_t)(p->scratch + i * p->nbytes) : src;
            it.dst = (uint64_t)(uintptr_t)(out + i * p->nbytes);
            it.src_len = (p->has_lz4 || ok) ? (uint32_t)p->nbytes : 0;
            it.dst_cap = 0;
        }
        if (!p->has_lz4 && !p->has_shuffle && ok) copies.push_back(lsio_cuda_segment{src, (uint64_t)(uintptr_t)(out + i * p->nbytes), p->nbytes});
    }
    LSIO_CUDA_TRY(cudaEventRecord(p->t0, st));
    if (p->has_lz4) {
        int rc = lsio_codecs_lz4_decompress_async(p->device, p->d_items_lz4, n, p->d_results, stream);
        if (rc) return rc;
    }
    LSIO_CUDA_TRY(cudaEventRecord(p->t1, st));
    if (p->has_shuffle) {
        int rc = lsio_codecs_unshuffle_async(p->device, p->d_items_shuf, p->d_tile_map, (uint32_t)n * p->tiles_per_row, p->elem, 0, stream);
        if (rc) return rc;
    }
    if (!copies.empty()) {
        int rc = lsio_cuda_pack_strip_crc32c(p->device, copies.data(), copies.size(), LSIO_CUDA_PACK_STORE, nullptr, stream);
        if (rc) return rc;
    }
    LSIO_CUDA_TRY(cudaEventRecord(p->done, st));
    p->n_inflight = n;
    p->inflight = true;
    return 0;
}

int64_t lsio_codec_pipeline_finish(lsio_codec_pipeline* p, int32_t* status) {
    if (!p) return -EINVAL;
    if (!p->inflight) return 0;
    DeviceGuard dg(p->device);
    cudaError_t e = cudaEventSynchronize(p->done);
    p->inflight = false;
    if (e != cudaSuccess) return -(int64_t)cuda_fail(e, "codec pipeline", __FILE__, __LINE__);
    int64_t bad = 0;
    for (size_t i = 0; i < p->n_inflight; i++) {
        int32_t s = p->host_status[i];
        if (s == LSIO_CODEC_OK && p->has_lz4) {
            s = p->results[i].status;
            if (s == LSIO_CODEC_OK && p->results[i].out_len != p->nbytes) s = LSIO_CODEC_ERR_OUTPUT;
        }
        if (status) status[i] = s;
        bad += s != LSIO_CODEC_OK;
    }
    return bad;
}

int lsio_codec_pipeline_last_ms(lsio_codec_pipeline* p, float* first_stage_ms, float* rest_ms) {
    if (!p || p->inflight) return -EINVAL;
    DeviceGuard dg(p->device);
    if (first_stage_ms) LSIO_CUDA_TRY(cudaEventElapsedTime(first_stage_ms, p->t0, p->t1));
    if (rest_ms) LSIO_CUDA_TRY(cudaEventElapsedTime(rest_ms, p->t1, p->done));
    return 0;
}

void lsio_codec_pipeline_destroy(lsio_codec_pipeline* p) {
    if (!p) return;
    if (p->inflight) lsio_codec_pipeline_finish(p, nullptr);
    free_pipeline(p);
}

}  // extern "C"
