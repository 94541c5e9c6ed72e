/*
 * lsio_codecs.h -- C ABI of the codec layer: chunks that the I/O path has landed in HBM are decoded
 * in HBM.  The reference only plans this layer and holds no code for it:
 *     planned_design.md:83-90   "Abstraction level 2: Parallel compute on chunks ... These crates will all
 *                                consume the output channel from the IO layer ... lsio_codecs:
 *                                Compression / decompression"
 *     README.md:68-70           "Wrap a few decompression algorithms"
 *     planned_design.md:100-104 Zarr as the first file format
 * so the entry points below follow the published formats Zarr stores chunks in:
 *     LZ4 block format            (numcodecs.LZ4 = 4-byte little-endian decoded size + one LZ4 block)
 *     byte shuffle                (numcodecs.Shuffle / Blosc SHUFFLE, any element size)
 *     Zarr v3 `crc32c` codec      (chunk bytes + their CRC-32C as 4 little-endian bytes): verified for free
 *                                 from the checksum the pack kernel fuses into the landing, see
 *                                 LSIO_CRC32C_RESIDUE
 * Plain pointers and sizes only; `src`/`dst` are CUDA device addresses on `device`.  Return values as
 * in lsio_cuda.h: 0, a negative errno-style code, or a positive cudaError_t.  No CPU fallback.
 */
#ifndef LSIO_CODECS_H
#define LSIO_CODECS_H

#include <stddef.h>
#include <stdint.h>

#include "lsio_io.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One chunk (or independently compressed block of a chunk) to decode. */
typedef struct lsio_codec_item {
    uint64_t src;     /* device address of the encoded bytes; no alignment needed */
    uint64_t dst;     /* device address of the output; no alignment needed; must not overlap src */
    uint32_t src_len; /* encoded bytes */
    uint32_t dst_cap; /* room at dst (lz4); ignored by shuffle/unshuffle, which write exactly src_len bytes */
} lsio_codec_item;

#define LSIO_CODEC_OK 0
#define LSIO_CODEC_ERR_INPUT 1  /* encoded stream truncated or overrun */
#define LSIO_CODEC_ERR_OUTPUT 2 /* decoded bytes would not fit dst_cap */
#define LSIO_CODEC_ERR_OFFSET 3 /* LZ4 match offset 0 or reaching before the start of the output */

typedef struct lsio_codec_result {
    uint32_t out_len; /* bytes written at dst */
    int32_t status;   /* LSIO_CODEC_* */
} lsio_codec_result;

/* LZ4 block decode of n independent blocks in one launch (one warp per block; the warp parses the
 * token stream from a shared-memory window and copies literals and matches cooperatively).
 * `items` and `results` are HOST arrays; the call uploads, launches on `stream`, downloads and
 * synchronises.  A failed block sets its own status only. */
int lsio_codecs_lz4_decompress(int device, const lsio_codec_item* items, size_t n, lsio_codec_result* results, void* stream);
/* Asynchronous form for pipelines: `items_dev` / `results_dev` are device-accessible arrays
 * (device or mapped pinned memory).  Enqueues exactly one kernel launch on `stream`. */
int lsio_codecs_lz4_decompress_async(int device, const lsio_codec_item* items_dev, size_t n, lsio_codec_result* results_dev, void* stream);

/* Byte unshuffle (decode) / shuffle (encode) of n buffers of `elem_size`-byte elements:
 * unshuffle: dst[i*elem + j] = src[j*cnt + i], cnt = src_len / elem_size; the src_len % elem_size
 * trailing bytes are copied unchanged (numcodecs.Shuffle).  HBM-bound: 2 * src_len bytes per buffer.
 * Host `items`; uploads, launches one kernel on `stream`, synchronises. */
int lsio_codecs_unshuffle(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, void* stream);
int lsio_codecs_shuffle(int device, const lsio_codec_item* items, size_t n, uint32_t elem_size, void* stream);
/* Asynchronous form: `items_dev` device-accessible; `tile_map_dev` = n_tiles pairs {item index, tile index
 * within the item} of uint32, one per 16 KiB tile of every item, in launch order -- built on the host with
 * lsio_codecs_shuffle_tile_map (returns the tile count; writes at most cap_tiles pairs; tile_map may be NULL to
 * only count).  Enqueues exactly one kernel launch on `stream`; forward != 0 shuffles, 0 unshuffles. */
uint32_t lsio_codecs_shuffle_tiles(uint32_t src_len);
int64_t lsio_codecs_shuffle_tile_map(const uint32_t* src_lens, size_t n, uint32_t* tile_map, size_t cap_tiles);
int lsio_codecs_unshuffle_async(int device, const lsio_codec_item* items_dev, const uint32_t* tile_map_dev, uint32_t n_tiles, uint32_t elem_size,
                                int forward, void* stream);

/* Zarr v3 `crc32c` codec: CRC-32C over (bytes || CRC-32C(bytes) as 4 LE bytes) is this constant whatever
 * the bytes are.  A chunk stored with that codec is therefore verified by comparing the CRC the pack
 * kernel already computed over the landed range (lsio_output.crc32c) with the residue: no extra pass. */
#define LSIO_CRC32C_RESIDUE 0x48674BC7u
int lsio_codecs_crc32c_trailer_ok(uint32_t crc32c_of_chunk_with_trailer);


/* ---- the decode side of a Zarr codec chain over a whole batch of completion-channel items ----------------
 * What a `lsio_codecs` crate would expose to the layers above (planned_design.md:83-90: "These crates will all
 * consume the output channel from the IO layer"): hand it the lsio_output records of a recv_batch and get one
 * dense (n, decoded_nbytes) array in HBM, chunk i in row i.  Stages are listed in ENCODE order, as Zarr metadata
 * lists codecs (filters, compressor, checksum); decoding runs them backwards. */
#define LSIO_CODEC_STAGE_SHUFFLE 1u /* param = element size in bytes (numcodecs.Shuffle) */
#define LSIO_CODEC_STAGE_LZ4 2u     /* param = header bytes before the LZ4 block (numcodecs.LZ4: 4) */
#define LSIO_CODEC_STAGE_CRC32C 3u  /* Zarr v3 crc32c codec: verified from lsio_output.crc32c, 4 bytes dropped */
typedef struct lsio_codec_stage {
    uint32_t kind;
    uint32_t param;
} lsio_codec_stage;

typedef struct lsio_codec_pipeline lsio_codec_pipeline;

/* `decoded_nbytes` (a multiple of 16) is what every chunk decodes to; `max_batch` bounds n of one decode call.
 * Owns pinned descriptor/result arrays and, when two data-moving stages are chained, a max_batch x
 * decoded_nbytes scratch buffer in HBM.  For double buffering create two pipelines and alternate. */
int lsio_codec_pipeline_create(int device, const lsio_codec_stage* stages, size_t n_stages, uint64_t decoded_nbytes, uint32_t max_batch,
                               lsio_codec_pipeline** out);
/* Enqueues the decode of `n` items on `stream` (at most 2 kernel launches, no host synchronisation) into
 * `out_dev` (n rows of decoded_nbytes, 16-byte aligned).  Every item must be an Ok chunk on the pipeline's
 * device; the chunks must stay alive until lsio_codec_pipeline_finish returns.  One decode in flight per pipeline. */
int lsio_codec_pipeline_decode(lsio_codec_pipeline* self, const lsio_output* outputs, size_t n, void* out_dev, void* stream);
/* Waits for the decode enqueued last.  status[i] (may be NULL) receives LSIO_CODEC_* per item, with
 * LSIO_CODEC_ERR_CHECKSUM for a crc32c trailer mismatch and LSIO_CODEC_ERR_OUTPUT when a chunk decoded to a
 * size other than decoded_nbytes.  Returns the number of bad items (0 = all rows valid) or a negative/CUDA code. */
#define LSIO_CODEC_ERR_CHECKSUM 4
#define LSIO_CODEC_ERR_ITEM 5 /* the item was an Err, not a chunk, or lives on another device */
int64_t lsio_codec_pipeline_finish(lsio_codec_pipeline* self, int32_t* status);
/* Measurement aid: device time of the decode finished last, split into its first launch (LZ4 when present) and the rest. */
int lsio_codec_pipeline_last_ms(lsio_codec_pipeline* self, float* first_stage_ms, float* rest_ms);
void lsio_codec_pipeline_destroy(lsio_codec_pipeline* self);

#ifdef __cplusplus
}
#endif
#endif
