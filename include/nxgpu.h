/*
 * nxgpu.h — C-ABI of the B200 DEFLATE engine that stands in for the POWER NX-GZIP
 * accelerator underneath libnxz's unchanged zlib-compatible host code.
 *
 * Two groups of entry points:
 *
 *  (1) The drop-in boundary: the six symbols libnxz's host code links against and
 *      that lib/gzip_vas.c + lib/crc32_power.c define on POWER.  Signatures are
 *      the reference's, so the reference's objects link against libnxgpu.so
 *      without modification (INTEGRATION.md shows the link line).
 *
 *  (2) The batch extension (additive, absent from the reference): array-of-jobs
 *      calls that fill a GPU, in host-pointer and device-pointer flavours.  Per-item
 *      semantics are defined by running that item through zlib / the reference
 *      (SURVEY.md §8b).
 *
 * Plain pointers and sizes only; no CUDA or torch types appear in any signature.
 * Every call fails loudly (negative return) when no sm_100a device is usable —
 * there is no CPU fallback inside this library.
 */
#ifndef NXGPU_H
#define NXGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* (1) Drop-in boundary                                                      */
/* ------------------------------------------------------------------------- */

/* Layout-compatible prefix of `struct nx_dev_t` (reference lib/nx_zlib.h:178-194).
 * The handle is allocated by the caller (lib/nx_zlib.c:562); the engine only
 * touches `paste_addr`, `fd` and `function`, exactly as lib/gzip_vas.c:94-142 does. */
struct nxgpu_dev_prefix {
	int lock, nx_errno, socket_id, nx_id, open_cnt, use_cnt, init_total_credits;
	int creator_pid;
	void *paste_addr;   /* non-NULL == usable (checked at lib/gzip_vas.c:294)   */
	int fd;             /* here: CUDA device ordinal the handle is bound to      */
	int function;
};
#ifndef NXGPU_NO_DROPIN_DECLS
typedef struct nx_dev_t *nx_devp_t;          /* opaque; cast to nxgpu_dev_prefix */
#endif

/* The 2048-byte, 2048-aligned job descriptor `nx_gzip_crb_cpb_t`
 * (reference inc_nx/nxu.h:286-616).  Handled here as raw bytes; offsets below are
 * asserted against the reference header by oracle/Makefile's `layout_check`. */
#ifndef NXGPU_NO_DROPIN_DECLS   /* define when the reference's own headers are in scope */
struct nx_gzip_crb_cpb_t;
typedef struct nx_gzip_crb_cpb_t nx_gzip_crb_cpb_t;
#endif

enum {
	NXGPU_CRB_FC = 0,            /* BE32, function code in the low byte          */
	NXGPU_CRB_CSB_ADDR = 8,
	NXGPU_CRB_SRC_DDE = 16,      /* nx_dde_t {BE32 count(low 8b), BE32 bc, BE64 addr} */
	NXGPU_CRB_DST_DDE = 32,
	NXGPU_CRB_CSB = 240,         /* byte0 bit7 = V, byte2 = CC, byte3 = CE, +4 tpbc */
	NXGPU_CPB = 256,
	NXGPU_CPB_IN_ADLER = 256 + 0,
	NXGPU_CPB_IN_CRC = 256 + 4,
	NXGPU_CPB_IN_HISTLEN = 256 + 8,   /* bits 0:11 histlen(qw), low 3 bits in_subc */
	NXGPU_CPB_IN_SFBT = 256 + 12,     /* sfbt(4b) | rembytecnt(16b) / dhtlen(12b)  */
	NXGPU_CPB_IN_DHT = 256 + 16,
	NXGPU_CPB_OUT_ADLER = 256 + 384,
	NXGPU_CPB_OUT_CRC = 256 + 388,
	NXGPU_CPB_OUT_TEBC = 256 + 392,   /* tebc bits 13:15 of the word, subc low 16  */
	NXGPU_CPB_OUT_SFBT = 256 + 396,
	NXGPU_CPB_OUT_SPBC_COMP = 256 + 400,
	NXGPU_CPB_OUT_LZCOUNT = 256 + 400,
	NXGPU_CPB_OUT_DHT = 256 + 400,
	NXGPU_CPB_OUT_SPBC_DECOMP = 256 + 688,
	NXGPU_CPB_OUT_SPBC_COMP_WITH_COUNT = 256 + 1664,
	NXGPU_CRB_CPB_SIZE = 2048
};

#ifndef NXGPU_NO_DROPIN_DECLS
/* replaces lib/gzip_vas.c:92 */
extern uint64_t tb_freq;
/* replaces lib/gzip_vas.c:144  (function must be NX_FUNC_COMP_GZIP == 2;
 * pri = GPU ordinal, or -1 for "any").  0 on success, -1 + errno on failure. */
int nx_function_begin(int function, int pri, nx_devp_t nxhandle);
/* replaces lib/gzip_vas.c:166 */
int nx_function_end(nx_devp_t nxhandle);
/* replaces lib/gzip_vas.c:203 — wait about `ticks` 512 MHz timebase ticks */
uint64_t nx_wait_ticks(uint64_t ticks, uint64_t accumulated_ticks, int do_sleep);
/* replaces lib/gzip_vas.c:281 — synchronous: on return 0 the CSB (V=1, CC, CE,
 * tpbc) and CPB-out are filled.  Returns -EAGAIN when the device is unusable. */
int nxu_run_job(nx_gzip_crb_cpb_t *c, nx_devp_t nxhandle);
/* replaces lib/crc32_power.c:71 — raw (no pre/post inversion) reflected CRC-32
 * update; p 16-byte aligned, len a multiple of 16 (lib/crc32_ppc.c:33-67). */
unsigned int __crc32_vpmsum(unsigned int crc, const void *p, unsigned long len);
#endif

/* ------------------------------------------------------------------------- */
/* (2) Batch extension                                                       */
/* ------------------------------------------------------------------------- */

typedef struct nxgpu_ctx nxgpu_ctx;     /* one CUDA device + stream + scratch  */

enum { NXGPU_MEM_HOST = 0, NXGPU_MEM_DEVICE = 1 };
enum { NXGPU_WRAP_RAW = 0, NXGPU_WRAP_ZLIB = 1, NXGPU_WRAP_GZIP = 2, NXGPU_WRAP_AUTO = 3,
       /* deflate_stream only: raw deflate that is NOT the end of the stream — every chunk, the last
        * one too, ends on the joiner and none carries BFINAL (a GPU's range of a multi-GPU stream) */
       NXGPU_WRAP_RAW_CONT = 5 };

/* status codes: 0 and the zlib-style negatives, plus NX completion codes where
 * a caller wants them (inc_nx/nxu.h:823-857) */
enum {
	NXGPU_OK = 0,
	NXGPU_E_NODEV = -100,       /* no usable sm_100 GPU / CUDA error            */
	NXGPU_E_ARG = -2,           /* == Z_STREAM_ERROR                            */
	NXGPU_E_DATA = -3,          /* == Z_DATA_ERROR                              */
	NXGPU_E_MEM = -4,           /* == Z_MEM_ERROR                               */
	NXGPU_E_BUF = -5            /* == Z_BUF_ERROR (target too small)            */
};

/* dev = CUDA ordinal, or -1 to honour LOCAL_RANK / default 0 */
int nxgpu_open(int dev, nxgpu_ctx **out);
void nxgpu_close(nxgpu_ctx *ctx);
/* textual reason of the last failure on this thread ("" if none) */
const char *nxgpu_last_error(void);
/* device allocation helpers so non-CUDA host languages can hold device buffers */
int nxgpu_dev_alloc(nxgpu_ctx *ctx, size_t bytes, void **dptr);
int nxgpu_dev_free(nxgpu_ctx *ctx, void *dptr);
int nxgpu_memcpy_h2d(nxgpu_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int nxgpu_memcpy_d2h(nxgpu_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
int nxgpu_host_alloc(size_t bytes, void **hptr);      /* pinned */
int nxgpu_host_free(void *hptr);
int nxgpu_sync(nxgpu_ctx *ctx);
/* CUDA-event timing of whatever is enqueued between the two calls on the
 * context's stream; returns elapsed milliseconds from nxgpu_timer_stop. */
int nxgpu_timer_start(nxgpu_ctx *ctx);
int nxgpu_timer_stop(nxgpu_ctx *ctx, float *ms);
/* number of kernels this context has launched so far */
uint64_t nxgpu_launch_count(nxgpu_ctx *ctx);
/* average device time (ms) and count of launches of the dominant kernel of the
 * given family since the last reset ("deflate", "inflate", "checksum") */
int nxgpu_kernel_time(nxgpu_ctx *ctx, const char *family, double *ms_total, uint64_t *launches);
void nxgpu_kernel_time_reset(nxgpu_ctx *ctx);

/* --- checksums: replaces crc32()/adler32() of lib/nx_crc.c:437, lib/nx_adler32.c:182
 * for large buffers, and the combine of lib/nx_crc.c:374 / lib/nx_adler32.c:154. */
typedef struct {
	const void *src;
	uint64_t len;
	uint32_t crc_seed;     /* zlib convention: crc32(0,NULL,0) == 0            */
	uint32_t adler_seed;   /* zlib convention: adler32(0,NULL,0) == 1          */
} nxgpu_cksum_item;
typedef struct { uint32_t crc32, adler32; } nxgpu_cksum_result;
int nxgpu_checksum_batch(nxgpu_ctx *ctx, const nxgpu_cksum_item *items, size_t n,
			 nxgpu_cksum_result *results, int mem);
int nxgpu_crc32(nxgpu_ctx *ctx, uint32_t seed, const void *src, uint64_t len, int mem, uint32_t *out);
int nxgpu_adler32(nxgpu_ctx *ctx, uint32_t seed, const void *src, uint64_t len, int mem, uint32_t *out);
uint32_t nxgpu_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t nxgpu_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);

/* --- deflate: the NX compress function codes (inc_nx/nxu.h:803-811) as a batch.
 * Each item is one independent job: `src_len` new bytes at `src`, preceded in
 * memory by `hist_len` (<=32768) bytes usable as dictionary, exactly like the
 * in_histlen priming of lib/nx_deflate.c:828-880.  One deflate block is written
 * per item (dynamic Huffman, tables built on the GPU; stored if incompressible).
 * Unless NXGPU_F_FINAL is set the block has BFINAL=0 and is followed by the empty
 * stored block (00 00 FF FF after 3 header bits) that lib/nx_deflate.c:220-243
 * uses as joiner, so the outputs of consecutive items concatenate bytewise. */
enum {
	NXGPU_F_FINAL = 1,        /* BFINAL=1, no joiner, padded to a byte          */
	NXGPU_F_FIXED = 2,        /* fixed Huffman (Z_FIXED)                        */
	NXGPU_F_NO_JOINER = 4,    /* leave the tail bit-unaligned (nxu_run_job use) */
	NXGPU_F_NO_HEADER = 8,    /* with FIXED or a caller's table: no block header in front ...        */
	NXGPU_F_NO_EOB = 16       /* ... / no end-of-block behind: a piece of a block other items complete */
};
typedef struct {
	const void *src;
	uint32_t src_len;
	uint32_t hist_len;
	void *dst;
	uint32_t dst_cap;
	uint32_t flags;
} nxgpu_deflate_item;
typedef struct {
	int32_t rc;
	uint32_t out_len;      /* bytes written                                    */
	uint32_t tebc;         /* valid bits in the last byte (0 == 8)             */
	uint32_t crc32;        /* of the item's src_len new bytes, seed 0          */
	uint32_t adler32;      /* idem, seed 1                                     */
	uint32_t n_tokens;
} nxgpu_deflate_result;
/* level 1..9 (zlib scale; 0 -> 6 like lib/nx_deflate.c:655-658) */
int nxgpu_deflate_batch(nxgpu_ctx *ctx, const nxgpu_deflate_item *items, size_t n,
			nxgpu_deflate_result *results, int level, int mem);
/* worst-case output bytes of one item of src_len bytes */
uint32_t nxgpu_deflate_bound(uint32_t src_len);

/* Whole-buffer deflate: cuts `src` into `chunk` byte pieces (default 262144 when
 * 0), runs them as one batch with 32 KiB priming, stitches the pieces on the
 * device into one RFC 1950/1951/1952 stream and adds header/trailer.
 * `chunk_offsets` (optional, n_chunks+1 entries) receives the byte offset of each
 * piece inside dst — the sync-point index SURVEY.md §8b asks for. */
typedef struct {
	uint64_t out_len;
	uint32_t crc32, adler32;
	uint32_t n_chunks;
	uint64_t n_tokens;
} nxgpu_stream_result;
int nxgpu_deflate_stream(nxgpu_ctx *ctx, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			 int level, int wrap, uint32_t chunk, uint64_t *chunk_offsets,
			 nxgpu_stream_result *res, int mem);
uint64_t nxgpu_deflate_stream_bound(uint64_t src_len, uint32_t chunk);
/* OR into `wrap` of nxgpu_deflate_stream: chunks do not look back into each other (true
 * Z_FULL_FLUSH semantics, the resetting flush of lib/nx_deflate.c:1690-1712) — costs a little
 * ratio, makes the member seekable: any zlib still inflates it, and nxgpu_inflate_stream inflates
 * all chunks of the index in parallel. */
#define NXGPU_STREAM_INDEPENDENT 0x100
/* Inflates ONE member as the n_chunks segments of `chunk_offsets` (n_chunks + 1 entries, as returned
 * by nxgpu_deflate_stream with the same `wrap` and `chunk`) in a single batch; verifies the
 * zlib/gzip trailer against the combined per-segment checksums (lib/nx_crc.c:374,
 * lib/nx_adler32.c:154).  NXGPU_E_DATA if the segments turn out to depend on each other. */
int nxgpu_inflate_stream(nxgpu_ctx *ctx, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			 int wrap, const uint64_t *chunk_offsets, uint32_t n_chunks, uint32_t chunk,
			 nxgpu_stream_result *res, int mem);

/* --- inflate: the NX decompress function code (inc_nx/nxu.h:812) as a batch of
 * independent members / sync-point segments. */
typedef struct {
	const void *src;
	uint32_t src_len;
	void *dst;
	uint32_t dst_cap;
	uint32_t wrap;         /* NXGPU_WRAP_*                                      */
	uint32_t hist_len;     /* bytes before dst usable as window (segments)     */
} nxgpu_inflate_item;
typedef struct {
	int32_t rc;            /* 0, NXGPU_E_DATA, NXGPU_E_BUF                      */
	uint32_t out_len;
	uint32_t in_used;      /* source bytes consumed incl. header and trailer   */
	uint32_t crc32;        /* of the output                                    */
	uint32_t adler32;
	uint32_t flags;        /* bit0: final block seen; bit1: trailer verified;
	                          bit2: source ended on a block boundary, no final block (rc = NXGPU_E_DATA) */
} nxgpu_inflate_result;
/* NXGPU_MEM_HOST: when the items' targets lie back to back in one host buffer the outputs return in a few large
 * copies, so the bytes of dst[out_len .. dst_cap) of every item are UNSPECIFIED after the call (as are all dst_cap
 * bytes of an item whose rc != 0); only dst[0 .. out_len) is the result.
 * In a batch of at most 32 items the longest streams (64 KiB of source or more) are decoded by many warps each (block
 * starts found by header search, blocks decoded speculatively and then for real: DESIGN.md 4.2) as long as that shortens
 * the call: a lone long stream always, sixteen streams of a few hundred KB not.  Results are those of the one-warp-pair
 * decode.  NXGPU_INFLATE_PAR_MIN=<bytes> sends every stream of that length through it (0 = none), a developer / test
 * switch like NXGPU_INFLATE_SOLO_MAX. */
int nxgpu_inflate_batch(nxgpu_ctx *ctx, const nxgpu_inflate_item *items, size_t n,
			nxgpu_inflate_result *results, int mem);

/* --- one member deflated by several GPUs of one box (BASELINE.json configs[3]) -------------------
 * One process (or thread) per GPU opens the same named team; rank r of nranks owns a contiguous range of the
 * stream.  nxgpu_team_deflate is collective: every rank passes its range (in rank order the ranges are the whole
 * stream; every range but the last a multiple of `chunk`), the ranks deflate in parallel as raw deflate joined by
 * the empty stored block of lib/nx_deflate.c:220-243, the compressed sizes are exchanged and exclusive-scanned ON
 * THE DEVICES (through a control block in pinned shared memory — no host round trip between the deflate kernel
 * and the copy), every GPU writes its range at its global offset straight into the destination, and rank 0 folds
 * the per-range checksums (lib/nx_crc.c:374, lib/nx_adler32.c:154) into the trailer.
 *   dst_mem NXGPU_MEM_HOST:   the member is assembled in shared host memory (nxgpu_team_dst() in every rank): each
 *                             GPU writes its part over its own PCIe link.
 *   dst_mem NXGPU_MEM_DEVICE: in rank 0's device buffer (nxgpu_team_dst() on rank 0), written by the peers over
 *                             NVLink P2P (CUDA IPC).
 * The destination is overwritten by the next collective call.  `name` must be unique per team on the box. */
typedef struct nxgpu_team nxgpu_team;
typedef struct {
	uint64_t out_len;           /* the whole member: header + every range + trailer          */
	uint32_t crc32, adler32;    /* of the whole uncompressed stream                          */
	uint64_t src_len;           /* uncompressed bytes of all ranks                           */
	uint64_t my_offset, my_size;/* where this rank's range went                              */
	float device_ms;            /* this rank: deflate + exchange + copy, on its stream       */
} nxgpu_team_result;
int nxgpu_team_open(nxgpu_ctx *ctx, const char *name, int rank, int nranks, uint64_t dst_cap, int dst_mem, nxgpu_team **team);
int nxgpu_team_deflate(nxgpu_team *team, const void *src, uint64_t src_len, int level, int wrap, uint32_t chunk,
		       int src_mem, nxgpu_team_result *res);
void *nxgpu_team_dst(nxgpu_team *team);
void nxgpu_team_close(nxgpu_team *team);

/* A buffer of concatenated gzip members (multi-member .gz files; what gunzip, samples/gunzip_nx.c and the gz* layer
 * read) inflated as one batch; the members are discovered on the device (candidate headers, a dry decoding run, the
 * chain from offset 0).  *out_len receives the total (also on NXGPU_E_BUF: the capacity needed), *n_members the count.
 * Every member's CRC-32 and ISIZE are verified.  Members must be smaller than 4 GiB on both sides. */
int nxgpu_gunzip_concat(nxgpu_ctx *ctx, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			uint64_t *out_len, uint32_t *n_members, int mem);

/* --- gz* reader over the batched inflate (replaces the read side of lib/nx_gzlib.c:220-325, whose
 * __gzread pulls 10 bytes per read() and stops after the first member).  zlib's calling convention: gzread returns the
 * bytes delivered, 0 at the end of the file, -1 on error; every member of a multi-member file is inflated (one batch on
 * the device) and its CRC-32 / ISIZE verified.  ctx may be NULL (a context on the default device is opened and closed
 * with the file).  Only mode "r" is bound; writing stays with the reference's gzwrite over nxu_run_job. */
typedef struct nxgpu_gzfile nxgpu_gzfile;
nxgpu_gzfile *nxgpu_gzopen(nxgpu_ctx *ctx, const char *path, const char *mode);
nxgpu_gzfile *nxgpu_gzdopen(nxgpu_ctx *ctx, int fd, const char *mode);
int nxgpu_gzread(nxgpu_gzfile *file, void *buf, unsigned len);
int nxgpu_gzeof(nxgpu_gzfile *file);
uint32_t nxgpu_gzmembers(nxgpu_gzfile *file);
int nxgpu_gzclose(nxgpu_gzfile *file);

/* --- makedata-style synthetic text (reference samples/makedata.c:35-70):
 * byte-for-byte the stream `makedata -s seed -b log2size < seedfile` writes
 * (host-side generator; inputs for bench and tests).  Returns bytes written. */
uint64_t nxgpu_makedata(int seed, int log2size, const void *seedfile, uint64_t seedfile_len,
			void *out, uint64_t out_cap);
/* bytes [from, to) of that stream, generated without holding the stream (the generator looks back at most 64 KiB):
 * shards of the 16 GiB multi-GPU workload.  Returns bytes written. */
uint64_t nxgpu_makedata_range(int seed, int log2size, const void *seedfile, uint64_t seedfile_len,
			      uint64_t from, uint64_t to, void *out);

/* --- DHT generation: what lib/nx_dhtgen.c:945 dhtgen() computes on the host on a
 * DHT-cache miss (call site lib/nx_dht.c:632) — 286 lit/len + 30 distance counts to a length-limited
 * (<= 15 bit) dynamic Huffman table, written as the RFC 1951 3.2.7 block header from HLIT on, the
 * bytes of cpb.in_dht.  nxgpu_dhtgen has dhtgen()'s own argument list (counts in host byte order;
 * *dht_num_valid_bits: valid bits of the last byte, 0 meaning 8; cpb_header: prepend the 16-byte CPB
 * prefix holding in_dhtlen).  The batch flavour takes n x 316 counters and returns n x 288 bytes +
 * n bit lengths.  Unlike the reference's one-pass heuristic limiter the code is an exact Kraft-complete
 * Huffman code; symbols with a zero count get no code (pre-fill with fill_zero_lzcounts(…, 1), as
 * lib/nx_dht.c:627 does, for a table that is reused on other data). */
int nxgpu_dhtgen(nxgpu_ctx *ctx, const uint32_t *lhist, int num_lhist, const uint32_t *dhist, int num_dhist,
		 char *dht, int *dht_num_bytes, int *dht_num_valid_bits, int cpb_header);
int nxgpu_dhtgen_batch(nxgpu_ctx *ctx, const uint32_t *counts, size_t n, uint8_t *dht, uint32_t *dht_bits, int mem);

/* --- job coalescing inside nxu_run_job (the reference submits one CRB per
 * paste, lib/gzip_vas.c:281-417).  Descriptors submitted concurrently from different threads are
 * run as one GPU batch; this reports how that went on device `dev` since process start:
 * batches served, descriptors served, largest batch. */
void nxgpu_job_stats(int dev, uint64_t *batches, uint64_t *jobs, uint64_t *max_batch);

#ifdef __cplusplus
}
#endif
#endif /* NXGPU_H */
