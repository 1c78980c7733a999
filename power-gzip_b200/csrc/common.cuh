// common.cuh — shared device/host definitions of the B200 DEFLATE engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>

namespace nxgpu {

constexpr int kNumSMs = 148;             // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kWindow = 32768;           // RFC 1951 maximum match distance
constexpr int kMinMatch = 5;             // shortest match the LZ77 stage emits: the chains hash 5 bytes
constexpr int kMaxMatch = 258;

// LZ77 token, one u32: literal = byte value; match = 1<<31 | (len-3)<<15 | (dist-1)
__host__ __device__ inline uint32_t tok_match(uint32_t len, uint32_t dist) { return 0x80000000u | ((len - 3) << 15) | (dist - 1); }
__host__ __device__ inline bool tok_is_match(uint32_t t) { return (t >> 31) != 0; }
__host__ __device__ inline uint32_t tok_len(uint32_t t) { return ((t >> 15) & 0xff) + 3; }
__host__ __device__ inline uint32_t tok_dist(uint32_t t) { return (t & 0x7fff) + 1; }

// length (3..258) -> code index 0..28 plus extra bits (RFC 1951 §3.2.5)
__host__ __device__ inline void len_code(uint32_t len, uint32_t &code, uint32_t &nextra, uint32_t &extra)
{
	uint32_t l = len - 3;
	if (l < 8) { code = l; nextra = 0; extra = 0; return; }
	if (l == 255) { code = 28; nextra = 0; extra = 0; return; }
#ifdef __CUDA_ARCH__
	uint32_t nb = 31 - __clz(l);
#else
	uint32_t nb = 31 - __builtin_clz(l);
#endif
	code = 4 * (nb - 1) + ((l >> (nb - 2)) & 3);
	nextra = nb - 2;
	extra = l & ((1u << nextra) - 1);
}
// distance (1..32768) -> code index 0..29 plus extra bits
__host__ __device__ inline void dist_code(uint32_t dist, uint32_t &code, uint32_t &nextra, uint32_t &extra)
{
	uint32_t d = dist - 1;
	if (d < 4) { code = d; nextra = 0; extra = 0; return; }
#ifdef __CUDA_ARCH__
	uint32_t nb = 31 - __clz(d);
#else
	uint32_t nb = 31 - __builtin_clz(d);
#endif
	code = 2 * nb + ((d >> (nb - 1)) & 1);
	nextra = nb - 1;
	extra = d & ((1u << nextra) - 1);
}

// ---- device job descriptors (host fills, kernels read) ----
struct DeflateJob {
	const uint8_t *src;     // first NEW byte; src[-hist_len .. -1] is dictionary
	uint32_t src_len;
	uint32_t hist_len;
	uint8_t *out;           // 16-byte aligned private output slot
	uint32_t out_cap;
	uint32_t flags;         // NXGPU_F_*
	const uint8_t *dht;     // optional caller-supplied dynamic header (bits from HLIT), nxu_run_job path
	uint32_t dht_bits;
	uint32_t *lzcount;      // optional 316 counters out (COUNT function codes)
	uint32_t pad_;
};
struct DeflateOut {
	int32_t rc;
	uint32_t out_len;
	uint32_t tebc;
	uint32_t n_tokens;
	uint32_t btype;         // 0 stored, 1 fixed, 2 dynamic
	uint32_t reserved[3];
};

constexpr uint32_t kWrapDry = 0x100;     // InflateJob::wrap flag: decode and count only, write nothing (member discovery)
constexpr uint32_t kWrapJob = 4;         // InflateJob::wrap: NX decompress job semantics (nxu_run_job), raw deflate
constexpr uint32_t kWrapNoHeader = 0x200; // InflateJob::wrap flag: the container header is behind us, src/start_bit point at a block header
constexpr uint32_t kWrapSkip = 0x400;    // InflateJob::wrap flag: not this launch's business (the parallel path runs it), write nothing
constexpr uint32_t kInflateMapStop = 8;  // InflateOut::flags: stopped at a block boundary that InflateJob::stop_map marks
constexpr int32_t kInflateRetry = -1000; // InflateOut::rc (internal): the parallel path disagreed with itself, run the job serially
struct InflateJob {
	const uint8_t *src;
	uint32_t src_len;
	uint32_t wrap;
	uint8_t *dst;           // dst[-hist_len..-1] window
	uint32_t dst_cap;
	uint32_t hist_len;
	// --- NX job mode only (wrap == kWrapJob): resume state of inc_nx/nxu.h:296-400 ---
	const uint8_t *dht;     // in_dht bits (from HLIT) when sfbt is 110x
	uint8_t *out_dht;       // 288 bytes: dynamic header of the block the job stopped in
	uint32_t dht_bits;
	uint32_t start_bit;     // bits of src[0] already consumed (0..7)
	uint32_t sfbt;          // 0xxx fresh block header; 100x stored, 101x fixed, 110x dynamic; bit0 = BFINAL
	uint32_t rembytecnt;    // stored bytes still to copy (sfbt 100x)
	uint32_t single_block;  // FC 0x12 / 0x16: suspend at the end of the first block that completes
	uint32_t pad_;
	// --- one stream decoded by many warps (inflate_par.cuh) ---
	const uint32_t *stop_map; // optional: one bit per source bit; a BFINAL=0 block that ends on a set bit ends the job (flags |= kInflateMapStop)
	uint64_t map_bit0;        // map index of bit 0 of src[0]
	const uint8_t *hist_ptr;  // optional: the 32 KiB in front of dst live here (last hist_len bytes valid) instead of at dst[-hist_len..]
};
struct InflateOut {
	int32_t rc;             // job mode: NX completion code (0/3 ok, 13 target full, 66/67/68 data)
	uint32_t out_len;
	uint32_t in_used;       // job mode: source bytes the engine READ (SPBC minus history, manual §2.4)
	uint32_t flags;
	uint32_t trailer_crc;   // from the stream (gzip) / adler (zlib)
	uint32_t trailer_isize;
	// --- NX job mode ---
	uint32_t sfbt;          // manual Table 5-3 / 6-4
	uint32_t subc;          // source bits read but not processed (16-bit field: Table 5-3 bounds it by 2285)
	uint32_t rembytecnt;
	uint32_t dhtlen;        // valid bits in out_dht
	uint32_t end_bit_lo, end_bit_hi;   // kInflateMapStop: map index of the boundary
};

struct CksumJob {
	const uint8_t *src;
	uint64_t len;
};

// deflate tuning per zlib-style level
struct LevelParams { int depth; int lazy; int nice; int d1; };   // d1 != 0: two-pass parse, d1 = depth of the shallow pass;
                                                                 // there `depth` is the base the deep pass scales with the data (deflate.cu)
__host__ __device__ inline LevelParams level_params(int level)
{
	switch (level) {          // chain depth, lazy threshold (0 = greedy), nice length, shallow-pass depth (0 = single pass)
	case 1: return { 1, 0, 32, 0 };          // one candidate per position: still 16 % smaller than zlib -1 on the benchmark text
	case 2: return { 2, 0, 64, 0 };
	case 3: return { 4, 0, 128, 0 };
	case 4: return { 4, 16, 128, 0 };
	case 5: return { 12, 32, 258, 2 };
	case 6: return { 32, 16, 258, 2 };
	case 7: return { 32, 64, 258, 3 };
	case 8: return { 48, 258, 258, 3 };
	case 9: return { 96, 258, 258, 4 };
	default: return { 32, 16, 258, 2 };
	}
}

#define NXGPU_CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { nxgpu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return NXGPU_E_NODEV; } } while (0)
void set_error(const char *fmt, ...);

// Stitching fused into the deflate kernel (nxgpu_deflate_stream): a chunk's CTA learns where its bytes go from
// its predecessor (chain[k] = end offset of chunk k + 1, 0 = not known yet), publishes its own end and copies its
// slot there itself — to device memory or straight into pinned host memory, so the device-to-host transfer of
// the compressed stream overlaps the compression of later chunks.
struct StreamOut {
	uint8_t *dst = nullptr;               // final stream (device pointer, may alias pinned host memory); nullptr = off
	uint64_t cap = 0;
	uint64_t base = 0;                    // bytes of container header in front of chunk 0
	uint64_t *offsets = nullptr;          // n + 1 entries: start of every chunk, end of the last
	unsigned long long *chain = nullptr;  // n entries, zeroed before the launch
};

// One-time setup that CUDA keeps PER DEVICE (cudaFuncSetAttribute, __device__ / __constant__ symbol uploads): a process may
// open contexts on several GPUs (nx_function_begin maps pri % ndev), so "done" is tracked per device ordinal, under a lock.
struct PerDeviceOnce {
	std::mutex m;
	bool done[64] = {};
	template <class F> cudaError_t run(F f)
	{
		int dev = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess)
			return e;
		std::lock_guard<std::mutex> g(m);
		if (dev < 0 || dev >= 64)
			return f();
		if (done[dev])
			return cudaSuccess;
		e = f();
		if (e == cudaSuccess)
			done[dev] = true;
		return e;
	}
};

// kernel launchers (defined in the .cu files)
size_t deflate_smem_bytes();
size_t deflate_scratch_words(uint32_t tok_stride);   // per-CTA token scratch (u32 words)
cudaError_t launch_deflate(const DeflateJob *jobs, DeflateOut *outs, uint32_t n_jobs, int level,
			   uint32_t *tok_scratch, uint32_t tok_stride, int grid, cudaStream_t s,
			   uint32_t *job_counter, const uint32_t *ready, uint32_t jobs_per_flag, const StreamOut *so = nullptr);
cudaError_t launch_dhtgen(const uint32_t *counts, uint32_t n, uint8_t *dht_out, uint32_t *dht_bits, cudaStream_t s);
cudaError_t launch_gzip_candidates(const uint8_t *src, uint64_t len, uint64_t *cand, uint32_t max_cand, uint32_t *count, cudaStream_t s);
cudaError_t launch_inflate(const InflateJob *jobs, InflateOut *outs, uint32_t n_jobs, uint32_t *counter, cudaStream_t s);
// one stream decoded by many warps (inflate_par.cuh): block-start candidates, speculative decode of every candidate with
// markers for the unknown window, chain + window resolution, the real decode of every chained piece
struct SpecOut { uint64_t end_bit; uint32_t out_len; uint32_t status; uint32_t max_back; uint32_t pad_; };
struct ChainMeta { uint64_t bit; uint64_t exp_end; uint32_t unit; uint32_t out_off; uint32_t exp_len; uint32_t last; };
struct ParPlan {
	InflateJob job;              // the descriptor as the caller built it (device pointers)
	uint32_t *map;               // one bit per source bit: a dynamic block header that decodes to complete codes starts here
	const uint64_t *cands;       // the set bits, ascending
	uint32_t n_cand;
	uint32_t pad_;
	uint16_t *rings;             // per candidate: the last 32 Ki symbols of its speculative output, markers for what lies in front of it
	SpecOut *spec;               // per candidate
	InflateJob *cjobs;           // the chain: descriptors of the pieces, in stream order
	InflateOut *couts;
	ChainMeta *meta;
	uint8_t *hists;              // per chain piece: the 32 KiB in front of it
	uint32_t *n_chain;
	InflateOut *head_out;        // piece 0 (from the descriptor's own start state to the first candidate boundary)
	InflateOut *final_out;       // where the caller expects the result
	InflateJob *retry_job;       // written by the last kernel: the descriptor again if the pieces disagreed, else a skip
};
cudaError_t launch_blockfind(const uint8_t *src, uint32_t src_len, uint64_t first_bit, uint32_t *map, uint64_t *surv, uint32_t surv_cap,
			     uint64_t *cand, uint32_t cand_cap, uint32_t *counts, cudaStream_t s);
cudaError_t launch_inflate_par(const ParPlan &plan, uint32_t *counter, cudaStream_t s);
// checksum.cu
cudaError_t checksum_init_tables();
size_t checksum_range_bytes();
size_t checksum_partial_bytes();
void checksum_fill_range(void *ranges, size_t idx, const void *src, uint64_t len, uint64_t after, uint32_t job);
cudaError_t launch_checksum_ranges(const void *d_ranges, uint32_t n_ranges, void *d_parts, int which, cudaStream_t s);
cudaError_t launch_checksum_combine(const void *d_ranges, const void *d_parts, const uint32_t *d_rs, uint32_t n_jobs,
				    const uint32_t *d_crc_seed, const uint32_t *d_adler_seed,
				    uint32_t *d_crc_out, uint32_t *d_adler_out, cudaStream_t s, uint32_t max_ranges_per_job = 1);
cudaError_t launch_ranges_from_inflate(const InflateJob *jobs, const InflateOut *outs, uint32_t n, void *d_ranges, uint32_t *d_rs, cudaStream_t s, uint32_t per_job = 1);
uint32_t host_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
// stitch.cu
struct BitPiece { const uint8_t *src; uint64_t nbits; uint64_t dst_bit; };
struct BitGroup { uint8_t *dst; uint64_t total_bits; uint32_t first_piece; uint32_t n_pieces; };
cudaError_t launch_bitconcat(const BitPiece *pieces, const BitGroup *groups, uint32_t n_groups, cudaStream_t s);
cudaError_t launch_scan_offsets(const DeflateOut *outs, uint32_t n, uint64_t base, uint64_t *offsets, cudaStream_t s);
cudaError_t launch_gather(const DeflateJob *jobs, const DeflateOut *outs, const uint64_t *offsets, uint32_t n,
			  uint8_t *dst, uint64_t dst_cap, cudaStream_t s);
cudaError_t launch_finish_stream(const DeflateOut *outs, const uint64_t *offsets, uint32_t n, uint8_t *dst, uint64_t dst_cap,
				 int wrap, const uint32_t *d_crc, const uint32_t *d_adler, uint64_t src_len, uint64_t *d_total, cudaStream_t s);

} // namespace nxgpu
