// deflate.cu — the NX compress function (inc_nx/nxu.h:803-811) and the
// dynamic-Huffman-table generation of lib/nx_dhtgen.c:945 (row a6) as ONE persistent sm_100a
// kernel: one CTA per SM, each CTA compresses one job (chunk) at a time entirely out of
// shared memory.
//
//   Stage    warp 0 copies the chunk into a 64 KiB shared-memory ring with TMA bulk copies
//            (cp.async.bulk + mbarrier, 4 KiB blocks) and threads 5-byte-hash chains through it
//            (14-bit table of 16-bit heads, u16 distance links), non-atomically and pipelined.
//   LZ77     warps 1-31 take 512-byte sub-blocks in order, sleep on the mbarrier of the block
//            that completes their chains, and parse one lane per position in windows of 32:
//            single pass with window skipping (levels 1-4) or shallow pass + deep pass over the
//            queued token starts (levels 5-9).  No CTA-wide barrier inside this stage.
//   Stitch   matches that run over a sub-block end are cut back to a token start of the next
//            sub-block; token counts are prefix-summed and the tokens copied to a flat stream
//            while the lit/len + dist histograms accumulate in shared memory.
//   Huffman  code lengths from a rank sort, a two-queue merge and a Kraft-sum length limiter;
//            the RFC 1951 §3.2.7 header uses a real code-length code.
//   Pack     code widths are prefix-summed across the CTA (warp shuffles) and the codes are
//            OR-ed into a shared-memory staging window that is flushed with coalesced stores.
//
// Tensor cores are not used: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {

namespace {

constexpr int kThreads = 1024;
constexpr uint32_t kRingMask = 0xFFFFu;   // 64 KiB data ring / 64 Ki-entry chain ring
constexpr int kHashBits = 14;
constexpr int kHashSize = 1 << kHashBits;
constexpr uint32_t kNone = 0xFFFFFFFFu;    // empty hash head
constexpr int kStageWords = 2048;
constexpr uint32_t kSub = 512;            // bytes per sub-block (one parser warp at a time)
constexpr uint32_t kBlk = 4096;           // bytes per TMA staging block
constexpr int kSlots = 4;                 // staging mbarriers (block b uses slot b % kSlots)
constexpr int kDoneSlots = 16;            // "chains of block b are built" mbarriers (slot b % kDoneSlots)
constexpr int kMirrorWords = 80;          // ring bytes [0, 320) are mirrored behind the ring

struct __align__(16) Smem {
	uint32_t ring32[16384 + kMirrorWords];   // 64 KiB input ring (position & 0xFFFF) + copy of its first bytes,
	                                  // so that reads of up to 320 bytes never have to wrap
	uint16_t prev[65536];             // 128 KiB: distance to the previous position with the same hash
	uint16_t head[kHashSize];         // 32 KiB: low 16 bits of the most recent position per hash
	uint64_t mbar[kSlots];            // TMA completion barriers of the staging blocks
	uint64_t done_bar[kDoneSlots];    // chain builder arrives once block b's chains are built; parsers sleep on it
	uint64_t hash_bar[kDoneSlots];    // split producer: warp 0 arrives once block b's hashes sit in prev[]; warp 1 turns them into links
	volatile uint32_t parse_pos[32];  // per parser warp: first position of the sub-block in flight
	uint32_t next_sub;                // next sub-block to hand out
	uint32_t ll_freq[288];
	uint32_t d_freq[32];
	uint32_t n_tok;
	uint32_t misc[8];
};
static_assert(sizeof(Smem) <= 232448, "shared memory budget");

// Huffman / pack scratch overlays regions that are dead once the LZ77 stage of a job is over.
struct HuffScratch {                      // lives in Smem::prev
	uint32_t w[640];
	uint16_t parent[640];
	uint16_t sorted_sym[320];
	uint32_t bl_count[16];
	uint32_t next_code[16];
	uint8_t ll_len[288];
	uint8_t d_len[32];
	uint8_t cl_len[20];
	uint16_t ll_code[288];
	uint16_t d_code[32];
	uint16_t cl_code[20];
	uint32_t cl_freq[20];
	uint16_t cl_sym[320];             // symbol | extra<<8
	uint32_t n_cl_sym;
	int nl, tot;
	uint32_t hdr_words[96];           // block header bits (<= 3 + 2283 bits)
	uint32_t hdr_bits;
	uint32_t warp_sums[32];
	uint32_t batch_total;
};
static_assert(sizeof(HuffScratch) <= 65536 * 2, "huff scratch");

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// optional cycle accounting per CTA (NXGPU_DEBUG_CYCLES=1): [0] chain build, [1] producer waits for data,
// [2] producer waits for ring space, [3] parser busy (warp 1), [4] parser waits for chains (warp 1),
// [5] huffman+pack, [6] whole job, [7] windows (warp 1)
__device__ unsigned long long *g_dbg = nullptr;
#define DBG_ADD(slot, v) do { if (g_dbg && lane_id() == 0) atomicAdd(&g_dbg[blockIdx.x * 8 + (slot)], (unsigned long long)(v)); } while (0)

// The ring is addressed by byte offset (position & 0xFFFF); thanks to the mirror behind it a read
// may run up to 320 bytes past the offset without masking.
__device__ __forceinline__ uint32_t ld32(const uint8_t *ring8, uint32_t aligned_off)
{
	return *reinterpret_cast<const uint32_t *>(ring8 + aligned_off);
}
__device__ __forceinline__ uint32_t load4(const uint8_t *ring8, uint32_t pos)
{
	const uint32_t o = pos & kRingMask, a = o & ~3u;
	return __funnelshift_r(ld32(ring8, a), ld32(ring8, a + 4), (o & 3) * 8);
}
// hash of the 5 bytes at pos (the chains only ever propose matches of 5 or more bytes)
__device__ __forceinline__ uint32_t hash5(const uint8_t *ring8, uint32_t pos)
{
	const uint32_t o = pos & kRingMask, a = o & ~3u;
	const uint32_t w0 = ld32(ring8, a), w1 = ld32(ring8, a + 4), w2 = ld32(ring8, a + 8);
	const uint32_t sh = (o & 3) * 8;
	const uint32_t lo = __funnelshift_r(w0, w1, sh);
	const uint32_t b4 = __funnelshift_r(w1, w2, sh) & 0xffu;
	return ((lo * 0x9E3779B1u) ^ (b4 * 0x85EBCA6Bu + (lo >> 15))) * 0x2545F491u >> (32 - kHashBits);
}

// ---- mbarrier + TMA bulk copy (global -> shared) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
		     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
// Waiting warps must not poll at full rate: the SM is issue-bound and a spinning warp takes an issue
// slot every few cycles (ncu: a third of all instructions were wait loops, with try_wait's suspend
// hint too).  Back off with real sleeps instead; short sleeps (<~0.5 us) return immediately.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
	uint32_t ns = 512;
	while (!mbar_try_wait(bar, parity)) {
		__nanosleep(ns);
		if (ns < 4096)
			ns *= 2;
	}
}
// L2 policy for data that streams through once (the input, the finished stream): evict first, so that the per-CTA
// scratch (tokens, per-position results: written and re-read within one chunk, 148 x ~0.3 MB) stays resident in the
// 126 MB L2 instead of being written back to DRAM behind every chunk (round 1: 2.1 x the algorithmic DRAM traffic)
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
	return pol;
}
// Scratch that is dead (the tokens once they are packed, the private output slot once it is copied into the stream) would
// still be written back to DRAM when the L2 evicts its dirty lines — 1.0 GB of the 2.3 GB a 1 GiB launch moved.
// discard.global.L2 drops the lines instead.  base must be 128-byte aligned and the whole lines private to this CTA.
__device__ __forceinline__ void l2_discard_lines(const void *base, uint32_t bytes)
{
	const char *p = static_cast<const char *>(base);
	for (uint32_t o = threadIdx.x * 128u; o < bytes; o += kThreads * 128u)
		asm volatile("discard.global.L2 [%0], 128;\n" ::"l"(p + o) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
		     ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_v4_evict_first(uint4 *p, uint4 v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// ---- warp 0: stage the input into the ring with TMA and thread the hash chains through it ----
// Position space starts at gbase (the 16-byte aligned address at or below the first dictionary
// byte); block b covers positions [b*kBlk, (b+1)*kBlk).  Staging block b overwrites the ring slots
// of positions 64 KiB earlier, so it is only issued once every parser is past them (parse_pos).
//
// Chain insert: every lane reads the bucket's head (the previous position with this hash) and
// then stores its own position, one warp instruction for 32 consecutive positions.  Loads and
// stores of one warp reach shared memory in program order, so nothing on the critical path waits
// for a load result: the links (position - previous) are written a few iterations later.  Lanes of
// ONE instruction that share a bucket all link to the older head and one of them becomes the new
// head: a chain can skip same-hash positions less than 32 bytes back, which costs well under 1 %
// of ratio (tools/lzsim.c, racy=32) and removes every atomic from the build.
__device__ void build_chains(Smem &S, uint32_t lo, uint32_t hi)
{
	const uint32_t lane = lane_id();
	const uint8_t *ring8 = reinterpret_cast<const uint8_t *>(S.ring32);
	constexpr int U = 4;
	for (uint32_t p0 = lo; p0 < hi; p0 += 32 * U) {
		uint32_t h[U], old[U];
#pragma unroll
		for (int u = 0; u < U; u++)
			h[u] = hash5(ring8, p0 + 32 * u + lane);
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			old[u] = 0;
			if (pos < hi) {
				old[u] = S.head[h[u]];
				S.head[h[u]] = (uint16_t)pos;
			}
			__syncwarp();
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			// heads keep 16 bits: the distance is exact for anything inside the window; an empty or stale
			// bucket yields some in-window position that the parser's byte comparison rejects
			const uint32_t d = (pos - old[u]) & 0xFFFFu;
			if (pos < hi)
				S.prev[pos & kRingMask] = (uint16_t)(d <= (uint32_t)kWindow ? d : 0);
		}
	}
}

// The chain build as a two-warp pipeline (level 1 is bound by it: 3.7 cycles per position on one warp): warp 0 hashes
// the staged positions and parks the 14-bit hash in the position's own prev[] slot, which nobody reads before the link
// is written; warp 1 (another scheduler) reads it back and does the head exchange + link.  Same links as build_chains.
__device__ void hash_range(Smem &S, uint32_t lo, uint32_t hi)
{
	const uint32_t lane = lane_id();
	const uint8_t *ring8 = reinterpret_cast<const uint8_t *>(S.ring32);
	constexpr int U = 4;
	for (uint32_t p0 = lo; p0 < hi; p0 += 32 * U) {
		uint32_t h[U];
#pragma unroll
		for (int u = 0; u < U; u++)
			h[u] = hash5(ring8, p0 + 32 * u + lane);
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			if (pos < hi)
				S.prev[pos & kRingMask] = (uint16_t)h[u];
		}
	}
}
__device__ void insert_range(Smem &S, uint32_t lo, uint32_t hi)
{
	const uint32_t lane = lane_id();
	constexpr int U = 4;
	for (uint32_t p0 = lo; p0 < hi; p0 += 32 * U) {
		uint32_t h[U], old[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			h[u] = pos < hi ? S.prev[pos & kRingMask] : 0;
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			old[u] = 0;
			if (pos < hi) {
				old[u] = S.head[h[u]];
				S.head[h[u]] = (uint16_t)pos;
			}
			__syncwarp();
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t pos = p0 + 32 * u + lane;
			const uint32_t d = (pos - old[u]) & 0xFFFFu;
			if (pos < hi)
				S.prev[pos & kRingMask] = (uint16_t)(d <= (uint32_t)kWindow ? d : 0);
		}
	}
}

// warp 1 of the split producer: block by block behind the hashing warp
__device__ void inserter(Smem &S, uint32_t P0, uint32_t PE)
{
	const uint32_t PEa = (PE + 15) & ~15u;
	const uint32_t hash_hi = PE >= 4 ? PE - 4 : 0;
	const uint32_t nblk = (PEa + kBlk - 1) / kBlk;
	uint32_t BF = P0;
	for (uint32_t b = 0; b < nblk; b++) {
		const uint32_t parity = (b / kDoneSlots) & 1;
		while (!mbar_try_wait(&S.hash_bar[b % kDoneSlots], parity))
			__nanosleep(64);
		__threadfence_block();
		const uint32_t blk_hi = min((b + 1) * kBlk, PEa);
		const uint32_t hi = (b + 1 == nblk) ? hash_hi : min(blk_hi - 4, hash_hi);
		long long t1 = clock64();
		if (hi > BF) {
			insert_range(S, BF, hi);
			BF = hi;
		}
		DBG_ADD(0, clock64() - t1);
		__threadfence_block();
		__syncwarp();
		if (lane_id() == 0)
			mbar_arrive(&S.done_bar[b % kDoneSlots]);
	}
}

__device__ void producer(Smem &S, const uint8_t *gbase, uint32_t P0, uint32_t PE, bool split)
{
	const uint32_t lane = lane_id();
	const uint32_t PEa = (PE + 15) & ~15u;
	const uint32_t hash_hi = PE >= 4 ? PE - 4 : 0;           // positions with 5 bytes available
	const uint32_t nblk = (PEa + kBlk - 1) / kBlk;
	uint32_t issued = 0, BF = P0;
	const uint64_t pol = l2_evict_first_policy();
	for (uint32_t b = 0; b < nblk; b++) {
		// keep up to two blocks in flight beyond b; only the block we need next may block on ring space
		while (issued < nblk && issued <= b + 2) {
			const uint32_t need = (issued + 1) * kBlk;
			long long t0 = clock64();
			bool ok;
			for (;;) {
				const uint32_t mp = __reduce_min_sync(0xffffffffu, S.parse_pos[lane]);
				ok = need <= mp + 32768u || mp == kNone;
				if (ok || issued > b)
					break;
				__nanosleep(500);
			}
			if (issued == b)
				DBG_ADD(2, clock64() - t0);
			if (!ok)
				break;
			if (lane == 0) {
				const uint32_t lo = issued * kBlk;
				const uint32_t bytes = min(kBlk, PEa - lo);
				uint64_t *bar = &S.mbar[issued % kSlots];
				mbar_expect_tx(bar, bytes);
				tma_load_1d(reinterpret_cast<uint8_t *>(S.ring32) + (lo & kRingMask), gbase + lo, bytes, bar, pol);
			}
			issued++;
		}
		{
			long long t0 = clock64();
			const uint32_t parity = (b / kSlots) & 1;
			while (!mbar_try_wait(&S.mbar[b % kSlots], parity))
				;
			DBG_ADD(1, clock64() - t0);
		}
		if (((b * kBlk) & kRingMask) == 0) {
			for (int i = lane; i < kMirrorWords; i += 32)
				S.ring32[16384 + i] = S.ring32[i];
			__syncwarp();
		}
		// positions of block b whose 5 bytes are staged: the last 4 wait for block b+1
		const uint32_t blk_hi = min((b + 1) * kBlk, PEa);
		const uint32_t hi = (b + 1 == nblk) ? hash_hi : min(blk_hi - 4, hash_hi);
		long long t1 = clock64();
		if (hi > BF) {
			if (split)
				hash_range(S, BF, hi);
			else
				build_chains(S, BF, hi);
			BF = hi;
		}
		if (!split)
			DBG_ADD(0, clock64() - t1);
		// block b done: chains exist for every position below (b+1)*kBlk - 4 (everything, for the last block) —
		// or, split, their hashes are parked for warp 1
		__threadfence_block();
		__syncwarp();
		if (lane == 0)
			mbar_arrive(split ? &S.hash_bar[b % kDoneSlots] : &S.done_bar[b % kDoneSlots]);
	}
}

// length of the common prefix of p and q, at most maxl.  P0..P3 are the first 16 bytes at p
// (the caller keeps them in registers for the whole window); 16 bytes per step after that.
__device__ __forceinline__ uint32_t first_diff(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3)
{
	if (x0) return (uint32_t)(__ffs(x0) - 1) >> 3;
	if (x1) return 4 + ((uint32_t)(__ffs(x1) - 1) >> 3);
	if (x2) return 8 + ((uint32_t)(__ffs(x2) - 1) >> 3);
	return 12 + ((uint32_t)(__ffs(x3) - 1) >> 3);
}
__device__ __forceinline__ uint32_t match_length(const uint8_t *ring8, uint32_t p, uint32_t q, uint32_t maxl,
						 uint32_t P0, uint32_t P1, uint32_t P2, uint32_t P3)
{
	const uint32_t qo = q & kRingMask, qa = qo & ~3u, qs = (qo & 3) * 8;
	const uint32_t w0 = ld32(ring8, qa), w1 = ld32(ring8, qa + 4), w2 = ld32(ring8, qa + 8), w3 = ld32(ring8, qa + 12);
	uint32_t qw0 = ld32(ring8, qa + 16);
	uint32_t x0 = P0 ^ __funnelshift_r(w0, w1, qs), x1 = P1 ^ __funnelshift_r(w1, w2, qs);
	uint32_t x2 = P2 ^ __funnelshift_r(w2, w3, qs), x3 = P3 ^ __funnelshift_r(w3, qw0, qs);
	uint32_t l;
	if (x0 | x1 | x2 | x3) {
		l = first_diff(x0, x1, x2, x3);
	} else {
		const uint32_t po = p & kRingMask, ps = (po & 3) * 8;
		uint32_t pa = (po & ~3u) + 16, qb = qa + 16;
		uint32_t pw0 = ld32(ring8, pa);
		l = 16;
		while (l < maxl) {
			const uint32_t pw1 = ld32(ring8, pa + 4), pw2 = ld32(ring8, pa + 8), pw3 = ld32(ring8, pa + 12), pw4 = ld32(ring8, pa + 16);
			const uint32_t qw1 = ld32(ring8, qb + 4), qw2 = ld32(ring8, qb + 8), qw3 = ld32(ring8, qb + 12), qw4 = ld32(ring8, qb + 16);
			x0 = __funnelshift_r(pw0, pw1, ps) ^ __funnelshift_r(qw0, qw1, qs);
			x1 = __funnelshift_r(pw1, pw2, ps) ^ __funnelshift_r(qw1, qw2, qs);
			x2 = __funnelshift_r(pw2, pw3, ps) ^ __funnelshift_r(qw2, qw3, qs);
			x3 = __funnelshift_r(pw3, pw4, ps) ^ __funnelshift_r(qw3, qw4, qs);
			if (x0 | x1 | x2 | x3) {
				l += first_diff(x0, x1, x2, x3);
				break;
			}
			l += 16; pa += 16; qb += 16; pw0 = pw4; qw0 = qw4;
		}
	}
	return l < maxl ? l : maxl;
}

// ---- parser warps: one sub-block [sub_lo, sub_hi) at a time, front to back ----
// The warp looks at a window of 32 consecutive positions (one lane each).  Every lane walks the
// hash chain of its position, nearest candidate first, keeping the longest match; a candidate is
// only compared in full when the 4 bytes that would make it longer than the best so far agree.
// As soon as the lane at the head of the token stream has its final answer the token is emitted
// and the head jumps over the match; once it leaves the window the next window starts where it
// landed, so positions covered by a match are never searched.  Tokens START inside the sub-block;
// the last match may run past its end (end_pos), and the stitch pass trims it to a token boundary
// of the next sub-block.
__device__ uint32_t parse_subblock(Smem &S, uint32_t sub_lo, uint32_t sub_hi, uint32_t valid_lo, uint32_t PE,
				   int depth, int nice, int lazy, uint32_t *tk, uint32_t &nwin, uint32_t &end_pos)
{
	const uint32_t lane = lane_id();
	const uint32_t lt = (1u << lane) - 1;
	const uint8_t *ring8 = reinterpret_cast<const uint8_t *>(S.ring32);
	uint32_t n = 0;
	uint32_t p = sub_lo;
	// Chain depth follows the data: a window costs about the same however far it advances, so the
	// budget of hops per byte stays level when the depth grows with the bytes the recent windows
	// advanced (long matches = many equally good candidates = the case where deeper chains pay).
	const int base_depth = depth;
	uint32_t adv = 32;                                        // running average of the advance per window
	while (p < sub_hi) {
		nwin++;
		depth = min(4 * base_depth, max(base_depth, (int)(base_depth * adv) >> 5));
		const uint32_t pos = p + lane;
		const uint32_t nlive = min(32u, sub_hi - p);
		const bool live = lane < nlive;
		const uint32_t maxl = live ? min((uint32_t)kMaxMatch, PE - pos) : 0;
		const uint32_t maxdist = min((uint32_t)kWindow, pos - valid_lo);
		uint32_t P0, P1, P2, P3;
		{
			const uint32_t o = pos & kRingMask, a = o & ~3u, sh = (o & 3) * 8;
			const uint32_t w0 = ld32(ring8, a), w1 = ld32(ring8, a + 4), w2 = ld32(ring8, a + 8), w3 = ld32(ring8, a + 12), w4 = ld32(ring8, a + 16);
			P0 = __funnelshift_r(w0, w1, sh); P1 = __funnelshift_r(w1, w2, sh);
			P2 = __funnelshift_r(w2, w3, sh); P3 = __funnelshift_r(w3, w4, sh);
		}
		const uint32_t myb = P0 & 0xffu;
		uint32_t bl = kMinMatch - 1, bd = 0, acc = 0;
		uint32_t d = (maxl >= (uint32_t)kMinMatch) ? S.prev[pos & kRingMask] : 0;
		uint32_t endw = __funnelshift_r(P0, P1, 8);              // the 4 bytes ending at offset bl = 4
		uint32_t head = 0;
		bool finished = false;
		for (int hop = 1; !finished; hop++) {
			if (d) {
				acc += d;
				if (acc > maxdist) {
					d = 0;
				} else {
					const uint32_t q = pos - acc;
					d = S.prev[q & kRingMask];
					if (load4(ring8, q + bl - 3) == endw) {
						const uint32_t len = match_length(ring8, pos, q, maxl, P0, P1, P2, P3);
						if (len > bl) {
							bl = len; bd = acc;
							if (len >= (uint32_t)nice || len >= maxl)
								d = 0;
							else
								endw = load4(ring8, pos + bl - 3);
						}
					}
				}
			}
			if (hop < depth) {
				if (hop & 1)
					continue;                              // look at the head every other hop
			} else {
				d = 0;
			}
			const uint32_t dm = __ballot_sync(0xffffffffu, d == 0);
			if ((dm | ((1u << head) - 1)) == 0xffffffffu)
				break;                                     // everything from the head on is final
			// serial steps while the lane at the head is final
			while (head < nlive && ((dm >> head) & 1)) {
				uint32_t L = __shfl_sync(0xffffffffu, bl, head);
				L = L >= (uint32_t)kMinMatch ? L : 0;
				if (L && lazy && L < (uint32_t)lazy && head < 31) {
					if (!((dm >> (head + 1)) & 1))
						break;
					if (__shfl_sync(0xffffffffu, bl, head + 1) > L)
						L = 0;
				}
				const uint32_t bdh = __shfl_sync(0xffffffffu, bd, head);
				const uint32_t byh = __shfl_sync(0xffffffffu, myb, head);
				if (lane == 0)
					tk[n] = L ? tok_match(L, bdh) : byh;
				n++;
				head += L ? L : 1;
			}
			if (head >= nlive)
				finished = true;
		}
		if (!finished) {
			// greedy / lazy choice per position, then the reach set of the head by pointer jumping
			const uint32_t len = (bl >= (uint32_t)kMinMatch) ? bl : 0;
			uint32_t nlen = __shfl_down_sync(0xffffffffu, len, 1);
			if (lane == 31)
				nlen = 0;                                  // no look-ahead across windows
			bool take = len != 0;
			if (take && lazy && len < (uint32_t)lazy && nlen > len)
				take = false;
			uint32_t J = live ? lane + (take ? len : 1) : 64;      // next token start (>= nlive: leaves the window)
			uint32_t Rl = live ? (1u << lane) : 0;
#pragma unroll
			for (int k = 0; k < 5; k++) {
				const uint32_t tJ = __shfl_sync(0xffffffffu, J, J & 31);
				const uint32_t tR = __shfl_sync(0xffffffffu, Rl, J & 31);
				if (J < nlive) { Rl |= tR; J = tJ; }
			}
			const uint32_t R = __shfl_sync(0xffffffffu, Rl, head);
			const uint32_t exitJ = __shfl_sync(0xffffffffu, J, head);
			if ((R >> lane) & 1)
				tk[n + __popc(R & lt)] = take ? tok_match(len, bd) : myb;
			n += __popc(R);
			head = exitJ;
		}
		p += head;
		adv = (3 * adv + min(head, 256u) + 2) >> 2;
	}
	end_pos = p;
	return n;
}

// ---- two-pass parser (levels 5 and up) ----
// Searching every position to full depth wastes most of the work: only the positions where a token
// starts (about one in ten) ever use their result.  So: pass 1 walks every position of the
// sub-block to a SHALLOW depth (d1 hops, one lane per position, 32 consecutive positions per window)
// and parses those results; the token starts of that parse — plus the position behind each start
// whose match is short enough for the lazy rule to look at — are queued, and pass 2 walks only the
// queued positions to the FULL depth, 32 of them per step, one per lane.  The final parse reads the
// per-position results (deep where there is one, shallow elsewhere) back from scratch.
// tools/lzsim.c (two=2): same or better ratio than full depth everywhere at a third of the hops.

// one lane's chain walk: longest match for `pos` among the first `depth` chain candidates
// (seed: a match already known for this position — the shallow pass's result when the deep pass walks the same
// chain again; only strictly longer matches replace it, so the 4-byte filter rejects the candidates that
// produced it without a full compare and the result is the same as walking from scratch)
__device__ __forceinline__ void walk_chain(const Smem &S, const uint8_t *ring8, uint32_t pos, uint32_t maxl, uint32_t maxdist,
					   int depth, uint32_t nice, uint32_t &bl, uint32_t &bd, uint32_t seed = 0,
					   uint32_t *cursor = nullptr, uint32_t resume = 0xffffffffu, bool probe_runs = false)
{
	// cursor: where this walk stopped in the chain (distance walked << 16 | next link), so that a later, deeper
	// walk of the same position (resume) continues there instead of repeating the first hops
	uint32_t P0, P1, P2, P3;
	{
		const uint32_t o = pos & kRingMask, a = o & ~3u, sh = (o & 3) * 8;
		const uint32_t w0 = ld32(ring8, a), w1 = ld32(ring8, a + 4), w2 = ld32(ring8, a + 8), w3 = ld32(ring8, a + 12), w4 = ld32(ring8, a + 16);
		P0 = __funnelshift_r(w0, w1, sh); P1 = __funnelshift_r(w1, w2, sh);
		P2 = __funnelshift_r(w2, w3, sh); P3 = __funnelshift_r(w3, w4, sh);
	}
	bl = kMinMatch - 1; bd = 0;
	uint32_t acc = 0;
	uint32_t d = (maxl >= (uint32_t)kMinMatch) ? S.prev[pos & kRingMask] : 0;
	uint32_t endw = __funnelshift_r(P0, P1, 8);                  // the 4 bytes ending at offset bl = 4
	if (tok_is_match(seed) && tok_len(seed) <= maxl) {
		bl = tok_len(seed); bd = tok_dist(seed);
		if (bl >= nice || bl >= maxl)
			d = 0;
		else
			endw = load4(ring8, pos + bl - 3);
	}
	if (resume != 0xffffffffu && d) {
		acc = resume >> 16;
		d = resume & 0xffffu;
	}
	// Runs (a 1-8 byte pattern repeated): the nearest candidates of such a position sit inside the same 32-position
	// insert instruction, where the racy chain build links everybody to the OLDER head — so the chains miss the first
	// 32-64 bytes of every run and short runs (20-120 repeats) came out 2.4 x zlib's size.  Those positions come out of
	// the shallow pass as literals, i.e. as token starts, so the DEEP pass sees every one of them: it looks at distances
	// 1-8 directly (the 8 bytes in front of the position are two more words away).  28 instructions per 32 queued
	// positions, about 0.5 % of a sub-block.  Tried instead and measured: the same probe in the shallow pass behind a
	// warp vote on the first links (-4 % with a vote that catches run starts, no gain with a stricter one), exact
	// chaining inside the insert instruction with match.any (level 6: 23 -> 11.5 GB/s) or with a read-back of the racy
	// store (-28 %), a repeat-distance probe with 3/4-byte matches (-6 % for +0.3 % ratio).
	if (probe_runs && maxl >= (uint32_t)kMinMatch && maxdist >= 4) {
		const uint32_t o = pos & kRingMask, a = o & ~3u, sh = (o & 3) * 8;
		const uint32_t wm1 = ld32(ring8, (a - 4) & kRingMask);
		const uint32_t before = __funnelshift_r(wm1, ld32(ring8, a), sh);                                  // bytes [pos-4, pos)
		const uint32_t before8 = __funnelshift_r(ld32(ring8, (a - 8) & kRingMask), wm1, sh);              // bytes [pos-8, pos-4)
		uint32_t rd = 0;
		if (maxdist >= 8) {
			if (before8 == P0) rd = 8;
			if (__funnelshift_r(before8, before, 8) == P0) rd = 7;
			if (__funnelshift_r(before8, before, 16) == P0) rd = 6;
			if (__funnelshift_r(before8, before, 24) == P0) rd = 5;
		}
		if (before == P0) rd = 4;
		if (__funnelshift_r(before, P0, 8) == P0) rd = 3;
		if (__funnelshift_r(before, P0, 16) == P0) rd = 2;
		if (__funnelshift_r(before, P0, 24) == P0) rd = 1;
		if (rd) {
			const uint32_t rl = match_length(ring8, pos, pos - rd, maxl, P0, P1, P2, P3);
			if (rl > bl) {
				bl = rl; bd = rd;
				if (bl >= nice || bl >= maxl)
					d = 0;
				else
					endw = load4(ring8, pos + bl - 3);
			}
		}
	}
	for (int hop = 0; hop < depth; hop++) {
		if (!__any_sync(0xffffffffu, d != 0))
			break;
		if (d) {
			acc += d;
			if (acc > maxdist) {
				d = 0;
			} else {
				const uint32_t q = pos - acc;
				d = S.prev[q & kRingMask];
				if (load4(ring8, q + bl - 3) == endw) {
					const uint32_t len = match_length(ring8, pos, q, maxl, P0, P1, P2, P3);
					if (len > bl) {
						bl = len; bd = acc;
						if (len >= nice || len >= maxl)
							d = 0;
						else
							endw = load4(ring8, pos + bl - 3);
					}
				}
			}
		}
	}
	if (cursor)
		*cursor = acc << 16 | d;
}

// greedy / lazy choice per position of one window, then pointer jumping: every lane ends up with
// the set of token starts reached from it (Rl) and where the stream leaves the window (J >= nlive)
__device__ __forceinline__ void window_parse(uint32_t lane, uint32_t nlive, uint32_t len, int lazy, bool &take, uint32_t &J, uint32_t &Rl)
{
	const bool live = lane < nlive;
	uint32_t nlen = __shfl_down_sync(0xffffffffu, len, 1);
	if (lane == 31)
		nlen = 0;                                              // no look-ahead across windows
	take = len != 0;
	if (take && lazy && len < (uint32_t)lazy && nlen > len)
		take = false;
	J = live ? lane + (take ? len : 1) : 64;
	Rl = live ? (1u << lane) : 0;
#pragma unroll
	for (int k = 0; k < 5; k++) {
		const uint32_t tJ = __shfl_sync(0xffffffffu, J, J & 31);
		const uint32_t tR = __shfl_sync(0xffffffffu, Rl, J & 31);
		if (J < nlive) { Rl |= tR; J = tJ; }
	}
}

__device__ uint32_t parse_subblock_2pass(Smem &S, uint32_t sub_lo, uint32_t sub_hi, uint32_t valid_lo, uint32_t PE,
					 int d1, int depth, int nice, int lazy, bool use_rep /* run probe */, bool skip_covered, uint32_t *tk, uint32_t *pres,
					 uint32_t &nwin, uint32_t &end_pos)
{
	const uint32_t lane = lane_id();
	const uint32_t lt = (1u << lane) - 1;
	const uint8_t *ring8 = reinterpret_cast<const uint8_t *>(S.ring32);
	const uint32_t npos = sub_hi - sub_lo;
	uint32_t qpos = 0, qtok = 0, qcur = 0, qn = 0;             // pass-2 queue: one position (its shallow result, its chain cursor) per lane
	uint32_t headA = 0;                                        // pass-1 parse: next token start (relative to sub_lo)
	uint32_t carry = 0;                                        // mark for lane 0 of the next window
	uint32_t cov_d = 0;                                        // distance of the match that carries the shallow parse to headA (0: a literal)

	// Chain depth of the deep pass follows the data: the cost of a sub-block is (queued positions) x
	// depth, and long-match data (few token starts per byte, many equally good candidates) is where
	// deeper chains pay, so the depth grows with the bytes scanned per queued position:
	// depth = base * bytes_per_queued / 12, within [base / 3, 2 * base].
	const int base_depth = depth;
	uint32_t fired_at = 0;                                     // bytes of the sub-block scanned when the queue last fired
	auto deep = [&](uint32_t count, uint32_t scanned) {
		const uint32_t bpq16 = ((scanned - fired_at) << 4) / max(count, 1u);      // bytes per queued position, x16
		fired_at = scanned;
		depth = max(base_depth / 3, min(2 * base_depth, (int)((uint32_t)base_depth * bpq16 / (12 * 16))));
		const bool act = lane < count;
		const uint32_t pos = act ? qpos : sub_lo;
		const uint32_t maxl = act ? min((uint32_t)kMaxMatch, PE - pos) : 0;
		uint32_t bl, bd;
		walk_chain(S, ring8, pos, maxl, min((uint32_t)kWindow, pos - valid_lo), max(1, depth - d1), (uint32_t)nice, bl, bd, act ? qtok : 0, nullptr, act ? qcur : 0, use_rep);   // the first d1 hops were walked by the shallow pass
		if (act && bd)
			__stcg(&pres[pos - sub_lo], tok_match(bl, bd));
	};

	// ---- pass 1 (+ pass 2 whenever 32 positions are queued) ----
	// The token the shallow parse is inside of may cover a whole window: no token starts there, so nothing there is
	// queued for the deep pass, and the only way the final parse can land in it is a deep match at an earlier start that
	// ends in it — and then the covering match's own tail (same distance, what is left of its length) is a valid match
	// for that position.  Such a window is not searched at all.
	auto covered = [&](uint32_t w0) { return skip_covered && cov_d && carry == 0 && headA >= w0 + 32 && w0 + 32 <= npos; };
	auto inherit = [&](uint32_t w0) {
		const uint32_t rem = headA - (w0 + lane);
		__stcg(&pres[w0 + lane], rem >= (uint32_t)kMinMatch ? tok_match(rem, cov_d) : (uint32_t)ring8[(sub_lo + w0 + lane) & kRingMask]);
	};
	// one window of 32 positions whose shallow results are in (mytok, mycur), one position per lane
	auto window = [&](uint32_t w0, uint32_t mytok, uint32_t mycur) {
		nwin++;
		const uint32_t pos = sub_lo + w0 + lane;
		const uint32_t nlive = min(32u, npos - w0);
		const bool live = lane < nlive;
		const uint32_t len = tok_is_match(mytok) ? tok_len(mytok) : 0;
		if (live)
			__stcg(&pres[w0 + lane], len ? mytok : (uint32_t)ring8[pos & kRingMask]);
		uint32_t M = carry;
		carry = 0;
		if (headA < w0 + nlive) {
			bool take; uint32_t J, Rl;
			window_parse(lane, nlive, len, lazy, take, J, Rl);
			const uint32_t h = headA - w0;
			const uint32_t R = __shfl_sync(0xffffffffu, Rl, h);
			headA = w0 + __shfl_sync(0xffffffffu, J, h);
			{
				// the token that leaves the window: the last start of the reach set
				const uint32_t last = 31 - __clz(R);
				const uint32_t ltake = __ballot_sync(0xffffffffu, take);
				cov_d = ((ltake >> last) & 1) ? tok_dist(__shfl_sync(0xffffffffu, mytok, last)) : 0;
			}
			// a start whose match is short enough for the lazy rule also needs the position behind it
			const uint32_t nx = __ballot_sync(0xffffffffu, ((R >> lane) & 1) && lazy && len < (uint32_t)lazy);
			M |= R | (nx << 1);
			carry = nx >> 31;
		}
		if (nlive < 32)
			M &= (1u << nlive) - 1;
		while (M) {
			const uint32_t cnt = __popc(M), room = 32 - qn, tk_n = min(cnt, room);
			// the window lane whose position this queue slot takes = the (lane - qn)-th set bit of M: binary search on the
			// prefix population counts (__fns is ~60 straight-line instructions, and it ran twice per round)
			const uint32_t want = lane - qn;
			uint32_t from = 0;
#pragma unroll
			for (int st = 16; st; st >>= 1)
				if ((uint32_t)__popc(M & (0xffffffffu >> (32 - st - from))) <= want)
					from += st;
			from &= 31;
			const uint32_t ftok = __shfl_sync(0xffffffffu, mytok, from);
			const uint32_t fcur = __shfl_sync(0xffffffffu, mycur, from);
			if (lane >= qn && lane < qn + tk_n) {
				qpos = sub_lo + w0 + from;
				qtok = ftok;
				qcur = fcur;
			}
			qn += tk_n;
			if (tk_n == cnt)
				M = 0;
			else
				M = __ballot_sync(0xffffffffu, ((M >> lane) & 1) && (uint32_t)__popc(M & lt) >= tk_n);    // keep the set bits behind the first tk_n
			if (qn == 32) {
				deep(32, min(w0 + 32, npos));
				qn = 0;
			}
		}
	};
	for (uint32_t w0 = 0; w0 < npos; w0 += 32) {
		if (covered(w0)) { inherit(w0); continue; }
		const uint32_t pos = sub_lo + w0 + lane;
		const uint32_t maxl = lane < min(32u, npos - w0) ? min((uint32_t)kMaxMatch, PE - pos) : 0;
		uint32_t bl, bd, mycur;
		walk_chain(S, ring8, pos, maxl, min((uint32_t)kWindow, pos - valid_lo), d1, (uint32_t)nice, bl, bd, 0, &mycur);
		window(w0, bd ? tok_match(bl, bd) : 0, mycur);
	}
	if (qn)
		deep(qn, npos);
	__threadfence_block();
	__syncwarp();

	// ---- final parse over the stored per-position results ----
	uint32_t n = 0, head = 0;
	uint32_t t_next = lane < npos ? __ldcg(&pres[lane]) : 0;          // results are read one window ahead (L2 latency)
	for (uint32_t w0 = 0; w0 < npos; w0 += 32) {
		const uint32_t nlive = min(32u, npos - w0);
		const uint32_t t = t_next;
		if (w0 + 32 + lane < npos)
			t_next = __ldcg(&pres[w0 + 32 + lane]);
		if (head >= w0 + nlive)
			continue;
		const bool live = lane < nlive;
		const uint32_t pos = sub_lo + w0 + lane;
		const uint32_t len = (live && tok_is_match(t)) ? tok_len(t) : 0;
		bool take; uint32_t J, Rl;
		window_parse(lane, nlive, len, lazy, take, J, Rl);
		const uint32_t h = head - w0;
		const uint32_t R = __shfl_sync(0xffffffffu, Rl, h);
		head = w0 + __shfl_sync(0xffffffffu, J, h);
		if ((R >> lane) & 1)
			tk[n + __popc(R & lt)] = take ? t : (uint32_t)ring8[pos & kRingMask];
		n += __popc(R);
	}
	end_pos = sub_lo + head;
	return n;
}

// ---- stitch helpers (after the parse): sub-block sb+1 was parsed from its nominal start although
// the last match of sub-block sb may cover its first bytes.  The match is cut back to the latest
// token start of sb+1 it reaches, and sb+1 drops the tokens in front of that start.
__device__ __forceinline__ uint32_t tok_span(uint32_t t) { return tok_is_match(t) ? tok_len(t) : 1; }

// tokens tk[0..cnt) start at `start` and end at `end`; returns the largest j (0..cnt, cnt = "end")
// whose start s_j <= e, and s_j itself
__device__ uint32_t scan_head(const uint32_t *tk, uint32_t cnt, uint32_t start, uint32_t end, uint32_t e, uint32_t &s_best)
{
	const uint32_t lane = lane_id();
	uint32_t j_best = 0;
	s_best = start;
	for (uint32_t base = 0; base < cnt; base += 32) {
		const uint32_t i = base + lane;
		const uint32_t span = i < cnt ? tok_span(tk[i]) : 0;
		uint32_t incl = span;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += y;
		}
		const uint32_t s_i = start + incl - span;
		const uint32_t q = __ballot_sync(0xffffffffu, i < cnt && s_i <= e);
		const uint32_t nq = __popc(q);
		if (nq) {
			j_best = base + nq - 1;
			s_best = __shfl_sync(0xffffffffu, s_i, nq - 1);
		}
		start += __shfl_sync(0xffffffffu, incl, 31);
		if (nq < 32)
			return j_best;
	}
	if (end <= e) {
		j_best = cnt;
		s_best = end;
	}
	return j_best;
}

// ---- Huffman: CTA-collective code-length construction ----
__constant__ uint8_t c_lext[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
__constant__ uint8_t c_dext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
__constant__ uint8_t c_clorder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// freq[n] (shared) -> len[n] (shared), all lengths <= maxbits, complete code (Kraft sum == 1)
// when at least two symbols are present.  In the spirit of lib/nx_dhtgen.c:418-595.
__device__ void build_lengths(HuffScratch &H, const uint32_t *freq, int n, int maxbits, uint8_t *len)
{
	const int t = threadIdx.x;
	uint32_t f = (t < n) ? freq[t] : 0;
	if (t < 16)
		H.bl_count[t] = 0;
	if (t < n)
		len[t] = 0;
	const int nl = __syncthreads_count(f > 0);
	if (nl == 0)
		return;
	if (f > 0) {
		int rank = 0;
		for (int j = 0; j < n; j++) {
			uint32_t fj = freq[j];
			rank += (fj > 0) && (fj < f || (fj == f && j < t));
		}
		H.sorted_sym[rank] = (uint16_t)t;
		H.w[rank] = f;
	}
	__syncthreads();
	if (nl == 1) {
		if (t == 0)
			len[H.sorted_sym[0]] = 1;
		__syncthreads();
		return;
	}
	if (t == 0) {
		// two-queue merge (leaves ascending in w[0..nl), internal nodes appended).  One thread, one dependent step per
		// merge: the two heads of either queue live in registers (kInf = absent), the weight behind them is loaded when a
		// head is taken, so a pick never waits for shared memory (it took ~200 cycles per merge with every weight re-read).
		constexpr uint32_t kInf = 0xffffffffu;
		int q1 = 0, q2 = nl, tot = nl;
		uint32_t wl = H.w[0], wl2 = nl > 1 ? H.w[1] : kInf;       // w[q1], w[q1 + 1]
		uint32_t wn = kInf, wn2 = kInf;                            // w[q2], w[q2 + 1]
		while ((nl - q1) + (tot - q2) > 1) {
			int a, b;
			uint32_t wa, wb;
			if (wl <= wn) { a = q1++; wa = wl; wl = wl2; wl2 = q1 + 1 < nl ? H.w[q1 + 1] : kInf; }
			else { a = q2++; wa = wn; wn = wn2; wn2 = q2 + 1 < tot ? H.w[q2 + 1] : kInf; }
			if (wl <= wn) { b = q1++; wb = wl; wl = wl2; wl2 = q1 + 1 < nl ? H.w[q1 + 1] : kInf; }
			else { b = q2++; wb = wn; wn = wn2; wn2 = q2 + 1 < tot ? H.w[q2 + 1] : kInf; }
			const uint32_t ws = wa + wb;
			H.w[tot] = ws;
			H.parent[a] = (uint16_t)tot;
			H.parent[b] = (uint16_t)tot;
			// the new node may be one of the two heads of its queue
			if (tot == q2) wn = ws;
			else if (tot == q2 + 1) wn2 = ws;
			tot++;
		}
		H.tot = tot;
	}
	__syncthreads();
	if (t < nl) {
		int d = 0, x = t;
		const int root = H.tot - 1;
		while (x != root) { x = H.parent[x]; d++; }
		atomicAdd(&H.bl_count[d < maxbits ? d : maxbits], 1u);
	}
	__syncthreads();
	if (t == 0) {
		// length limiting: keep the Kraft sum at exactly 1 (each pass: -1 leaf at maxbits, split one shorter code)
		uint32_t total = 0;
		for (int l = maxbits; l >= 1; l--)
			total += H.bl_count[l] << (maxbits - l);
		while (total > (1u << maxbits)) {
			H.bl_count[maxbits]--;
			for (int l = maxbits - 1; l >= 1; l--)
				if (H.bl_count[l]) { H.bl_count[l]--; H.bl_count[l + 1] += 2; break; }
			total--;
		}
	}
	__syncthreads();
	if (t < nl) {
		// rank 0 is the rarest symbol: it takes the longest remaining length
		uint32_t cum = 0;
		int l = maxbits;
		for (; l >= 1; l--) {
			cum += H.bl_count[l];
			if ((uint32_t)t < cum)
				break;
		}
		len[H.sorted_sym[t]] = (uint8_t)l;
	}
	__syncthreads();
}

// canonical codes (RFC 1951 §3.2.2), stored bit-reversed for LSB-first emission
__device__ void assign_codes(HuffScratch &H, const uint8_t *len, int n, uint16_t *code)
{
	const int t = threadIdx.x;
	if (t < 16)
		H.bl_count[t] = 0;
	__syncthreads();
	if (t < n && len[t])
		atomicAdd(&H.bl_count[len[t]], 1u);
	__syncthreads();
	if (t == 0) {
		uint32_t c = 0;
		H.bl_count[0] = 0;
		for (int l = 1; l <= 15; l++) {
			c = (c + H.bl_count[l - 1]) << 1;
			H.next_code[l] = c;
		}
	}
	__syncthreads();
	if (t < n) {
		const int l = len[t];
		uint32_t c = 0;
		if (l) {
			c = H.next_code[l];
			for (int j = 0; j < t; j++)
				c += (len[j] == l);
			c = __brev(c) >> (32 - l);
		}
		code[t] = (uint16_t)c;
	}
	__syncthreads();
}

__device__ __forceinline__ void put_bits(uint32_t *words, uint32_t &bitpos, uint32_t v, uint32_t n)
{
	// single-thread writer into zero-initialised words
	const uint32_t w = bitpos >> 5, o = bitpos & 31;
	words[w] |= v << o;
	if (o + n > 32)
		words[w + 1] |= v >> (32 - o);
	bitpos += n;
}

// Builds ll/d code lengths + codes and the dynamic block header (into H.hdr_words).
// Returns the total bit cost of a dynamic block (header + symbols + extra bits + EOB), CTA-uniform.
__device__ uint32_t build_dynamic(Smem &S, HuffScratch &H, uint32_t bfinal)
{
	const int t = threadIdx.x;
	build_lengths(H, S.ll_freq, 286, 15, H.ll_len);
	build_lengths(H, S.d_freq, 30, 15, H.d_len);
	if (t == 0) {
		// drop the dummy counts that only kept the trees complete
		if (S.misc[0] != 0xFFFFu) S.d_freq[S.misc[0]] = 0;
		if (S.misc[1] != 0xFFFFu) S.d_freq[S.misc[1]] = 0;
		if (S.misc[2]) S.ll_freq[0] -= 1;
	}
	__syncthreads();
	assign_codes(H, H.ll_len, 286, H.ll_code);
	assign_codes(H, H.d_len, 30, H.d_code);
	if (t < 20)
		H.cl_freq[t] = 0;
	for (int i = t; i < 96; i += kThreads)
		H.hdr_words[i] = 0;
	__syncthreads();
	if (t == 0) {
		// run-length encode the code lengths with symbols 16/17/18 (lib/nx_dhtgen.c:709-915 does the same job)
		int hlit = 286, hdist = 30;
		while (hlit > 257 && H.ll_len[hlit - 1] == 0) hlit--;
		while (hdist > 1 && H.d_len[hdist - 1] == 0) hdist--;
		const int n = hlit + hdist;
		uint32_t ns = 0;
		for (int i = 0; i < n;) {
			const int v = i < hlit ? H.ll_len[i] : H.d_len[i - hlit];
			int k = i + 1;
			while (k < n && (k < hlit ? H.ll_len[k] : H.d_len[k - hlit]) == v) k++;
			int run = k - i;
			if (v == 0) {
				while (run >= 11) { int r = run > 138 ? 138 : run; H.cl_sym[ns++] = (uint16_t)(18 | ((r - 11) << 8)); H.cl_freq[18]++; run -= r; }
				if (run >= 3) { H.cl_sym[ns++] = (uint16_t)(17 | ((run - 3) << 8)); H.cl_freq[17]++; run = 0; }
				while (run-- > 0) { H.cl_sym[ns++] = 0; H.cl_freq[0]++; }
			} else {
				H.cl_sym[ns++] = (uint16_t)v; H.cl_freq[v]++; run--;
				while (run >= 3) { int r = run > 6 ? 6 : run; H.cl_sym[ns++] = (uint16_t)(16 | ((r - 3) << 8)); H.cl_freq[16]++; run -= r; }
				while (run-- > 0) { H.cl_sym[ns++] = (uint16_t)v; H.cl_freq[v]++; }
			}
			i = k;
		}
		H.n_cl_sym = ns;
		H.nl = hlit | (hdist << 16);
		// a code-length code needs two symbols to be complete
		int nz = 0;
		for (int i = 0; i < 19; i++) nz += H.cl_freq[i] != 0;
		if (nz < 2) { if (H.cl_freq[0] == 0) H.cl_freq[0] = 1; else H.cl_freq[1] = 1; }
	}
	__syncthreads();
	build_lengths(H, H.cl_freq, 19, 7, H.cl_len);
	assign_codes(H, H.cl_len, 19, H.cl_code);
	if (t == 0) {
		const int hlit = H.nl & 0xffff, hdist = H.nl >> 16;
		int hclen = 19;
		while (hclen > 4 && H.cl_len[c_clorder[hclen - 1]] == 0) hclen--;
		uint32_t bp = 0;
		put_bits(H.hdr_words, bp, bfinal | (2u << 1), 3);
		put_bits(H.hdr_words, bp, hlit - 257, 5);
		put_bits(H.hdr_words, bp, hdist - 1, 5);
		put_bits(H.hdr_words, bp, hclen - 4, 4);
		for (int i = 0; i < hclen; i++)
			put_bits(H.hdr_words, bp, H.cl_len[c_clorder[i]], 3);
		for (uint32_t i = 0; i < H.n_cl_sym; i++) {
			const uint32_t s = H.cl_sym[i] & 0xff, x = H.cl_sym[i] >> 8;
			put_bits(H.hdr_words, bp, H.cl_code[s], H.cl_len[s]);
			if (s == 16) put_bits(H.hdr_words, bp, x, 2);
			else if (s == 17) put_bits(H.hdr_words, bp, x, 3);
			else if (s == 18) put_bits(H.hdr_words, bp, x, 7);
		}
		H.hdr_bits = bp;
	}
	__syncthreads();
	// symbol cost
	uint32_t bits = 0;
	if (t < 286)
		bits = S.ll_freq[t] * (H.ll_len[t] + (t > 256 ? c_lext[t - 257] : 0));
	else if (t >= 288 && t < 318)
		bits = S.d_freq[t - 288] * (H.d_len[t - 288] + c_dext[t - 288]);
	for (int o = 16; o; o >>= 1)
		bits += __shfl_xor_sync(0xffffffffu, bits, o);
	if (lane_id() == 0)
		H.warp_sums[threadIdx.x >> 5] = bits;
	__syncthreads();
	uint32_t total = H.hdr_bits;
	for (int i = 0; i < 10; i++)
		total += H.warp_sums[i];
	__syncthreads();
	return total;
}

__device__ uint32_t fixed_cost_and_tables(Smem &S, HuffScratch &H, bool install)
{
	const int t = threadIdx.x;
	uint32_t bits = 0;
	if (t < 288) {
		const uint8_t l = t < 144 ? 8 : t < 256 ? 9 : t < 280 ? 7 : 8;
		if (install)
			H.ll_len[t] = l;
		if (t < 286)
			bits = S.ll_freq[t] * (l + (t > 256 ? c_lext[t - 257] : 0));
	} else if (t < 320) {
		if (install)
			H.d_len[t - 288] = 5;
		if (t - 288 < 30)
			bits = S.d_freq[t - 288] * (5 + c_dext[t - 288]);
	}
	for (int o = 16; o; o >>= 1)
		bits += __shfl_xor_sync(0xffffffffu, bits, o);
	__syncthreads();
	if (lane_id() == 0)
		H.warp_sums[threadIdx.x >> 5] = bits;
	__syncthreads();
	uint32_t total = 3;
	for (int i = 0; i < 10; i++)
		total += H.warp_sums[i];
	__syncthreads();
	if (install) {
		assign_codes(H, H.ll_len, 288, H.ll_code);
		assign_codes(H, H.d_len, 32, H.d_code);
	}
	return total;
}

// ---- bit packer: tokens -> Huffman bit stream (staging window in the dead input ring) ----
struct Packer {
	uint32_t *stage;        // kStageWords words, shared
	uint32_t *out32;        // global, 16-byte aligned
	uint32_t words_out;     // full words already flushed
	uint32_t carry;         // valid bits in stage[0]
};

// OR `n` bits of v at absolute stage bit offset `bp` (multi-thread safe)
__device__ __forceinline__ void or_bits(uint32_t *stage, uint32_t bp, uint64_t v, uint32_t n)
{
	if (n == 0)
		return;
	const uint32_t w = bp >> 5, o = bp & 31;
	const uint64_t lo = v << o;
	atomicOr(&stage[w], (uint32_t)lo);
	if (o + n > 32)
		atomicOr(&stage[w + 1], (uint32_t)(lo >> 32));
	if (o + n > 64)
		atomicOr(&stage[w + 2], (uint32_t)(v >> (64 - o)));
}

// after `nbits` new bits were OR-ed behind the carry: flush the full words, keep the tail
__device__ void packer_flush(Packer &P, uint32_t nbits)
{
	const uint32_t total = P.carry + nbits;
	const uint32_t full = total >> 5;
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < full; i += kThreads)
		P.out32[P.words_out + i] = P.stage[i];
	const uint32_t tail = P.stage[full];
	__syncthreads();
	// (only the words this round touched: everything behind them is still zero)
	for (uint32_t i = threadIdx.x; i < full + 3 && i < (uint32_t)kStageWords; i += kThreads)
		P.stage[i] = (i == 0) ? tail : 0;
	__syncthreads();
	P.words_out += full;
	P.carry = total & 31;
}

__device__ void copy_bits_to_stage(Packer &P, const uint32_t *words, uint32_t nbits)
{
	// thread-parallel: word i of the source lands at bit offset carry + 32 i
	const uint32_t nw = (nbits + 31) >> 5;
	for (uint32_t i = threadIdx.x; i < nw; i += kThreads) {
		uint32_t n = (i == nw - 1 && (nbits & 31)) ? (nbits & 31) : 32;
		uint32_t v = words[i];
		if (n < 32)
			v &= (1u << n) - 1;
		or_bits(P.stage, P.carry + 32 * i, v, n);
	}
}

__device__ void encode_tokens(Smem &S, HuffScratch &H, Packer &P, const uint32_t *tok, uint32_t ntok)
{
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	uint32_t t_next = threadIdx.x < ntok ? __ldcg(&tok[threadIdx.x]) : 0;     // tokens are read one round ahead (L2 latency)
	for (uint32_t base = 0; base < ntok; base += kThreads) {
		const uint32_t i = base + threadIdx.x;
		uint64_t v = 0;
		uint32_t n = 0;
		const uint32_t t = t_next;
		if (i + kThreads < ntok)
			t_next = __ldcg(&tok[i + kThreads]);
		if (i < ntok) {
			if (!tok_is_match(t)) {
				v = H.ll_code[t];
				n = H.ll_len[t];
			} else {
				uint32_t lc, le, lx, dc, de, dx;
				len_code(tok_len(t), lc, le, lx);
				dist_code(tok_dist(t), dc, de, dx);
				v = H.ll_code[257 + lc];
				n = H.ll_len[257 + lc];
				v |= (uint64_t)lx << n; n += le;
				v |= (uint64_t)H.d_code[dc] << n; n += H.d_len[dc];
				v |= (uint64_t)dx << n; n += de;
			}
		}
		// CTA-wide exclusive prefix sum of code widths (warp shuffles + one shared hop)
		uint32_t incl = n;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += y;
		}
		if (lane == 31)
			H.warp_sums[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t ws = H.warp_sums[lane], wi = ws;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= (uint32_t)o)
					wi += y;
			}
			H.warp_sums[lane] = wi - ws;
			if (lane == 31)
				H.batch_total = wi;
		}
		__syncthreads();
		const uint32_t off = H.warp_sums[warp] + incl - n;
		const uint32_t batch_bits = H.batch_total;
		or_bits(P.stage, P.carry + off, v, n);
		packer_flush(P, batch_bits);
	}
}

__device__ void put_tail_bits(Packer &P, uint64_t v, uint32_t n)
{
	if (threadIdx.x == 0)
		or_bits(P.stage, P.carry, v, n);
	packer_flush(P, n);
}

// ---- stored blocks (BTYPE=00) straight from global memory ----
__device__ uint32_t emit_stored(const DeflateJob &J, bool final_flag)
{
	uint32_t o = 0, done = 0;
	const uint32_t n = J.src_len;
	do {
		const uint32_t len = min(65535u, n - done);
		const bool last = done + len == n;
		if (threadIdx.x == 0) {
			J.out[o] = (last && final_flag) ? 1 : 0;
			J.out[o + 1] = (uint8_t)len; J.out[o + 2] = (uint8_t)(len >> 8);
			J.out[o + 3] = (uint8_t)~len; J.out[o + 4] = (uint8_t)(~len >> 8);
		}
		o += 5;
		for (uint32_t i = threadIdx.x; i < len; i += kThreads)
			J.out[o + i] = J.src[done + i];
		o += len;
		done += len;
	} while (done < n);
	return o;
}

// per-CTA scratch in global memory (L2-resident): flat token stream | per-position token slots of
// the sub-blocks | tokens per sub-block | first flat index per sub-block
__host__ __device__ inline uint32_t scratch_nsub(uint32_t tok_stride) { return tok_stride / kSub + 2; }
constexpr int kMeta = 8;     // per sub-block: tokens, end position, tokens dropped in front, kept tokens, flat offset, tail count, tail tokens[2]
enum { M_CNT = 0, M_END = 1, M_SKIP = 2, M_NEW = 3, M_OFF = 4, M_TAILN = 5, M_TAIL0 = 6, M_TAIL1 = 7 };
// (a multiple of 32 words, so that every CTA's scratch starts on a 128-byte line: dead token lines are discarded from the L2)
__host__ __device__ inline size_t scratch_words(uint32_t tok_stride) { return (2 * (size_t)tok_stride + kMeta * (size_t)scratch_nsub(tok_stride) + 32 * (size_t)(kSub + 32) + 31) & ~(size_t)31; }

// ---- fused stitch (StreamOut): wait for the predecessor's end offset, publish ours, copy the slot there ----
__device__ void stream_out(Smem &S, const StreamOut &so, const DeflateJob &J, uint32_t job, uint32_t n_jobs, uint32_t len)
{
	if (threadIdx.x == 0) {
		unsigned long long start = so.base;
		if (job > 0) {
			volatile unsigned long long *prev = so.chain + (job - 1);
			unsigned long long v;
			while ((v = *prev) == 0)
				__nanosleep(200);                 // chunk job-1 is on another SM and, like us, only waits on lower chunks
			start = v - 1;
		}
		*reinterpret_cast<volatile unsigned long long *>(so.chain + job) = start + len + 1;
		so.offsets[job] = start;
		if (job + 1 == n_jobs)
			so.offsets[n_jobs] = start + len;
		S.misc[4] = (uint32_t)start;
		S.misc[5] = (uint32_t)(start >> 32);
	}
	__syncthreads();
	const uint64_t off = (uint64_t)S.misc[4] | (uint64_t)S.misc[5] << 32;
	if (len == 0 || off + len > so.cap)
		return;
	const uint8_t *src = J.out;                       // 16-byte aligned slot
	uint8_t *d = so.dst + off;
	// head bytes until d is 16-byte aligned, 16-byte vectors assembled from the aligned source, tail bytes
	const uint32_t head = min(len, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15));
	if (threadIdx.x < head)
		d[threadIdx.x] = src[threadIdx.x];
	const uint32_t nvec = (len - head) >> 4;
	const uint32_t sh = (head & 3) * 8;
	const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src) + (head >> 2);
	uint4 *d128 = reinterpret_cast<uint4 *>(d + head);
	const uint64_t pol = l2_evict_first_policy();
	for (uint32_t v = threadIdx.x; v < nvec; v += kThreads) {
		const uint32_t *p = s32 + 4 * v;
		const uint32_t a = p[0], b = p[1], c = p[2], e = p[3], f = sh ? p[4] : 0;
		st_v4_evict_first(d128 + v, sh ? make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh))
					       : make_uint4(a, b, c, e), pol);
	}
	const uint32_t done = head + nvec * 16;
	if (threadIdx.x < len - done)
		d[done + threadIdx.x] = src[done + threadIdx.x];
}

__global__ void __launch_bounds__(kThreads, 1)
deflate_kernel(const DeflateJob *__restrict__ jobs, DeflateOut *__restrict__ outs, uint32_t n_jobs,
	       int depth, int lazy, int nice, uint32_t *tok_scratch, uint32_t tok_stride, uint32_t parser_mask,
	       uint32_t *job_counter, const volatile uint32_t *ready, uint32_t jobs_per_flag, int d1, StreamOut so)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	Smem &S = *reinterpret_cast<Smem *>(smem_raw);
	HuffScratch &H = *reinterpret_cast<HuffScratch *>(S.prev);
	const uint32_t warp = threadIdx.x >> 5;
	uint32_t *tok = tok_scratch + (size_t)blockIdx.x * scratch_words(tok_stride);
	uint32_t *tokpos = tok + tok_stride;
	uint32_t *meta = tokpos + tok_stride;
	uint32_t *pres = meta + kMeta * (size_t)scratch_nsub(tok_stride);     // per parser warp: result per position of its sub-block

	for (;;) {
		// jobs are handed out in order; when the input is still being uploaded (host-pointer streams)
		// ready[k] turns non-zero once the slice holding jobs [k*jobs_per_flag, (k+1)*jobs_per_flag) has landed
		if (threadIdx.x == 0) {
			uint32_t j = atomicAdd(job_counter, 1u);
			if (j < n_jobs && ready) {
				// bounded: if the upload of this slice never completes (a failed copy on the host side) the chunk is
				// reported as failed after ~30 s instead of spinning for ever
				uint32_t spins = 0;
				while (ready[j / jobs_per_flag] == 0 && ++spins < 30000000u)
					__nanosleep(1000);
				__threadfence();
				if (spins >= 30000000u) {
					DeflateOut o = {};
					o.rc = NXGPU_E_NODEV;
					outs[j] = o;
					if (so.dst) so.chain[j] = 1;          // an empty chunk at offset 0: nobody behind it waits for ever either
					j = 0xffffffffu;                      // this CTA stops
				}
			}
			S.misc[7] = j;
		}
		__syncthreads();
		const uint32_t job = S.misc[7];
		if (job >= n_jobs)
			break;
		const long long tjob0 = clock64();
		const DeflateJob J = jobs[job];
		const uintptr_t first = reinterpret_cast<uintptr_t>(J.src) - J.hist_len;
		const uint8_t *gbase = reinterpret_cast<const uint8_t *>(first & ~(uintptr_t)15);
		const uint32_t P0 = (uint32_t)(first & 15);      // first dictionary byte
		const uint32_t PS = P0 + J.hist_len;              // first byte to compress
		const uint32_t PE = PS + J.src_len;
		const uint32_t n_sub = (J.src_len + kSub - 1) / kSub;
		const uint32_t nblk = (((PE + 15) & ~15u) + kBlk - 1) / kBlk;   // staging blocks, as in producer()

		// ---- init ----
		for (int i = threadIdx.x; i < kHashSize; i += kThreads)
			S.head[i] = 0;
		for (int i = threadIdx.x; i < 288; i += kThreads)
			S.ll_freq[i] = 0;
		if (threadIdx.x < 32) {
			S.d_freq[threadIdx.x] = 0;
			S.parse_pos[threadIdx.x] = (((parser_mask >> threadIdx.x) & 1) && threadIdx.x && n_sub) ? PS : kNone;
		}
		if (threadIdx.x == 0) {
			S.n_tok = 0;
			S.next_sub = 0;
			for (int i = 0; i < kSlots; i++)
				mbar_init(&S.mbar[i], 1);
			for (int i = 0; i < kDoneSlots; i++) {
				mbar_init(&S.done_bar[i], 1);
				mbar_init(&S.hash_bar[i], 1);
			}
			asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
		}
		// the ring was last written by ordinary stores (bit-packer staging); TMA writes come next
		asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
		__syncthreads();

		// ---- LZ77: warp 0 stages + builds chains, the parser warps take sub-blocks in order ----
		const bool split = (parser_mask & 1) != 0;        // bit 0 (warp 0 is never a parser): chain build split over warps 0 and 1
		const bool run_probe = (d1 & 0x200) != 0;
		if (warp == 0) {
			producer(S, gbase, P0, PE, split);
		} else if (warp == 1 && split) {
			inserter(S, P0, PE);
		} else if ((parser_mask >> warp) & 1) {
			uint32_t nwin = 0;
			long long busy = 0, waited = 0;
			for (;;) {
				uint32_t sb = 0;
				if (lane_id() == 0)
					sb = atomicAdd(&S.next_sub, 1u);
				sb = __shfl_sync(0xffffffffu, sb, 0);
				if (sb >= n_sub)
					break;
				const uint32_t sub_lo = PS + sb * kSub;
				const uint32_t sub_hi = min(PE, sub_lo + kSub);
				__syncwarp();
				if (lane_id() == 0)
					S.parse_pos[warp] = sub_lo;
				const long long t0 = clock64();
				{
					// chains for every position below sub_hi exist once block bw is done; the wait
					// suspends the warp in hardware instead of polling (polling steals issue slots
					// from the producer).  All unfinished sub-blocks lie within 16 KiB of each other,
					// so the barrier slot cannot be more than one phase away.
					// (data is needed a full match beyond sub_hi: the last match may run that far)
					const uint32_t bw = min(nblk - 1, (min(PE, sub_hi + kMaxMatch) + 3) / kBlk);
					uint64_t *bar = &S.done_bar[bw % kDoneSlots];
					const uint32_t parity = (bw / kDoneSlots) & 1;
					mbar_wait_sleep(bar, parity);
				}
				__threadfence_block();
				const long long t1 = clock64();
				uint32_t end_pos;
				const bool probe = run_probe;
				const uint32_t cnt = d1 & 0xff
					? parse_subblock_2pass(S, sub_lo, sub_hi, P0, PE, d1 & 0xff, depth, nice, lazy, probe, (d1 & 0x400) != 0, tokpos + (size_t)sb * kSub,
							       pres + (size_t)warp * (kSub + 32), nwin, end_pos)
					: parse_subblock(S, sub_lo, sub_hi, P0, PE, depth, nice, lazy, tokpos + (size_t)sb * kSub, nwin, end_pos);
				if (lane_id() == 0) {
					meta[sb * kMeta + M_CNT] = cnt;
					meta[sb * kMeta + M_END] = end_pos;
				}
				waited += t1 - t0;
				busy += clock64() - t1;
			}
			__syncwarp();
			if (lane_id() == 0)
				S.parse_pos[warp] = kNone;
			if (warp == 2) { DBG_ADD(3, busy); DBG_ADD(4, waited); DBG_ADD(7, nwin); }
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			for (int i = 0; i < kSlots; i++)
				asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&S.mbar[i])) : "memory");
			for (int i = 0; i < kDoneSlots; i++) {
				asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&S.done_bar[i])) : "memory");
				asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&S.hash_bar[i])) : "memory");
			}
		}

		// ---- stitch: per sub-block, tokens to drop in front and the cut-back last token ----
		for (uint32_t sb = warp; sb < n_sub; sb += kThreads / 32) {
			const uint32_t sub_lo = PS + sb * kSub;
			const uint32_t cnt = meta[sb * kMeta + M_CNT], end = meta[sb * kMeta + M_END];
			const uint32_t *tk = tokpos + (size_t)sb * kSub;
			uint32_t s_keep = 0, skip = 0;
			if (sb > 0)
				skip = scan_head(tk, cnt, sub_lo, end, meta[(sb - 1) * kMeta + M_END], s_keep);
			uint32_t tail_n = 0, tail0 = 0, tail1 = 0;
			if (skip < cnt) {
				// the last token: cut back to the latest token start of the next sub-block it reaches
				uint32_t cut = end;
				if (sb + 1 < n_sub)
					scan_head(tokpos + (size_t)(sb + 1) * kSub, meta[(sb + 1) * kMeta + M_CNT], min(PE, sub_lo + kSub),
						  meta[(sb + 1) * kMeta + M_END], end, cut);
				const uint32_t last = tk[cnt - 1];
				const uint32_t a = end - tok_span(last);              // where the last token starts
				const uint32_t L = cut - a;
				if (L == tok_span(last)) {
					tail_n = 1; tail0 = last;
				} else if (L >= 3) {
					tail_n = 1; tail0 = tok_match(L, tok_dist(last));
				} else {
					tail_n = L; tail0 = J.src[a - PS]; tail1 = L == 2 ? J.src[a - PS + 1] : 0;   // 1 or 2 literals
				}
			}
			if (lane_id() == 0) {
				meta[sb * kMeta + M_SKIP] = skip;
				meta[sb * kMeta + M_NEW] = skip < cnt ? cnt - skip - 1 + tail_n : 0;
				meta[sb * kMeta + M_TAILN] = tail_n;
				meta[sb * kMeta + M_TAIL0] = tail0;
				meta[sb * kMeta + M_TAIL1] = tail1;
			}
		}
		__syncthreads();

		// ---- flat token stream + histograms: scan the kept counts, then copy ----
		uint32_t ntok = 0;
		for (uint32_t base = 0; base < n_sub; base += kThreads) {
			const uint32_t i = base + threadIdx.x;
			const uint32_t c = i < n_sub ? meta[i * kMeta + M_NEW] : 0;
			uint32_t incl = c;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane_id() >= (uint32_t)o)
					incl += y;
			}
			if (lane_id() == 31)
				H.warp_sums[warp] = incl;
			__syncthreads();
			if (warp == 0) {
				const uint32_t ws = H.warp_sums[lane_id()];
				uint32_t wi = ws;
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
					if (lane_id() >= (uint32_t)o)
						wi += y;
				}
				H.warp_sums[lane_id()] = wi - ws;
				if (lane_id() == 31)
					H.batch_total = wi;
			}
			__syncthreads();
			if (i < n_sub)
				meta[i * kMeta + M_OFF] = ntok + H.warp_sums[warp] + incl - c;
			ntok += H.batch_total;
			__syncthreads();
		}
		for (uint32_t sb = warp; sb < n_sub; sb += kThreads / 32) {
			const uint32_t *m = meta + sb * kMeta;
			const uint32_t cnt = m[M_CNT], skip = m[M_SKIP], nnew = m[M_NEW], off = m[M_OFF], tail_n = m[M_TAILN];
			const uint32_t body = nnew - tail_n;                          // tokens copied unchanged
			const uint32_t *src = tokpos + (size_t)sb * kSub + skip;
			(void)cnt;
			for (uint32_t k = lane_id(); k < nnew; k += 32) {
				const uint32_t t = k < body ? src[k] : (k == body ? m[M_TAIL0] : m[M_TAIL1]);
				tok[off + k] = t;
				if (tok_is_match(t)) {
					uint32_t lc, le, lx, dc, de, dx;
					len_code(tok_len(t), lc, le, lx);
					dist_code(tok_dist(t), dc, de, dx);
					atomicAdd(&S.ll_freq[257 + lc], 1u);
					atomicAdd(&S.d_freq[dc], 1u);
				} else {
					atomicAdd(&S.ll_freq[t], 1u);
				}
			}
		}
		if (threadIdx.x == 0)
			S.ll_freq[256] = (J.flags & NXGPU_F_NO_EOB) ? 0 : 1;   // EOB
		__syncthreads();
		const long long thuf0 = clock64();

		// ---- Huffman tables, block type decision ----
		const bool is_final = (J.flags & NXGPU_F_FINAL) != 0;
		const bool no_joiner = (J.flags & NXGPU_F_NO_JOINER) != 0;
		const bool force_fixed = (J.flags & NXGPU_F_FIXED) != 0;
		const bool preset = J.dht != nullptr;
		// a piece of a block that other CTAs write the rest of (one large nxu_run_job descriptor cut into
		// pieces, nxgpu_job.cu): no block header in front and / or no end-of-block behind; caller's or fixed table only
		const bool no_header = (J.flags & NXGPU_F_NO_HEADER) != 0 && (preset || force_fixed);
		const bool no_eob = (J.flags & NXGPU_F_NO_EOB) != 0;
		if (J.lzcount) {
			for (int i = threadIdx.x; i < 316; i += kThreads)
				J.lzcount[i] = i < 286 ? S.ll_freq[i] : S.d_freq[i - 286];
		}
		__syncthreads();
		// every tree needs two coded symbols to be complete (same trick as zlib's build_tree)
		// (the dummies only shape the tree; build_dynamic removes them again before costing)
		if (threadIdx.x == 0) {
			S.misc[0] = S.misc[1] = 0xFFFFu; S.misc[2] = 0;
			if (!preset && !force_fixed) {
				int nz = 0;
				for (int i = 0; i < 30; i++) nz += S.d_freq[i] != 0;
				if (nz == 0) { S.d_freq[0] = 1; S.d_freq[1] = 1; S.misc[0] = 0; S.misc[1] = 1; }
				else if (nz == 1) { const int f = S.d_freq[0] ? 1 : 0; S.d_freq[f] = 1; S.misc[0] = f; }
				nz = 0;
				for (int i = 0; i < 286; i++) nz += S.ll_freq[i] != 0;
				if (nz == 1) { S.ll_freq[0] += 1; S.misc[2] = 1; }    // only EOB present (empty input)
			}
		}
		__syncthreads();

		uint32_t btype, body_bits;
		if (preset) {
			// caller-supplied table (nxu_run_job DHT function codes): lens = dht[0..316), header bits follow
			if (threadIdx.x < 286) H.ll_len[threadIdx.x] = J.dht[threadIdx.x];
			else if (threadIdx.x >= 288 && threadIdx.x < 318) H.d_len[threadIdx.x - 288] = J.dht[threadIdx.x - 2];
			for (int i = threadIdx.x; i < 96; i += kThreads) H.hdr_words[i] = 0;
			__syncthreads();
			assign_codes(H, H.ll_len, 286, H.ll_code);
			assign_codes(H, H.d_len, 30, H.d_code);
			if (threadIdx.x == 0) {
				uint32_t bp = 0;
				if (!no_header)
					put_bits(H.hdr_words, bp, (is_final ? 1u : 0u) | (2u << 1), 3);
				for (uint32_t b = 0; b < J.dht_bits && !no_header; b += 8) {
					uint32_t nb = min(8u, J.dht_bits - b);
					put_bits(H.hdr_words, bp, J.dht[320 + (b >> 3)] & ((1u << nb) - 1), nb);
				}
				H.hdr_bits = bp;
			}
			__syncthreads();
			uint32_t bits = 0;
			if (threadIdx.x < 286)
				bits = S.ll_freq[threadIdx.x] * (H.ll_len[threadIdx.x] + (threadIdx.x > 256 ? c_lext[threadIdx.x - 257] : 0));
			else if (threadIdx.x >= 288 && threadIdx.x < 318)
				bits = S.d_freq[threadIdx.x - 288] * (H.d_len[threadIdx.x - 288] + c_dext[threadIdx.x - 288]);
			// a needed symbol without a code is the NX "missing code" error (CC=66)
			bool missing = false;
			if (threadIdx.x < 286) missing = S.ll_freq[threadIdx.x] && !H.ll_len[threadIdx.x];
			else if (threadIdx.x >= 288 && threadIdx.x < 318) missing = S.d_freq[threadIdx.x - 288] && !H.d_len[threadIdx.x - 288];
			const int any_missing = __syncthreads_or(missing);
			for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
			if (lane_id() == 0) H.warp_sums[warp] = bits;
			__syncthreads();
			body_bits = H.hdr_bits;
			for (int i = 0; i < 10; i++) body_bits += H.warp_sums[i];
			__syncthreads();
			btype = 2;
			if (any_missing) {
				if (threadIdx.x == 0) { DeflateOut o = {}; o.rc = 66; o.n_tokens = ntok; outs[job] = o; }
				__syncthreads();
				if (so.dst) stream_out(S, so, J, job, n_jobs, 0);
				continue;
			}
		} else if (force_fixed) {
			body_bits = fixed_cost_and_tables(S, H, true) - (no_header ? 3 : 0);
			btype = 1;
		} else {
			const uint32_t dyn_bits = build_dynamic(S, H, is_final ? 1u : 0u);
			const uint32_t fix_bits = fixed_cost_and_tables(S, H, false);
			if (fix_bits <= dyn_bits) {
				body_bits = fixed_cost_and_tables(S, H, true);
				btype = 1;
			} else {
				body_bits = dyn_bits;
				btype = 2;
			}
		}
		// tail: joiner = empty stored block (3 header bits, pad, 00 00 FF FF) as in lib/nx_deflate.c:220-243
		uint32_t total_bits = body_bits;
		if (!is_final && !no_joiner)
			total_bits = ((total_bits + 3 + 7) & ~7u) + 32;
		uint32_t total_bytes = (total_bits + 7) >> 3;
		const uint32_t stored_bytes = J.src_len + 5 * ((J.src_len + 65534) / 65535 + (J.src_len == 0));
		const bool allow_stored = !preset && !force_fixed && !no_joiner;
		if (allow_stored && stored_bytes < total_bytes) {
			btype = 0;
			total_bytes = stored_bytes;
			total_bits = stored_bytes * 8;
		}
		if (total_bytes > J.out_cap) {
			if (threadIdx.x == 0) { DeflateOut o = {}; o.rc = NXGPU_E_BUF; o.out_len = total_bytes; o.n_tokens = ntok; outs[job] = o; }
			__syncthreads();
			if (so.dst) stream_out(S, so, J, job, n_jobs, 0);
			continue;
		}

		if (btype == 0) {
			emit_stored(J, is_final);
		} else {
			Packer P;
			P.stage = S.ring32;
			P.out32 = reinterpret_cast<uint32_t *>(J.out);
			P.words_out = 0;
			P.carry = 0;
			for (int i = threadIdx.x; i < kStageWords; i += kThreads)
				P.stage[i] = 0;
			__syncthreads();
			if (btype == 2) {
				copy_bits_to_stage(P, H.hdr_words, H.hdr_bits);
				packer_flush(P, H.hdr_bits);
			} else {
				if (!no_header)
					put_tail_bits(P, (is_final ? 1u : 0u) | (1u << 1), 3);
			}
			encode_tokens(S, H, P, tok, ntok);
			if (!no_eob)
				put_tail_bits(P, H.ll_code[256], H.ll_len[256]);
			if (!is_final && !no_joiner) {
				const uint32_t used = (P.words_out * 32 + P.carry);
				const uint32_t pad = (8 - ((used + 3) & 7)) & 7;
				put_tail_bits(P, 0, 3 + pad);
				put_tail_bits(P, 0xFFFF0000ull, 32);
			}
			// last partial word, byte by byte
			__syncthreads();
			if (threadIdx.x == 0) {
				const uint32_t tailbytes = (P.carry + 7) >> 3;
				const uint32_t v = P.stage[0];
				for (uint32_t b = 0; b < tailbytes; b++)
					J.out[P.words_out * 4 + b] = (uint8_t)(v >> (8 * b));
			}
		}
		// the tokens of this chunk are dead now (every thread is past the barriers of the packer)
		l2_discard_lines(tok, (ntok * 4 + 127) & ~127u);
		for (uint32_t sb = threadIdx.x; sb < n_sub; sb += kThreads) {
			const uint32_t lines = (meta[sb * kMeta + M_CNT] * 4 + 127) >> 7;
			const char *tp = reinterpret_cast<const char *>(tokpos + (size_t)sb * kSub);
			for (uint32_t l = 0; l < lines; l++)
				asm volatile("discard.global.L2 [%0], 128;\n" ::"l"(tp + 128 * l) : "memory");
		}
		if (warp == 2) { DBG_ADD(5, clock64() - thuf0); DBG_ADD(6, clock64() - tjob0); }
		if (threadIdx.x == 0) {
			DeflateOut o = {};
			o.rc = 0;
			o.out_len = total_bytes;
			o.tebc = total_bits & 7;
			o.n_tokens = ntok;
			o.btype = btype;
			outs[job] = o;
		}
		__syncthreads();
		if (so.dst) {
			stream_out(S, so, J, job, n_jobs, total_bytes);
			__syncthreads();
			// the slot has been copied into the stream: drop its lines too (slots are 128-byte aligned and padded)
			if ((reinterpret_cast<uintptr_t>(J.out) & 127) == 0)
				l2_discard_lines(J.out, (total_bytes + 127) & ~127u);
		}
	}
}

// ---- DHT generation on its own (lib/nx_dhtgen.c:945 dhtgen(), called from lib/nx_dht.c:632 on a cache miss) ----
// One CTA per histogram: 286 lit/len + 30 distance counts in, the RFC 1951 §3.2.7 dynamic header from
// HLIT on (the bytes of cpb.in_dht, inc_nx/nxu.h:300-310) and its length in bits out.  Same
// length-limited Huffman construction the deflate kernel uses for its own blocks (build_dynamic).
__global__ void __launch_bounds__(kThreads, 1)
dhtgen_kernel(const uint32_t *__restrict__ counts, uint32_t n, uint8_t *__restrict__ dht_out, uint32_t *__restrict__ dht_bits)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	Smem &S = *reinterpret_cast<Smem *>(smem_raw);
	HuffScratch &H = *reinterpret_cast<HuffScratch *>(S.prev);
	for (uint32_t b = blockIdx.x; b < n; b += gridDim.x) {
		const uint32_t *c = counts + (size_t)b * 316;
		for (int i = threadIdx.x; i < 288; i += kThreads)
			S.ll_freq[i] = i < 286 ? c[i] : 0;
		if (threadIdx.x < 32)
			S.d_freq[threadIdx.x] = threadIdx.x < 30 ? c[286 + threadIdx.x] : 0;
		__syncthreads();
		if (threadIdx.x == 0) {
			if (S.ll_freq[256] == 0)
				S.ll_freq[256] = 1;                           // every block ends with an end-of-block symbol
			// every tree needs two coded symbols to be complete (as in deflate_kernel)
			S.misc[0] = S.misc[1] = 0xFFFFu; S.misc[2] = 0;
			int nz = 0;
			for (int i = 0; i < 30; i++) nz += S.d_freq[i] != 0;
			if (nz == 0) { S.d_freq[0] = 1; S.d_freq[1] = 1; S.misc[0] = 0; S.misc[1] = 1; }
			else if (nz == 1) { const int f = S.d_freq[0] ? 1 : 0; S.d_freq[f] = 1; S.misc[0] = f; }
			nz = 0;
			for (int i = 0; i < 286; i++) nz += S.ll_freq[i] != 0;
			if (nz == 1) { S.ll_freq[0] += 1; S.misc[2] = 1; }
		}
		__syncthreads();
		build_dynamic(S, H, 0);
		// drop the 3 block-header bits in front: cpb.in_dht starts at HLIT
		const uint32_t nbits = H.hdr_bits - 3;
		for (uint32_t i = threadIdx.x; i < 288; i += kThreads) {
			uint32_t v = 0;
			if (8 * i < nbits) {
				const uint32_t bit = 8 * i + 3, w = bit >> 5, o = bit & 31;
				v = H.hdr_words[w] >> o;
				if (o > 24)
					v |= H.hdr_words[w + 1] << (32 - o);
				v &= 0xff;
				if (nbits - 8 * i < 8)
					v &= (1u << (nbits - 8 * i)) - 1;
			}
			dht_out[(size_t)b * 288 + i] = (uint8_t)v;
		}
		if (threadIdx.x == 0)
			dht_bits[b] = nbits;
		__syncthreads();
	}
}

} // namespace

cudaError_t launch_dhtgen(const uint32_t *counts, uint32_t n, uint8_t *dht_out, uint32_t *dht_bits, cudaStream_t s)
{
	static PerDeviceOnce once;
	cudaError_t e0 = once.run([] { return cudaFuncSetAttribute(dhtgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)); });
	if (e0 != cudaSuccess)
		return e0;
	dhtgen_kernel<<<n < (uint32_t)kNumSMs ? n : kNumSMs, kThreads, sizeof(Smem), s>>>(counts, n, dht_out, dht_bits);
	return cudaGetLastError();
}

size_t deflate_smem_bytes() { return sizeof(Smem); }
size_t deflate_scratch_words(uint32_t tok_stride) { return scratch_words(tok_stride); }

cudaError_t launch_deflate(const DeflateJob *jobs, DeflateOut *outs, uint32_t n_jobs, int level,
			   uint32_t *tok_scratch, uint32_t tok_stride, int grid, cudaStream_t s,
			   uint32_t *job_counter, const uint32_t *ready, uint32_t jobs_per_flag, const StreamOut *so)
{
	static PerDeviceOnce once;
	cudaError_t e0 = once.run([] { return cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)); });
	if (e0 != cudaSuccess)
		return e0;
	LevelParams lp = level_params(level);
	if (const char *ov = getenv("NXGPU_LZ_PARAMS")) {         // developer override: "depth,lazy,nice"
		int a, b, c;
		int d = 0;
		const int got = sscanf(ov, "%d,%d,%d,%d", &a, &b, &c, &d);
		if (got >= 3 && a >= 1) { lp.depth = a; lp.lazy = b; lp.nice = c; lp.d1 = got == 4 ? d : 0; }
	}
	static const bool dbg = getenv("NXGPU_DEBUG_CYCLES") != nullptr;
	unsigned long long *d_dbg = nullptr;
	if (dbg) {
		cudaMalloc(&d_dbg, (size_t)grid * 64);
		cudaMemsetAsync(d_dbg, 0, (size_t)grid * 64, s);
		cudaMemcpyToSymbolAsync(g_dbg, &d_dbg, sizeof(d_dbg), 0, cudaMemcpyHostToDevice, s);
	}
	// warp 0 stages the input and hashes it, warp 1 links the chains, warps 2-31 parse (NXGPU_PRODUCER_SPLIT=0: warp 0 does both
	// halves of the build and warp 1 parses — the A/B switch behind profiles/)
	static const bool split = !(getenv("NXGPU_PRODUCER_SPLIT") && atoi(getenv("NXGPU_PRODUCER_SPLIT")) == 0);
	uint32_t parser_mask = split ? 0xFFFFFFFCu : 0xFFFFFFFEu;
	// The single-pass levels are bound by the chain build, not by the parsers: warps 4 and 5 — the first parsers on the
	// schedulers of warps 0 and 1 — stay idle there and leave those issue slots to the build (level 1: 64.9 -> 67.0 GB/s;
	// at level 6, which is parser-bound, the same mask costs 2 %).
	if (split && lp.d1 == 0)
		parser_mask = 0xFFFFFFCCu;
	if (const char *pm = getenv("NXGPU_PARSER_MASK"))
		parser_mask = (uint32_t)strtoul(pm, nullptr, 0) & (split ? 0xFFFFFFFCu : 0xFFFFFFFEu);
	if (parser_mask == 0)
		parser_mask = split ? 4 : 2;
	if (split)
		parser_mask |= 1;                                 // bit 0 = split (warp 0 never parses)
	cudaError_t me = cudaMemsetAsync(job_counter, 0, sizeof(uint32_t), s);
	if (me != cudaSuccess)
		return me;
	// the run probe (bit 9) rides in the high bits of d1 (NXGPU_RUN_PROBE=0: developer switch)
	static const bool use_rep = !(getenv("NXGPU_RUN_PROBE") && atoi(getenv("NXGPU_RUN_PROBE")) == 0);
	// bit 10: windows wholly covered by the current shallow token are not searched (NXGPU_SKIP_COVERED=0: developer switch)
	static const bool skip_cov = !(getenv("NXGPU_SKIP_COVERED") && atoi(getenv("NXGPU_SKIP_COVERED")) == 0);
	// bit 11: the shallow pass searches every second position (NXGPU_STRIDE2=0: developer switch)
	const int d1f = lp.d1 | (use_rep ? 0x200 : 0) | (skip_cov ? 0x400 : 0);
	deflate_kernel<<<grid, kThreads, sizeof(Smem), s>>>(jobs, outs, n_jobs, lp.depth, lp.lazy, lp.nice, tok_scratch, tok_stride, parser_mask,
							    job_counter, ready, jobs_per_flag ? jobs_per_flag : 1, d1f, so ? *so : StreamOut());
	cudaError_t e = cudaGetLastError();
	if (dbg) {
		std::vector<unsigned long long> h((size_t)grid * 8);
		cudaStreamSynchronize(s);
		cudaMemcpy(h.data(), d_dbg, (size_t)grid * 64, cudaMemcpyDeviceToHost);
		unsigned long long t[8] = { 0 };
		for (int b = 0; b < grid; b++) for (int k = 0; k < 8; k++) t[k] += h[(size_t)b * 8 + k];
		const double st = t[4] ? (double)t[4] : 1.0;
		(void)st;
		const double nj = n_jobs;
		fprintf(stderr, "[nxgpu cycles/job] level %d (depth %d lazy %d nice %d d1 %d): chain link/build %.0f wait-data %.0f wait-space %.0f | parser(w2) busy %.0f wait %.0f windows %.0f | huff+pack %.0f total %.0f\n",
			level, lp.depth, lp.lazy, lp.nice, lp.d1, t[0] / nj, t[1] / nj, t[2] / nj, t[3] / nj, t[4] / nj, t[7] / nj, t[5] / nj, t[6] / nj);
		d_dbg = nullptr;
		cudaMemcpyToSymbol(g_dbg, &d_dbg, sizeof(d_dbg));
	}
	return e;
}

} // namespace nxgpu
