// deflate.cu — the NX compress function (inc_nx/nxu.h:803-811) and the
// dynamic-Huffman-table generation of lib/nx_dhtgen.c:945 (row a6) as ONE persistent sm_100a
// kernel: one CTA per SM, each CTA compresses one job (chunk) at a time entirely out of
// shared memory.
//
//   LZ77     exact hash chains (4-byte hash, 13 bits) over a 64 KiB shared-memory ring of the
//            input (cp.async-staged, 16-byte vectorised) with a 64 Ki-entry u16 chain ring.
//            Warp 0 inserts positions in stream order, 32 per iteration, resolving intra-warp
//            predecessors with match.any; all 32 warps then search the chains for EVERY
//            position of the step in parallel (one lane per position); warp 1 turns the per-
//            position best matches into the greedy/lazy token stream with an in-warp
//            pointer-jumping reachability scan (no serial token walk).
//   Huffman  lit/len + dist histograms accumulate in shared memory during the parse; code
//            lengths come from a rank sort, a two-queue merge and a Kraft-sum length limiter;
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
constexpr int kWin = 30;                  // matcher warps = windows of 32 positions per step
constexpr int kStep = kWin * 32;          // positions matched per CTA step
constexpr uint32_t kRingMask = 0xFFFFu;   // 64 KiB data ring / 64 Ki-entry chain ring
constexpr int kHashBits = 12;
constexpr int kHashSize = 1 << kHashBits;
constexpr uint32_t kNone = 0xFFFFFFFFu;    // empty hash head
constexpr int kStageWords = 2048;

struct __align__(16) Smem {
	uint32_t ring32[16384];           // 64 KiB input ring (position & 0xFFFF)
	uint16_t prev[65536];             // 128 KiB: distance to the previous position with the same hash
	uint32_t head[kHashSize];         // 16 KiB: most recent position per hash (absolute, kNone = empty)
	// matcher -> parser hand-off for the current step: per position, where the token stream leaves
	// the window (low 10 bits) and how many tokens it emits on the way (high 6 bits)
	uint16_t JC[kStep];
	volatile uint32_t win_flag[32];   // == step + 1 once the window's JC entries are written
	// parser -> matcher: per window the entry lane (>= 32: jumped over) and first token index
	volatile uint32_t ent[2][32];
	volatile uint32_t off[2][32];
	uint32_t ll_freq[288];
	uint32_t d_freq[32];
	int grp_ctr[2];
	uint32_t n_tok;
	uint32_t misc[13];
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

// optional cycle accounting per CTA (NXGPU_DEBUG_CYCLES=1): [0] build, [1] parse, [2] match (warp 2),
// [3] barrier wait (warp 2), [4] steps, [5] huffman+pack, [6] whole job
__device__ unsigned long long *g_dbg = nullptr;
#define DBG_ADD(slot, v) do { if (g_dbg && lane_id() == 0) atomicAdd(&g_dbg[blockIdx.x * 8 + (slot)], (unsigned long long)(v)); } while (0)

__device__ __forceinline__ uint32_t load4(const uint32_t *ring32, uint32_t pos)
{
	uint32_t a = (pos >> 2) & 0x3FFF, b = (a + 1) & 0x3FFF;
	return __funnelshift_r(ring32[a], ring32[b], (pos & 3) * 8);
}
__device__ __forceinline__ uint32_t hash4(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - kHashBits); }
__device__ __forceinline__ uint32_t ring_byte(const uint32_t *ring32, uint32_t pos)
{
	return reinterpret_cast<const uint8_t *>(ring32)[pos & kRingMask];
}

// ---- stage [lo, hi) (multiples of 16, relative to gbase) of the input into the ring ----
__device__ void stage_input(Smem &S, const uint8_t *gbase, uint32_t lo, uint32_t hi, uint32_t valid_lo, uint32_t valid_hi)
{
	for (uint32_t p = lo + threadIdx.x * 16; p < hi; p += kThreads * 16) {
		uint8_t *dst = reinterpret_cast<uint8_t *>(S.ring32) + (p & kRingMask);
		if (p >= valid_lo && p + 16 <= valid_hi) {
			uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gbase + p));
		} else {
			// edge block: touch only bytes that belong to the caller
			for (int k = 0; k < 16; k++) {
				uint32_t q = p + k;
				dst[k] = (q >= valid_lo && q < valid_hi) ? gbase[q] : 0;
			}
		}
	}
	asm volatile("cp.async.commit_group;\n" ::);
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;\n" ::); }

// ---- warp 0: insert positions [lo, hi) into the hash chains, in stream order ----
// One shared-memory atomic exchange per position: the returned value is the previous head, and
// when several lanes of the warp hit the same bucket the hardware applies them one after the
// other, so each lane receives the position of the lane applied just before it.  If that order
// was ascending (checked below) the links are exactly those of a sequential insert; otherwise
// the warp repairs the bucket with match.any (exact, just slower).
__device__ void build_chains(Smem &S, uint32_t lo, uint32_t hi)
{
	const uint32_t lane = lane_id();
	for (uint32_t p0 = lo; p0 < hi; p0 += 32) {
		const uint32_t pos = p0 + lane;
		const bool valid = pos < hi;
		const uint32_t h = hash4(load4(S.ring32, pos));
		uint32_t old = kNone;
		if (valid)
			old = atomicExch(&S.head[h], pos);
		const bool same_iter = valid && old != kNone && old >= p0;
		if (__ballot_sync(0xffffffffu, same_iter && old > pos)) {
			// out-of-order application: rebuild this iteration's links exactly
			const uint32_t key = valid ? h : (0x80000000u | lane);
			const uint32_t grp = __match_any_sync(0xffffffffu, key);
			const uint32_t outside = __ballot_sync(0xffffffffu, valid && !same_iter);
			const int src = __ffs(grp & outside) - 1;            // the lane that saw the pre-iteration head
			const uint32_t pre = __shfl_sync(0xffffffffu, old, src < 0 ? 0 : src);
			const uint32_t lower = grp & ((1u << lane) - 1);
			old = lower ? p0 + (31 - __clz(lower)) : (src < 0 ? kNone : pre);
			if (valid && (grp >> lane) == 1u)
				S.head[h] = pos;
		}
		if (valid) {
			const uint32_t d = pos - old;
			S.prev[pos & kRingMask] = (uint16_t)((old != kNone && d <= (uint32_t)kWindow) ? d : 0);
		}
		__syncwarp();
	}
}

__device__ __forceinline__ uint32_t match_length(const uint32_t *ring32, uint32_t p, uint32_t q, uint32_t maxl)
{
	uint32_t pa = p >> 2, qa = q >> 2;
	const uint32_t ps = (p & 3) * 8, qs = (q & 3) * 8;
	uint32_t pw0 = ring32[pa & 0x3FFF], qw0 = ring32[qa & 0x3FFF];
	uint32_t l = 0;
	while (l < maxl) {
		uint32_t pw1 = ring32[(pa + 1) & 0x3FFF], qw1 = ring32[(qa + 1) & 0x3FFF];
		uint32_t x = __funnelshift_r(pw0, pw1, ps) ^ __funnelshift_r(qw0, qw1, qs);
		if (x) {
			l += (__ffs(x) - 1) >> 3;
			break;
		}
		l += 4; pa++; qa++; pw0 = pw1; qw0 = qw1;
	}
	return l < maxl ? l : maxl;
}

// ---- matcher warps: best match for each of the 32 positions of one window ----
// Phase 1 (one lane per position): walk the hash chain, comparing at most kCap bytes per
// candidate — a candidate that reaches kCap ends the first pass, like zlib's nice_length.
// Phase 2 (whole warp): consecutive positions whose capped match has the SAME distance lie
// inside one long repeat; only the first of them is extended (32 lanes x 4 bytes per pass) and
// the followers derive their length from it.  Without this every position inside a 258-byte
// repeat re-compares up to 258 bytes on a single lane.
// Phase 3: positions inside such a repeat keep walking their chain, but only a candidate that
// also matches at offset `bl` (beyond the inherited length) is compared in full.
constexpr uint32_t kCap = 32;

// bytes [0, capl) of p and q, capl <= 32; loads are issued in two independent batches so a
// compare costs about two shared-memory round trips instead of one per word
__device__ __forceinline__ uint32_t match_len_cap(const uint32_t *ring32, uint32_t p, uint32_t q, uint32_t capl)
{
	const uint32_t pa = p >> 2, qa = q >> 2;
	const uint32_t ps = (p & 3) * 8, qs = (q & 3) * 8;
	uint32_t pw[5], qw[5];
#pragma unroll
	for (int i = 0; i < 5; i++) {
		pw[i] = ring32[(pa + i) & 0x3FFF];
		qw[i] = ring32[(qa + i) & 0x3FFF];
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const uint32_t x = __funnelshift_r(pw[i], pw[i + 1], ps) ^ __funnelshift_r(qw[i], qw[i + 1], qs);
		if (x)
			return min(capl, 4u * i + ((uint32_t)(__ffs(x) - 1) >> 3));
	}
	if (capl <= 16)
		return capl;
	uint32_t pv[5], qv[5];
	pv[0] = pw[4]; qv[0] = qw[4];
#pragma unroll
	for (int i = 1; i < 5; i++) {
		pv[i] = ring32[(pa + 4 + i) & 0x3FFF];
		qv[i] = ring32[(qa + 4 + i) & 0x3FFF];
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const uint32_t x = __funnelshift_r(pv[i], pv[i + 1], ps) ^ __funnelshift_r(qv[i], qv[i + 1], qs);
		if (x)
			return min(capl, 16u + 4u * i + ((uint32_t)(__ffs(x) - 1) >> 3));
	}
	return capl;
}

struct WinResult {       // what a matcher lane keeps in registers until its window is emitted
	uint32_t token;      // literal byte or tok_match(len, dist)
	uint32_t reach;      // positions of the window visited when the stream enters at this lane
};

__device__ WinResult match_window(Smem &S, uint32_t g, uint32_t p0, uint32_t step_hi, uint32_t PE, uint32_t valid_lo,
				  uint32_t step_seq, int depth, int nice, int lazy)
{
	const uint32_t lane = lane_id();
	const uint32_t nice_eff = min((uint32_t)nice, kCap);
	const uint32_t pos = p0 + lane;
	const bool live = pos < step_hi;
	const uint32_t maxl = live ? min((uint32_t)kMaxMatch, PE - pos) : 0;
	uint32_t bl = kMinMatch - 1, bd = 0;
	uint32_t acc = 0;
	int hop = 0;
	const uint32_t maxdist = min((uint32_t)kWindow, pos - valid_lo);
	uint32_t d = (maxl >= (uint32_t)kMinMatch) ? S.prev[pos & kRingMask] : 0;   // next chain link (prefetched)
	if (d) {
		const uint32_t capl = min(maxl, kCap);
		// filter: the 4 bytes ending at offset bl must match (for bl == 3 that is the 4-gram itself,
		// which weeds out hash collisions; later it is the tail that a longer match needs)
		uint32_t endw = load4(S.ring32, pos + bl - 3);
		for (; hop < depth; hop++) {
			if (d == 0) { hop = depth; break; }
			acc += d;
			if (acc > maxdist) { hop = depth; d = 0; break; }
			const uint32_t q = pos - acc;
			d = S.prev[q & kRingMask];                         // independent of the check below
			if (load4(S.ring32, q + bl - 3) != endw)
				continue;
			const uint32_t len = match_len_cap(S.ring32, pos, q, capl);
			if (len > bl) {
				bl = len; bd = acc;
				if (len >= nice_eff || len >= capl) { hop++; break; }
				endw = load4(S.ring32, pos + bl - 3);
			}
		}
	}
	// ---- phase 2 ----
	const bool capped = (bl == kCap) && (maxl > kCap);
	{
		const uint32_t d_prev = __shfl_up_sync(0xffffffffu, bd, 1);
		const bool c_prev = __shfl_up_sync(0xffffffffu, (int)capped, 1) != 0 && lane > 0;
		const bool head = capped && !(c_prev && d_prev == bd);
		const uint32_t heads_all = __ballot_sync(0xffffffffu, head);
		uint32_t heads = heads_all;
		const uint32_t myhead = capped ? 31 - __clz(heads_all & ((2u << lane) - 1)) : 32;
		while (heads) {
			const int h = __ffs(heads) - 1;
			heads &= heads - 1;
			const uint32_t hp = p0 + h;
			const uint32_t hd = __shfl_sync(0xffffffffu, bd, h);
			const uint32_t hmax = min((uint32_t)kMaxMatch + 31, PE - hp);
			uint32_t L = kCap;
			for (uint32_t base = kCap; base < hmax; base += 128) {
				const uint32_t off = base + 4 * lane;
				const uint32_t x = load4(S.ring32, hp + off) ^ load4(S.ring32, hp - hd + off);
				uint32_t e = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4;
				const uint32_t room = off < hmax ? min(4u, hmax - off) : 0;
				e = min(e, room);
				const uint32_t stop = __ballot_sync(0xffffffffu, e < 4);
				if (stop == 0) { L = base + 128; continue; }
				const int f = __ffs(stop) - 1;
				L = base + 4 * f + __shfl_sync(0xffffffffu, e, f);
				break;
			}
			if (myhead == (uint32_t)h)
				bl = min(maxl, L - (lane - h));
		}
	}
	// ---- phase 3 ----
	if (capped && bl < maxl && bl < (uint32_t)nice) {
		uint32_t endb = ring_byte(S.ring32, pos + bl);
		for (; hop < depth; hop++) {
			if (d == 0)
				break;
			acc += d;
			if (acc > maxdist)
				break;
			const uint32_t q = pos - acc;
			d = S.prev[q & kRingMask];
			if (ring_byte(S.ring32, q + bl) != endb)
				continue;
			const uint32_t len = match_length(S.ring32, pos, q, maxl);
			if (len > bl) {
				bl = len; bd = acc;
				if (len >= (uint32_t)nice || len >= maxl)
					break;
				endb = ring_byte(S.ring32, pos + bl);
			}
		}
	}

	// ---- greedy / lazy choice per position, then reach sets by pointer jumping ----
	const uint32_t len = (bl >= (uint32_t)kMinMatch) ? bl : 0;
	uint32_t nlen = __shfl_down_sync(0xffffffffu, len, 1);
	if (lane == 31)
		nlen = 0;                                  // no look-ahead across windows
	bool take = len != 0;
	if (take && lazy && len < (uint32_t)lazy && nlen > len)
		take = false;
	uint32_t J = live ? lane + (take ? len : 1) : 64;       // next token start relative to p0 (>= 32: leaves the window)
	uint32_t Rl = live ? (1u << lane) : 0;
#pragma unroll
	for (int k = 0; k < 5; k++) {
		const uint32_t tJ = __shfl_sync(0xffffffffu, J, J & 31);
		const uint32_t tR = __shfl_sync(0xffffffffu, Rl, J & 31);
		if (J < 32) { Rl |= tR; J = tJ; }
	}
	// the parser warp only needs, per possible entry lane, the exit and the token count
	S.JC[g * 32 + lane] = (uint16_t)(J | ((uint32_t)__popc(Rl) << 10));
	__threadfence_block();
	__syncwarp();
	if (lane == 0)
		S.win_flag[g] = step_seq;
	WinResult r;
	r.token = take ? tok_match(len, bd) : ring_byte(S.ring32, pos);
	r.reach = Rl;
	return r;
}

// a matcher warp writes out the tokens of the window it matched in the PREVIOUS step: by now the
// parser has published where the stream entered that window (lane, or >= 32 if it was jumped over)
// and the index of its first token
__device__ __forceinline__ void emit_window(Smem &S, const WinResult &r, uint32_t entry, uint32_t off, uint32_t *tok)
{
	if (entry >= 32)
		return;
	const uint32_t lane = lane_id();
	const uint32_t R = __shfl_sync(0xffffffffu, r.reach, entry);
	if ((R >> lane) & 1) {
		const uint32_t t = r.token;
		if (tok_is_match(t)) {
			uint32_t lc, le, lx, dc, de, dx;
			len_code(tok_len(t), lc, le, lx);
			dist_code(tok_dist(t), dc, de, dx);
			atomicAdd(&S.ll_freq[257 + lc], 1u);
			atomicAdd(&S.d_freq[dc], 1u);
		} else {
			atomicAdd(&S.ll_freq[t], 1u);
		}
		tok[off + __popc(R & ((1u << lane) - 1))] = t;
	}
}

// ---- warp 1: follow the token stream through the windows of one step as the matchers finish them;
//      publishes per window the entry lane and the index of its first token ----
__device__ void parse_step(Smem &S, uint32_t step_lo, uint32_t step_hi, uint32_t step_seq, uint32_t &cur, uint32_t &ntok)
{
	const uint32_t nwin = (step_hi - step_lo + 31) / 32;
	const uint32_t par = step_seq & 1;
	for (uint32_t g = 0; g < nwin; g++) {
		const uint32_t p0 = step_lo + g * 32;
		while (S.win_flag[g] != step_seq)
			;
		__threadfence_block();
		uint32_t entry = 64;
		const uint32_t off = ntok;
		if (cur < p0 + 32 && cur < step_hi) {
			entry = cur - p0;
			const uint32_t jc = S.JC[g * 32 + entry];
			cur = p0 + (jc & 1023);
			ntok += jc >> 10;
		}
		if (lane_id() == 0) {
			S.ent[par][g] = entry;
			S.off[par][g] = off;
		}
	}
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
		// two-queue merge (leaves ascending in w[0..nl), internal nodes appended)
		int q1 = 0, q2 = nl, tot = nl;
		while ((nl - q1) + (tot - q2) > 1) {
			int a, b;
			if (q1 < nl && (q2 >= tot || H.w[q1] <= H.w[q2])) a = q1++; else a = q2++;
			if (q1 < nl && (q2 >= tot || H.w[q1] <= H.w[q2])) b = q1++; else b = q2++;
			H.w[tot] = H.w[a] + H.w[b];
			H.parent[a] = (uint16_t)tot;
			H.parent[b] = (uint16_t)tot;
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
	for (uint32_t i = threadIdx.x; i < (uint32_t)kStageWords; i += kThreads)
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
	for (uint32_t base = 0; base < ntok; base += kThreads) {
		const uint32_t i = base + threadIdx.x;
		uint64_t v = 0;
		uint32_t n = 0;
		if (i < ntok) {
			const uint32_t t = tok[i];
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

__global__ void __launch_bounds__(kThreads, 1)
deflate_kernel(const DeflateJob *__restrict__ jobs, DeflateOut *__restrict__ outs, uint32_t n_jobs,
	       int depth, int lazy, int nice, uint32_t *tok_scratch, uint32_t tok_stride)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	Smem &S = *reinterpret_cast<Smem *>(smem_raw);
	HuffScratch &H = *reinterpret_cast<HuffScratch *>(S.prev);
	const uint32_t warp = threadIdx.x >> 5;
	uint32_t *tok = tok_scratch + (size_t)blockIdx.x * tok_stride;

	for (uint32_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
		const long long tjob0 = clock64();
		const DeflateJob J = jobs[job];
		const uintptr_t first = reinterpret_cast<uintptr_t>(J.src) - J.hist_len;
		const uint8_t *gbase = reinterpret_cast<const uint8_t *>(first & ~(uintptr_t)15);
		const uint32_t P0 = (uint32_t)(first & 15);      // first dictionary byte
		const uint32_t PS = P0 + J.hist_len;              // first byte to compress
		const uint32_t PE = PS + J.src_len;
		const uint32_t PEa = (PE + 15) & ~15u;
		const uint32_t hash_hi = PE >= 3 ? PE - 3 : 0;    // positions with 4 bytes available
		const uint32_t nsteps = (J.src_len + kStep - 1) / kStep;

		// ---- init ----
		for (int i = threadIdx.x; i < kHashSize; i += kThreads)
			S.head[i] = kNone;
		for (int i = threadIdx.x; i < 288; i += kThreads)
			S.ll_freq[i] = 0;
		if (threadIdx.x < 32)
			S.d_freq[threadIdx.x] = 0;
		if (threadIdx.x < 32)
			S.win_flag[threadIdx.x] = 0;
		if (threadIdx.x == 0) { S.grp_ctr[0] = 0; S.grp_ctr[1] = 0; S.n_tok = 0; }
		uint32_t DF = min(PEa, (PS + 3 * kStep + 15) & ~15u);   // data staged through here
		stage_input(S, gbase, 0, DF, P0, PE);
		stage_wait();
		__syncthreads();
		uint32_t BF = P0;                                        // warp-0 state: chains built up to here
		uint32_t cur = PS, ntok = 0;                             // warp-1 state: next token start, tokens so far
		WinResult held = { 0, 0 };                               // matcher state: last step's window, not yet emitted
		bool have_held = false;
		const uint32_t g = warp - 2;                             // matcher warps 2..31 own window g of every step
		if (warp == 0) {
			uint32_t hi = min(PS + kStep, hash_hi);
			if (hi > BF) { build_chains(S, BF, hi); BF = hi; }
		}
		__syncthreads();

		// ---- LZ77 pipeline: warp 0 build(s+1) | warp 1 parse(s) | warps 2.. emit(s-1), match(s) | stage(s+3) ----
		for (uint32_t s = 0; s < nsteps; s++) {
			const uint32_t step_lo = PS + s * kStep;
			const uint32_t step_hi = min(step_lo + kStep, PE);
			const uint32_t want = min(PEa, (PS + (s + 4) * kStep + 15) & ~15u);
			if (want > DF) { stage_input(S, gbase, DF, want, P0, PE); DF = want; }
			long long tc1 = clock64();
			if (warp == 0) {
				uint32_t hi = min(PS + (s + 2) * kStep, hash_hi);
				if (hi > BF) { build_chains(S, BF, hi); BF = hi; }
				DBG_ADD(0, clock64() - tc1);
			} else if (warp == 1) {
				parse_step(S, step_lo, step_hi, s + 1, cur, ntok);
				DBG_ADD(1, clock64() - tc1);
			} else {
				if (have_held)
					emit_window(S, held, S.ent[s & 1][g], S.off[s & 1][g], tok);
				const uint32_t p0 = step_lo + g * 32;
				have_held = p0 < step_hi;
				if (have_held)
					held = match_window(S, g, p0, step_hi, PE, P0, s + 1, depth, nice, lazy);
			}
			long long tc2 = clock64();
			stage_wait();
			__syncthreads();
			if (warp == 2) { DBG_ADD(2, tc2 - tc1); DBG_ADD(3, clock64() - tc2); DBG_ADD(4, 1); }
		}
		if (warp >= 2 && have_held)
			emit_window(S, held, S.ent[nsteps & 1][g], S.off[nsteps & 1][g], tok);
		if (threadIdx.x == 32)
			S.n_tok = ntok;
		if (threadIdx.x == 0)
			S.ll_freq[256] = 1;                               // EOB
		__syncthreads();
		ntok = S.n_tok;
		const long long thuf0 = clock64();

		// ---- Huffman tables, block type decision ----
		const bool is_final = (J.flags & NXGPU_F_FINAL) != 0;
		const bool no_joiner = (J.flags & NXGPU_F_NO_JOINER) != 0;
		const bool force_fixed = (J.flags & NXGPU_F_FIXED) != 0;
		const bool preset = J.dht != nullptr;
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
				put_bits(H.hdr_words, bp, (is_final ? 1u : 0u) | (2u << 1), 3);
				for (uint32_t b = 0; b < J.dht_bits; b += 8) {
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
				continue;
			}
		} else if (force_fixed) {
			body_bits = fixed_cost_and_tables(S, H, true);
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
				put_tail_bits(P, (is_final ? 1u : 0u) | (1u << 1), 3);
			}
			encode_tokens(S, H, P, tok, ntok);
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
	}
}

} // namespace

size_t deflate_smem_bytes() { return sizeof(Smem); }

cudaError_t launch_deflate(const DeflateJob *jobs, DeflateOut *outs, uint32_t n_jobs, int level,
			   uint32_t *tok_scratch, uint32_t tok_stride, int grid, cudaStream_t s)
{
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
		if (e != cudaSuccess)
			return e;
		configured = true;
	}
	const LevelParams lp = level_params(level);
	static const bool dbg = getenv("NXGPU_DEBUG_CYCLES") != nullptr;
	unsigned long long *d_dbg = nullptr;
	if (dbg) {
		cudaMalloc(&d_dbg, (size_t)grid * 64);
		cudaMemsetAsync(d_dbg, 0, (size_t)grid * 64, s);
		cudaMemcpyToSymbolAsync(g_dbg, &d_dbg, sizeof(d_dbg), 0, cudaMemcpyHostToDevice, s);
	}
	deflate_kernel<<<grid, kThreads, sizeof(Smem), s>>>(jobs, outs, n_jobs, lp.depth, lp.lazy, lp.nice, tok_scratch, tok_stride);
	cudaError_t e = cudaGetLastError();
	if (dbg) {
		std::vector<unsigned long long> h((size_t)grid * 8);
		cudaStreamSynchronize(s);
		cudaMemcpy(h.data(), d_dbg, (size_t)grid * 64, cudaMemcpyDeviceToHost);
		unsigned long long t[8] = { 0 };
		for (int b = 0; b < grid; b++) for (int k = 0; k < 8; k++) t[k] += h[(size_t)b * 8 + k];
		const double st = t[4] ? (double)t[4] : 1.0;
		fprintf(stderr, "[nxgpu cycles/step] level %d: build %.0f parse %.0f match(w2) %.0f barrier-wait(w2) %.0f | per job: huff+pack %.0f total %.0f (steps/job %.1f)\n",
			level, t[0] / st, t[1] / st, t[2] / st, t[3] / st, (double)t[5] / n_jobs, (double)t[6] / n_jobs, st / n_jobs);
		d_dbg = nullptr;
		cudaMemcpyToSymbol(g_dbg, &d_dbg, sizeof(d_dbg));
	}
	return e;
}

} // namespace nxgpu
