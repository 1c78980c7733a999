// inflate.cu — the NX decompress function (inc_nx/nxu.h:812) as sm_100a kernels.
//
// inflate_kernel: one warp per independent member / sync-point segment, thousands in flight.  The compressed bytes are
// staged into a small shared-memory ring with coalesced loads; lane 0 walks the Huffman stream through shared-memory
// lookup tables (10-bit lit/len, 9-bit distance, canonical fall-back for longer codes) and queues 32 symbols
// (fast_walk), the 32 lanes judge them and form the tokens (walk_batch); then all lanes materialise the queue — a
// shuffle prefix sum gives every symbol its output offset, literals and every match that does not read this batch's own
// output are written byte-parallel (the owner of a byte: one warp-wide OR + a population count), the few short-distance
// matches follow in order.  The window is the output buffer itself (L1/L2-resident), so no history copies are needed
// (the reference's host side copies 32 KiB per job, lib/nx_inflate.c:1633-1687).
//
// inflate_solo_kernel: a few streams, each on a warp pair — a walker and a copier with a token queue between them, the
// window in a shared-memory ring (DuoQueue, duo_copier, ring_fill).
//
// inflate_par.cuh: ONE stream over many warp pairs (block-start search, speculative decode with window markers, chain).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kLitBits = 10, kDistBits = 9;
constexpr uint32_t kInWords = 128;         // per-warp staging ring for the compressed input (512 B)
constexpr uint32_t kReadAhead = 8;         // job mode: source bytes the engine has read behind the point where it stopped by itself
constexpr uint32_t kInAhead = 60;          // the warp tops the ring up (64 words) while fewer words than this lie ahead: a batch consumes at most
                                           // 48 + 3, and the three words of the bit window in front of the read position stay in the ring
                                           // (fast_walk re-reads them): 59 + 64 + 3 <= kInWords

struct WarpTables {
	uint32_t lit[1 << kLitBits];      // ent_make(): bits consumed (code + extra) | codelen<<5 | type<<9 | value<<11   (0 = slow path); type 2
	                                  // literal (value = byte), 3 length (value = base - 3), 1 end of block: bit 10 = the fast loop handles it
	uint32_t dist[1 << kDistBits];    // bits consumed | codelen<<5 | (base - 1)<<11
	uint16_t lit_sorted[288];
	uint16_t dist_sorted[32];
	uint16_t lit_count[16], dist_count[16];
	uint8_t lens[320];
	uint32_t q[32];                   // decoded symbols: literal, or tok_match(len, dist); from fast_walk: the bit window at the symbol
	uint32_t q2[32];                  // fast_walk: the bit window at the symbol's distance code
	uint32_t in[kInWords];            // compressed input, word w of the member at in[w % kInWords]
};

// table entry: the number of bits the symbol part consumes sits in the low five bits, where a funnel shift in wrap mode
// reads it without an extraction
__device__ __forceinline__ uint32_t ent_make(uint32_t tot, uint32_t cl, uint32_t type, uint32_t value) { return tot | (cl << 5) | (type << 9) | (value << 11); }
__device__ __forceinline__ uint32_t ent_tot(uint32_t e) { return e & 31; }
__device__ __forceinline__ uint32_t ent_cl(uint32_t e) { return (e >> 5) & 15; }
__device__ __forceinline__ uint32_t ent_type(uint32_t e) { return (e >> 9) & 3; }
__device__ __forceinline__ uint32_t ent_val(uint32_t e) { return e >> 11; }

__constant__ uint16_t k_len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
__constant__ uint8_t k_len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
__constant__ uint16_t k_dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
__constant__ uint8_t k_dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
__constant__ uint8_t k_clorder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// LSB-first bit reader (lane 0 only).  The compressed bytes come out of a small shared-memory ring
// that the whole warp fills with coalesced loads between symbol batches (stage_input), so the
// decode loop never waits for global memory; when the ring runs dry inside a long header lane 0
// fills it on its own.  Positions are 32-bit words counted from the 4-byte aligned address at or
// below the first source byte; bytes outside [src, src+len) read as zero.
// The bit window is three 32-bit registers (current word, next word, one prefetched from the ring)
// and a bit offset: peeking is one funnel shift, dropping one add and a rarely taken advance — no
// 64-bit shifts on the symbol path.
struct BitReader {
	const uint32_t *base32;   // aligned base
	uint32_t *ring;           // shared, kInWords words
	uint32_t skip;            // first valid byte (0..3) relative to base32
	uint64_t end;             // one past the last valid byte, relative to base32
	uint32_t end_word;        // end / 4: no overrun is possible while wpos <= end_word
	uint32_t wpos;            // words fetched so far: the window holds words wpos-3, wpos-2, wpos-1
	uint32_t staged;          // words [.., staged) are in the ring
	uint32_t w0, w1, w2;
	uint32_t bo;              // bits of w0 already consumed (< 32)

	// word w of the source with the bytes outside the member zeroed (any lane)
	static __device__ __forceinline__ uint32_t load_word(const uint32_t *base32, uint32_t skip, uint64_t end, uint32_t w)
	{
		const uint64_t b0 = (uint64_t)w * 4;
		if (b0 >= end)
			return 0;
		uint32_t v = base32[w];
		if (b0 + 4 > end)
			v &= (1u << ((uint32_t)(end - b0) * 8)) - 1;
		if (b0 < skip)
			v &= ~0u << (skip * 8);
		return v;
	}
	// lane 0 on its own: 8 more words (only when the warp-wide staging did not reach far enough)
	static __device__ __noinline__ void fill_single(const uint32_t *base32, uint32_t skip, uint64_t end, uint32_t *ring, uint32_t staged)
	{
		uint32_t w[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
			w[k] = load_word(base32, skip, end, staged + k);
#pragma unroll
		for (int k = 0; k < 8; k++)
			ring[(staged + k) % kInWords] = w[k];
	}
	__device__ __forceinline__ uint32_t next_word()
	{
		if (wpos == staged) {
			fill_single(base32, skip, end, ring, staged);
			staged += 8;
		}
		return ring[wpos++ % kInWords];
	}
	__device__ __forceinline__ void setup(const uint8_t *s, uint32_t n, uint32_t *ring_)
	{
		const uintptr_t a = reinterpret_cast<uintptr_t>(s);
		base32 = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		skip = (uint32_t)(a & 3);
		end = (uint64_t)skip + n;
		end_word = (uint32_t)(end >> 2);
		ring = ring_;
		wpos = staged = 0; w0 = w1 = w2 = 0; bo = 0;
	}
	// start reading at byte `start` of the member (lane 0); setup() must have run
	__device__ __forceinline__ void seek(uint32_t start)
	{
		const uint64_t p = (uint64_t)skip + start;
		wpos = staged = (uint32_t)(p >> 2);
		w0 = next_word(); w1 = next_word(); w2 = next_word();
		bo = (uint32_t)(p & 3) * 8;
	}
	__device__ __forceinline__ void init(const uint8_t *s, uint32_t n, uint32_t start, uint32_t *ring_)
	{
		setup(s, n, ring_);
		seek(start);
	}
	__device__ __forceinline__ uint32_t peek32() const { return __funnelshift_r(w0, w1, bo); }
	__device__ __forceinline__ uint32_t peek(uint32_t n) const { return peek32() & ((1u << n) - 1); }   // n < 32
	__device__ __forceinline__ void drop(uint32_t n)                                                 // n <= 32
	{
		bo += n;
		if (bo >= 32) {
			w0 = w1; w1 = w2; w2 = next_word();
			bo -= 32;
		}
	}
	__device__ __forceinline__ uint32_t get(uint32_t n) { uint32_t v = peek(n); drop(n); return v; }
	__device__ __forceinline__ void align_byte() { drop((8 - (bo & 7)) & 7); }
	// bits consumed, counted from base32 / from the first byte of the member
	__device__ __forceinline__ uint64_t bits_abs() const { return (uint64_t)(wpos - 3) * 32 + bo; }
	__device__ __forceinline__ uint64_t bits_used() const { return bits_abs() - skip * 8; }
	__device__ __forceinline__ bool overrun() const { return bits_abs() > end * 8; }
	// first byte not yet consumed, at a byte boundary (stored blocks are copied straight from memory)
	__device__ __forceinline__ uint32_t byte_pos() const { return (uint32_t)(bits_abs() >> 3) - skip; }
};

// canonical decode, one bit at a time (codes longer than the primary table)
__device__ int slow_decode(BitReader &br, const uint16_t *count, const uint16_t *sorted)
{
	int code = 0, first = 0, index = 0;
#pragma unroll 1
	for (int l = 1; l <= 15; l++) {
		code |= (int)br.get(1);
		int c = count[l];
		if (code - c < first)
			return sorted[index + (code - first)];
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}

// warp-cooperative: lens[n] -> primary LUT + canonical arrays.  Returns false if over-subscribed.
// strict: also refuse what zlib's inftrees.c refuses — an incomplete set unless its longest code is one bit
__device__ bool build_table(const uint8_t *lens, int n, uint32_t *lut, int lut_bits, uint16_t *count,
			    uint16_t *sorted, bool is_dist, bool strict)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1;
	// per-length counts and per-symbol rank inside its length class
	uint32_t my_cnt = 0;                 // lane l (1..15) holds count[l]
	uint32_t ranks[9];
#pragma unroll
	for (int g = 0; g < 9; g++) {
		if (g * 32 >= n)
			break;
		const int s = g * 32 + lane;
		const int l = s < n ? lens[s] : 0;
		uint32_t r = 0;
		for (int L = 1; L <= 15; L++) {
			const uint32_t m = __ballot_sync(0xffffffffu, l == L);
			const uint32_t base = __shfl_sync(0xffffffffu, my_cnt, L);
			if (l == L)
				r = base + __popc(m & lt);
			if ((int)lane == L)
				my_cnt += __popc(m);
		}
		ranks[g] = r;
	}
	// Kraft check + first code / offset per length (every lane computes the same small scan)
	uint32_t next_code = 0, offs = 0, my_next = 0, my_offs = 0;
	int left = 1, max_len = 0;
	bool over = false;
	for (int L = 1; L <= 15; L++) {
		const uint32_t c = __shfl_sync(0xffffffffu, my_cnt, L);
		left = (left << 1) - (int)c;
		if (left < 0)
			over = true;
		if ((int)lane == L) { my_next = next_code; my_offs = offs; }
		next_code = (next_code + c) << 1;
		offs += c;
		if (c)
			max_len = L;
	}
	if (lane < 16)
		count[lane] = (lane == 0) ? 0 : (uint16_t)my_cnt;
	if (over)
		return false;
	if (strict && left > 0 && max_len > 1)
		return false;
	for (int i = lane; i < (1 << lut_bits); i += 32)
		lut[i] = 0;
	__syncwarp();
#pragma unroll
	for (int g = 0; g < 9; g++) {
		if (g * 32 >= n)
			break;
		const int s = g * 32 + lane;
		const int l = s < n ? lens[s] : 0;
		const uint32_t nc = __shfl_sync(0xffffffffu, my_next, l & 31);
		const uint32_t of = __shfl_sync(0xffffffffu, my_offs, l & 31);
		if (l) {
			const uint32_t r = ranks[g];
			sorted[of + r] = (uint16_t)s;
			if (l <= lut_bits) {
				const uint32_t code = __brev(nc + r) >> (32 - l);
				uint32_t e;
				if (is_dist) {
					e = s < 30 ? ent_make(l + k_dist_extra[s], l, 0, k_dist_base[s] - 1) : 0;
				} else if (s < 256) {
					e = ent_make(l, l, 2, s);
				} else if (s == 256) {
					e = ent_make(l, l, 1, 0);
				} else if (s < 286) {
					e = ent_make(l + k_len_extra[s - 257], l, 3, k_len_base[s - 257] - 3);
				} else {
					e = 0;
				}
				if (e)
					for (uint32_t k = code; k < (1u << lut_bits); k += (1u << l))
						lut[k] = e;
			}
		}
	}
	__syncwarp();
	return true;
}

// dynamic block header from HLIT on (RFC 1951 3.2.7), lane 0 only: code lengths into lens[0..hlit+hdist)
__device__ int parse_dyn_header(BitReader &br, uint8_t *lens, int &hlit, int &hdist, bool strict)
{
	int rc = 0;
	const uint32_t v = br.get(14);
	hlit = (int)(v & 31) + 257; hdist = (int)((v >> 5) & 31) + 1;
	const int hclen = (int)(v >> 10) + 4;
	if (hlit > 286 || hdist > 30) rc = NXGPU_E_DATA;
	// code-length code: 19 symbols, <= 7 bits, decoded canonically
	uint8_t cl[19];
	for (int i = 0; i < 19; i++) cl[i] = 0;
	for (int i = 0; i < hclen; i++) cl[k_clorder[i]] = (uint8_t)br.get(3);
	uint16_t ccount[16], csorted[19], coffs[16];
	for (int i = 0; i < 16; i++) ccount[i] = 0;
	for (int i = 0; i < 19; i++) ccount[cl[i]]++;
	ccount[0] = 0;
	int left = 1;
	for (int l = 1; l <= 7; l++) { left = (left << 1) - ccount[l]; if (left < 0) rc = NXGPU_E_DATA; }
	if (strict && left > 0) rc = NXGPU_E_DATA;       // zlib: an incomplete code-length code is "invalid code lengths set"
	coffs[1] = 0;
	for (int l = 1; l < 15; l++) coffs[l + 1] = coffs[l] + ccount[l];
	for (int i = 0; i < 19; i++) if (cl[i]) csorted[coffs[cl[i]]++] = (uint16_t)i;
	int n = 0;
	while (!rc && n < hlit + hdist) {
		const int sym = slow_decode(br, ccount, csorted);
		if (sym < 0) { rc = NXGPU_E_DATA; break; }
		if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
		int rep, val = 0;
		if (sym == 16) { if (n == 0) { rc = NXGPU_E_DATA; break; } val = lens[n - 1]; rep = 3 + (int)br.get(2); }
		else if (sym == 17) rep = 3 + (int)br.get(3);
		else rep = 11 + (int)br.get(7);
		if (n + rep > hlit + hdist) { rc = NXGPU_E_DATA; break; }
		while (rep--) lens[n++] = (uint8_t)val;
	}
	if (!rc && lens[256] == 0) rc = NXGPU_E_DATA;
	return rc;
}

// Lane 0's table walk, written for latency.  A lone warp has nobody to hide behind: it issues an instruction every two
// to three cycles at best, a dependent ALU result takes 4-5, a shared-memory load ~25 and a branch whose condition
// comes out of a load stalls until it is there (ncu source page; the first version of this loop — three-word window,
// funnel shift per look-up, seven branches, 80 instructions — spent 263 cycles per symbol).  So:
//  * the bit window is 64 bits kept shifted left by two, so a table index is one AND away (byte offset of a 4-byte
//    entry); it is topped up with a ring word, by predication, whenever 30 bits or fewer are left;
//  * a table entry carries the bits its part of the symbol consumes (code + extra bits) in its low five bits, which is
//    where a wrap-mode funnel shift reads its shift amount: AND -> LDS -> funnel shift is the whole chain per code;
//  * every symbol does both look-ups, a literal consumes nothing at the second;
//  * lane 0 neither forms tokens nor judges symbols: it queues the two 32-bit windows the look-ups were made from (q, q2)
//    for every slot that is left, and the loop is counted — a symbol the tables do not resolve (long code, end of
//    block, more than the 31 bits one top-up guarantees) is followed by bounded garbage.  The other lanes redo the
//    look-ups, all slots at once, find the first such symbol, add up the bits in front of it and form the tokens
//    (walk_batch).
template <int kOff> __device__ __forceinline__ uint32_t lds_at(uint32_t saddr)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(kOff));
	return v;
}
template <int kOff> __device__ __forceinline__ void sts_at(uint32_t saddr, uint32_t v)
{
	asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(saddr), "n"(kOff), "r"(v) : "memory");
}
// tsa: 32-bit shared address of the warp's WarpTables (computed once by the caller: deriving it from a generic pointer
// costs an S2UR and address arithmetic inside the loop)
__device__ __forceinline__ void fast_walk(const BitReader &br, uint32_t tsa, uint32_t qn)
{
	constexpr int kLit = (int)offsetof(WarpTables, lit), kDist = (int)offsetof(WarpTables, dist), kQ = (int)offsetof(WarpTables, q),
		      kQ2 = (int)offsetof(WarpTables, q2), kIn = (int)offsetof(WarpTables, in);
	uint32_t cnt = 32 - br.bo;                                  // valid bits in the window
	const uint32_t first = br.w0 >> br.bo;
	uint32_t lo = first << 2, hi = first >> 30;
	uint32_t next = br.w1;                                       // the ring word behind the window
	uint32_t wo = ((br.wpos - 1) << 2) & (4 * kInWords - 4);     // byte offset in the ring of the word behind `next`
	uint32_t qa = tsa + 4 * qn;
	const uint32_t qe = tsa + 4 * 32;
#pragma unroll 2
	do {
		{
			const uint32_t sh = cnt + 2;
			lo |= __funnelshift_lc(0, next, sh);                 // next << sh; nothing once sh > 31
			const uint32_t add_hi = __funnelshift_lc(next, 0, sh); // next >> (32 - sh)
			if (cnt <= 30) {
				hi |= add_hi;
				cnt += 32;
				next = lds_at<kIn>(tsa + wo);
				wo = (wo + 4) & (4 * kInWords - 4);
			}
		}
		const uint32_t e = lds_at<kLit>(tsa + (lo & ((4u << kLitBits) - 4)));
		const uint32_t lo1 = __funnelshift_r(lo, hi, e);            // shifts by e & 31
		const uint32_t hi1 = __funnelshift_r(hi, 0, e);
		const uint32_t d = lds_at<kDist>(tsa + (lo1 & ((4u << kDistBits) - 4)));
		sts_at<kQ>(qa, lo);
		sts_at<kQ2>(qa, lo1);
		qa += 4;
		const uint32_t dsel = (e & 0x600) == 0x600 ? d : 0u;          // a length: the distance code follows
		lo = __funnelshift_r(lo1, hi1, dsel);
		hi = __funnelshift_r(hi1, 0, dsel);
		cnt -= (e & 31) + (dsel & 31);
	} while (qa != qe);
}

// One batch of up to 32 symbols into T.q[0 .. qn), as tokens (all lanes call this; the read position lives in lane 0).
// status: 0 = the queue is full or can be refilled, 1 = end of block, 2 = the source ran out inside a symbol (lane 0:
// sym_at = where that symbol starts, in bits from the first source byte), 3 = a code the tables do not know.
constexpr uint32_t kWalkEob = 1, kWalkSrcEnd = 2, kWalkBadCode = 3;
__device__ __forceinline__ uint32_t walk_batch(BitReader &br, WarpTables &T, uint32_t tsa, uint32_t lane, uint32_t &status, uint64_t &sym_at)
{
	constexpr int kIn = (int)offsetof(WarpTables, in);
	uint32_t qn = 0;
	status = 0;
	// overruns can only happen once the last words of the source are in the window: a batch consumes at most
	// 32 x 48 bits = 48 words.  Until then the ring cannot run dry either: the staging left >= 60 words ahead.
	const bool careful = __shfl_sync(0xffffffffu, (int)(br.wpos + 52 > br.end_word), 0) != 0;
	for (;;) {
		if (!careful) {
			if (lane == 0)
				fast_walk(br, tsa, qn);
			__syncwarp();
			// every lane judges the symbol in its slot
			const uint32_t lo = T.q[lane], lo1 = T.q2[lane];
			const uint32_t e = T.lit[(lo >> 2) & ((1u << kLitBits) - 1)];
			const uint32_t d = T.dist[(lo1 >> 2) & ((1u << kDistBits) - 1)];
			const bool is_len = ent_type(e) == 3;
			const uint32_t used = ent_tot(e) + (is_len ? ent_tot(d) : 0u);
			// resolved by the tables (literal / length with a short distance code) inside the bits a top-up guarantees
			const bool ok = (e & 0x400) && !(is_len && d == 0) && used <= 31;
			const uint32_t bad = __ballot_sync(0xffffffffu, lane >= qn && !ok);
			const uint32_t fb = bad ? (uint32_t)__ffs(bad) - 1 : 32u;
			const bool mine = lane >= qn && lane < fb;
			uint32_t sum = mine ? used : 0u;
			for (int o = 16; o; o >>= 1)
				sum += __shfl_xor_sync(0xffffffffu, sum, o);
			if (mine) {
				const uint32_t cl = ent_cl(e), dl = ent_cl(d);
				const uint32_t v = ent_val(e) + ((lo >> (cl + 2)) & ~(~0u << (ent_tot(e) - cl)));
				const uint32_t dm1 = ent_val(d) + ((lo1 >> (dl + 2)) & ~(~0u << (ent_tot(d) - dl)));
				T.q[lane] = is_len ? (0x80000000u | (v << 15) | dm1) : v;
			}
			if (lane == 0) {
				// the symbols in front of the first unresolved one are consumed
				const uint64_t at = br.bits_abs() + sum;
				const uint32_t w = (uint32_t)(at >> 5);
				br.w0 = lds_at<kIn>(tsa + ((w << 2) & (4 * kInWords - 4)));
				br.w1 = lds_at<kIn>(tsa + (((w + 1) << 2) & (4 * kInWords - 4)));
				br.w2 = lds_at<kIn>(tsa + (((w + 2) << 2) & (4 * kInWords - 4)));
				br.wpos = w + 3;
				br.bo = (uint32_t)at & 31;
			}
			qn = fb;
			if (qn == 32)
				break;
		}
		// the general code: one symbol (careful: as many as fit).  (Every lane has read its queue slot: the ballot and the
		// shuffles above sit between those reads and lane 0's write below.)
		__syncwarp();
		uint32_t st = 0;
		if (lane == 0) {
			do {
				const uint32_t s_wpos = br.wpos, s_bo = br.bo;     // where this symbol starts
				bool err = false;
				uint32_t tokv = 0;
				int kind = 0;                            // 0 literal, 1 match, 2 end of block
				const uint32_t w = br.peek32();
				const uint32_t e = T.lit[w & ((1u << kLitBits) - 1)];
				const uint32_t cl = ent_cl(e);
				uint32_t len = 0;
				if (cl) {
					const uint32_t tc = ent_type(e);
					if (tc == 2) {
						br.drop(cl);
						tokv = ent_val(e);
					} else if (tc == 3) {
						kind = 1;
						const uint32_t nextra = ent_tot(e) - cl;
						len = ent_val(e) + 3 + ((w >> cl) & ((1u << nextra) - 1));
						br.drop(cl + nextra);
					} else {
						kind = 2;
						br.drop(cl);
					}
				} else {
					const int sym = slow_decode(br, T.lit_count, T.lit_sorted);
					if (sym < 0 || sym >= 286) { err = true; kind = 2; }
					else if (sym < 256) { tokv = (uint32_t)sym; }
					else if (sym == 256) { kind = 2; }
					else { kind = 1; len = k_len_base[sym - 257] + br.get(k_len_extra[sym - 257]); }
				}
				if (kind == 1) {
					const uint32_t wd = br.peek32();
					const uint32_t d = T.dist[wd & ((1u << kDistBits) - 1)];
					const uint32_t dl = ent_cl(d);
					uint32_t dist = 1;
					if (dl) {
						const uint32_t dextra = ent_tot(d) - dl;
						dist = ent_val(d) + 1 + ((wd >> dl) & ((1u << dextra) - 1));
						br.drop(dl + dextra);
					} else {
						const int ds = slow_decode(br, T.dist_count, T.dist_sorted);
						if (ds < 0 || ds >= 30) err = true;
						else dist = k_dist_base[ds] + br.get(k_dist_extra[ds]);
					}
					tokv = tok_match(len, dist);
				}
				if (careful && br.overrun()) {
					// the symbol needs bits the source does not have
					sym_at = (uint64_t)(s_wpos - 3) * 32 + s_bo - br.skip * 8;
					st = kWalkSrcEnd;
					break;
				}
				if (err) { st = kWalkBadCode; break; }
				if (kind == 2) { st = kWalkEob; break; }
				T.q[qn++] = tokv;
			} while (careful && qn < 32);
		}
		__syncwarp();
		qn = __shfl_sync(0xffffffffu, qn, 0);
		st = __shfl_sync(0xffffffffu, st, 0);
		status = st;
		if (st || careful || qn == 32)
			break;
	}
	return qn;
}

// A batch of tokens into a ring of the last 32 Ki output symbols in shared memory: ring[(base + b) % 32 Ki] for the
// batch's bytes b in [0, total).  All lanes call; t / is_m / mylen / incl: the lane's token, its length and the inclusive
// prefix sum.  E: uint8_t (the real window) or uint16_t (speculation: bytes and window markers).
//  * Lane `lane` moves bytes b0 + lane and b0 + 32 + lane of every round of 64: the owner of a byte comes out of one
//    warp-wide OR and a population count.  (The first version searched it with five dependent shuffle steps per byte,
//    600 cycles per round for a lone warp; a contiguous span per lane was tried too and lost to divergence: every lane
//    sits in a different token, 4100 cycles per batch of ~500 bytes.)
//  * A match that reads this batch's own output (distance < bytes produced up to and including it) is skipped by the
//    spans and copied afterwards, in order, by the whole warp.
//  * A match whose source lies so far back that it shares ring slots with this batch's output (distance > 32 Ki - total)
//    is copied first, in order, before anything else is written.
constexpr uint32_t kWinBytes = 32768;
constexpr uint32_t kRingMask = 32768 - 1;
// ring element i of a ring at shared address rsa (a generic pointer would cost an S2UR + address arithmetic per access)
template <typename E> __device__ __forceinline__ uint32_t ring_ld(uint32_t rsa, uint32_t i)
{
	uint32_t v;
	if (sizeof(E) == 1)
		asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(rsa + (i & kRingMask)) : "memory");
	else
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(rsa + 2 * (i & kRingMask)) : "memory");
	return v;
}
template <typename E> __device__ __forceinline__ void ring_st(uint32_t rsa, uint32_t i, uint32_t v)
{
	if (sizeof(E) == 1)
		asm volatile("st.shared.u8 [%0], %1;" ::"r"(rsa + (i & kRingMask)), "r"(v) : "memory");
	else
		asm volatile("st.shared.u16 [%0], %1;" ::"r"(rsa + 2 * (i & kRingMask)), "r"(v) : "memory");
}
template <typename E>
__device__ __forceinline__ void ring_fill(E *ring, uint32_t base, uint32_t lane, uint32_t t, bool is_m, uint32_t mylen, uint32_t incl, uint32_t total)
{
	const uint32_t rsa = (uint32_t)__cvta_generic_to_shared(ring);
	const uint32_t my_at = base + incl - mylen;
	uint32_t fm = __ballot_sync(0xffffffffu, is_m && tok_dist(t) > 32768 - total);
	uint32_t mm = __ballot_sync(0xffffffffu, is_m && tok_dist(t) < incl);
	const uint32_t skip = fm | mm;
	while (fm) {
		const int src_lane = __ffs(fm) - 1;
		fm &= fm - 1;
		const uint32_t mt = __shfl_sync(0xffffffffu, t, src_lane);
		const uint32_t wp = __shfl_sync(0xffffffffu, my_at, src_lane);
		const uint32_t len = tok_len(mt), dist = tok_dist(mt);
		for (uint32_t k = lane; k < len; k += 32)
			ring_st<E>(rsa, wp + k, ring_ld<E>(rsa, wp + k - dist));
		__syncwarp();
	}
	{
		// byte b of the batch belongs to the token whose start is the last one at or in front of b: per round of 32 bytes the
		// tokens that start inside it are OR-reduced into a bit mask (one REDUX), and a population count gives every lane
		// its owner.  Two rounds per iteration, all loads in front of all stores.
		const uint32_t start = incl - mylen;                    // lanes behind the last token: start = total, owns nothing live
		uint32_t before = 0;                                     // tokens that start in front of the round
		const uint32_t upto = 0xffffffffu >> (31 - lane);        // bits 0 .. lane
		for (uint32_t b0 = 0; b0 < total; b0 += 64) {
			const uint32_t ra = start - b0, rb = start - b0 - 32;
			const uint32_t ma = __reduce_or_sync(0xffffffffu, ra < 32 ? 1u << ra : 0u);
			const uint32_t mb = __reduce_or_sync(0xffffffffu, rb < 32 ? 1u << rb : 0u);
			const uint32_t oa = (before + __popc(ma & upto) - 1) & 31;
			before += __popc(ma);
			const uint32_t ob = (before + __popc(mb & upto) - 1) & 31;
			before += __popc(mb);
			const uint32_t ta = __shfl_sync(0xffffffffu, t, oa), tb = __shfl_sync(0xffffffffu, t, ob);
			const uint32_t ba = b0 + lane, bb = ba + 32;
			const bool ca = ba < total && tok_is_match(ta) && !((skip >> oa) & 1);
			const bool cb = bb < total && tok_is_match(tb) && !((skip >> ob) & 1);
			uint32_t xa = ta, xb = tb;
			if (ca) xa = ring_ld<E>(rsa, base + ba - tok_dist(ta));
			if (cb) xb = ring_ld<E>(rsa, base + bb - tok_dist(tb));
			if (ba < total && (ca || !tok_is_match(ta))) ring_st<E>(rsa, base + ba, xa);
			if (bb < total && (cb || !tok_is_match(tb))) ring_st<E>(rsa, base + bb, xb);
		}
	}
	__syncwarp();
	while (mm) {
		const int src_lane = __ffs(mm) - 1;
		mm &= mm - 1;
		const uint32_t mt = __shfl_sync(0xffffffffu, t, src_lane);
		const uint32_t wp = __shfl_sync(0xffffffffu, my_at, src_lane);
		const uint32_t len = tok_len(mt), dist = tok_dist(mt);
		if (dist >= len || dist >= 32) {
			// each 32-byte pass reads bytes that earlier passes (or earlier symbols) wrote
			for (uint32_t k = 0; k < len; k += 32) {
				if (k + lane < len)
					ring_st<E>(rsa, wp + k + lane, ring_ld<E>(rsa, wp + k + lane - dist));
				if (dist < len)
					__syncwarp();
			}
		} else {
			// short period: the pattern in front of the match repeats
			for (uint32_t k = lane; k < len; k += 32)
				ring_st<E>(rsa, wp + k, ring_ld<E>(rsa, wp - dist + (k % dist)));
		}
		__syncwarp();
	}
}

// ---- two warps per stream: a walker and a copier ----
// Walking the bit stream occupies one lane, copying the bytes all 32, and for a warp that is alone with its stream the
// two phases take about the same time (ncu: ~145 cycles per symbol to walk and judge, ~100 to copy).  With a second warp
// they overlap: the walker (inflate_one<kWin, true>) hands every batch of tokens and every stored run to the copier
// through a two-slot queue in shared memory; the copier owns the window ring and the output position, checks the
// target space and the distances, and reports the first thing that is wrong.  The walker looks at that verdict before
// every batch and at the end, where the copier's verdict comes first: it concerns an earlier part of the stream.
struct DuoQueue {
	uint32_t head, tail;              // commands pushed / taken
	uint32_t cerr;                    // copier: 0, or the completion code of the first batch it had to refuse
	uint32_t out_final, done;
	uint32_t pad_[3];
	struct Cmd { uint32_t kind, a, b, pad_; uint32_t tok[32]; } cmd[2];   // kind 1: a tokens; 2: b stored bytes at source offset a; 3: end
};
constexpr uint32_t kDuoTokens = 1, kDuoStored = 2, kDuoEnd = 3;
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// walker warp, all lanes: lane l's token in t (kDuoTokens), or a stored run
__device__ __forceinline__ void duo_push(DuoQueue &Q, uint32_t &head, uint32_t lane, uint32_t kind, uint32_t a, uint32_t b, uint32_t t)
{
	while (head - ld_vol(&Q.tail) >= 2)
		__nanosleep(40);
	DuoQueue::Cmd &c = Q.cmd[head & 1];
	c.tok[lane] = t;
	if (lane == 0) { c.kind = kind; c.a = a; c.b = b; }
	__threadfence_block();
	__syncwarp();
	head++;
	if (lane == 0)
		st_vol(&Q.head, head);
}

// copier warp: everything inflate_one<true> does with a batch once it is decoded
__device__ void duo_copier(const InflateJob &J, DuoQueue &Q, uint8_t *win)
{
	const uint32_t lane = threadIdx.x & 31;
	const bool job = J.wrap == kWrapJob;
	const uint32_t wofs = J.hist_len;
	{
		const uint32_t h = J.hist_len < kWinBytes ? J.hist_len : kWinBytes;
		const uint8_t *hp = J.hist_ptr ? J.hist_ptr + kWinBytes : J.dst;
		for (uint32_t i = lane; i < h; i += 32)
			win[(wofs - h + i) & (kWinBytes - 1)] = hp[(int32_t)i - (int32_t)h];
		__syncwarp();
	}
	uint32_t out = 0, err = 0, tail = 0;
	for (;;) {
		while (ld_vol(&Q.head) == tail)
			__nanosleep(40);
		__threadfence_block();
		const DuoQueue::Cmd &c = Q.cmd[tail & 1];
		const uint32_t kind = ld_vol(&c.kind), a = ld_vol(&c.a), b = ld_vol(&c.b);
		const uint32_t t = kind == kDuoTokens && lane < a ? ld_vol(&c.tok[lane]) : 0;
		__syncwarp();
		tail++;
		if (lane == 0)
			st_vol(&Q.tail, tail);                       // the slot is free again
		if (kind == kDuoEnd)
			break;
		if (err)
			continue;
		if (kind == kDuoStored) {
			if (b > J.dst_cap - out) {
				err = job ? 13 : NXGPU_E_BUF;
			} else {
				for (uint32_t i = lane; i < b; i += 32) {
					const uint8_t v = J.src[a + i];
					J.dst[out + i] = v;
					win[(wofs + out + i) & (kWinBytes - 1)] = v;
				}
				out += b;
				__syncwarp();
			}
		} else {
			const bool is_m = lane < a && tok_is_match(t);
			const uint32_t mylen = lane < a ? (is_m ? tok_len(t) : 1) : 0;
			uint32_t incl = mylen;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= (uint32_t)o)
					incl += y;
			}
			const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
			const uint32_t my_out = out + incl - mylen;
			if (total > J.dst_cap - out) {
				err = job ? 13 : NXGPU_E_BUF;
			} else if (__any_sync(0xffffffffu, is_m && tok_dist(t) > my_out + J.hist_len)) {
				err = job ? 67 : NXGPU_E_DATA;
			} else {
				ring_fill<uint8_t>(win, wofs + out, lane, t, is_m, mylen, incl, total);
				uint8_t *const dq = J.dst + out;
				for (uint32_t i = lane; i < total; i += 32)
					dq[i] = win[(wofs + out + i) & (kWinBytes - 1)];
				out += total;
			}
		}
		if (err && lane == 0)
			st_vol(&Q.cerr, err);
	}
	if (lane == 0) {
		st_vol(&Q.out_final, out);
		__threadfence_block();
		st_vol(&Q.done, 1);
	}
}

// kWin: the last 32 KiB of output are mirrored in a shared-memory ring (win), and every match source is read from
// there instead of from global memory.  A warp that is alone with its stream (a lone uncompress() through
// nxu_run_job, a handful of large members) otherwise pays an L2 round trip per materialise step and per
// short-distance match: 54 MB/s.  With thousands of members in flight the other warps hide that latency and the
// 32 KiB per warp would cost occupancy, so the batch kernel keeps reading the window from L1/L2.
template <bool kWin, bool kDuo = false>
__device__ void inflate_one(const InflateJob &J, InflateOut &O, WarpTables &T, uint8_t *win, DuoQueue *Q = nullptr)
{
	uint32_t duo_head = 0;               // kDuo: commands pushed so far
	const uint32_t lane = threadIdx.x & 31;
	const bool job = J.wrap == kWrapJob;     // NX decompress-job semantics: stop at the source end and report where
	const uint64_t total_bits = (uint64_t)J.src_len * 8;
	uint32_t tsa = (uint32_t)__cvta_generic_to_shared(&T);      // fast_walk addresses the tables and queues through this
	asm volatile("" : "+r"(tsa));
	BitReader br;
	br.setup(J.src, J.src_len, T.in);   // every lane knows the geometry; the read position lives in lane 0
	int rc = 0;                      // uniform after each broadcast
	uint32_t out = 0;
	// window ring: output position p (counted from the first history byte) lives at win[p % 32 KiB]
	const uint32_t wofs = J.hist_len;
	if (kWin && !kDuo) {
		const uint32_t h = J.hist_len < kWinBytes ? J.hist_len : kWinBytes;
		const uint8_t *hp = J.hist_ptr ? J.hist_ptr + kWinBytes : J.dst;      // one past the last history byte
		for (uint32_t i = lane; i < h; i += 32)
			win[(wofs - h + i) & (kWinBytes - 1)] = hp[(int32_t)i - (int32_t)h];
		__syncwarp();
	}
	// dry run (kWrapDry): walk the Huffman stream and count, write nothing — finds where a member ends and how
	// long its output is (nxgpu_gunzip_concat discovers the members of a concatenated file this way)
	const bool dry = (J.wrap & kWrapDry) != 0;
	uint32_t start = 0, wrap = J.wrap & 0xff;
	const bool no_header = (J.wrap & kWrapNoHeader) != 0;
	bool map_stop = false;           // a block ended where J.stop_map says another warp takes over
	uint32_t tr_crc = 0, tr_isize = 0, flags = 0;
	// job mode: set when the source ran out (or the final EOB was seen); lane 0 holds the details
	bool suspended = false;
	bool self_stop = false;          // stopped at the end of a BFINAL=0 block with nothing decoded behind it
	uint32_t o_sfbt = 0, o_subc = 0, o_rem = 0;
	uint64_t dht_from = 0;           // where the current dynamic header starts (bit offset in dht_src)
	uint32_t dht_len = 0;
	bool dht_saved = false;          // the current dynamic table came from J.dht, not from the stream
	// resume state consumed by the first loop iteration
	uint32_t resume = job ? (J.sfbt & 0xe) : 0;
	if (resume != 0x8 && resume != 0xa && resume != 0xc)
		resume = 0;

	// ---- container header (lane 0), lib/nx_inflate.c:329-730 does this on the host ----
	if (lane == 0) {
		const uint8_t *s = J.src;
		const uint32_t n = J.src_len;
		if (wrap == NXGPU_WRAP_AUTO) {
			if (n >= 2 && s[0] == 0x1f && s[1] == 0x8b) wrap = NXGPU_WRAP_GZIP;
			else if (n >= 2 && (s[0] & 0x0f) == 8 && (((uint32_t)s[0] << 8 | s[1]) % 31) == 0) wrap = NXGPU_WRAP_ZLIB;
			else wrap = NXGPU_WRAP_RAW;
		}
		if (no_header) {
			start = 0;
		} else if (wrap == NXGPU_WRAP_GZIP) {
			if (n < 18 || s[0] != 0x1f || s[1] != 0x8b || s[2] != 8) rc = NXGPU_E_DATA;
			else {
				const uint32_t flg = s[3];
				uint32_t p = 10;
				if (flg & 4) { if (p + 2 <= n) p += 2 + (s[p] | (uint32_t)s[p + 1] << 8); else p = n + 1; }
				if (flg & 8) { while (p < n && s[p]) p++; p++; }
				if (flg & 16) { while (p < n && s[p]) p++; p++; }
				if (flg & 2) p += 2;
				if (p > n) rc = NXGPU_E_DATA;
				start = p;
			}
		} else if (wrap == NXGPU_WRAP_ZLIB) {
			if (n < 6 || (s[0] & 0x0f) != 8 || (((uint32_t)s[0] << 8 | s[1]) % 31) || (s[1] & 0x20)) rc = NXGPU_E_DATA;
			start = 2;
		}
		if (!rc) {
			br.seek(start);
			if ((job || no_header) && J.start_bit)
				br.drop(J.start_bit & 7);
		}
	}
	rc = __shfl_sync(0xffffffffu, rc, 0);

	bool final_block = false;
	bool block_ended = false;
	while (!rc && !final_block && !suspended) {
		if (J.stop_map && block_ended) {
			uint64_t at = 0;
			if (lane == 0)
				at = J.map_bit0 + br.bits_used();
			at = __shfl_sync(0xffffffffu, at, 0);
			if ((J.stop_map[at >> 5] >> (at & 31)) & 1) {
				map_stop = true;
				self_stop = true;
				break;
			}
		}
		if (job && block_ended) {
			// A block with BFINAL=0 just ended.  Manual Table 5-3, SFBT 1110: the engine suspends here by
			// itself for the single-block function codes (inc_nx/nxu.h:813,815), and also when the source
			// ends exactly on the block boundary (SUBC=0, "the next byte will contain a block header").
			if (J.single_block != 0 || __shfl_sync(0xffffffffu, (int)(br.bits_used() == total_bits), 0) != 0) {
				self_stop = true;
				break;
			}
		}
		// ---- block header (lane 0) ----
		uint32_t btype = 0, stored_len = 0, stored_at = 0;
		int hlit = 0, hdist = 0;
		if (lane == 0) {
			if (resume) {
				// continue inside the block the previous job stopped in (inc_nx/nxu.h:330-372)
				final_block = J.sfbt & 1;
				if (resume == 0x8) {
					btype = 0;
					stored_len = J.rembytecnt;
					stored_at = (uint32_t)((br.bits_used() + 7) >> 3);
				} else if (resume == 0xa) {
					btype = 1;
				} else {
					btype = 2;
					BitReader dr;
					dr.init(J.dht, (J.dht_bits + 7) >> 3, 0, T.lit);   // the LUT is not built yet: borrow it as the ring
					rc = parse_dyn_header(dr, T.lens, hlit, hdist, false);
					if (rc || dr.bits_used() > J.dht_bits) rc = 68;
					dht_saved = true; dht_from = 0; dht_len = J.dht_bits;
				}
			} else {
				const uint64_t blk = br.bits_used();
				const uint32_t h = br.get(3);
				final_block = h & 1;
				btype = h >> 1;
				if (btype == 0) {
					br.align_byte();
					const uint32_t v = br.peek32();
					br.drop(32);
					if (((v ^ (v >> 16)) & 0xffff) != 0xffff) rc = NXGPU_E_DATA;
					stored_len = v & 0xffff;
					// give the buffered bytes back: stored data is copied straight from memory
					stored_at = br.byte_pos();
				} else if (btype == 2) {
					dht_from = br.bits_used();
					rc = parse_dyn_header(br, T.lens, hlit, hdist, !job);
					dht_len = (uint32_t)(br.bits_used() - dht_from);
					dht_saved = false;
				} else if (btype == 3) {
					rc = NXGPU_E_DATA;
				}
				if (br.overrun()) {
					if (job) {
						// the header is incomplete: hand all of it back (manual Table 5-3, 111x)
						const uint32_t f = blk < total_bits ? (J.src[blk >> 3] >> (blk & 7)) & 1 : 0;
						o_sfbt = 0xe | f;
						o_subc = (uint32_t)(total_bits - blk);
						suspended = true;
						final_block = false;
						rc = 0;
					} else {
						rc = NXGPU_E_DATA;
						if (blk == total_bits)
							flags |= 4;          // the source ended exactly on a block boundary, no final block seen
					}
				} else if (rc && job) {
					rc = 68;
				}
			}
		}
		resume = 0;
		rc = __shfl_sync(0xffffffffu, rc, 0);
		btype = __shfl_sync(0xffffffffu, btype, 0);
		final_block = __shfl_sync(0xffffffffu, (int)final_block, 0) != 0;
		suspended = __shfl_sync(0xffffffffu, (int)suspended, 0) != 0;
		if (rc || suspended)
			break;

		if (btype == 0) {
			stored_len = __shfl_sync(0xffffffffu, stored_len, 0);
			stored_at = __shfl_sync(0xffffffffu, stored_at, 0);
			uint32_t n = stored_len;
			if (stored_at + stored_len > J.src_len) {
				if (!job) { rc = NXGPU_E_DATA; break; }
				n = J.src_len > stored_at ? J.src_len - stored_at : 0;   // copy what is there, resume later
			}
			if (kDuo) {
				// the copier owns the target and the window
				if (ld_vol(&Q->cerr)) { rc = (int)ld_vol(&Q->cerr); break; }
				if (n)
					duo_push(*Q, duo_head, lane, kDuoStored, stored_at, n, 0);
			} else {
				if (n > J.dst_cap - out) { rc = job ? 13 : NXGPU_E_BUF; break; }
				for (uint32_t i = lane; i < n && !dry; i += 32) {
					const uint8_t v = J.src[stored_at + i];
					J.dst[out + i] = v;
					if (kWin) win[(wofs + out + i) & (kWinBytes - 1)] = v;
				}
				out += n;
			}
			if (n < stored_len) {
				o_sfbt = 0x8 | (final_block ? 1u : 0u);
				o_rem = stored_len - n;
				o_subc = 0;
				suspended = true;
				final_block = false;
				break;
			}
			if (lane == 0)
				br.seek(stored_at + stored_len);
			__syncwarp();
			block_ended = true;
			continue;
		}

		// ---- decode tables ----
		if (btype == 1) {
			for (int i = lane; i < 288; i += 32)
				T.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
			T.lens[288 + lane] = 5;
			hlit = 288; hdist = 30;
		} else {
			hlit = __shfl_sync(0xffffffffu, hlit, 0);
			hdist = __shfl_sync(0xffffffffu, hdist, 0);
		}
		__syncwarp();
		// members are judged like zlib judges them; NX job descriptors keep the engine's lenient rule (over-subscription only).
		// The fixed code's 30 five-bit distance codes are an incomplete set by construction.
		bool ok = build_table(T.lens, hlit, T.lit, kLitBits, T.lit_count, T.lit_sorted, false, !job && btype == 2);
		ok = build_table(T.lens + hlit, hdist, T.dist, kDistBits, T.dist_count, T.dist_sorted, true, !job && btype == 2) && ok;
		if (!ok) { rc = job ? 68 : NXGPU_E_DATA; break; }

		// ---- symbols ----
		bool block_done = false;
		while (!block_done && !rc && !suspended) {
			// ---- stage the compressed input: one coalesced 128-byte load per lane-0 shortfall ----
			{
				const uint32_t wp = __shfl_sync(0xffffffffu, br.wpos, 0);
				uint32_t st = __shfl_sync(0xffffffffu, br.staged, 0);
				if (st - wp < kInAhead && (uint64_t)st * 4 < br.end) {
					const uint32_t w0 = BitReader::load_word(br.base32, br.skip, br.end, st + lane);
					const uint32_t w1 = BitReader::load_word(br.base32, br.skip, br.end, st + 32 + lane);
					T.in[(st + lane) % kInWords] = w0;
					T.in[(st + 32 + lane) % kInWords] = w1;
					br.staged = st + 64;
					__syncwarp();
				}
			}
			uint32_t wst;
			uint64_t sym_at = 0;
			const uint32_t qn = walk_batch(br, T, tsa, lane, wst, sym_at);
			if (wst == kWalkSrcEnd) {
				if (job) {
					if (lane == 0) {
						o_sfbt = (btype == 1 ? 0xau : 0xcu) | (final_block ? 1u : 0u);
						o_subc = (uint32_t)(total_bits - sym_at);
					}
					suspended = true;
				} else {
					rc = NXGPU_E_DATA;
				}
			} else if (wst == kWalkBadCode) {
				rc = job ? 66 : NXGPU_E_DATA;
			} else if (wst == kWalkEob) {
				block_done = true;
			}
			if (rc)
				break;
			// ---- materialise the queue ----
			const uint32_t t = lane < qn ? T.q[lane] : 0;
			if (kDuo) {
				if (ld_vol(&Q->cerr)) { rc = (int)ld_vol(&Q->cerr); break; }
				if (qn)
					duo_push(*Q, duo_head, lane, kDuoTokens, qn, 0, t);
				continue;
			}
			const bool is_m = lane < qn && tok_is_match(t);
			const uint32_t mylen = lane < qn ? (is_m ? tok_len(t) : 1) : 0;
			uint32_t incl = mylen;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= (uint32_t)o)
					incl += y;
			}
			const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
			const uint32_t my_out = out + incl - mylen;
			if (total > J.dst_cap - out) { rc = job ? 13 : NXGPU_E_BUF; break; }
			if (dry) {
				if (__any_sync(0xffffffffu, is_m && tok_dist(t) > my_out + J.hist_len)) { rc = NXGPU_E_DATA; break; }
				out += total;
				continue;
			}
			const bool bad = is_m && tok_dist(t) > my_out + J.hist_len;
			if (__any_sync(0xffffffffu, bad)) { rc = job ? 67 : NXGPU_E_DATA; break; }
			// A match whose source ends in front of this batch's output (distance >= the bytes the batch
			// has produced up to and including it) depends on nothing written in this batch.  Those and
			// the literals are materialised byte-parallel: output byte b belongs to the first symbol whose
			// inclusive prefix exceeds b (binary search over the lanes' prefixes by shuffle), so all the
			// loads of a batch are in flight together.  The rest (short distances) follow one by one.
			if (kWin) {
				// the batch goes into the window ring, then out to the target in one coalesced sweep
				ring_fill<uint8_t>(win, wofs + out, lane, t, is_m, mylen, incl, total);
				uint8_t *const dq = J.dst + out;
				for (uint32_t i = lane; i < total; i += 32)
					dq[i] = win[(wofs + out + i) & (kWinBytes - 1)];
				out += total;
				continue;
			}
			uint32_t mm = __ballot_sync(0xffffffffu, is_m && tok_dist(t) < incl);
			uint8_t *const dq = J.dst + out;
			{
				// the owner of every byte of a round of 32: one warp-wide OR of the token starts inside the round and a
				// population count (see ring_fill); two rounds per iteration, all loads in front of all stores
				const uint32_t start = incl - mylen;
				uint32_t before = 0;
				const uint32_t upto = 0xffffffffu >> (31 - lane);
				for (uint32_t b0 = 0; b0 < total; b0 += 64) {
					const uint32_t ra = start - b0, rb = start - b0 - 32;
					const uint32_t ma = __reduce_or_sync(0xffffffffu, ra < 32 ? 1u << ra : 0u);
					const uint32_t mb = __reduce_or_sync(0xffffffffu, rb < 32 ? 1u << rb : 0u);
					const uint32_t oa = (before + __popc(ma & upto) - 1) & 31;
					before += __popc(ma);
					const uint32_t ob = (before + __popc(mb & upto) - 1) & 31;
					before += __popc(mb);
					const uint32_t ta = __shfl_sync(0xffffffffu, t, oa), tb = __shfl_sync(0xffffffffu, t, ob);
					const uint32_t ba = b0 + lane, bb = ba + 32;
					const bool ca = ba < total && tok_is_match(ta) && !((mm >> oa) & 1);
					const bool cb = bb < total && tok_is_match(tb) && !((mm >> ob) & 1);
					uint32_t xa = ta, xb = tb;
					if (ca) xa = dq[(int32_t)ba - (int32_t)tok_dist(ta)];
					if (cb) xb = dq[(int32_t)bb - (int32_t)tok_dist(tb)];
					if (ba < total && (ca || !tok_is_match(ta))) dq[ba] = (uint8_t)xa;
					if (bb < total && (cb || !tok_is_match(tb))) dq[bb] = (uint8_t)xb;
				}
			}
			__syncwarp();
			while (mm) {
				const int src_lane = __ffs(mm) - 1;
				mm &= mm - 1;
				const uint32_t mt = __shfl_sync(0xffffffffu, t, src_lane);
				const uint32_t mo = __shfl_sync(0xffffffffu, my_out, src_lane);
				const uint32_t len = tok_len(mt), dist = tok_dist(mt);
				uint8_t *d = J.dst + mo;
				if (dist >= len || dist >= 32) {
					// each 32-byte pass reads bytes that earlier passes (or earlier symbols) wrote
					for (uint32_t k = 0; k < len; k += 32) {
						if (k + lane < len)
							d[k + lane] = *(volatile const uint8_t *)(d + k + lane - dist);
						if (dist < len)
							__syncwarp();
					}
				} else {
					// short period: the pattern d[-dist..-1] repeats
					for (uint32_t k = lane; k < len; k += 32)
						d[k] = *(volatile const uint8_t *)(d - dist + (k % dist));
				}
				__syncwarp();
			}
			out += total;
		}
		if (suspended)
			final_block = false;
		block_ended = true;
	}

	if (kDuo) {
		// the copier finishes what it was given; its verdict is about an earlier part of the stream than anything found here
		duo_push(*Q, duo_head, lane, kDuoEnd, 0, 0, 0);
		while (!ld_vol(&Q->done))
			__nanosleep(40);
		__threadfence_block();
		const uint32_t ce = ld_vol(&Q->cerr);
		if (ce) rc = (int)ce;
		out = ld_vol(&Q->out_final);
	}
	if (job) {
		// ---- NX completion state (lib/nx_inflate.c:1372-1609 reads these) ----
		const bool in_dyn = suspended && (__shfl_sync(0xffffffffu, o_sfbt, 0) & 0xe) == 0xc;
		dht_len = __shfl_sync(0xffffffffu, dht_len, 0);
		const uint32_t from_lo = __shfl_sync(0xffffffffu, (uint32_t)dht_from, 0);
		const bool saved = __shfl_sync(0xffffffffu, (int)dht_saved, 0) != 0;
		if (in_dyn && J.out_dht) {
			// hand the dynamic header back so that the next job can resume inside this block
			const uint8_t *hs = saved ? J.dht : J.src;
			const uint32_t hbytes = saved ? (J.dht_bits + 7) >> 3 : J.src_len;
			const uint32_t nb = (dht_len + 7) >> 3;
			for (uint32_t i = lane; i < 288; i += 32) {
				uint32_t v = 0;
				if (i < nb) {
					const uint32_t bit = from_lo + 8 * i, by = bit >> 3, sh = bit & 7;
					const uint32_t a = by < hbytes ? hs[by] : 0, b = by + 1 < hbytes ? hs[by + 1] : 0;
					v = ((a | (b << 8)) >> sh) & 0xff;
					if (i == nb - 1 && (dht_len & 7))
						v &= (1u << (dht_len & 7)) - 1;
				}
				J.out_dht[i] = (uint8_t)v;
			}
		}
		if (lane == 0) {
			uint32_t src_read = J.src_len;      // the source ran out: all of it was read
			if (!rc && !suspended) {
				// The engine stopped by itself - final EOB (sfbt 0000) or the end of a BFINAL=0 block (1110) -
				// with source possibly left.  "SPBC indicates the number of compressed source bytes read by the
				// accelerator, SUBC the number of source bits that the accelerator discarded because they were
				// past the stream end" (manual §2.4): a gzip trailer reads as SUBC 64..71, a zlib one as 32..39
				// (inc_nx/nxu.h:454-465).  The engine has read at most kReadAhead bytes behind the byte holding
				// the last processed bit; the host finds the stream end at spbc - histlen - subc/8
				// (lib/nx_inflate.c:1452-1472).  SUBC is a 16-bit field: counting everything supplied wraps it.
				const uint64_t used = br.bits_used();
				uint64_t rd = ((used + 7) >> 3) + kReadAhead;
				if (rd > J.src_len) rd = J.src_len;
				src_read = (uint32_t)rd;
				o_sfbt = self_stop ? 0xe : 0;
				o_subc = (uint32_t)(rd * 8 - used);
				if (!self_stop) flags |= 1;
			}
			O.rc = rc;
			O.out_len = out;
			O.in_used = src_read;
			O.flags = flags | (wrap << 8);
			O.trailer_crc = 0;
			O.trailer_isize = 0;
			O.sfbt = o_sfbt;
			O.subc = o_subc;
			O.rembytecnt = o_rem;
			O.dhtlen = in_dyn ? dht_len : 0;
			if (map_stop) {
				const uint64_t at = J.map_bit0 + br.bits_used();
				O.flags |= kInflateMapStop;
				O.end_bit_lo = (uint32_t)at; O.end_bit_hi = (uint32_t)(at >> 32);
			}
		}
		return;
	}

	// ---- trailer (lane 0) ----
	uint32_t in_used = 0;
	if (lane == 0) {
		if (!rc && map_stop) {
			const uint64_t at = J.map_bit0 + br.bits_used();
			flags |= kInflateMapStop;
			O.end_bit_lo = (uint32_t)at; O.end_bit_hi = (uint32_t)(at >> 32);
		} else if (!rc) {
			flags |= 1;
			uint32_t p = (uint32_t)((br.bits_used() + 7) >> 3);
			const uint8_t *s = J.src;
			if (wrap == NXGPU_WRAP_GZIP) {
				if (p + 8 > J.src_len) rc = NXGPU_E_DATA;
				else {
					tr_crc = s[p] | (uint32_t)s[p + 1] << 8 | (uint32_t)s[p + 2] << 16 | (uint32_t)s[p + 3] << 24;
					tr_isize = s[p + 4] | (uint32_t)s[p + 5] << 8 | (uint32_t)s[p + 6] << 16 | (uint32_t)s[p + 7] << 24;
					p += 8;
				}
			} else if (wrap == NXGPU_WRAP_ZLIB) {
				if (p + 4 > J.src_len) rc = NXGPU_E_DATA;
				else {
					tr_crc = (uint32_t)s[p] << 24 | (uint32_t)s[p + 1] << 16 | (uint32_t)s[p + 2] << 8 | s[p + 3];
					p += 4;
				}
			}
			if (p > J.src_len) rc = NXGPU_E_DATA;
			in_used = p;
		}
		O.rc = rc;
		O.out_len = out;
		O.in_used = in_used;
		O.flags = flags | (wrap << 8);
		O.trailer_crc = tr_crc;
		O.trailer_isize = tr_isize;
	}
}

template <int kMinCtas>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kMinCtas)
inflate_kernel(const InflateJob *__restrict__ jobs, InflateOut *__restrict__ outs, uint32_t n_jobs, uint32_t *next_job)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	WarpTables &T = reinterpret_cast<WarpTables *>(smem_raw)[threadIdx.x >> 5];
	const uint32_t lane = threadIdx.x & 31;
	for (;;) {
		// dynamic work distribution: members differ a lot in cost
		uint32_t j = 0;
		if (lane == 0)
			j = atomicAdd(next_job, 1u);
		j = __shfl_sync(0xffffffffu, j, 0);
		if (j >= n_jobs)
			break;
		const InflateJob J = jobs[j];
		if (J.wrap & kWrapSkip)
			continue;
		inflate_one<false>(J, outs[j], T, nullptr);
		__syncwarp();
	}
}

// a few large streams: a walker warp and a copier warp per CTA, the 32 KiB window in shared memory (see inflate_one<kWin, kDuo>)
constexpr size_t kDuoTables = (sizeof(WarpTables) + 15) & ~(size_t)15;
constexpr size_t kDuoQueueAt = kDuoTables, kDuoWinAt = (kDuoTables + sizeof(DuoQueue) + 15) & ~(size_t)15;
// one stream on the two warps of a CTA (64 threads, both warps call)
__device__ __forceinline__ void inflate_duo(const InflateJob &J, InflateOut &O, uint8_t *smem_raw)
{
	WarpTables &T = *reinterpret_cast<WarpTables *>(smem_raw);
	DuoQueue &Q = *reinterpret_cast<DuoQueue *>(smem_raw + kDuoQueueAt);
	uint8_t *win = smem_raw + kDuoWinAt;
	if (threadIdx.x < 8)
		reinterpret_cast<uint32_t *>(&Q)[threadIdx.x] = 0;
	__syncthreads();
	if (threadIdx.x < 32)
		inflate_one<true, true>(J, O, T, win, &Q);
	else
		duo_copier(J, Q, win);
	__syncthreads();
}

__global__ void __launch_bounds__(64)
inflate_solo_kernel(const InflateJob *__restrict__ jobs, InflateOut *__restrict__ outs, uint32_t n_jobs, uint32_t *next_job)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	__shared__ uint32_t s_job;
	for (;;) {
		if (threadIdx.x == 0)
			s_job = atomicAdd(next_job, 1u);
		__syncthreads();
		const uint32_t j = s_job;
		__syncthreads();
		if (j >= n_jobs)
			break;
		const InflateJob J = jobs[j];
		if (J.wrap & kWrapSkip)
			continue;
		if ((J.wrap & kWrapDry) != 0) {
			if (threadIdx.x < 32)
				inflate_one<false>(J, outs[j], *reinterpret_cast<WarpTables *>(smem_raw), nullptr);
			continue;
		}
		inflate_duo(J, outs[j], smem_raw);
	}
}

#include "inflate_par.cuh"

// every offset where a gzip member COULD start: 1f 8b 08 and a flag byte without reserved bits (RFC 1952 §2.3);
// the list is unordered, false positives (the pattern inside compressed or stored data) are weeded out by decoding
__global__ void gzip_candidates_kernel(const uint8_t *__restrict__ src, uint64_t len, uint64_t *__restrict__ cand, uint32_t max_cand,
				       uint32_t *__restrict__ count)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p + 18 <= len; p += stride) {
		if (src[p] == 0x1f && src[p + 1] == 0x8b && src[p + 2] == 8 && (src[p + 3] & 0xe0) == 0) {
			const uint32_t k = atomicAdd(count, 1u);
			if (k < max_cand)
				cand[k] = p;
		}
	}
}

} // namespace

cudaError_t launch_gzip_candidates(const uint8_t *src, uint64_t len, uint64_t *cand, uint32_t max_cand, uint32_t *count, cudaStream_t s)
{
	cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	gzip_candidates_kernel<<<kNumSMs * 8, 256, 0, s>>>(src, len, cand, max_cand, count);
	return cudaGetLastError();
}

template <int kMinCtas>
static cudaError_t launch_inflate_t(const InflateJob *jobs, InflateOut *outs, uint32_t n_jobs, uint32_t *counter, cudaStream_t s)
{
	static PerDeviceOnce once;
	const size_t smem = sizeof(WarpTables) * kWarpsPerCta;
	cudaError_t e = once.run([smem] { return cudaFuncSetAttribute(inflate_kernel<kMinCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
	if (e != cudaSuccess)
		return e;
	e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	// persistent grid: as many CTAs as fit an SM (registers and shared memory allow kMinCtas)
	int per_sm = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_kernel<kMinCtas>, kWarpsPerCta * 32, smem);
	if (per_sm < 1)
		per_sm = 1;
	uint32_t want = (n_jobs + kWarpsPerCta - 1) / kWarpsPerCta;
	uint32_t grid = (uint32_t)(kNumSMs * per_sm);
	if (want < grid)
		grid = want ? want : 1;
	inflate_kernel<kMinCtas><<<grid, kWarpsPerCta * 32, smem, s>>>(jobs, outs, n_jobs, counter);
	return cudaGetLastError();
}

// Up to this many streams per launch run on the solo kernel (5 CTAs of one warp fit an SM with 40 KiB each): beyond that
// the batch kernel's 28 warps per SM hide the window's L2 latency by themselves.
static cudaError_t launch_inflate_solo(const InflateJob *jobs, InflateOut *outs, uint32_t n_jobs, uint32_t *counter, cudaStream_t s)
{
	static PerDeviceOnce once;
	const size_t smem = kDuoWinAt + kWinBytes;
	cudaError_t e = once.run([smem] { return cudaFuncSetAttribute(inflate_solo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
	if (e != cudaSuccess)
		return e;
	e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	const uint32_t grid = n_jobs < (uint32_t)(kNumSMs * 5) ? n_jobs : (uint32_t)(kNumSMs * 5);
	inflate_solo_kernel<<<grid ? grid : 1, 64, smem, s>>>(jobs, outs, n_jobs, counter);
	return cudaGetLastError();
}

cudaError_t launch_blockfind(const uint8_t *src, uint32_t src_len, uint64_t first_bit, uint32_t *map, uint64_t *surv, uint32_t surv_cap,
			     uint64_t *cand, uint32_t cand_cap, uint32_t *counts, cudaStream_t s)
{
	cudaError_t e = cudaMemsetAsync(map, 0, ((size_t)src_len / 4 + 4) * 4, s);
	if (e != cudaSuccess)
		return e;
	e = cudaMemsetAsync(counts, 0, 2 * sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	const uint32_t n_words = src_len / 4 + 2;
	uint32_t grid = (n_words + 255) / 256;
	if (grid > (uint32_t)kNumSMs * 8)
		grid = kNumSMs * 8;
	blockfind_filter_kernel<<<grid, 256, 0, s>>>(src, src_len, first_bit, surv, surv_cap, counts);
	uint32_t g2 = (surv_cap + 127) / 128;
	if (g2 > (uint32_t)kNumSMs * 8)
		g2 = kNumSMs * 8;
	blockfind_verify_kernel<<<g2 ? g2 : 1, 128, 0, s>>>(src, src_len, surv, surv_cap, map, cand, cand_cap, counts);
	return cudaGetLastError();
}

cudaError_t launch_inflate_par(const ParPlan &plan, uint32_t *counter, cudaStream_t s)
{
	static PerDeviceOnce once;
	const size_t smem_spec = kDuoWinAt + kRingSyms * 2, smem_chain = kDuoWinAt + kWinBytes;
	cudaError_t e = once.run([=] {
		cudaError_t r = cudaFuncSetAttribute(inflate_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_spec);
		if (r != cudaSuccess)
			return r;
		return cudaFuncSetAttribute(inflate_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain);
	});
	if (e != cudaSuccess)
		return e;
	e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	inflate_spec_kernel<<<plan.n_cand + 1, 64, smem_spec, s>>>(plan);
	inflate_link_kernel<<<1, 32, 0, s>>>(plan);
	if (plan.job.wrap & kWrapDry) {
		inflate_dry_finish_kernel<<<1, 32, 0, s>>>(plan);
		e = cudaGetLastError();
		if (e != cudaSuccess)
			return e;
		return launch_inflate_solo(plan.retry_job, plan.final_out, 1, counter, s);
	}
	if (plan.n_cand) {
		const uint32_t gw = plan.n_cand < (uint32_t)kNumSMs * 2 ? plan.n_cand : (uint32_t)kNumSMs * 2;
		inflate_windows_kernel<<<dim3(gw, 4), 1024, 0, s>>>(plan);
		const uint32_t gc = plan.n_cand < (uint32_t)kNumSMs * 5 ? plan.n_cand : (uint32_t)kNumSMs * 5;
		inflate_chain_kernel<<<gc, 64, smem_chain, s>>>(plan, counter);
	}
	inflate_finish_kernel<<<1, 32, 0, s>>>(plan);
	e = cudaGetLastError();
	if (e != cudaSuccess)
		return e;
	// the descriptor again, serially, should the pieces have disagreed (retry_job is a skip otherwise)
	return launch_inflate_solo(plan.retry_job, plan.final_out, 1, counter, s);
}

cudaError_t launch_inflate(const InflateJob *jobs, InflateOut *outs, uint32_t n_jobs, uint32_t *counter, cudaStream_t s)
{
	const char *sm = getenv("NXGPU_INFLATE_SOLO_MAX");          // developer / test switch: 0 = never
	const uint32_t solo_max = sm ? (uint32_t)atoi(sm) : (uint32_t)(kNumSMs * 5);
	if (n_jobs <= solo_max)
		return launch_inflate_solo(jobs, outs, n_jobs, counter, s);
	static const int occ = getenv("NXGPU_INFLATE_OCC") ? atoi(getenv("NXGPU_INFLATE_OCC")) : 7;   // developer switch
	if (occ <= 4)
		return launch_inflate_t<4>(jobs, outs, n_jobs, counter, s);
	if (occ == 5)
		return launch_inflate_t<5>(jobs, outs, n_jobs, counter, s);
	if (occ == 6)
		return launch_inflate_t<6>(jobs, outs, n_jobs, counter, s);
	return launch_inflate_t<7>(jobs, outs, n_jobs, counter, s);
}

} // namespace nxgpu
