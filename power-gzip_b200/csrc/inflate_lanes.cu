// inflate_lanes.cu — batched inflate of MANY independent members (BASELINE.json configs[2]: 100 000 x 64 KiB gzip members):
// ONE LANE PER MEMBER.  The Huffman walk of a deflate stream is serial, so a warp that decodes one member keeps 31 lanes
// idle during the walk (inflate.cu: 59 warp instructions per symbol, 8.6 lanes active).  Here every lane of a warp is a
// complete scalar inflater working on its own member: one warp instruction advances 32 members at once.
//
//   * Decode tables live in shared memory, INTERLEAVED by lane: word i of lane L sits at word i*32 + L, i.e. always in
//     bank L — 32 data-dependent look-ups of one instruction never conflict.
//     Per lane (562 words, 70.3 KiB per warp, 3 warps per SM): 9-bit literal/length root table and 8-bit distance root
//     table (16-bit entries), the canonical `sorted symbols` + per-length counts for the codes longer than the roots.
//   * Bit reader: 64-bit buffer per lane, refilled 32 bits at a time from the lane's own compressed stream (L1-cached).
//   * Output: literals are byte stores; matches are copied a 32-bit word at a time (one sliding source word, a funnel
//     shift, one aligned store) once the destination is aligned; distances below 4 go byte by byte.
//   * A lane that finishes its member takes the next one from a global counter; block headers are parsed and their tables
//     built by the lane itself (all lanes start with a header at the same time, so the first build runs in lockstep).
//
// Same results as inflate.cu's warp-per-member kernel (which stays in charge of NX job descriptors, dry runs and small
// batches): InflateOut per member, container header/trailer parsed here, checksums by checksum.cu afterwards.
// Malformed streams are rejected exactly where zlib rejects them (inftrees.c: over-subscribed and incomplete sets — an
// incomplete set only with a single one-bit code —, missing end-of-block code, too many symbols, invalid stored
// lengths, distances beyond the window, invalid codes).
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {
namespace {

constexpr int kWarps = 3;                 // warps per CTA, one CTA per SM
constexpr int kLitRoot = 9, kDistRoot = 8;
// per-lane regions, in 32-bit words
constexpr int W_LIT = 0;                  // 512 u16 entries
constexpr int W_DIST = 256;               // 256 u16 entries
constexpr int W_SORT = 384;               // 320 u16: literal/length symbols sorted by code, then the distance symbols at +288
constexpr int W_CNT = 544;                // u16 count[16] for lit (8 words) + dist (8 words)
constexpr int W_LONG = 560;               // per table: canonical `first` | `index` << 16 after the root length (where the long-code walk resumes)
constexpr int kLaneWords = 562;
// scratch while a header is parsed (the regions they alias are rebuilt afterwards)
constexpr int W_LENS = W_LIT;             // 320 u8 code lengths = 80 words
constexpr int W_CL = W_SORT;              // 128 u8: 7-bit table of the code-length code = 32 words

struct Tab {
	uint32_t *base;                       // this warp's block + lane
	__device__ __forceinline__ uint32_t &w(int i) const { return base[i * 32]; }
	__device__ __forceinline__ uint32_t get16(int region, uint32_t e) const
	{
		const uint32_t v = base[(region + (e >> 1)) * 32];
		return (e & 1) ? v >> 16 : v & 0xffffu;
	}
	__device__ __forceinline__ void set16(int region, uint32_t e, uint32_t x) const
	{
		uint32_t &v = base[(region + (e >> 1)) * 32];
		v = (e & 1) ? (v & 0xffffu) | (x << 16) : (v & 0xffff0000u) | (x & 0xffffu);
	}
	__device__ __forceinline__ uint32_t get8(int region, uint32_t e) const { return (base[(region + (e >> 2)) * 32] >> ((e & 3) * 8)) & 0xffu; }
	__device__ __forceinline__ void set8(int region, uint32_t e, uint32_t x) const
	{
		uint32_t &v = base[(region + (e >> 2)) * 32];
		const uint32_t sh = (e & 3) * 8;
		v = (v & ~(0xffu << sh)) | (x << sh);
	}
};

// LSB-first bit reader over a byte range of global memory; words beyond the range read as zero
struct Bits {
	const uint32_t *in32;
	uint32_t wpos, wend;                  // next word to load, one past the last word holding member bytes
	uint32_t skip8;                       // bits of the first word in front of the first byte
	uint64_t bb;
	uint32_t bn;
	__device__ __forceinline__ void init(const uint8_t *p, uint32_t nbytes)
	{
		const uintptr_t a = reinterpret_cast<uintptr_t>(p);
		in32 = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		const uint32_t sk = (uint32_t)(a & 3);
		skip8 = sk * 8;
		wend = (sk + nbytes + 3) >> 2;
		wpos = 0; bb = 0; bn = 0;
		if (wend) { bb = __ldg(in32) >> skip8; bn = 32 - skip8; wpos = 1; }
	}
	__device__ __forceinline__ void refill()
	{
		if (bn < 32) {
			const uint32_t v = wpos < wend ? __ldg(in32 + wpos) : 0;
			wpos++;
			bb |= (uint64_t)v << bn;
			bn += 32;
		}
	}
	__device__ __forceinline__ uint32_t peek() const { return (uint32_t)bb; }
	__device__ __forceinline__ void drop(uint32_t n) { bb >>= n; bn -= n; }
	__device__ __forceinline__ uint32_t get(uint32_t n) { refill(); const uint32_t v = (uint32_t)bb & ((1u << n) - 1); drop(n); return v; }    // n <= 16
	// bits consumed since init
	__device__ __forceinline__ uint64_t used() const { return (uint64_t)wpos * 32 - bn - skip8; }
};

// count[1..15] and the symbols sorted by (length, symbol) from the code lengths at lens_off (u8, region W_LENS);
// scratch counters live in the distance root table's words (it is filled afterwards).
// returns 0 ok, 1 incomplete, -1 over-subscribed; maxlen = longest code
__device__ int count_and_sort(const Tab &T, int lens_off, int n, int cnt_word, int sort_off, int root, int long_word, int &maxlen)
{
	for (int i = 0; i < 32; i++) T.w(W_DIST + i) = 0;
	for (int s = 0; s < n; s++)
		T.w(W_DIST + T.get8(W_LENS, lens_off + s)) += 1;
	int left = 1;
	maxlen = 0;
	uint32_t o = 0;
	for (int i = 1; i < 16; i++) {
		const uint32_t c = T.w(W_DIST + i);
		left = (left << 1) - (int)c;
		if (left < 0) return -1;
		T.w(W_DIST + 16 + i) = o; o += c;
		if (c) maxlen = i;
	}
	for (int i = 0; i < 8; i++)
		T.w(cnt_word + i) = (i ? T.w(W_DIST + 2 * i) : 0) | T.w(W_DIST + 2 * i + 1) << 16;
	uint32_t first = 0, index = 0;
	for (int l = 1; l <= root; l++) {
		const uint32_t c = T.w(W_DIST + l);
		index += c; first = (first + c) << 1;
	}
	T.w(long_word) = first | index << 16;
	for (int s = 0; s < n; s++) {
		const uint32_t l = T.get8(W_LENS, lens_off + s);
		if (l) {
			const uint32_t at = T.w(W_DIST + 16 + l);
			T.w(W_DIST + 16 + l) = at + 1;
			T.set16(W_SORT, sort_off + at, s);
		}
	}
	return left > 0 ? 1 : 0;
}

// root table from the sorted symbols: entry = codelen | payload << 4 for codes up to `root` bits, 0 for longer ones
__device__ void fill_root(const Tab &T, int region, int root, int cnt_word, int sort_off, bool is_dist)
{
	for (int i = 0; i < (1 << root) / 2; i++)
		T.w(region + i) = 0;
	uint32_t code = 0, idx = 0;
	for (int l = 1; l <= root; l++) {
		const uint32_t c = T.get16(cnt_word, l);
		for (uint32_t k = 0; k < c; k++, idx++, code++) {
			const uint32_t s = T.get16(W_SORT, sort_off + idx);
			uint32_t payload;
			if (is_dist) {
				payload = s;                                  // distance symbol 0..29 (30, 31 are invalid codes)
			} else if (s <= 256) {
				payload = s << 1;                             // literal or end of block
			} else {
				// length symbol: extra bits and base - 3 (RFC 1951 §3.2.5)
				const uint32_t k2 = s - 257;
				uint32_t extra, base;
				if (k2 < 8) { extra = 0; base = k2; }
				else if (k2 == 28) { extra = 0; base = 255; }
				else if (k2 > 28) { extra = 7; base = 0; }     // symbols 286, 287: invalid
				else { extra = (k2 >> 2) - 1; base = ((4 | (k2 & 3)) << extra); }
				payload = 1 | extra << 1 | base << 4;
			}
			const uint32_t e = (uint32_t)l | payload << 4;
			const uint32_t rev = __brev(code) >> (32 - l);
			for (uint32_t x = rev; x < (1u << root); x += (1u << l))
				T.set16(region, x, e);
		}
		code <<= 1;
	}
}

// canonical decode of a code LONGER than the root table: resumes the canonical walk behind length `root`
// (code = the root bits MSB-first, first / index as left by lengths 1..root); returns the symbol or -1 (no such code)
__device__ __forceinline__ int decode_long(const Tab &T, const Bits &B, int root, int cnt_word, int sort_off, int long_word, uint32_t &len_out)
{
	const uint32_t bits = B.peek();
	const uint32_t fi = T.w(long_word);
	int code = (int)(__brev(bits) >> (32 - root)) << 1;
	int first = (int)(fi & 0xffffu), index = (int)(fi >> 16);
	for (int l = root + 1; l <= 15; l++) {
		code |= (int)((bits >> (l - 1)) & 1);
		const int c = (int)T.get16(cnt_word, l);
		if (code - c < first) {
			len_out = (uint32_t)l;
			return (int)T.get16(W_SORT, sort_off + index + (code - first));
		}
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}

// o[0..len) = o[-dist..): the LZ77 copy of one lane.  Every load is a round trip to L2 (the bytes were written by this
// lane a moment ago), so the loop moves 16 or 32 bytes per trip with all loads of a step in flight, and a short
// distance is first widened: once k*dist bytes of a period exist, copying from k*dist back is the same copy.
__device__ __forceinline__ uint32_t ldv(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ void copy_match(uint8_t *o, uint32_t dist, uint32_t len)
{
	if (dist < 16) {
		uint32_t wide = dist;
		while (wide < 16) wide += dist;
		const uint32_t pre = min(len, wide - dist);
		for (uint32_t k = 0; k < pre; k++)
			o[k] = *(volatile const uint8_t *)(o + k - dist);
		o += pre; len -= pre; dist = wide;
	}
	// bytes, then words, until the destination is 16-byte aligned (dist >= 16: a source word never overlaps its destination)
	while (len && (reinterpret_cast<uintptr_t>(o) & 3)) { *o = *(volatile const uint8_t *)(o - dist); o++; len--; }
	if (len < 4) {
		while (len) { *o = *(volatile const uint8_t *)(o - dist); o++; len--; }
		return;
	}
	const uint8_t *sp = o - dist;
	const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3) * 8;
	const uint32_t *s32 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)3);
	uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
	uint32_t nw = len >> 2;
	while (nw && (reinterpret_cast<uintptr_t>(o32) & 15)) {
		const uint32_t a = ldv(s32), b = sh ? ldv(s32 + 1) : 0;
		*o32 = sh ? __funnelshift_r(a, b, sh) : a;
		o32++; s32++; nw--;
	}
	if (dist >= 32) {
		while (nw >= 8) {
			uint32_t w[9];
#pragma unroll
			for (int k = 0; k < 8; k++) w[k] = ldv(s32 + k);
			w[8] = sh ? ldv(s32 + 8) : 0;
			uint4 x, y;
			if (sh) {
				x = make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
				y = make_uint4(__funnelshift_r(w[4], w[5], sh), __funnelshift_r(w[5], w[6], sh), __funnelshift_r(w[6], w[7], sh), __funnelshift_r(w[7], w[8], sh));
			} else {
				x = make_uint4(w[0], w[1], w[2], w[3]); y = make_uint4(w[4], w[5], w[6], w[7]);
			}
			reinterpret_cast<uint4 *>(o32)[0] = x;
			reinterpret_cast<uint4 *>(o32)[1] = y;
			o32 += 8; s32 += 8; nw -= 8;
		}
	}
	while (nw >= 4) {
		uint32_t w[5];
#pragma unroll
		for (int k = 0; k < 4; k++) w[k] = ldv(s32 + k);
		w[4] = sh ? ldv(s32 + 4) : 0;
		const uint4 x = sh ? make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh))
				   : make_uint4(w[0], w[1], w[2], w[3]);
		*reinterpret_cast<uint4 *>(o32) = x;
		o32 += 4; s32 += 4; nw -= 4;
	}
	while (nw) {
		const uint32_t a = ldv(s32), b = sh ? ldv(s32 + 1) : 0;
		*o32 = sh ? __funnelshift_r(a, b, sh) : a;
		o32++; s32++; nw--;
	}
	o = reinterpret_cast<uint8_t *>(o32);
	len &= 3;
	while (len) { *o = *(volatile const uint8_t *)(o - dist); o++; len--; }
}

__device__ __forceinline__ void put_result(InflateOut &O, int rc, uint32_t out_len, uint32_t in_used, uint32_t flags, uint32_t tcrc, uint32_t tsize)
{
	O.rc = rc; O.out_len = out_len; O.in_used = in_used; O.flags = flags; O.trailer_crc = tcrc; O.trailer_isize = tsize;
	O.sfbt = 0; O.subc = 0; O.rembytecnt = 0; O.dhtlen = 0;
}

enum { ST_NEED_JOB = 0, ST_HEADER = 1, ST_HUFF = 2, ST_FINISH = 3, ST_IDLE = 4 };

__global__ void __launch_bounds__(kWarps * 32, 1)
inflate_lanes_kernel(const InflateJob *__restrict__ jobs, InflateOut *__restrict__ outs, uint32_t n_jobs, uint32_t *next_job)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	Tab T;
	T.base = smem + (size_t)warp * kLaneWords * 32 + lane;

	Bits B;
	B.in32 = nullptr; B.wpos = B.wend = 0; B.skip8 = 0; B.bb = 0; B.bn = 0;
	uint32_t state = ST_NEED_JOB;
	uint32_t job = 0;
	const uint8_t *src = nullptr;
	uint32_t src_len = 0, start = 0, wrap = 0;
	uint8_t *out = nullptr;
	uint32_t op = 0, ocap = 0, hist = 0;
	int rc = 0;
	bool final_block = false;
	bool dist_ok = true, lit_long = false, dist_long = false;

	for (;;) {
		// ---- a new member for every idle lane ----
		if (state == ST_NEED_JOB) {
			job = atomicAdd(next_job, 1u);
			if (job >= n_jobs) {
				state = ST_IDLE;
			} else {
				const InflateJob J = jobs[job];
				src = J.src; src_len = J.src_len; out = J.dst; ocap = J.dst_cap; hist = J.hist_len; wrap = J.wrap;
				op = 0; rc = 0; final_block = false; start = 0;
				const uint8_t *s = src;
				const uint32_t n = src_len;
				// container header (lib/nx_inflate.c:329-730 does this on the host)
				if (wrap == NXGPU_WRAP_AUTO) {
					if (n >= 2 && s[0] == 0x1f && s[1] == 0x8b) wrap = NXGPU_WRAP_GZIP;
					else if (n >= 2 && (s[0] & 0x0f) == 8 && (((uint32_t)s[0] << 8 | s[1]) % 31) == 0) wrap = NXGPU_WRAP_ZLIB;
					else wrap = NXGPU_WRAP_RAW;
				}
				if (wrap == NXGPU_WRAP_GZIP) {
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
				if (rc) {
					state = ST_FINISH;
				} else {
					B.init(src + start, src_len - start);
					state = ST_HEADER;
				}
			}
		}
		if (__all_sync(0xffffffffu, state == ST_IDLE))
			break;

		// ---- block header: stored blocks are copied here, Huffman blocks get their tables ----
		if (state == ST_HEADER) {
			const uint32_t h = B.get(3);
			final_block = h & 1;
			const uint32_t btype = h >> 1;
			if (btype == 0) {
				B.drop(B.bn & 7);                       // to the byte boundary
				const uint32_t len = B.get(16), nlen = B.get(16);
				if ((len ^ nlen) != 0xffffu) {
					rc = NXGPU_E_DATA; state = ST_FINISH;
				} else {
					const uint64_t at = (B.used() >> 3) + start;            // byte offset in the member of the stored data
					if (at + len > src_len) { rc = NXGPU_E_DATA; state = ST_FINISH; }
					else if (len > ocap - op) { rc = NXGPU_E_BUF; state = ST_FINISH; }
					else {
						const uint8_t *sp = src + at;
						uint8_t *dp = out + op;
						for (uint32_t i = 0; i < len; i++) dp[i] = sp[i];
						op += len;
						B.init(src + at + len, src_len - (uint32_t)(at + len));
						start = (uint32_t)(at + len);
						state = final_block ? ST_FINISH : ST_HEADER;
					}
				}
			} else if (btype == 3) {
				rc = NXGPU_E_DATA; state = ST_FINISH;
			} else {
				int hlit = 288, hdist = 32;            // fixed code: 32 five-bit distance codes (30 and 31 are invalid symbols)
				if (btype == 1) {
					for (int i = 0; i < 80; i++) {
						const int s0 = 4 * i;
						const uint32_t v = s0 < 144 ? 0x08080808u : s0 < 256 ? 0x09090909u : s0 < 280 ? 0x07070707u : s0 < 288 ? 0x08080808u : 0x05050505u;
						T.w(W_LENS + i) = v;
					}
				} else {
					const uint32_t v = B.get(14);
					hlit = (int)(v & 31) + 257; hdist = (int)((v >> 5) & 31) + 1;
					const int hclen = (int)(v >> 10) + 4;
					if (hlit > 286 || hdist > 30) rc = NXGPU_E_DATA;
					// code-length code: 19 symbols of up to 7 bits -> a 7-bit table of (len | sym << 3)
					uint32_t cl_lo = 0, cl_hi = 0, cl_x = 0;            // 19 x 3-bit lengths packed: symbols 0..9, 10..18
					for (int i = 0; i < hclen; i++) {
						static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
						const uint32_t l = B.get(3);
						const int sy = order[i];
						if (sy < 10) cl_lo |= l << (3 * sy); else cl_hi |= l << (3 * (sy - 10));
					}
					(void)cl_x;
					uint64_t ccnt = 0;                                  // 8 x 8-bit counters
					for (int sy = 0; sy < 19; sy++) {
						const uint32_t l = sy < 10 ? (cl_lo >> (3 * sy)) & 7 : (cl_hi >> (3 * (sy - 10))) & 7;
						if (l) ccnt += 1ull << (8 * l);
					}
					int left = 1;
					uint32_t nc[8];
					uint32_t code = 0;
					int cmax = 0;
#pragma unroll
					for (int l = 1; l <= 7; l++) {
						const int c = (int)((ccnt >> (8 * l)) & 255);
						left = (left << 1) - c;
						nc[l] = code;
						code = (code + c) << 1;
						if (c) cmax = l;
					}
					// zlib (inftrees.c, type CODES): over-subscribed or incomplete -> "invalid code lengths set"
					if (left < 0 || left > 0) rc = NXGPU_E_DATA;
					if (!rc) {
						for (int i = 0; i < 32; i++) T.w(W_CL + i) = 0;
						for (int sy = 0; sy < 19; sy++) {
							const uint32_t l = sy < 10 ? (cl_lo >> (3 * sy)) & 7 : (cl_hi >> (3 * (sy - 10))) & 7;
							if (!l) continue;
							uint32_t cd = 0;
#pragma unroll
							for (int q = 1; q <= 7; q++) if (l == (uint32_t)q) { cd = nc[q]; nc[q]++; }
							const uint32_t rev = __brev(cd) >> (32 - l);
							for (uint32_t x = rev; x < 128; x += (1u << l))
								T.set8(W_CL, x, l | (uint32_t)sy << 3);
						}
						(void)cmax;
						int n = 0;
						uint32_t prev = 0;
						while (n < hlit + hdist) {
							B.refill();
							const uint32_t e = T.get8(W_CL, B.peek() & 127);
							const uint32_t l = e & 7, sy = e >> 3;
							if (!l) { rc = NXGPU_E_DATA; break; }
							B.drop(l);
							if (sy < 16) { T.set8(W_LENS, n, sy); prev = sy; n++; continue; }
							uint32_t rep, val = 0;
							if (sy == 16) { if (n == 0) { rc = NXGPU_E_DATA; break; } val = prev; rep = 3 + B.get(2); }
							else if (sy == 17) rep = 3 + B.get(3);
							else rep = 11 + B.get(7);
							if (n + (int)rep > hlit + hdist) { rc = NXGPU_E_DATA; break; }
							for (uint32_t k = 0; k < rep; k++) T.set8(W_LENS, n + k, val);
							n += (int)rep;
							prev = val;
						}
						if (!rc && T.get8(W_LENS, 256) == 0) rc = NXGPU_E_DATA;          // missing end-of-block code
					}
				}
				if (!rc) {
					// distance lengths behind the literal/length ones: keep them where they are, build from (lens_off = hlit)
					int lmax = 0, dmax = 0;
					const int lr = count_and_sort(T, 0, hlit, W_CNT, 0, kLitRoot, W_LONG, lmax);
					const int dr = count_and_sort(T, hlit, hdist, W_CNT + 8, 288, kDistRoot, W_LONG + 1, dmax);
					// zlib: over-subscribed always fails; incomplete only passes when the longest code is 1 bit (or there is no code at all)
					if (lr < 0 || dr < 0 || (lr > 0 && lmax > 1) || (dr > 0 && dmax > 1)) rc = NXGPU_E_DATA;
					dist_ok = dmax > 0;
					lit_long = lmax > kLitRoot; dist_long = dmax > kDistRoot;
					if (!rc) {
						fill_root(T, W_LIT, kLitRoot, W_CNT, 0, false);
						fill_root(T, W_DIST, kDistRoot, W_CNT + 8, 288, true);
					}
				}
				state = rc ? ST_FINISH : ST_HUFF;
			}
		}

		// ---- symbols: a burst per pass so that headers and new members of other lanes do not interleave with every symbol ----
		if (state == ST_HUFF) {
#pragma unroll 1
			for (int it = 0; it < 256 && state == ST_HUFF; it++) {
				B.refill();
				uint32_t e = T.get16(W_LIT, B.peek() & ((1u << kLitRoot) - 1));
				uint32_t l = e & 15;
				uint32_t sym_or_payload = e >> 4;
				if (l == 0) {
					// a code longer than the root table (or no such code)
					uint32_t ll = 0;
					const int s = lit_long ? decode_long(T, B, kLitRoot, W_CNT, 0, W_LONG, ll) : -1;
					if (s < 0) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }
					l = ll;
					if (s <= 256) sym_or_payload = (uint32_t)s << 1;
					else {
						const uint32_t k2 = (uint32_t)s - 257;
						uint32_t extra, base;
						if (k2 < 8) { extra = 0; base = k2; }
						else if (k2 == 28) { extra = 0; base = 255; }
						else if (k2 > 28) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }
						else { extra = (k2 >> 2) - 1; base = ((4 | (k2 & 3)) << extra); }
						sym_or_payload = 1 | extra << 1 | base << 4;
					}
				}
				B.drop(l);
				if (!(sym_or_payload & 1)) {
					const uint32_t s = sym_or_payload >> 1;
					if (s < 256) {
						if (op >= ocap) { rc = NXGPU_E_BUF; state = ST_FINISH; break; }
						out[op++] = (uint8_t)s;
						continue;
					}
					state = final_block ? ST_FINISH : ST_HEADER;     // end of block
					break;
				}
				// length + distance
				const uint32_t lextra = (sym_or_payload >> 1) & 7;
				if (lextra == 7) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }      // symbols 286 / 287
				uint32_t len = (sym_or_payload >> 4) + 3 + (B.peek() & ((1u << lextra) - 1));
				B.drop(lextra);
				B.refill();
				uint32_t d = T.get16(W_DIST, B.peek() & ((1u << kDistRoot) - 1));
				uint32_t dl = d & 15, dsym = d >> 4;
				if (dl == 0) {
					uint32_t ll = 0;
					const int s = (dist_ok && dist_long) ? decode_long(T, B, kDistRoot, W_CNT + 8, 288, W_LONG + 1, ll) : -1;
					if (s < 0) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }
					dl = ll; dsym = (uint32_t)s;
				}
				if (dsym >= 30) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }
				B.drop(dl);
				uint32_t dist;
				if (dsym < 4) {
					dist = dsym + 1;
				} else {
					const uint32_t dextra = (dsym >> 1) - 1;
					B.refill();
					dist = ((2 | (dsym & 1)) << dextra) + 1 + (B.peek() & ((1u << dextra) - 1));
					B.drop(dextra);
				}
				if (dist > op + hist) { rc = NXGPU_E_DATA; state = ST_FINISH; break; }
				if (len > ocap - op) { rc = NXGPU_E_BUF; state = ST_FINISH; break; }
				uint8_t *o = out + op;
				op += len;
				copy_match(o, dist, len);
			}
		}

		// ---- member done: trailer, result ----
		if (state == ST_FINISH) {
			uint32_t in_used = 0, tcrc = 0, tsize = 0, flags = 0;
			if (!rc) {
				flags |= 1;
				const uint64_t usedbits = B.used();
				uint64_t p = start + ((usedbits + 7) >> 3);
				if (p > src_len) rc = NXGPU_E_DATA;                              // the stream ran past its own end
				const uint8_t *s = src;
				if (!rc && wrap == NXGPU_WRAP_GZIP) {
					if (p + 8 > src_len) rc = NXGPU_E_DATA;
					else {
						tcrc = s[p] | (uint32_t)s[p + 1] << 8 | (uint32_t)s[p + 2] << 16 | (uint32_t)s[p + 3] << 24;
						tsize = s[p + 4] | (uint32_t)s[p + 5] << 8 | (uint32_t)s[p + 6] << 16 | (uint32_t)s[p + 7] << 24;
						p += 8;
					}
				} else if (!rc && wrap == NXGPU_WRAP_ZLIB) {
					if (p + 4 > src_len) rc = NXGPU_E_DATA;
					else {
						tcrc = (uint32_t)s[p] << 24 | (uint32_t)s[p + 1] << 16 | (uint32_t)s[p + 2] << 8 | s[p + 3];
						p += 4;
					}
				}
				in_used = rc ? 0 : (uint32_t)p;
			}
			put_result(outs[job], rc, op, in_used, flags | (wrap << 8), tcrc, tsize);
			state = ST_NEED_JOB;
		}
	}
}

} // namespace

size_t inflate_lanes_smem() { return (size_t)kWarps * kLaneWords * 32 * 4; }

cudaError_t launch_inflate_lanes(const InflateJob *jobs, InflateOut *outs, uint32_t n_jobs, uint32_t *counter, cudaStream_t s)
{
	static PerDeviceOnce once;
	const size_t smem = inflate_lanes_smem();
	cudaError_t e = once.run([smem] { return cudaFuncSetAttribute(inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
	if (e != cudaSuccess)
		return e;
	e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
	if (e != cudaSuccess)
		return e;
	uint32_t grid = (n_jobs + kWarps * 32 - 1) / (kWarps * 32);
	if (grid > (uint32_t)kNumSMs)
		grid = kNumSMs;
	inflate_lanes_kernel<<<grid ? grid : 1, kWarps * 32, smem, s>>>(jobs, outs, n_jobs, counter);
	return cudaGetLastError();
}

} // namespace nxgpu
