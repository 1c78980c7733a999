// checksum.cu — CRC-32 and Adler-32 for sm_100a, replacing the
// POWER vpmsum CRC of lib/crc32_power.c:71 and the scalar loops of lib/nx_adler32.c:81, with
// the GF(2) / modular combines of lib/nx_crc.c:374 and lib/nx_adler32.c:154 done on the device.
//
// Pass 1 (checksum_ranges_kernel): the input is cut into ranges; one CTA (1024 threads, one per SM) streams a range
// in 16 KiB tiles with fully coalesced 16-byte loads straight into registers: thread t owns the 16-byte piece t of
// every tile.  CRC-32 is linear over GF(2), so each of the thread's four 32-bit word columns is its own accumulator:
// acc <- Shift_16KiB(acc) ^ word — four independent chains per thread, ONE fixed shift operator (4 x 256-entry table,
// 4 look-ups per word = 1 per byte, the same count as slice-by-4) and no shared-memory staging of the data at all.
// The shift table has a private copy per lane (entry (k, v) of lane L at word (k*256 + v)*32 + L: always bank L), so
// the data-dependent look-ups never conflict.  Adler-32 rides along with two dp4a per word (byte sum and the
// position-weighted sum of the piece; tile and piece offsets are folded in once per range).  At the end of a range
// the 4096 word columns are folded with fixed shift operators (3 Horner steps per thread, then a 10-level tree over
// the threads: shuffles inside a warp, shared memory across warps).
// Pass 2 (checksum_combine_kernel): one warp per job folds its ranges: every range is shifted by
// the bytes that follow it (x^(8n) mod P by repeated squaring) and XOR-ed; seeds are applied last.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {

namespace {

constexpr uint32_t kPoly = 0xEDB88320u;
constexpr uint32_t kBase = 65521u;
constexpr int kCkThreads = 1024;
constexpr int kPiece = 16;                          // bytes per thread per tile: one 16-byte vector load
constexpr int kTile = kCkThreads * kPiece;          // 16 KiB
constexpr int kLevels = 10;                         // log2(1024)

struct CkTables {
	uint32_t word[4][256];        // append 4 zero bytes (one slice-by-4 step without data)
	uint32_t gap[4][256];         // append kTile zero bytes
	uint32_t lvl[kLevels][4][256];// append kPiece * 2^k zero bytes
};
struct CkConst {
	uint32_t x2n[32];             // x^(2^n) mod P
	uint32_t inv_hi[128];         // x^(-8 * 256 * i)
	uint32_t inv_lo[256];         // x^(-8 * i)
};
__device__ CkTables g_tables;
__constant__ CkConst c_ck;

__host__ __device__ inline uint32_t multmodp(uint32_t a, uint32_t b)
{
	if (a == 0)
		return 0;
	uint32_t m = 1u << 31, p = 0;
	for (;;) {
		if (a & m) {
			p ^= b;
			if ((a & (m - 1)) == 0)
				break;
		}
		m >>= 1;
		b = (b & 1) ? (b >> 1) ^ kPoly : b >> 1;
	}
	return p;
}

// x^(n * 2^k) mod P
__device__ inline uint32_t x2nmodp_dev(uint64_t n, unsigned k)
{
	uint32_t p = 1u << 31;
	while (n) {
		if (n & 1)
			p = multmodp(c_ck.x2n[k & 31], p);
		n >>= 1;
		k++;
	}
	return p;
}

struct __align__(16) CkSmem {
	// the tile-shift table, one private copy per lane: entry (k, v) of lane L sits at word ((k*256 + v) * 32 + L),
	// i.e. always in bank L, so the four data-dependent look-ups per word never conflict (a single shared
	// copy costs ~3.5 wavefronts per look-up)
	uint32_t gap32[4 * 256 * 32];              // 128 KiB
	uint32_t red[kCkThreads / 32][4];
};

__device__ __forceinline__ uint32_t apply4(const uint32_t (*t)[256], uint32_t c)
{
	return t[0][c & 255] ^ t[1][(c >> 8) & 255] ^ t[2][(c >> 16) & 255] ^ t[3][c >> 24];
}
// shift a register by one tile; g = the lane's own table copy (gap32 + lane)
__device__ __forceinline__ uint32_t shift_tile(const uint32_t *g, uint32_t c)
{
	return g[(c & 255) << 5] ^ g[(256 + ((c >> 8) & 255)) << 5] ^ g[(512 + ((c >> 16) & 255)) << 5] ^ g[(768 + (c >> 24)) << 5];
}

struct Range { const uint8_t *src; uint64_t len; uint64_t after; uint32_t job; uint32_t pad_; };
struct Partial { uint32_t crc, s1, s2, pad_; };

// the thread's 16 bytes of the tile at tile_off; bytes outside [valid_lo, valid_hi) read as zero
__device__ __forceinline__ uint4 load_piece(const uint8_t *abase, uint64_t tile_off, uint64_t valid_lo, uint64_t valid_hi)
{
	const uint64_t p = tile_off + (uint64_t)threadIdx.x * kPiece;
	if (p >= valid_lo && p + kPiece <= valid_hi) {
		uint4 v;
		// streaming data, read once: do not keep it in L1
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(abase + p));
		return v;
	}
	uint32_t w[4] = { 0, 0, 0, 0 };
	if (p < valid_hi && p + kPiece > valid_lo) {
		for (int b = 0; b < kPiece; b++) {
			const uint64_t q = p + b;
			if (q >= valid_lo && q < valid_hi)
				w[b >> 2] |= (uint32_t)abase[q] << (8 * (b & 3));
		}
	}
	return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool kCrc, bool kAdler>
__global__ void __launch_bounds__(kCkThreads, 1)
checksum_ranges_kernel(const Range *__restrict__ ranges, uint32_t n_ranges, Partial *__restrict__ parts)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	CkSmem &S = *reinterpret_cast<CkSmem *>(smem_raw);
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	if (kCrc)
		for (int i = t; i < 4 * 256 * 32; i += kCkThreads)
			S.gap32[i] = (&g_tables.gap[0][0])[i >> 5];
	__syncthreads();
	const uint32_t *g = S.gap32 + lane;

	for (uint32_t r = blockIdx.x; r < n_ranges; r += gridDim.x) {
		const Range R = ranges[r];
		const uintptr_t a = reinterpret_cast<uintptr_t>(R.src);
		const uint8_t *abase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
		const uint64_t lead = a & 15, vhi = lead + R.len;
		const uint64_t ntiles = (vhi + kTile - 1) / kTile;
		uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;         // one CRC accumulator per 32-bit word column of the piece
		uint32_t cumA = 0;                               // Adler: byte sum so far (< 2^32 for ranges below 16 GiB)
		unsigned long long sumB = 0, P = 0;              // position-weighted sums inside the pieces; sum over tiles of the byte sum in front
		uint4 nxt = ntiles ? load_piece(abase, 0, lead, vhi) : make_uint4(0, 0, 0, 0);
		uint4 nxt2 = ntiles > 1 ? load_piece(abase, kTile, lead, vhi) : make_uint4(0, 0, 0, 0);
		// tiles 1 .. ntiles-2 lie wholly inside the range: their loads need no bounds checks
		const uint8_t *pp = abase + (uint64_t)t * kPiece + 2 * (uint64_t)kTile;
		const uint32_t nt = (uint32_t)ntiles;
		for (uint32_t k = 0; k < nt; k++) {
			const uint4 v = nxt;
			nxt = nxt2;
			if (k + 3 < nt) {
				// two tiles ahead: ~32 KiB in flight per SM; streaming data, read once: do not keep it in L1
				asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(nxt2.x), "=r"(nxt2.y), "=r"(nxt2.z), "=r"(nxt2.w) : "l"(pp));
			} else if (k + 2 < nt) {
				nxt2 = load_piece(abase, (uint64_t)(k + 2) * kTile, lead, vhi);
			}
			pp += kTile;
			if (kCrc) {
				c0 = shift_tile(g, c0) ^ v.x;
				c1 = shift_tile(g, c1) ^ v.y;
				c2 = shift_tile(g, c2) ^ v.z;
				c3 = shift_tile(g, c3) ^ v.w;
			}
			if (kAdler) {
				uint32_t a_ = __dp4a(v.x, 0x01010101u, 0u);
				a_ = __dp4a(v.y, 0x01010101u, a_);
				a_ = __dp4a(v.z, 0x01010101u, a_);
				a_ = __dp4a(v.w, 0x01010101u, a_);
				// byte i of the piece counts 16 - i times
				uint32_t b_ = __dp4a(v.x, 0x0d0e0f10u, 0u);
				b_ = __dp4a(v.y, 0x090a0b0cu, b_);
				b_ = __dp4a(v.z, 0x05060708u, b_);
				b_ = __dp4a(v.w, 0x01020304u, b_);
				P += cumA;
				cumA += a_;
				sumB += b_;
			}
		}
		// ---- fold: the thread's four columns (Horner, one 4-byte shift each), then the 1024 pieces as a tree ----
		uint32_t crc = 0;
		if (kCrc) {
			crc = apply4(g_tables.word, c0) ^ c1;
			crc = apply4(g_tables.word, crc) ^ c2;
			crc = apply4(g_tables.word, crc) ^ c3;
#pragma unroll
			for (int lv = 0; lv < 5; lv++) {
				const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << lv);
				if ((lane & ((2 << lv) - 1)) == 0)
					crc = apply4(g_tables.lvl[lv], crc) ^ other;
			}
		}
		uint32_t s1 = 0, s2 = 0;
		if (kAdler) {
			// S2_t = sumB + c_t * cumA + kTile * P,  c_t = bytes behind this piece inside a tile
			const uint32_t ct = (uint32_t)(kCkThreads - 1 - t) * kPiece;
			const uint32_t A = cumA % kBase;
			s1 = A;
			s2 = (uint32_t)((sumB % kBase + (uint64_t)ct * A + (uint64_t)(kTile % kBase) * (P % kBase)) % kBase);
			for (int o = 16; o; o >>= 1) {
				s1 += __shfl_xor_sync(0xffffffffu, s1, o);
				s2 += __shfl_xor_sync(0xffffffffu, s2, o);
			}
		}
		if (lane == 0) {
			S.red[warp][0] = crc;
			S.red[warp][1] = s1 % kBase;
			S.red[warp][2] = s2 % kBase;
		}
		__syncthreads();
		if (warp == 0) {
			uint32_t c = S.red[lane][0];
			uint32_t x1 = S.red[lane][1];
			uint32_t x2 = S.red[lane][2];
			if (kCrc) {
#pragma unroll
				for (int lv = 5; lv < kLevels; lv++) {
					const uint32_t other = __shfl_down_sync(0xffffffffu, c, 1 << (lv - 5));
					if ((lane & ((2 << (lv - 5)) - 1)) == 0)
						c = apply4(g_tables.lvl[lv], c) ^ other;
				}
			}
			for (int o = 16; o; o >>= 1) {
				x1 += __shfl_xor_sync(0xffffffffu, x1, o);
				x2 += __shfl_xor_sync(0xffffffffu, x2, o);
			}
			if (lane == 0) {
				// a word entered its accumulator unshifted: one more 4-byte step makes it the register "after the word";
				// then undo the zero padding behind the data (z bytes)
				const uint64_t z = ntiles * kTile - vhi;
				Partial p;
				p.crc = kCrc ? multmodp(multmodp(c_ck.inv_hi[z >> 8], c_ck.inv_lo[z & 255]), apply4(g_tables.word, c)) : 0;
				x1 %= kBase; x2 %= kBase;
				// S2 was taken over the padded message: true b = S2 - z * S1
				const uint32_t zz = (uint32_t)(z % kBase);
				p.s1 = x1;
				p.s2 = (uint32_t)((x2 + (uint64_t)kBase * kBase - (uint64_t)zz * x1) % kBase);
				p.pad_ = 0;
				parts[r] = p;
			}
		}
		__syncthreads();
	}
}

// kGroup threads per job (a warp for batches of small jobs, a whole CTA for a job cut into many
// ranges): job j owns ranges [rs[j], rs[j+1]).  Every range's partial is shifted by the bytes that
// follow it inside the job (x^(8n) mod P by square-and-multiply, ~30 GF(2) multiplications — the
// expensive part, hence one thread per range) and the group folds the results.
template <int kGroup>
__global__ void __launch_bounds__(kGroup >= 256 ? kGroup : 256)
checksum_combine_kernel(const Range *__restrict__ ranges, const Partial *__restrict__ parts,
			const uint32_t *__restrict__ rs, uint32_t n_jobs,
			const uint32_t *__restrict__ crc_seed, const uint32_t *__restrict__ adler_seed,
			uint32_t *__restrict__ crc_out, uint32_t *__restrict__ adler_out)
{
	__shared__ uint32_t red_c[32], red_s1[32], red_s2[32];
	__shared__ unsigned long long red_t[32];
	const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) / kGroup, lid = threadIdx.x % kGroup, lane = threadIdx.x & 31;
	const bool live = j < n_jobs;
	const uint32_t r0 = live ? rs[j] : 0, r1 = live ? rs[j + 1] : 0;
	uint32_t c = 0, s1 = 0, s2 = 0;
	uint64_t total = 0;
	for (uint32_t r = r0 + lid; r < r1; r += kGroup) {
		const Range R = ranges[r];
		const Partial p = parts[r];
		c ^= multmodp(x2nmodp_dev(R.after, 3), p.crc);
		s1 = (s1 + p.s1) % kBase;
		s2 = (uint32_t)((s2 + p.s2 + (uint64_t)(R.after % kBase) * p.s1) % kBase);
		total += R.len;
	}
	for (int o = 16; o; o >>= 1) {
		c ^= __shfl_xor_sync(0xffffffffu, c, o);
		s1 = (s1 + __shfl_xor_sync(0xffffffffu, s1, o)) % kBase;
		s2 = (s2 + __shfl_xor_sync(0xffffffffu, s2, o)) % kBase;
		total += __shfl_xor_sync(0xffffffffu, total, o);
	}
	if (kGroup > 32) {
		// one job per CTA: fold the warps through shared memory
		if (lane == 0) { red_c[threadIdx.x >> 5] = c; red_s1[threadIdx.x >> 5] = s1; red_s2[threadIdx.x >> 5] = s2; red_t[threadIdx.x >> 5] = total; }
		__syncthreads();
		if (threadIdx.x >= 32)
			return;
		const bool has = lane < kGroup / 32;
		c = has ? red_c[lane] : 0; s1 = has ? red_s1[lane] : 0; s2 = has ? red_s2[lane] : 0; total = has ? red_t[lane] : 0;
		for (int o = 16; o; o >>= 1) {
			c ^= __shfl_xor_sync(0xffffffffu, c, o);
			s1 = (s1 + __shfl_xor_sync(0xffffffffu, s1, o)) % kBase;
			s2 = (s2 + __shfl_xor_sync(0xffffffffu, s2, o)) % kBase;
			total += __shfl_xor_sync(0xffffffffu, total, o);
		}
	}
	if (live && lane == 0 && lid == 0) {
		if (crc_out) {
			// crc32(seed, data) = ~( shift(~seed, len) ^ raw0(data) )
			const uint32_t seed = crc_seed ? crc_seed[j] : 0;
			crc_out[j] = ~(multmodp(x2nmodp_dev(total, 3), ~seed) ^ c);
		}
		if (adler_out) {
			const uint32_t seed = adler_seed ? adler_seed[j] : 1;
			const uint32_t a0 = seed & 0xffff, b0 = seed >> 16;
			const uint32_t a = (a0 + s1) % kBase;
			const uint32_t b = (uint32_t)((b0 + (uint64_t)(total % kBase) * a0 + s2) % kBase);
			adler_out[j] = a | (b << 16);
		}
	}
}

// ranges straight from inflate results: per_job ranges per job (1 for batches of small members; a few long outputs are
// cut so that every SM takes a share — one 64 MiB output as a single range kept one CTA busy for 2 ms)
__global__ void ranges_from_inflate_kernel(const InflateJob *jobs, const InflateOut *outs, uint32_t n, Range *ranges, uint32_t *rs, uint32_t per_job)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i = t / per_job, k = t % per_job;
	if (i < n) {
		const uint64_t len = outs[i].rc == 0 || outs[i].rc == NXGPU_E_BUF ? outs[i].out_len : 0;
		uint64_t piece = (len + per_job - 1) / per_job;
		piece = (piece + kTile - 1) / kTile * kTile;
		const uint64_t from = piece * k < len ? piece * k : len;
		const uint64_t to = from + piece < len ? from + piece : len;
		Range R;
		R.src = jobs[i].dst + from; R.len = to - from; R.after = len - to; R.job = i; R.pad_ = 0;
		ranges[t] = R;
	}
	if (k == 0 && i <= n)
		rs[i] = i * per_job;
}

PerDeviceOnce g_tables_once;

uint32_t host_x8n(uint64_t n, const uint32_t *x2n)
{
	uint32_t p = 1u << 31;
	unsigned k = 3;
	while (n) {
		if (n & 1) p = multmodp(x2n[k & 31], p);
		n >>= 1; k++;
	}
	return p;
}
uint32_t host_xpow_bits(uint64_t e, const uint32_t *x2n)
{
	uint32_t p = 1u << 31;
	unsigned k = 0;
	while (e) {
		if (e & 1) p = multmodp(x2n[k & 31], p);
		e >>= 1; k++;
	}
	return p;
}
void make_shift_table(uint32_t op, uint32_t out[4][256])
{
	for (int k = 0; k < 4; k++)
		for (int b = 0; b < 256; b++)
			out[k][b] = multmodp(op, (uint32_t)b << (8 * k));
}

} // namespace

static cudaError_t upload_tables()
{
	static CkTables T;            // the caller holds g_tables_once's lock
	static CkConst C;
	uint32_t p = 1u << 30;
	C.x2n[0] = p;
	for (int n = 1; n < 32; n++)
		C.x2n[n] = p = multmodp(p, p);
	make_shift_table(host_x8n(4, C.x2n), T.word);
	make_shift_table(host_x8n(kTile, C.x2n), T.gap);
	for (int lv = 0; lv < kLevels; lv++)
		make_shift_table(host_x8n((uint64_t)kPiece << lv, C.x2n), T.lvl[lv]);
	const uint64_t ord = 0xFFFFFFFFull;          // order of x modulo the (primitive) CRC-32 polynomial
	for (int i = 0; i < 128; i++)
		C.inv_hi[i] = host_xpow_bits((ord - (8ull * 256 * i) % ord) % ord, C.x2n);
	for (int i = 0; i < 256; i++)
		C.inv_lo[i] = host_xpow_bits((ord - (8ull * i) % ord) % ord, C.x2n);
	cudaError_t e = cudaMemcpyToSymbol(g_tables, &T, sizeof(T));
	if (e != cudaSuccess) return e;
	e = cudaMemcpyToSymbol(c_ck, &C, sizeof(C));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(checksum_ranges_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CkSmem));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(checksum_ranges_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CkSmem));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(checksum_ranges_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CkSmem));
	if (e != cudaSuccess) return e;
	return cudaSuccess;
}

cudaError_t checksum_init_tables() { return g_tables_once.run(upload_tables); }

size_t checksum_range_bytes() { return sizeof(Range); }
size_t checksum_partial_bytes() { return sizeof(Partial); }

void checksum_fill_range(void *ranges, size_t idx, const void *src, uint64_t len, uint64_t after, uint32_t job)
{
	Range R;
	R.src = static_cast<const uint8_t *>(src); R.len = len; R.after = after; R.job = job; R.pad_ = 0;
	static_cast<Range *>(ranges)[idx] = R;
}

// which: 1 = crc only, 2 = adler only, 3 = both
cudaError_t launch_checksum_ranges(const void *d_ranges, uint32_t n_ranges, void *d_parts, int which, cudaStream_t s)
{
	if (n_ranges == 0)
		return cudaSuccess;
	uint32_t grid = n_ranges < (uint32_t)kNumSMs ? n_ranges : (uint32_t)kNumSMs;      // one CTA per SM (128 KiB of shared memory, 1024 threads)
	const Range *r = static_cast<const Range *>(d_ranges);
	Partial *p = static_cast<Partial *>(d_parts);
	if (which == 1)
		checksum_ranges_kernel<true, false><<<grid, kCkThreads, sizeof(CkSmem), s>>>(r, n_ranges, p);
	else if (which == 2)
		checksum_ranges_kernel<false, true><<<grid, kCkThreads, sizeof(CkSmem), s>>>(r, n_ranges, p);
	else
		checksum_ranges_kernel<true, true><<<grid, kCkThreads, sizeof(CkSmem), s>>>(r, n_ranges, p);
	return cudaGetLastError();
}

cudaError_t launch_checksum_combine(const void *d_ranges, const void *d_parts, const uint32_t *d_rs, uint32_t n_jobs,
				    const uint32_t *d_crc_seed, const uint32_t *d_adler_seed,
				    uint32_t *d_crc_out, uint32_t *d_adler_out, cudaStream_t s, uint32_t max_ranges_per_job)
{
	if (n_jobs == 0)
		return cudaSuccess;
	const Range *r = static_cast<const Range *>(d_ranges);
	const Partial *p = static_cast<const Partial *>(d_parts);
	if (max_ranges_per_job > 64) {
		// a big buffer cut into many ranges: a CTA per job, a thread per range
		checksum_combine_kernel<1024><<<n_jobs, 1024, 0, s>>>(r, p, d_rs, n_jobs, d_crc_seed, d_adler_seed, d_crc_out, d_adler_out);
	} else {
		const uint32_t warps_per_cta = 8;
		const uint32_t grid = (n_jobs + warps_per_cta - 1) / warps_per_cta;
		checksum_combine_kernel<32><<<grid, warps_per_cta * 32, 0, s>>>(r, p, d_rs, n_jobs, d_crc_seed, d_adler_seed, d_crc_out, d_adler_out);
	}
	return cudaGetLastError();
}

cudaError_t launch_ranges_from_inflate(const InflateJob *jobs, const InflateOut *outs, uint32_t n, void *d_ranges, uint32_t *d_rs, cudaStream_t s, uint32_t per_job)
{
	ranges_from_inflate_kernel<<<((n + 1) * per_job + 255) / 256, 256, 0, s>>>(jobs, outs, n, static_cast<Range *>(d_ranges), d_rs, per_job);
	return cudaGetLastError();
}

uint32_t host_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
	uint32_t x2n[32];
	uint32_t p = 1u << 30;
	x2n[0] = p;
	for (int n = 1; n < 32; n++)
		x2n[n] = p = multmodp(p, p);
	if (len2 == 0)
		return crc1;
	return multmodp(host_x8n(len2, x2n), crc1) ^ crc2;
}

} // namespace nxgpu
