// checksum.cu — CRC-32 and Adler-32 for sm_100a, replacing the
// POWER vpmsum CRC of lib/crc32_power.c:71 and the scalar loops of lib/nx_adler32.c:81, with
// the GF(2) / modular combines of lib/nx_crc.c:374 and lib/nx_adler32.c:154 done on the device.
//
// Pass 1 (checksum_ranges_kernel): the input is cut into ranges; one CTA streams a range through
// shared memory in 32 KiB tiles (cp.async, 16-byte vectors, double buffered).  Thread t owns the
// 64-byte strip t of every tile, so its running CRC register simply skips the other 511 strips
// with one "append 32704 zero bytes" operator (4 table look-ups) per tile; slice-by-4 tables sit
// in shared memory.  Strips are padded to 80 bytes in shared memory so the LDS.128 reads of a
// quarter-warp hit 8 distinct bank groups.  Adler-32 rides along with two dp4a per word.  A
// shuffle/shared-memory tree folds the 512 registers with fixed shift operators.
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
constexpr int kCkThreads = 512;
constexpr int kStrip = 64;
constexpr int kStripPad = 80;
constexpr int kTile = kCkThreads * kStrip;          // 32 KiB
constexpr int kLevels = 9;                          // log2(512)

struct CkTables {
	uint32_t slice[4][256];       // slice-by-4
	uint32_t gap[4][256];         // append (kTile - kStrip) zero bytes
	uint32_t lvl[kLevels][4][256];// append kStrip * 2^k zero bytes
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
	uint8_t tile[2][kCkThreads * kStripPad];   // 2 x 40 KiB
	// slice-by-4 tables, one private copy per lane: entry (k, v) of lane L sits at word ((k*256 + v) * 32 + L),
	// i.e. always in bank L, so the four data-dependent look-ups per word never conflict (a single shared
	// copy cost ~3.5 wavefronts per look-up and bounded the kernel at a third of the HBM peak)
	uint32_t slice32[4 * 256 * 32];            // 128 KiB
	uint32_t gap[4][256];
	uint32_t red[kCkThreads / 32][4];
};

__device__ __forceinline__ uint32_t apply4(const uint32_t (*t)[256], uint32_t c)
{
	return t[0][c & 255] ^ t[1][(c >> 8) & 255] ^ t[2][(c >> 16) & 255] ^ t[3][c >> 24];
}
// crc register after one more little-endian word; sl = the lane's own table copy (slice32 + lane)
__device__ __forceinline__ uint32_t crc_word(const uint32_t *sl, uint32_t c, uint32_t w)
{
	c ^= w;
	return sl[(3 * 256 + (c & 255)) << 5] ^ sl[(2 * 256 + ((c >> 8) & 255)) << 5] ^ sl[(1 * 256 + ((c >> 16) & 255)) << 5] ^ sl[(c >> 24) << 5];
}

struct Range { const uint8_t *src; uint64_t len; uint64_t after; uint32_t job; uint32_t pad_; };
struct Partial { uint32_t crc, s1, s2, pad_; };

__device__ void stage_tile(CkSmem &S, int buf, const uint8_t *abase, uint64_t tile_off, uint64_t valid_lo, uint64_t valid_hi)
{
	// thread t stages its own 64-byte strip (4 x 16 B); out-of-range bytes become zero
	uint8_t *dst = S.tile[buf] + threadIdx.x * kStripPad;
	const uint64_t o = tile_off + (uint64_t)threadIdx.x * kStrip;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint64_t p = o + 16 * k;
		if (p >= valid_lo && p + 16 <= valid_hi) {
			uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + 16 * k);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(abase + p));
		} else {
			uint32_t w[4] = { 0, 0, 0, 0 };
			for (int b = 0; b < 16; b++) {
				const uint64_t q = p + b;
				if (q >= valid_lo && q < valid_hi)
					w[b >> 2] |= (uint32_t)abase[q] << (8 * (b & 3));
			}
			*reinterpret_cast<uint4 *>(dst + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
		}
	}
	asm volatile("cp.async.commit_group;\n" ::);
}

template <bool kCrc, bool kAdler>
__global__ void __launch_bounds__(kCkThreads, 1)
checksum_ranges_kernel(const Range *__restrict__ ranges, uint32_t n_ranges, Partial *__restrict__ parts)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	CkSmem &S = *reinterpret_cast<CkSmem *>(smem_raw);
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	for (int i = t; i < 1024; i += kCkThreads)
		(&S.gap[0][0])[i] = (&g_tables.gap[0][0])[i];
	if (kCrc)
		for (int i = t; i < 4 * 256 * 32; i += kCkThreads)
			S.slice32[i] = (&g_tables.slice[0][0])[i >> 5];
	__syncthreads();
	const uint32_t *sl = S.slice32 + lane;

	for (uint32_t r = blockIdx.x; r < n_ranges; r += gridDim.x) {
		const Range R = ranges[r];
		const uintptr_t a = reinterpret_cast<uintptr_t>(R.src);
		const uint8_t *abase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
		const uint64_t lead = a & 15, vhi = lead + R.len;
		const uint64_t ntiles = (vhi + kTile - 1) / kTile;
		uint32_t crc = 0;
		uint32_t cumA = 0, sumB = 0, P = 0;      // Adler partial sums of this thread, mod 65521
		if (ntiles)
			stage_tile(S, 0, abase, 0, lead, vhi);
		for (uint64_t k = 0; k < ntiles; k++) {
			if (k + 1 < ntiles) {
				stage_tile(S, (int)((k + 1) & 1), abase, (k + 1) * kTile, lead, vhi);
				asm volatile("cp.async.wait_group 1;\n" ::);
			} else {
				asm volatile("cp.async.wait_group 0;\n" ::);
			}
			// each thread reads only what it staged itself: no CTA barrier needed
			const uint4 *src = reinterpret_cast<const uint4 *>(S.tile[k & 1] + t * kStripPad);
			if (kCrc && k)
				crc = apply4(S.gap, crc);
			uint32_t a_ = 0, b_ = 0;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const uint4 v = src[q];
				const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int j = 0; j < 4; j++) {
					if (kCrc)
						crc = crc_word(sl, crc, w[j]);
					if (kAdler) {
						b_ += 4 * a_;
						b_ = __dp4a(w[j], 0x01020304u, b_);
						a_ = __dp4a(w[j], 0x01010101u, a_);
					}
				}
			}
			if (kAdler) {
				// a_ <= 16320, b_ <= 64*255*... < 2^21: reduce once per strip
				P = (P + cumA) % kBase;
				cumA = (cumA + a_) % kBase;
				sumB = (sumB + b_) % kBase;
			}
		}
		// ---- fold the 512 strip registers: neighbour pairs, shift doubles each level ----
		if (kCrc) {
#pragma unroll
			for (int lv = 0; lv < 5; lv++) {
				const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << lv);
				if ((lane & ((2 << lv) - 1)) == 0)
					crc = apply4(g_tables.lvl[lv], crc) ^ other;
			}
		}
		uint32_t s1 = 0, s2 = 0;
		if (kAdler) {
			// S2_t = sumB + c_t * cumA + kTile * P,  c_t = bytes after this strip inside a tile
			const uint32_t ct = (uint32_t)(kCkThreads - 1 - t) * kStrip;
			s1 = cumA;
			s2 = (uint32_t)((sumB + (uint64_t)ct * cumA + (uint64_t)(kTile % kBase) * P) % kBase);
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
			uint32_t c = lane < kCkThreads / 32 ? S.red[lane][0] : 0;
			uint32_t x1 = lane < kCkThreads / 32 ? S.red[lane][1] : 0;
			uint32_t x2 = lane < kCkThreads / 32 ? S.red[lane][2] : 0;
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
				// undo the zero padding behind the data (z bytes) and account for the lead-in
				const uint64_t z = ntiles * kTile - vhi;
				Partial p;
				p.crc = kCrc ? multmodp(multmodp(c_ck.inv_hi[z >> 8], c_ck.inv_lo[z & 255]), c) : 0;
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

// ranges straight from inflate results: one range per job
__global__ void ranges_from_inflate_kernel(const InflateJob *jobs, const InflateOut *outs, uint32_t n, Range *ranges, uint32_t *rs)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		Range R;
		R.src = jobs[i].dst; R.len = outs[i].rc == 0 || outs[i].rc == NXGPU_E_BUF ? outs[i].out_len : 0; R.after = 0; R.job = i; R.pad_ = 0;
		ranges[i] = R;
	}
	if (i <= n)
		rs[i] = i;
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
	for (uint32_t n = 0; n < 256; n++) {
		uint32_t c = n;
		for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ kPoly : c >> 1;
		T.slice[0][n] = c;
	}
	for (int k = 1; k < 4; k++)
		for (int n = 0; n < 256; n++)
			T.slice[k][n] = (T.slice[k - 1][n] >> 8) ^ T.slice[0][T.slice[k - 1][n] & 255];
	uint32_t p = 1u << 30;
	C.x2n[0] = p;
	for (int n = 1; n < 32; n++)
		C.x2n[n] = p = multmodp(p, p);
	make_shift_table(host_x8n(kTile - kStrip, C.x2n), T.gap);
	for (int lv = 0; lv < kLevels; lv++)
		make_shift_table(host_x8n((uint64_t)kStrip << lv, C.x2n), T.lvl[lv]);
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
	uint32_t grid = n_ranges < (uint32_t)kNumSMs ? n_ranges : (uint32_t)kNumSMs;      // one CTA per SM (212 KiB of shared memory)
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

cudaError_t launch_ranges_from_inflate(const InflateJob *jobs, const InflateOut *outs, uint32_t n, void *d_ranges, uint32_t *d_rs, cudaStream_t s)
{
	ranges_from_inflate_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(jobs, outs, n, static_cast<Range *>(d_ranges), d_rs);
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
