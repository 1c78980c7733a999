// stitch.cu — joins per-chunk deflate outputs into one stream on the device: exclusive scan of
// the chunk sizes, gather into the final buffer at each chunk's offset, container trailer.
// This is the "stitch" of SURVEY.md §7 step 6; the reference's host code does the serial
// equivalent as it walks next_out (lib/nx_deflate.c:1051-1075, trailer :428-470).
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {
namespace {

// single CTA: offsets[i] = base + sum_{j<i} out_len[j]; offsets[n] = total
__global__ void __launch_bounds__(1024)
scan_offsets_kernel(const DeflateOut *__restrict__ outs, uint32_t n, uint64_t base, uint64_t *__restrict__ offsets)
{
	__shared__ uint64_t warp_tot[32];
	__shared__ uint64_t carry_s;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		carry_s = base;
	__syncthreads();
	for (uint32_t b = 0; b < n; b += 1024) {
		const uint32_t i = b + threadIdx.x;
		const uint64_t v = (i < n && outs[i].rc == 0) ? outs[i].out_len : 0;
		uint64_t incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			const uint64_t y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += y;
		}
		if (lane == 31)
			warp_tot[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint64_t w = warp_tot[lane], wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				const uint64_t y = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= (uint32_t)o)
					wi += y;
			}
			warp_tot[lane] = wi - w;
		}
		__syncthreads();
		const uint64_t carry = carry_s;
		if (i < n)
			offsets[i] = carry + warp_tot[warp] + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023)
			carry_s = carry + warp_tot[warp] + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		offsets[n] = carry_s;
}

// one CTA per chunk (grid-strided): copy the chunk's bytes to dst + offsets[i]
__global__ void __launch_bounds__(256)
gather_kernel(const DeflateJob *__restrict__ jobs, const DeflateOut *__restrict__ outs, const uint64_t *__restrict__ offsets,
	      uint32_t n, uint8_t *__restrict__ dst, uint64_t dst_cap)
{
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
		if (outs[i].rc != 0)
			continue;
		const uint32_t len = outs[i].out_len;
		const uint64_t off = offsets[i];
		if (off + len > dst_cap)
			continue;
		const uint8_t *src = jobs[i].out;            // 16-byte aligned slot
		uint8_t *d = dst + off;
		// head bytes until d is 4-byte aligned, then words assembled from the (aligned) source, then tail
		const uint32_t head = min(len, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(d) & 3)) & 3));
		if (threadIdx.x < head)
			d[threadIdx.x] = src[threadIdx.x];
		const uint32_t nwords = (len - head) >> 2;
		const uint32_t sh = (head & 3) * 8;
		const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
		uint32_t *d32 = reinterpret_cast<uint32_t *>(d + head);
		for (uint32_t w = threadIdx.x; w < nwords; w += blockDim.x) {
			// word w of the destination = source bytes [head + 4w, head + 4w + 4)
			const uint32_t lo = s32[w + (head >> 2)];
			const uint32_t hi = sh ? s32[w + (head >> 2) + 1] : 0;
			d32[w] = sh ? __funnelshift_r(lo, hi, sh) : lo;
		}
		const uint32_t done = head + nwords * 4;
		if (threadIdx.x < len - done)
			d[done + threadIdx.x] = src[done + threadIdx.x];
	}
}

// gzip / zlib trailer behind the last chunk, total length to *d_total
__global__ void finish_stream_kernel(const uint64_t *__restrict__ offsets, uint32_t n, uint8_t *__restrict__ dst, uint64_t dst_cap,
				     int wrap, const uint32_t *d_crc, const uint32_t *d_adler, uint64_t src_len, uint64_t *d_total)
{
	uint64_t p = offsets[n];
	if (wrap == NXGPU_WRAP_GZIP && p + 8 <= dst_cap) {
		const uint32_t c = *d_crc, l = (uint32_t)src_len;
		for (int k = 0; k < 4; k++) { dst[p + k] = (uint8_t)(c >> (8 * k)); dst[p + 4 + k] = (uint8_t)(l >> (8 * k)); }
		p += 8;
	} else if (wrap == NXGPU_WRAP_ZLIB && p + 4 <= dst_cap) {
		const uint32_t a = *d_adler;
		for (int k = 0; k < 4; k++) dst[p + k] = (uint8_t)(a >> (24 - 8 * k));
		p += 4;
	}
	*d_total = p;
}

// Joins the bit strings of the pieces of one deflate block (a large nxu_run_job descriptor is cut into
// pieces that separate CTAs compress with the same table, nxgpu_job.cu) into one: piece i holds `nbits`
// valid bits in a 16-byte aligned slot and starts at bit `dst_bit` of the group's output.  One CTA per
// group; every thread assembles whole output words (no atomics, coalesced stores).
__global__ void __launch_bounds__(1024)
bitconcat_kernel(const BitPiece *__restrict__ pieces, const BitGroup *__restrict__ groups)
{
	const BitGroup G = groups[blockIdx.x];
	const BitPiece *P = pieces + G.first_piece;
	uint32_t *out32 = reinterpret_cast<uint32_t *>(G.dst);
	const uint64_t nwords = (G.total_bits + 31) >> 5;
	for (uint64_t j = threadIdx.x; j < nwords; j += blockDim.x) {
		const uint64_t bit0 = j << 5;
		// the piece holding bit0: pieces are few (<= 64) and in order
		uint32_t lo = 0, hi = G.n_pieces;
		while (hi - lo > 1) {
			const uint32_t mid = (lo + hi) >> 1;
			if (P[mid].dst_bit <= bit0) lo = mid; else hi = mid;
		}
		uint32_t i = lo, word = 0, filled = 0;
		while (filled < 32 && i < G.n_pieces) {
			const uint64_t off = bit0 + filled - P[i].dst_bit;
			const uint64_t avail = P[i].nbits - off;
			if (avail == 0) { i++; continue; }
			const uint32_t take = (uint32_t)(avail < 32 - filled ? avail : 32 - filled);
			const uint32_t *s32 = reinterpret_cast<const uint32_t *>(P[i].src);
			const uint32_t w0 = s32[off >> 5], w1 = s32[(off >> 5) + 1];      // slots are padded: reading one word on is safe
			uint32_t v = __funnelshift_r(w0, w1, (uint32_t)(off & 31));
			if (take < 32)
				v &= (1u << take) - 1;
			word |= v << filled;
			filled += take;
			if (take == avail)
				i++;
		}
		out32[j] = word;
	}
}

} // namespace

cudaError_t launch_bitconcat(const BitPiece *pieces, const BitGroup *groups, uint32_t n_groups, cudaStream_t s)
{
	if (n_groups == 0)
		return cudaSuccess;
	bitconcat_kernel<<<n_groups, 1024, 0, s>>>(pieces, groups);
	return cudaGetLastError();
}

cudaError_t launch_scan_offsets(const DeflateOut *outs, uint32_t n, uint64_t base, uint64_t *offsets, cudaStream_t s)
{
	scan_offsets_kernel<<<1, 1024, 0, s>>>(outs, n, base, offsets);
	return cudaGetLastError();
}

cudaError_t launch_gather(const DeflateJob *jobs, const DeflateOut *outs, const uint64_t *offsets, uint32_t n,
			  uint8_t *dst, uint64_t dst_cap, cudaStream_t s)
{
	if (n == 0)
		return cudaSuccess;
	const uint32_t grid = n < (uint32_t)(kNumSMs * 8) ? n : (uint32_t)(kNumSMs * 8);
	gather_kernel<<<grid, 256, 0, s>>>(jobs, outs, offsets, n, dst, dst_cap);
	return cudaGetLastError();
}

cudaError_t launch_finish_stream(const DeflateOut *outs, const uint64_t *offsets, uint32_t n, uint8_t *dst, uint64_t dst_cap,
				 int wrap, const uint32_t *d_crc, const uint32_t *d_adler, uint64_t src_len, uint64_t *d_total, cudaStream_t s)
{
	(void)outs;
	finish_stream_kernel<<<1, 1, 0, s>>>(offsets, n, dst, dst_cap, wrap, d_crc, d_adler, src_len, d_total);
	return cudaGetLastError();
}

} // namespace nxgpu
