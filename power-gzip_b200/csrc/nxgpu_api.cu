// nxgpu_api.cu — host side of the batch extension declared in include/nxgpu.h: context,
// scratch management, job marshalling, launches.  No compute happens on the host: if the
// device or a launch fails every entry point returns a negative code and nxgpu_last_error()
// says why.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <string>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {

static thread_local char t_err[512];
void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_err, sizeof(t_err), fmt, ap);
	va_end(ap);
}

} // namespace nxgpu

#include "ctx.cuh"
using namespace nxgpu;

namespace {

int fam_index(const char *f)
{
	if (!strcmp(f, "deflate")) return 0;
	if (!strcmp(f, "inflate")) return 1;
	if (!strcmp(f, "checksum")) return 2;
	return -1;
}

} // namespace
void nxgpu::timer_begin(nxgpu_ctx *c, int fam)
{
	KernelTimer &t = c->timers[fam];
	if (!c->timing)
		return;
	if (t.used + 2 > t.ev.size()) {
		cudaEvent_t a, b;
		cudaEventCreate(&a); cudaEventCreate(&b);
		t.ev.push_back(a); t.ev.push_back(b);
	}
	cudaEventRecord(t.ev[t.used], c->stream);
}
void nxgpu::timer_end(nxgpu_ctx *c, int fam)
{
	KernelTimer &t = c->timers[fam];
	c->launches++;
	if (!c->timing)
		return;
	cudaEventRecord(t.ev[t.used + 1], c->stream);
	t.used += 2;
	t.launches++;
}
namespace {
void timer_collect(nxgpu_ctx *c)
{
	for (int f = 0; f < 3; f++) {
		KernelTimer &t = c->timers[f];
		for (size_t i = 0; i + 1 < t.used; i += 2) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]) == cudaSuccess)
				t.ms_total += ms;
		}
		t.used = 0;
	}
}

__global__ void scatter_outputs_kernel(const DeflateJob *__restrict__ jobs, const DeflateOut *__restrict__ outs,
				       uint8_t *const *__restrict__ dsts, const uint32_t *__restrict__ caps, uint32_t n)
{
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
		if (outs[i].rc != 0 || outs[i].out_len > caps[i])
			continue;
		const uint32_t len = outs[i].out_len;
		const uint8_t *s = jobs[i].out;
		uint8_t *d = dsts[i];
		for (uint32_t k = threadIdx.x; k < len; k += blockDim.x)
			d[k] = s[k];
	}
}

__global__ void write_bytes_kernel(uint8_t *dst, uint64_t v, int n)
{
	for (int k = 0; k < n; k++)
		dst[k] = (uint8_t)(v >> (8 * k));
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t kSliceChunks = 2 * kNumSMs;   // host-pointer streams are uploaded and compressed in slices of this many chunks (whole waves of the persistent grid)

} // namespace

extern "C" {

const char *nxgpu_last_error(void) { return t_err; }

int nxgpu_open(int dev, nxgpu_ctx **out)
{
	t_err[0] = 0;
	if (!out)
		return NXGPU_E_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
		return NXGPU_E_NODEV;
	}
	if (dev < 0) {
		const char *lr = getenv("LOCAL_RANK");
		dev = lr ? atoi(lr) % ndev : 0;
	}
	if (dev >= ndev) {
		set_error("device %d out of range (%d present)", dev, ndev);
		return NXGPU_E_NODEV;
	}
	NXGPU_CUDA_OK(cudaSetDevice(dev));
	cudaDeviceProp prop;
	NXGPU_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
	if (prop.major < 10) {
		set_error("device %d is sm_%d%d; this engine is built for sm_100a only", dev, prop.major, prop.minor);
		return NXGPU_E_NODEV;
	}
	nxgpu_ctx *c = new nxgpu_ctx();
	c->dev = dev;
	struct OpenGuard { nxgpu_ctx *c; ~OpenGuard() { if (c) nxgpu_close(c); } } open_guard{ c };    // no leak on the error returns below
	NXGPU_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	NXGPU_CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	NXGPU_CUDA_OK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
	NXGPU_CUDA_OK(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
	NXGPU_CUDA_OK(cudaEventCreate(&c->t0));
	NXGPU_CUDA_OK(cudaEventCreate(&c->t1));
	NXGPU_CUDA_OK(checksum_init_tables());
	open_guard.c = nullptr;
	*out = c;
	return 0;
}

void nxgpu_close(nxgpu_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->dev);
	cudaStreamSynchronize(c->stream);
	DevBuf *db[] = { &c->d_jobs, &c->d_outs, &c->d_tok, &c->d_slots, &c->d_ranges, &c->d_parts, &c->d_rs, &c->d_seeds,
			 &c->d_cks, &c->d_in, &c->d_out, &c->d_offsets, &c->d_misc, &c->d_dst_ptrs, &c->d_dht, &c->d_lz, &c->d_ctr, &c->d_flags, &c->d_ijobs, &c->d_iouts, &c->d_cat, &c->d_catdesc, &c->d_chain, &c->d_par1, &c->d_par2 };
	for (DevBuf *b : db) b->release();
	PinBuf *pb[] = { &c->h_jobs, &c->h_outs, &c->h_misc, &c->h_stage, &c->h_ones, &c->h_cat, &c->h_par };
	for (PinBuf *b : pb) b->release();
	for (ParSlot &ps : c->par) {
		ps.d1.release(); ps.d2.release(); ps.h.release();
		if (ps.ev) cudaEventDestroy(ps.ev);
		if (ps.st) cudaStreamDestroy(ps.st);
	}
	for (int f = 0; f < 3; f++)
		for (cudaEvent_t e : c->timers[f].ev) cudaEventDestroy(e);
	cudaEventDestroy(c->t0); cudaEventDestroy(c->t1);
	cudaEventDestroy(c->ev_main);
	cudaEventDestroy(c->ev_copy);
	cudaStreamDestroy(c->copy_stream);
	cudaStreamDestroy(c->stream);
	delete c;
}

int nxgpu_dev_alloc(nxgpu_ctx *c, size_t bytes, void **dptr)
{
	if (!c || !dptr) return NXGPU_E_ARG;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	NXGPU_CUDA_OK(cudaMalloc(dptr, bytes ? bytes : 1));
	return 0;
}
int nxgpu_dev_free(nxgpu_ctx *c, void *dptr)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	NXGPU_CUDA_OK(cudaFree(dptr));
	return 0;
}
int nxgpu_memcpy_h2d(nxgpu_ctx *c, void *dptr, const void *hptr, size_t bytes)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	NXGPU_CUDA_OK(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}
int nxgpu_memcpy_d2h(nxgpu_ctx *c, void *hptr, const void *dptr, size_t bytes)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	NXGPU_CUDA_OK(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}
int nxgpu_host_alloc(size_t bytes, void **hptr)
{
	if (!hptr) return NXGPU_E_ARG;
	NXGPU_CUDA_OK(cudaMallocHost(hptr, bytes ? bytes : 1));
	return 0;
}
int nxgpu_host_free(void *hptr)
{
	NXGPU_CUDA_OK(cudaFreeHost(hptr));
	return 0;
}
int nxgpu_sync(nxgpu_ctx *c)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}
int nxgpu_timer_start(nxgpu_ctx *c)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaEventRecord(c->t0, c->stream));
	return 0;
}
int nxgpu_timer_stop(nxgpu_ctx *c, float *ms)
{
	if (!c || !ms) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaEventRecord(c->t1, c->stream));
	NXGPU_CUDA_OK(cudaEventSynchronize(c->t1));
	NXGPU_CUDA_OK(cudaEventElapsedTime(ms, c->t0, c->t1));
	return 0;
}
uint64_t nxgpu_launch_count(nxgpu_ctx *c) { return c ? c->launches : 0; }
int nxgpu_kernel_time(nxgpu_ctx *c, const char *family, double *ms_total, uint64_t *launches)
{
	if (!c) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	const int f = family ? fam_index(family) : -1;
	if (!c || f < 0) return NXGPU_E_ARG;
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	timer_collect(c);
	if (ms_total) *ms_total = c->timers[f].ms_total;
	if (launches) *launches = c->timers[f].launches;
	return 0;
}
void nxgpu_kernel_time_reset(nxgpu_ctx *c)
{
	if (!c) return;
	cudaStreamSynchronize(c->stream);
	timer_collect(c);
	for (int f = 0; f < 3; f++) { c->timers[f].ms_total = 0; c->timers[f].launches = 0; }
}

uint32_t nxgpu_deflate_bound(uint32_t src_len)
{
	// stored blocks are the worst case: 5 bytes per 65535 plus the joiner
	return src_len + 5 * (src_len / 65535 + 1) + 16;
}
uint64_t nxgpu_deflate_stream_bound(uint64_t src_len, uint32_t chunk)
{
	if (chunk == 0) chunk = 262144;
	const uint64_t n = (src_len + chunk - 1) / chunk + 1;
	return src_len + n * (5 * (uint64_t)(chunk / 65535 + 1) + 16) + 32;
}

uint32_t nxgpu_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) { return host_crc32_combine(crc1, crc2, len2); }
uint32_t nxgpu_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
	const uint64_t B = 65521, rem = len2 % B;
	uint64_t s1 = a1 & 0xffff, s2 = (rem * s1) % B;
	s1 += (a2 & 0xffff) + B - 1;
	s2 += ((a1 >> 16) & 0xffff) + ((a2 >> 16) & 0xffff) + B - rem;
	return (uint32_t)((s1 % B) | ((s2 % B) << 16));
}

/* ------------------------------ checksums ------------------------------ */

// Device-resident inputs: items[i].src are device pointers.  Cuts big items into ranges so that
// the whole GPU works on one buffer; results land in d_cks (crc[n], adler[n]).
extern "C++" int nxgpu::checksum_device(nxgpu_ctx *c, const nxgpu_cksum_item *items, size_t n, int which)
{
	uint64_t total = 0;
	for (size_t i = 0; i < n; i++) total += items[i].len;
	// range size: at least 256 KiB, and about 8 ranges per SM for one large buffer
	uint64_t rsz = total / (uint64_t)(kNumSMs * 8);
	rsz = (rsz + 32767) / 32768 * 32768;
	if (rsz < 262144) rsz = 262144;
	size_t nr = 0;
	for (size_t i = 0; i < n; i++) nr += items[i].len ? (items[i].len + rsz - 1) / rsz : 1;
	const size_t rb = checksum_range_bytes(), pb = checksum_partial_bytes();
	int rc;
	const size_t hbytes = nr * rb + (n + 1) * 4 + n * 8;
	if ((rc = c->h_misc.reserve(hbytes))) return rc;
	if ((rc = c->d_ranges.reserve(hbytes))) return rc;
	if ((rc = c->d_parts.reserve(nr * pb))) return rc;
	if ((rc = c->d_cks.reserve(n * 8 + 16))) return rc;
	uint8_t *h = static_cast<uint8_t *>(c->h_misc.p);
	uint32_t *rs = reinterpret_cast<uint32_t *>(h + nr * rb);
	uint32_t *seeds = rs + (n + 1);
	size_t r = 0;
	for (size_t i = 0; i < n; i++) {
		rs[i] = (uint32_t)r;
		const uint64_t len = items[i].len;
		if (len == 0) { checksum_fill_range(h, r++, items[i].src, 0, 0, (uint32_t)i); }
		for (uint64_t o = 0; o < len; o += rsz) {
			const uint64_t l = len - o < rsz ? len - o : rsz;
			checksum_fill_range(h, r++, static_cast<const uint8_t *>(items[i].src) + o, l, len - o - l, (uint32_t)i);
		}
		seeds[i] = items[i].crc_seed;
		seeds[n + i] = items[i].adler_seed;
	}
	rs[n] = (uint32_t)r;
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_ranges.p, h, hbytes, cudaMemcpyHostToDevice, c->stream));
	uint8_t *d = static_cast<uint8_t *>(c->d_ranges.p);
	const uint32_t *d_rs = reinterpret_cast<const uint32_t *>(d + nr * rb);
	const uint32_t *d_seeds = d_rs + (n + 1);
	uint32_t *d_crc = static_cast<uint32_t *>(c->d_cks.p), *d_adler = d_crc + n;
	timer_begin(c, 2);
	NXGPU_CUDA_OK(launch_checksum_ranges(d, (uint32_t)nr, c->d_parts.p, which, c->stream));
	timer_end(c, 2);
	uint32_t max_rpj = 1;
	for (size_t i = 0; i < n; i++) max_rpj = std::max(max_rpj, rs[i + 1] - rs[i]);
	NXGPU_CUDA_OK(launch_checksum_combine(d, c->d_parts.p, d_rs, (uint32_t)n, d_seeds, d_seeds + n,
					      (which & 1) ? d_crc : nullptr, (which & 2) ? d_adler : nullptr, c->stream, max_rpj));
	c->launches++;
	return 0;
}

int nxgpu_checksum_batch(nxgpu_ctx *c, const nxgpu_cksum_item *items, size_t n, nxgpu_cksum_result *results, int mem)
{
	if (!c || (!items && n) || (!results && n)) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	if (n == 0) return 0;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	std::vector<nxgpu_cksum_item> dev_items;
	const nxgpu_cksum_item *use = items;
	if (mem == NXGPU_MEM_HOST) {
		uint64_t total = 0;
		for (size_t i = 0; i < n; i++) total += align_up(items[i].len, 16);
		if ((rc = c->d_in.reserve(total + 16))) return rc;
		dev_items.assign(items, items + n);
		uint64_t o = 0;
		for (size_t i = 0; i < n; i++) {
			uint8_t *dp = static_cast<uint8_t *>(c->d_in.p) + o;
			if (items[i].len && !items[i].src) return NXGPU_E_ARG;
			if (items[i].len)
				NXGPU_CUDA_OK(cudaMemcpyAsync(dp, items[i].src, items[i].len, cudaMemcpyHostToDevice, c->stream));
			dev_items[i].src = dp;
			o += align_up(items[i].len, 16);
		}
		use = dev_items.data();
	}
	if ((rc = checksum_device(c, use, n, 3))) return rc;
	if ((rc = c->h_outs.reserve(n * 8))) return rc;
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->h_outs.p, c->d_cks.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	const uint32_t *h = static_cast<const uint32_t *>(c->h_outs.p);
	for (size_t i = 0; i < n; i++) { results[i].crc32 = h[i]; results[i].adler32 = h[n + i]; }
	return 0;
}

int nxgpu_crc32(nxgpu_ctx *c, uint32_t seed, const void *src, uint64_t len, int mem, uint32_t *out)
{
	nxgpu_cksum_item it = { src, len, seed, 1 };
	nxgpu_cksum_result r;
	int rc = nxgpu_checksum_batch(c, &it, 1, &r, mem);
	if (!rc && out) *out = r.crc32;
	return rc;
}
int nxgpu_adler32(nxgpu_ctx *c, uint32_t seed, const void *src, uint64_t len, int mem, uint32_t *out)
{
	nxgpu_cksum_item it = { src, len, 0, seed };
	nxgpu_cksum_result r;
	int rc = nxgpu_checksum_batch(c, &it, 1, &r, mem);
	if (!rc && out) *out = r.adler32;
	return rc;
}

/* ------------------------------- deflate ------------------------------- */

// Core: all pointers in `jobs_h` are device pointers except `out`, which this routine assigns to
// private 16-byte aligned slots.  Leaves DeflateOut[n] in d_outs and per-item crc/adler in d_cks.
extern "C++" int nxgpu::deflate_device(nxgpu_ctx *c, DeflateJob *jobs_h, size_t n, int level, bool want_cksum, const StreamOut *so)
{
	int rc;
	if (level <= 0) level = 6;       // lib/nx_deflate.c:655-658 maps level 0 to 6 as well
	if (level > 9) level = 9;
	size_t slot_total = 0;
	uint32_t max_len = 0;
	// worst case per item: stored blocks; an nxu_run_job item (no joiner, caller's table, no stored
	// fallback) can come out at up to 15 bits per literal before the caller sees CC=64 and re-wraps it
	auto slot_cap = [](const DeflateJob &j) -> uint32_t {
		return (j.flags & NXGPU_F_NO_JOINER) ? 2 * j.src_len + 1024 : nxgpu_deflate_bound(j.src_len);
	};
	for (size_t i = 0; i < n; i++) {
		slot_total += align_up(slot_cap(jobs_h[i]) + 16, 128);      // whole 128-byte lines per slot: the kernel discards them from the L2 once copied
		if (jobs_h[i].src_len > max_len) max_len = jobs_h[i].src_len;
	}
	const int grid = (int)(n < (size_t)kNumSMs ? n : (size_t)kNumSMs);
	const uint32_t tok_stride = (uint32_t)align_up((size_t)max_len + 64, 512);
	if ((rc = c->d_slots.reserve(slot_total + 64))) return rc;
	if ((rc = c->d_tok.reserve((size_t)grid * deflate_scratch_words(tok_stride) * 4))) return rc;
	if ((rc = c->d_jobs.reserve(n * sizeof(DeflateJob)))) return rc;
	if ((rc = c->d_outs.reserve(n * sizeof(DeflateOut)))) return rc;
	size_t o = 0;
	for (size_t i = 0; i < n; i++) {
		jobs_h[i].out = static_cast<uint8_t *>(c->d_slots.p) + o;
		jobs_h[i].out_cap = slot_cap(jobs_h[i]);
		o += align_up(jobs_h[i].out_cap + 16, 128);
	}
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_jobs.p, jobs_h, n * sizeof(DeflateJob), cudaMemcpyHostToDevice, c->stream));
	// one persistent launch; jobs are claimed in order.  For host-pointer streams the kernel is launched
	// right away and each CTA waits for the ready flag of the slice its chunk lives in, so the uploads on
	// the copy stream overlap the compression with no launch boundaries in between.
	if ((rc = c->d_ctr.reserve(64))) return rc;
	timer_begin(c, 0);
	NXGPU_CUDA_OK(launch_deflate(static_cast<const DeflateJob *>(c->d_jobs.p), static_cast<DeflateOut *>(c->d_outs.p),
				     (uint32_t)n, level, static_cast<uint32_t *>(c->d_tok.p), tok_stride, grid, c->stream,
				     static_cast<uint32_t *>(c->d_ctr.p), c->ready_flags, c->jobs_per_flag, so));
	timer_end(c, 0);
	if (want_cksum) {
		std::vector<nxgpu_cksum_item> it(n);
		for (size_t i = 0; i < n; i++) { it[i].src = jobs_h[i].src; it[i].len = jobs_h[i].src_len; it[i].crc_seed = 0; it[i].adler_seed = 1; }
		if ((rc = checksum_device(c, it.data(), n, 3))) return rc;
	}
	return 0;
}

int nxgpu_deflate_batch(nxgpu_ctx *c, const nxgpu_deflate_item *items, size_t n, nxgpu_deflate_result *results,
			int level, int mem)
{
	if (!c || (!items && n) || (!results && n)) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	if (n == 0) return 0;
	// per-item sizes are 32-bit; the worst-case slot (2 x source + 1 KiB) must stay below 4 GiB too
	for (size_t i = 0; i < n; i++)
		if (items[i].src_len > 0x7fff0000u) { set_error("item %zu: src_len %u exceeds the 2 GiB item limit", i, items[i].src_len); return NXGPU_E_ARG; }
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	if ((rc = c->h_jobs.reserve(n * sizeof(DeflateJob)))) return rc;
	DeflateJob *jh = static_cast<DeflateJob *>(c->h_jobs.p);
	memset(jh, 0, n * sizeof(DeflateJob));
	if (mem == NXGPU_MEM_HOST) {
		// upload [src - hist_len, src + src_len) of every item; contiguous items share one copy
		uint64_t total = 0;
		for (size_t i = 0; i < n; i++) total += align_up((uint64_t)items[i].hist_len + items[i].src_len, 16) + 16;
		if ((rc = c->d_in.reserve(total))) return rc;
		uint64_t o = 0;
		const uint8_t *span_lo = nullptr, *span_hi = nullptr;   // host span already uploaded
		uint8_t *span_dev = nullptr;
		for (size_t i = 0; i < n; i++) {
			const uint8_t *lo = static_cast<const uint8_t *>(items[i].src) - items[i].hist_len;
			const uint8_t *hi = static_cast<const uint8_t *>(items[i].src) + items[i].src_len;
			if (items[i].hist_len > 32768) return NXGPU_E_ARG;
			if (span_lo && lo >= span_lo && lo <= span_hi && hi >= span_hi) {
				// continues the previous span: upload only the new tail
				const size_t add = hi - span_hi;
				if (add)
					NXGPU_CUDA_OK(cudaMemcpyAsync(span_dev + (span_hi - span_lo), span_hi, add, cudaMemcpyHostToDevice, c->stream));
				span_hi = hi;
				o = (span_dev - static_cast<uint8_t *>(c->d_in.p)) + (span_hi - span_lo);
			} else {
				o = align_up(o, 16);
				span_dev = static_cast<uint8_t *>(c->d_in.p) + o;
				span_lo = lo; span_hi = hi;
				if (hi > lo)
					NXGPU_CUDA_OK(cudaMemcpyAsync(span_dev, lo, hi - lo, cudaMemcpyHostToDevice, c->stream));
				o += hi - lo;
			}
			jh[i].src = span_dev + (static_cast<const uint8_t *>(items[i].src) - span_lo);
			jh[i].src_len = items[i].src_len; jh[i].hist_len = items[i].hist_len; jh[i].flags = items[i].flags;
		}
	} else {
		for (size_t i = 0; i < n; i++) {
			if (items[i].hist_len > 32768) return NXGPU_E_ARG;
			jh[i].src = static_cast<const uint8_t *>(items[i].src);
			jh[i].src_len = items[i].src_len; jh[i].hist_len = items[i].hist_len; jh[i].flags = items[i].flags;
		}
	}
	if ((rc = deflate_device(c, jh, n, level, true))) return rc;
	// results
	if ((rc = c->h_outs.reserve(n * sizeof(DeflateOut) + n * 8))) return rc;
	DeflateOut *oh = static_cast<DeflateOut *>(c->h_outs.p);
	uint32_t *ck = reinterpret_cast<uint32_t *>(oh + n);
	NXGPU_CUDA_OK(cudaMemcpyAsync(oh, c->d_outs.p, n * sizeof(DeflateOut), cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaMemcpyAsync(ck, c->d_cks.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
	if (mem == NXGPU_MEM_DEVICE) {
		// device copy of slot -> caller's dst
		if ((rc = c->d_dst_ptrs.reserve(n * 12))) return rc;
		if ((rc = c->h_stage.reserve(n * 12))) return rc;
		uint8_t **dp = static_cast<uint8_t **>(c->h_stage.p);
		uint32_t *caps = reinterpret_cast<uint32_t *>(dp + n);
		for (size_t i = 0; i < n; i++) { dp[i] = static_cast<uint8_t *>(items[i].dst); caps[i] = items[i].dst_cap; }
		NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_dst_ptrs.p, dp, n * 12, cudaMemcpyHostToDevice, c->stream));
		const uint32_t grid = (uint32_t)(n < (size_t)(kNumSMs * 8) ? n : (size_t)(kNumSMs * 8));
		scatter_outputs_kernel<<<grid, 256, 0, c->stream>>>(static_cast<const DeflateJob *>(c->d_jobs.p),
			static_cast<const DeflateOut *>(c->d_outs.p), static_cast<uint8_t *const *>(c->d_dst_ptrs.p),
			reinterpret_cast<const uint32_t *>(static_cast<uint8_t **>(c->d_dst_ptrs.p) + n), (uint32_t)n);
		NXGPU_CUDA_OK(cudaGetLastError());
		c->launches++;
	}
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (size_t i = 0; i < n; i++) {
		results[i].rc = oh[i].rc;
		results[i].out_len = oh[i].out_len;
		results[i].tebc = oh[i].tebc;
		results[i].crc32 = ck[i];
		results[i].adler32 = ck[n + i];
		results[i].n_tokens = oh[i].n_tokens;
		if (oh[i].rc == 0 && oh[i].out_len > items[i].dst_cap)
			results[i].rc = NXGPU_E_BUF;
	}
	if (mem == NXGPU_MEM_HOST) {
		for (size_t i = 0; i < n; i++)
			if (results[i].rc == 0 && results[i].out_len)
				NXGPU_CUDA_OK(cudaMemcpyAsync(items[i].dst, jh[i].out, results[i].out_len, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	}
	return 0;
}

// Everything of nxgpu_deflate_stream up to (not including) the first host synchronisation: the kernels and copies are
// enqueued on c->stream and the totals stay on the device (e->d_off[n] = end of the last chunk, e->d_off[n+1] = stream
// length with trailer, e->d_cks[0..1] = crc32 / adler32 of the input).  nxgpu_team.cu chains the cross-GPU exchange
// behind this without going through the host.
extern "C++" int nxgpu::deflate_stream_enqueue(nxgpu_ctx *c, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
					       int level, int wrap, uint32_t chunk, int mem, StreamEnq *e)
{
	if (!c || (!src && src_len) || !dst || !e) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	// whatever path leaves this function, the next launch on the context must not wait on this call's upload flags
	struct FlagGuard { nxgpu_ctx *c; ~FlagGuard() { c->ready_flags = nullptr; } } flag_guard{ c };
	// true Z_FULL_FLUSH semantics: no chunk looks back into the previous one, so the chunks of the
	// index can be inflated in parallel (nxgpu_inflate_stream)
	const bool independent = (wrap & NXGPU_STREAM_INDEPENDENT) != 0;
	wrap &= ~NXGPU_STREAM_INDEPENDENT;
	const bool cont = wrap == NXGPU_WRAP_RAW_CONT;
	if (cont) wrap = NXGPU_WRAP_RAW;
	if (wrap != NXGPU_WRAP_RAW && wrap != NXGPU_WRAP_ZLIB && wrap != NXGPU_WRAP_GZIP) return NXGPU_E_ARG;
	if (chunk == 0) chunk = 262144;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	const size_t n = src_len ? (size_t)((src_len + chunk - 1) / chunk) : 1;
	const uint8_t *dsrc = static_cast<const uint8_t *>(src);
	uint8_t *ddst = static_cast<uint8_t *>(dst);
	if (mem == NXGPU_MEM_HOST) {
		if ((rc = c->d_in.reserve(src_len + 16))) return rc;
		if ((rc = c->d_out.reserve(dst_cap + 16))) return rc;
		const size_t per_slice = kSliceChunks;
		const size_t n_slices = (n + per_slice - 1) / per_slice;
		if ((rc = c->d_flags.reserve((n_slices + 1) * 4))) return rc;
		if ((rc = c->h_ones.reserve(16))) return rc;
		*static_cast<uint32_t *>(c->h_ones.p) = 1;
		NXGPU_CUDA_OK(cudaMemsetAsync(c->d_flags.p, 0, (n_slices + 1) * 4, c->stream));
		// the copy stream must not overtake earlier work on d_in or the flag reset
		NXGPU_CUDA_OK(cudaEventRecord(c->ev_main, c->stream));
		NXGPU_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
		for (size_t lo = 0, k = 0; lo < n && src_len; lo += per_slice, k++) {
			const size_t cnt = n - lo < per_slice ? n - lo : per_slice;
			const uint64_t b0 = (uint64_t)lo * chunk;
			const uint64_t b1 = (lo + cnt == n) ? src_len : (uint64_t)(lo + cnt) * chunk;
			NXGPU_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t *>(c->d_in.p) + b0, static_cast<const uint8_t *>(src) + b0, b1 - b0,
						      cudaMemcpyHostToDevice, c->copy_stream));
			NXGPU_CUDA_OK(cudaMemcpyAsync(static_cast<uint32_t *>(c->d_flags.p) + k, c->h_ones.p, 4, cudaMemcpyHostToDevice, c->copy_stream));
		}
		NXGPU_CUDA_OK(cudaEventRecord(c->ev_copy, c->copy_stream));
		if (src_len) {
			c->ready_flags = static_cast<const uint32_t *>(c->d_flags.p);
			c->jobs_per_flag = (uint32_t)per_slice;
		}
		dsrc = static_cast<const uint8_t *>(c->d_in.p);
		ddst = static_cast<uint8_t *>(c->d_out.p);
	}
	if ((rc = c->h_jobs.reserve(n * sizeof(DeflateJob)))) return rc;
	DeflateJob *jh = static_cast<DeflateJob *>(c->h_jobs.p);
	memset(jh, 0, n * sizeof(DeflateJob));
	for (size_t i = 0; i < n; i++) {
		const uint64_t o = (uint64_t)i * chunk;
		jh[i].src = dsrc + o;
		jh[i].src_len = (uint32_t)(src_len - o < chunk ? src_len - o : chunk);
		jh[i].hist_len = independent ? 0 : (uint32_t)(o < 32768 ? o : 32768);
		jh[i].flags = (i + 1 == n && !cont) ? NXGPU_F_FINAL : 0;
	}
	// header
	uint64_t hdr = 0; int hdr_len = 0;
	if (wrap == NXGPU_WRAP_GZIP) {
		// 1f 8b 08 00 mtime=0 xfl=0 os=3, the blank header of lib/nx_deflate.c:473-489
		hdr = 0x1full | 0x8bull << 8 | 0x08ull << 16; hdr_len = 8;
	} else if (wrap == NXGPU_WRAP_ZLIB) {
		const int lv = level <= 0 ? 6 : level;
		const uint32_t flevel = lv < 2 ? 0 : lv < 6 ? 1 : lv == 6 ? 2 : 3;
		uint32_t h = (0x78u << 8) | (flevel << 6);
		h += 31 - (h % 31);
		hdr = (h >> 8) | ((h & 0xff) << 8); hdr_len = 2;
	}
	const uint64_t hdr_total = wrap == NXGPU_WRAP_GZIP ? 10 : hdr_len;
	if (dst_cap < hdr_total) { c->ready_flags = nullptr; return NXGPU_E_BUF; }
	// the stitch is fused into the kernel: every chunk's CTA copies its bytes to their final place as soon as its
	// predecessor has published where that is.  When the caller's host buffer is pinned (cudaHostAlloc /
	// cudaHostRegister) the kernel writes straight into it, so the device-to-host transfer of the compressed stream
	// overlaps the compression of later chunks.
	static const bool fused = !(getenv("NXGPU_FUSED_STITCH") && atoi(getenv("NXGPU_FUSED_STITCH")) == 0);   // developer switch
	bool zero_copy = false;
	if (mem == NXGPU_MEM_HOST && fused) {
		cudaPointerAttributes pa;
		if (cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
			ddst = static_cast<uint8_t *>(pa.devicePointer);
			zero_copy = true;
		} else {
			cudaGetLastError();
		}
	}
	if ((rc = c->d_offsets.reserve((n + 2) * 8))) { c->ready_flags = nullptr; return rc; }
	if ((rc = c->d_chain.reserve((n + 1) * 8))) { c->ready_flags = nullptr; return rc; }
	uint64_t *d_off = static_cast<uint64_t *>(c->d_offsets.p);
	StreamOut so;
	if (fused) {
		NXGPU_CUDA_OK(cudaMemsetAsync(c->d_chain.p, 0, (n + 1) * 8, c->stream));
		so.dst = ddst; so.cap = dst_cap; so.base = hdr_total; so.offsets = d_off;
		so.chain = static_cast<unsigned long long *>(c->d_chain.p);
	}
	rc = deflate_device(c, jh, n, level, false, fused ? &so : nullptr);
	if (c->ready_flags) {
		// everything behind the kernel (checksums, a later call's uploads) must see the whole input
		c->ready_flags = nullptr;
		cudaStreamWaitEvent(c->stream, c->ev_copy, 0);
	}
	if (rc) return rc;
	// whole-stream checksums: chunks are the ranges of one job
	nxgpu_cksum_item whole = { dsrc, src_len, 0, 1 };
	if ((rc = checksum_device(c, &whole, 1, 3))) return rc;
	if (hdr_len) {
		write_bytes_kernel<<<1, 1, 0, c->stream>>>(ddst, hdr, hdr_len);
		if (wrap == NXGPU_WRAP_GZIP)
			write_bytes_kernel<<<1, 1, 0, c->stream>>>(ddst + 8, 0x0300ull, 2);
		c->launches += 1;
	}
	const DeflateJob *dj = static_cast<const DeflateJob *>(c->d_jobs.p);
	const DeflateOut *dout = static_cast<const DeflateOut *>(c->d_outs.p);
	if (!fused) {
		NXGPU_CUDA_OK(launch_scan_offsets(dout, (uint32_t)n, hdr_total, d_off, c->stream));
		NXGPU_CUDA_OK(launch_gather(dj, dout, d_off, (uint32_t)n, ddst, dst_cap, c->stream));
	}
	const uint32_t *d_crc = static_cast<const uint32_t *>(c->d_cks.p);
	NXGPU_CUDA_OK(launch_finish_stream(dout, d_off, (uint32_t)n, ddst, dst_cap, wrap, d_crc, d_crc + 1, src_len, d_off + n + 1, c->stream));
	c->launches += fused ? 1 : 3;
	e->n = n; e->d_off = d_off; e->d_cks = static_cast<uint32_t *>(c->d_cks.p); e->ddst = ddst; e->zero_copy = zero_copy; e->wrap = wrap;
	return 0;
}

// The tail of nxgpu_deflate_stream: one synchronisation, per-chunk status, totals and (for pageable host targets) the copy back.
extern "C++" int nxgpu::deflate_stream_collect(nxgpu_ctx *c, const StreamEnq &e, void *dst, uint64_t dst_cap, int mem,
					       uint64_t *chunk_offsets, nxgpu_stream_result *res)
{
	NXGPU_LOCK(c);
	int rc;
	const size_t n = e.n;
	uint64_t *d_off = e.d_off;
	uint8_t *ddst = e.ddst;
	const bool zero_copy = e.zero_copy;
	const int wrap = e.wrap;
	// results back
	if ((rc = c->h_outs.reserve(n * sizeof(DeflateOut) + (n + 2) * 8 + 16))) return rc;
	DeflateOut *oh = static_cast<DeflateOut *>(c->h_outs.p);
	uint64_t *offh = reinterpret_cast<uint64_t *>(oh + n);
	uint32_t *ckh = reinterpret_cast<uint32_t *>(offh + n + 2);
	NXGPU_CUDA_OK(cudaMemcpyAsync(oh, c->d_outs.p, n * sizeof(DeflateOut), cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaMemcpyAsync(offh, d_off, (n + 2) * 8, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaMemcpyAsync(ckh, c->d_cks.p, 8, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	uint64_t ntok = 0;
	for (size_t i = 0; i < n; i++) {
		if (oh[i].rc) { set_error("chunk %zu failed rc=%d", i, oh[i].rc); return oh[i].rc < 0 ? oh[i].rc : NXGPU_E_DATA; }
		ntok += oh[i].n_tokens;
	}
	const uint64_t total = offh[n + 1];
	const uint64_t need = offh[n] + (wrap == NXGPU_WRAP_GZIP ? 8 : wrap == NXGPU_WRAP_ZLIB ? 4 : 0);
	if (need > dst_cap) { set_error("output needs %llu bytes, capacity %llu", (unsigned long long)need, (unsigned long long)dst_cap); return NXGPU_E_BUF; }
	if (mem == NXGPU_MEM_HOST && !zero_copy) {
		NXGPU_CUDA_OK(cudaMemcpyAsync(dst, ddst, total, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	}
	if (chunk_offsets)
		for (size_t i = 0; i <= n; i++) chunk_offsets[i] = offh[i];
	res->out_len = total;
	res->crc32 = ckh[0];
	res->adler32 = ckh[1];
	res->n_chunks = (uint32_t)n;
	res->n_tokens = ntok;
	return 0;
}

int nxgpu_deflate_stream(nxgpu_ctx *c, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			 int level, int wrap, uint32_t chunk, uint64_t *chunk_offsets, nxgpu_stream_result *res, int mem)
{
	if (!c || (!src && src_len) || !dst || !res) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	StreamEnq e;
	int rc = deflate_stream_enqueue(c, src, src_len, dst, dst_cap, level, wrap, chunk, mem, &e);
	if (rc) return rc;
	return deflate_stream_collect(c, e, dst, dst_cap, mem, chunk_offsets, res);
}

/* --------------------------- DHT generation --------------------------- */

// n histograms of 316 counters (286 lit/len then 30 distances, host byte order) -> n dynamic headers of
// up to 288 bytes (bits from HLIT on, LSB first, exactly the bytes of cpb.in_dht) + their bit lengths.
int nxgpu_dhtgen_batch(nxgpu_ctx *c, const uint32_t *counts, size_t n, uint8_t *dht, uint32_t *dht_bits, int mem)
{
	if (!c || !counts || !dht || !dht_bits) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	if (n == 0) return 0;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	const uint32_t *d_counts = counts;
	uint8_t *d_dht = dht;
	uint32_t *d_bits = dht_bits;
	if (mem == NXGPU_MEM_HOST) {
		if ((rc = c->d_lz.reserve(n * 316 * 4))) return rc;
		if ((rc = c->d_dht.reserve(n * 288 + n * 4 + 16))) return rc;
		NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_lz.p, counts, n * 316 * 4, cudaMemcpyHostToDevice, c->stream));
		d_counts = static_cast<const uint32_t *>(c->d_lz.p);
		d_dht = static_cast<uint8_t *>(c->d_dht.p);
		d_bits = reinterpret_cast<uint32_t *>(d_dht + align_up(n * 288, 16));
	}
	NXGPU_CUDA_OK(launch_dhtgen(d_counts, (uint32_t)n, d_dht, d_bits, c->stream));
	c->launches++;
	if (mem == NXGPU_MEM_HOST) {
		NXGPU_CUDA_OK(cudaMemcpyAsync(dht, d_dht, n * 288, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaMemcpyAsync(dht_bits, d_bits, n * 4, cudaMemcpyDeviceToHost, c->stream));
	}
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

// The signature of the reference's dhtgen() (lib/nx_dhtgen.c:945-954) plus the context.
int nxgpu_dhtgen(nxgpu_ctx *c, const uint32_t *lhist, int num_lhist, const uint32_t *dhist, int num_dhist,
		 char *dht, int *dht_num_bytes, int *dht_num_valid_bits, int cpb_header)
{
	if (!c || !lhist || !dhist || !dht || !dht_num_bytes || !dht_num_valid_bits) return NXGPU_E_ARG;
	if (num_lhist < 257 || num_lhist > 286 || num_dhist < 0 || num_dhist > 30) return NXGPU_E_ARG;
	uint32_t counts[316];
	memset(counts, 0, sizeof(counts));
	memcpy(counts, lhist, (size_t)num_lhist * 4);
	memcpy(counts + 286, dhist, (size_t)num_dhist * 4);
	uint8_t out[288];
	uint32_t bits = 0;
	const int rc = nxgpu_dhtgen_batch(c, counts, 1, out, &bits, NXGPU_MEM_HOST);
	if (rc) return rc;
	const int nbytes = (int)((bits + 7) / 8);
	char *body = dht;
	if (cpb_header) {
		// the 16 bytes in front of in_dht: in_dhtlen lives in the low 12 bits of the fourth word (inc_nx/nxu.h:296-310)
		memset(dht, 0, 16);
		dht[14] = (char)((bits >> 8) & 0x0f);
		dht[15] = (char)(bits & 0xff);
		body = dht + 16;
	}
	memcpy(body, out, (size_t)nbytes);
	*dht_num_bytes = nbytes;
	*dht_num_valid_bits = (int)(bits % 8);      // 0 is encoded as 8 bits
	return 0;
}

/* ------------------------------- inflate ------------------------------- */

} // extern "C"
namespace nxgpu {
// A descriptor takes the parallel path when it is long enough to hold several deflate blocks (zlib closes a block every
// 16 Ki symbols, 20-60 KiB of compressed text) and the launch does not fill the GPU anyway.  The path costs two block
// decodes plus ~0.4 ms; one warp needs ~2 ms per block (measured: a 109 KB source 7.7 ms, a 424 KB one 5.2 ms in parallel).
void inflate_par_select(InflateJob *jobs, size_t n, std::vector<std::pair<size_t, InflateJob>> &picked, bool dry_too)
{
	picked.clear();
	const char *e = getenv("NXGPU_INFLATE_PAR_MIN");          // bytes of source: every stream this long goes, 0 = never (developer / test switch)
	const uint64_t par_min = e ? strtoull(e, nullptr, 0) : 64 * 1024;
	if (par_min == 0 || n > 32)
		return;
	std::vector<size_t> elig;
	for (size_t i = 0; i < n; i++) {
		const InflateJob &j = jobs[i];
		if (j.src_len < par_min || j.single_block || (j.wrap & (kWrapSkip | kWrapNoHeader)) || ((j.wrap & kWrapDry) && !dry_too) || j.stop_map || j.hist_ptr)
			continue;
		elig.push_back(i);
	}
	if (elig.empty())
		return;
	// The streams of one launch run side by side, a warp pair each, so the launch takes as long as its longest stream
	// (20-45 MB of source per second).  The many-warp decodes run beside the launch on streams of their own, up to
	// kParSlots at a time; each costs the host ~0.5 ms (block search, one synchronisation, sort) and ~3 ms + source / 2 GB/s
	// on the device.  The k longest streams are taken out of the launch while max(launch, decodes) shrinks: a lone long
	// stream always goes, sixteen threads with a few hundred KB each stay where they are.
	std::sort(elig.begin(), elig.end(), [&](size_t a, size_t b) { return jobs[a].src_len > jobs[b].src_len; });
	uint32_t longest_other = 0;                                // the longest stream that is not eligible at all
	for (size_t i = 0; i < n; i++)
		if (std::find(elig.begin(), elig.end(), i) == elig.end() && !(jobs[i].wrap & kWrapSkip) && jobs[i].src_len > longest_other)
			longest_other = jobs[i].src_len;
	auto serial_ms = [](uint32_t src) { return src / 30e3; };
	double best = 1e30, dev_ms = 0;
	size_t best_k = 0;
	if (e)
		best_k = elig.size();                                  // the switch is set: every eligible stream goes (tests)
	for (size_t k = 0; k <= elig.size() && !e; k++) {
		const uint32_t rest = std::max(longest_other, k < elig.size() ? jobs[elig[k]].src_len : 0u);
		const double par = k ? 3.0 + 0.5 * k + dev_ms / (k < (size_t)kParSlots ? k : kParSlots) : 0.0;
		const double t = std::max(serial_ms(rest), par);
		if (t < best) { best = t; best_k = k; }
		if (k < elig.size())
			dev_ms += jobs[elig[k]].src_len / 2e6;
	}
	for (size_t k = 0; k < best_k; k++) {
		picked.emplace_back(elig[k], jobs[elig[k]]);
		jobs[elig[k]].wrap |= kWrapSkip;
	}
	std::sort(picked.begin(), picked.end(), [](const std::pair<size_t, InflateJob> &x, const std::pair<size_t, InflateJob> &y) { return x.first < y.first; });
}

// *serial = true: the stream has nothing to split at, nothing was launched, the caller runs it on one warp
static int inflate_parallel_one(nxgpu_ctx *c, ParSlot &S, const InflateJob &job, InflateOut *d_final, bool *serial)
{
	*serial = false;
	int rc;
	const uint32_t n = job.src_len;
	const size_t map_bytes = align_up((size_t)n + 64, 256);
	const uint32_t surv_cap = n / 32 + 1024;                  // the filter passes one bit offset in ~1200
	const uint32_t cand_cap = surv_cap < 65536 ? surv_cap : 65536;
	const size_t p1 = map_bytes + (size_t)surv_cap * 8 + (size_t)cand_cap * 8 + 256 + 2 * sizeof(InflateJob);
	if ((rc = S.d1.reserve(p1))) return rc;
	if ((rc = S.h.reserve((size_t)cand_cap * 8 + 512 + sizeof(InflateJob)))) return rc;
	uint8_t *b1 = static_cast<uint8_t *>(S.d1.p);
	uint32_t *map = reinterpret_cast<uint32_t *>(b1);
	uint64_t *surv = reinterpret_cast<uint64_t *>(b1 + map_bytes);
	uint64_t *cand = surv + surv_cap;
	uint32_t *counts = reinterpret_cast<uint32_t *>(cand + cand_cap);
	InflateJob *d_job = reinterpret_cast<InflateJob *>(reinterpret_cast<uint8_t *>(counts) + 128);
	uint8_t *hp = static_cast<uint8_t *>(S.h.p);
	uint32_t *h_counts = reinterpret_cast<uint32_t *>(hp);
	InflateJob *h_job = reinterpret_cast<InflateJob *>(hp + 128);
	uint64_t *h_cand = reinterpret_cast<uint64_t *>(hp + 256 + sizeof(InflateJob) - sizeof(InflateJob) % 8 + 8);
	*h_job = job;
	NXGPU_CUDA_OK(cudaMemcpyAsync(d_job, h_job, sizeof(InflateJob), cudaMemcpyHostToDevice, S.st));
	const bool is_job = (job.wrap & 0xff) == kWrapJob;
	NXGPU_CUDA_OK(launch_blockfind(job.src, n, is_job ? job.start_bit : 0, map, surv, surv_cap, cand, cand_cap, counts, S.st));
	NXGPU_CUDA_OK(cudaMemcpyAsync(h_counts, counts, 8, cudaMemcpyDeviceToHost, S.st));
	const uint32_t first = cand_cap < 1024 ? cand_cap : 1024;     // the list usually fits one small copy: one synchronisation
	NXGPU_CUDA_OK(cudaMemcpyAsync(h_cand, cand, (size_t)first * 8, cudaMemcpyDeviceToHost, S.st));
	NXGPU_CUDA_OK(cudaStreamSynchronize(S.st));
	c->launches += 2;
	const uint32_t n_cand = h_counts[1];
	static const bool trace = getenv("NXGPU_TRACE") != nullptr;
	if (trace)
		fprintf(stderr, "nxgpu inflate: %u source bytes, %u of %u bit offsets pass the header filter, %u block-start candidates\n", n, h_counts[0], n * 8, n_cand);
	if (h_counts[0] > surv_cap || n_cand > cand_cap || n_cand == 0) {
		// nothing to split at (one huge block, stored data) or more look-alikes than the lists hold: one warp
		*serial = true;
		return 0;
	}
	if (n_cand > first) {
		NXGPU_CUDA_OK(cudaMemcpyAsync(h_cand + first, cand + first, (size_t)(n_cand - first) * 8, cudaMemcpyDeviceToHost, S.st));
		NXGPU_CUDA_OK(cudaStreamSynchronize(S.st));
	}
	std::sort(h_cand, h_cand + n_cand);
	NXGPU_CUDA_OK(cudaMemcpyAsync(cand, h_cand, (size_t)n_cand * 8, cudaMemcpyHostToDevice, S.st));
	const size_t nc = n_cand;
	const size_t p2 = nc * 65536 + nc * 32768 + align_up(nc * sizeof(SpecOut), 256) + align_up(nc * sizeof(InflateJob), 256) +
			  align_up(nc * sizeof(InflateOut), 256) + align_up(nc * sizeof(ChainMeta), 256) + 1024;
	if (S.d2.reserve(p2)) {
		// no room for the rings of this many pieces: one warp
		cudaGetLastError();
		*serial = true;
		return 0;
	}
	uint8_t *b2 = static_cast<uint8_t *>(S.d2.p);
	ParPlan P;
	memset(&P, 0, sizeof(P));
	P.job = job;
	P.map = map;
	P.cands = cand;
	P.n_cand = n_cand;
	P.rings = reinterpret_cast<uint16_t *>(b2); b2 += nc * 65536;
	P.hists = b2; b2 += nc * 32768;
	P.spec = reinterpret_cast<SpecOut *>(b2); b2 += align_up(nc * sizeof(SpecOut), 256);
	P.cjobs = reinterpret_cast<InflateJob *>(b2); b2 += align_up(nc * sizeof(InflateJob), 256);
	P.couts = reinterpret_cast<InflateOut *>(b2); b2 += align_up(nc * sizeof(InflateOut), 256);
	P.meta = reinterpret_cast<ChainMeta *>(b2); b2 += align_up(nc * sizeof(ChainMeta), 256);
	P.n_chain = reinterpret_cast<uint32_t *>(b2);
	P.head_out = reinterpret_cast<InflateOut *>(b2 + 64);
	P.final_out = d_final;
	P.retry_job = d_job + 1;
	NXGPU_CUDA_OK(launch_inflate_par(P, counts + 16, S.st));
	c->launches += 6;
	return 0;
}

// the descriptors inflate_par_select() picked, on c->stream behind the launch they were taken out of (d_outs: that launch's
// results).  Streams that turn out to have nothing to split at run together in one more launch, a warp each.
int inflate_parallel(nxgpu_ctx *c, const std::vector<std::pair<size_t, InflateJob>> &picked, InflateOut *d_outs, bool inputs_marked)
{
	std::vector<InflateJob> serial;
	std::vector<size_t> serial_at;
	// every decode has its own stream and scratch (up to kParSlots in flight): they start behind the uploads of the inputs
	// (c->ev_main, recorded by the caller in front of its own launch when inputs_marked: the launch and these decodes then run
	// side by side), and the context's stream goes on behind all of them
	timer_begin(c, 1);
	if (!inputs_marked)
		NXGPU_CUDA_OK(cudaEventRecord(c->ev_main, c->stream));
	const size_t used = picked.size() < (size_t)kParSlots ? picked.size() : (size_t)kParSlots;
	for (size_t k = 0; k < used; k++) {
		ParSlot &S = c->par[k];
		if (!S.st) {
			NXGPU_CUDA_OK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
			NXGPU_CUDA_OK(cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming));
		}
		NXGPU_CUDA_OK(cudaStreamWaitEvent(S.st, c->ev_main, 0));
	}
	int rc = 0;
	for (size_t i = 0; i < picked.size() && !rc; i++) {
		ParSlot &S = c->par[i % kParSlots];
		if (i >= (size_t)kParSlots)
			NXGPU_CUDA_OK(cudaStreamSynchronize(S.st));          // its scratch is about to be reused
		bool s = false;
		rc = inflate_parallel_one(c, S, picked[i].second, d_outs + picked[i].first, &s);
		if (s) { serial.push_back(picked[i].second); serial_at.push_back(picked[i].first); }
	}
	for (size_t k = 0; k < used; k++) {
		NXGPU_CUDA_OK(cudaEventRecord(c->par[k].ev, c->par[k].st));
		NXGPU_CUDA_OK(cudaStreamWaitEvent(c->stream, c->par[k].ev, 0));
	}
	timer_end(c, 1);
	if (rc)
		return rc;
	if (serial.empty())
		return 0;
	// a job array as long as the original launch's, everything skipped but these (results land in their own slots)
	const size_t n = serial_at.back() + 1;
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));          // h_par / d_par1 are about to be reused
	if ((rc = c->h_par.reserve(n * sizeof(InflateJob)))) return rc;
	if ((rc = c->d_par1.reserve(n * sizeof(InflateJob)))) return rc;
	InflateJob *h = static_cast<InflateJob *>(c->h_par.p);
	memset(h, 0, n * sizeof(InflateJob));
	for (size_t i = 0; i < n; i++) h[i].wrap = kWrapSkip;
	for (size_t k = 0; k < serial.size(); k++) h[serial_at[k]] = serial[k];
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_par1.p, h, n * sizeof(InflateJob), cudaMemcpyHostToDevice, c->stream));
	timer_begin(c, 1);
	NXGPU_CUDA_OK(launch_inflate(static_cast<const InflateJob *>(c->d_par1.p), d_outs, (uint32_t)n, static_cast<uint32_t *>(c->d_misc.p), c->stream));
	timer_end(c, 1);
	c->launches++;
	return 0;
}
} // namespace nxgpu
extern "C" {

int nxgpu_inflate_batch(nxgpu_ctx *c, const nxgpu_inflate_item *items, size_t n, nxgpu_inflate_result *results, int mem)
{
	if (!c || (!items && n) || (!results && n)) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	if (n == 0) return 0;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	if ((rc = c->h_jobs.reserve(n * sizeof(InflateJob)))) return rc;
	InflateJob *jh = static_cast<InflateJob *>(c->h_jobs.p);
	memset(jh, 0, n * sizeof(InflateJob));
	bool dst_contig = true;
	if (mem == NXGPU_MEM_HOST) {
		uint64_t in_total = 0, out_total = 0;
		for (size_t i = 0; i < n; i++) {
			in_total += align_up(items[i].src_len, 16);
			out_total += align_up((uint64_t)items[i].hist_len + items[i].dst_cap, 16);
			if (i && static_cast<uint8_t *>(items[i].dst) != static_cast<uint8_t *>(items[i - 1].dst) + items[i - 1].dst_cap)
				dst_contig = false;
			if (items[i].hist_len) dst_contig = false;
		}
		// sources that already sit back to back in pinned host memory (a file read into a pinned buffer) go up as they
		// are; anything else is packed through pinned staging first: one H2D instead of n small ones either way
		bool src_direct = true;
		const uint8_t *s0 = static_cast<const uint8_t *>(items[0].src);
		for (size_t i = 1; i < n && src_direct; i++)
			src_direct = static_cast<const uint8_t *>(items[i].src) >= static_cast<const uint8_t *>(items[i - 1].src) + items[i - 1].src_len;
		const uint64_t src_span = src_direct ? (static_cast<const uint8_t *>(items[n - 1].src) - s0) + items[n - 1].src_len : 0;
		if (src_direct) {
			cudaPointerAttributes pa;
			src_direct = src_span <= 2 * in_total + 4096 && cudaPointerGetAttributes(&pa, s0) == cudaSuccess && pa.type == cudaMemoryTypeHost;
			cudaGetLastError();
		}
		if ((rc = c->d_in.reserve((src_direct ? src_span : in_total) + 16))) return rc;
		if ((rc = c->d_out.reserve(out_total + 16))) return rc;
		if (!src_direct && (rc = c->h_stage.reserve(in_total + 16))) return rc;
		uint64_t io = 0, oo = 0;
		uint8_t *hs = static_cast<uint8_t *>(c->h_stage.p);
		for (size_t i = 0; i < n; i++) {
			if (src_direct) {
				jh[i].src = static_cast<uint8_t *>(c->d_in.p) + (static_cast<const uint8_t *>(items[i].src) - s0);
			} else {
				memcpy(hs + io, items[i].src, items[i].src_len);
				jh[i].src = static_cast<uint8_t *>(c->d_in.p) + io;
			}
			jh[i].src_len = items[i].src_len; jh[i].wrap = items[i].wrap;
			jh[i].hist_len = items[i].hist_len; jh[i].dst_cap = items[i].dst_cap;
			if (dst_contig) {
				jh[i].dst = static_cast<uint8_t *>(c->d_out.p) + (static_cast<uint8_t *>(items[i].dst) - static_cast<uint8_t *>(items[0].dst));
			} else {
				jh[i].dst = static_cast<uint8_t *>(c->d_out.p) + oo + items[i].hist_len;
				if (items[i].hist_len)
					NXGPU_CUDA_OK(cudaMemcpyAsync(jh[i].dst - items[i].hist_len, static_cast<uint8_t *>(items[i].dst) - items[i].hist_len,
								      items[i].hist_len, cudaMemcpyHostToDevice, c->stream));
			}
			io += align_up(items[i].src_len, 16);
			oo += align_up((uint64_t)items[i].hist_len + items[i].dst_cap, 16);
		}
		if (src_direct)
			NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_in.p, s0, src_span, cudaMemcpyHostToDevice, c->stream));
		else
			NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_in.p, hs, in_total, cudaMemcpyHostToDevice, c->stream));
	} else {
		for (size_t i = 0; i < n; i++) {
			jh[i].src = static_cast<const uint8_t *>(items[i].src); jh[i].src_len = items[i].src_len; jh[i].wrap = items[i].wrap;
			jh[i].dst = static_cast<uint8_t *>(items[i].dst); jh[i].dst_cap = items[i].dst_cap; jh[i].hist_len = items[i].hist_len;
		}
	}
	if ((rc = c->d_jobs.reserve(n * sizeof(InflateJob)))) return rc;
	if ((rc = c->d_outs.reserve(n * sizeof(InflateOut)))) return rc;
	if ((rc = c->d_misc.reserve(64))) return rc;
	const size_t rb = checksum_range_bytes(), pb = checksum_partial_bytes();
	// A large host-memory batch with one contiguous target runs as up to four groups of members: the device-to-host
	// copy of a group's output goes to the copy stream as soon as the group is inflated and overlaps the next group
	// (a group is at least one full wave of warps, so the kernel loses nothing)
	const size_t wave = 4 * 7 * (size_t)kNumSMs;
	const size_t n_groups = (mem == NXGPU_MEM_HOST && dst_contig && n >= 2 * wave) ? std::min<size_t>(4, n / wave) : 1;
	// a few long outputs: every SM takes a share of each checksum
	uint32_t K = 1;
	if (n <= 32)
		for (size_t i = 0; i < n; i++)
			if (items[i].dst_cap >= (4u << 20)) K = kNumSMs;
	if ((rc = c->d_ranges.reserve(n * K * rb + (n + n_groups + 1) * 4))) return rc;
	if ((rc = c->d_parts.reserve(n * K * pb))) return rc;
	if ((rc = c->d_cks.reserve(n * 8 + 16))) return rc;
	std::vector<std::pair<size_t, InflateJob>> par;
	inflate_par_select(jh, n, par);
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_jobs.p, jh, n * sizeof(InflateJob), cudaMemcpyHostToDevice, c->stream));
	const InflateJob *dj = static_cast<const InflateJob *>(c->d_jobs.p);
	InflateOut *dout = static_cast<InflateOut *>(c->d_outs.p);
	uint8_t *d_rng = static_cast<uint8_t *>(c->d_ranges.p);
	uint8_t *d_prt = static_cast<uint8_t *>(c->d_parts.p);
	uint32_t *d_rs_all = reinterpret_cast<uint32_t *>(d_rng + n * K * rb);
	uint32_t *d_crc = static_cast<uint32_t *>(c->d_cks.p), *d_adler = d_crc + n;
	for (size_t g = 0; g < n_groups; g++) {
		const size_t g0 = g * n / n_groups, g1 = (g + 1) * n / n_groups, ng = g1 - g0;
		if (!par.empty())
			NXGPU_CUDA_OK(cudaEventRecord(c->ev_main, c->stream));            // the inputs are up: the many-warp decodes may start
		timer_begin(c, 1);
		NXGPU_CUDA_OK(launch_inflate(dj + g0, dout + g0, (uint32_t)ng, static_cast<uint32_t *>(c->d_misc.p), c->stream));
		timer_end(c, 1);
		if (!par.empty() && (rc = inflate_parallel(c, par, dout, true))) return rc;      // (n <= 32: one group)
		// crc32 / adler32 of every output, lengths taken from the device results
		uint32_t *d_rs = d_rs_all + g0 + g;
		NXGPU_CUDA_OK(launch_ranges_from_inflate(dj + g0, dout + g0, (uint32_t)ng, d_rng + g0 * K * rb, d_rs, c->stream, K));
		timer_begin(c, 2);
		NXGPU_CUDA_OK(launch_checksum_ranges(d_rng + g0 * K * rb, (uint32_t)(ng * K), d_prt + g0 * K * pb, 3, c->stream));
		timer_end(c, 2);
		NXGPU_CUDA_OK(launch_checksum_combine(d_rng + g0 * K * rb, d_prt + g0 * K * pb, d_rs, (uint32_t)ng, nullptr, nullptr, d_crc + g0, d_adler + g0, c->stream, K));
		c->launches += 2;
		if (mem == NXGPU_MEM_HOST && dst_contig) {
			uint8_t *h0 = static_cast<uint8_t *>(items[g0].dst);
			const uint64_t off = h0 - static_cast<uint8_t *>(items[0].dst);
			const uint64_t span = (static_cast<uint8_t *>(items[g1 - 1].dst) - h0) + items[g1 - 1].dst_cap;
			cudaStream_t cs = n_groups > 1 ? c->copy_stream : c->stream;
			if (n_groups > 1) {
				NXGPU_CUDA_OK(cudaEventRecord(c->ev_main, c->stream));
				NXGPU_CUDA_OK(cudaStreamWaitEvent(cs, c->ev_main, 0));
			}
			NXGPU_CUDA_OK(cudaMemcpyAsync(h0, static_cast<uint8_t *>(c->d_out.p) + off, span, cudaMemcpyDeviceToHost, cs));
		}
	}
	if (n_groups > 1) {
		NXGPU_CUDA_OK(cudaEventRecord(c->ev_copy, c->copy_stream));
		NXGPU_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
	}
	if ((rc = c->h_outs.reserve(n * sizeof(InflateOut) + n * 8))) return rc;
	InflateOut *oh = static_cast<InflateOut *>(c->h_outs.p);
	uint32_t *ck = reinterpret_cast<uint32_t *>(oh + n);
	NXGPU_CUDA_OK(cudaMemcpyAsync(oh, dout, n * sizeof(InflateOut), cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaMemcpyAsync(ck, d_crc, n * 8, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (size_t i = 0; i < n; i++) {
		nxgpu_inflate_result &r = results[i];
		r.rc = oh[i].rc; r.out_len = oh[i].out_len; r.in_used = oh[i].in_used;
		r.crc32 = ck[i]; r.adler32 = ck[n + i];
		r.flags = oh[i].flags & 5;
		const uint32_t wrap = oh[i].flags >> 8;
		if (r.rc == 0 && wrap == NXGPU_WRAP_GZIP) {
			// the check lib/nx_inflate.c:763-848 does on the host
			if (oh[i].trailer_crc != r.crc32 || oh[i].trailer_isize != r.out_len) r.rc = NXGPU_E_DATA; else r.flags |= 2;
		} else if (r.rc == 0 && wrap == NXGPU_WRAP_ZLIB) {
			if (oh[i].trailer_crc != r.adler32) r.rc = NXGPU_E_DATA; else r.flags |= 2;
		}
	}
	if (mem == NXGPU_MEM_HOST && !dst_contig) {
		for (size_t i = 0; i < n; i++)
			if (results[i].out_len)
				NXGPU_CUDA_OK(cudaMemcpyAsync(items[i].dst, jh[i].dst, results[i].out_len, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	}
	return 0;
}

/* One big member, inflated as its sync-point segments in ONE batch (the index
 * nxgpu_deflate_stream returns, consumed by inflate).  Legal when the segments do not look back
 * into each other (NXGPU_STREAM_INDEPENDENT); a primed stream is detected (a distance reaches in
 * front of a segment) and refused with NXGPU_E_DATA — inflate it as one member instead. */
int nxgpu_inflate_stream(nxgpu_ctx *c, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap, int wrap,
			 const uint64_t *chunk_offsets, uint32_t n_chunks, uint32_t chunk, nxgpu_stream_result *res, int mem)
{
	if (!c || !src || (!dst && dst_cap) || !res || !chunk_offsets || n_chunks == 0) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	if (wrap != NXGPU_WRAP_RAW && wrap != NXGPU_WRAP_ZLIB && wrap != NXGPU_WRAP_GZIP) return NXGPU_E_ARG;
	if (chunk == 0) chunk = 262144;
	const uint64_t trailer = wrap == NXGPU_WRAP_GZIP ? 8 : wrap == NXGPU_WRAP_ZLIB ? 4 : 0;
	for (uint32_t i = 0; i < n_chunks; i++)
		if (chunk_offsets[i] > chunk_offsets[i + 1] || chunk_offsets[i + 1] - chunk_offsets[i] > 0xffffff00ull) return NXGPU_E_ARG;
	if (chunk_offsets[n_chunks] + trailer > src_len) { set_error("index runs past the end of the stream"); return NXGPU_E_DATA; }
	if ((uint64_t)(n_chunks - 1) * chunk > dst_cap) return NXGPU_E_BUF;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	const uint8_t *dsrc = static_cast<const uint8_t *>(src);
	uint8_t *ddst = static_cast<uint8_t *>(dst);
	if (mem == NXGPU_MEM_HOST) {
		if ((rc = c->d_in.reserve(src_len + 16))) return rc;
		if ((rc = c->d_out.reserve(dst_cap + 16))) return rc;
		NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_in.p, src, src_len, cudaMemcpyHostToDevice, c->stream));
		dsrc = static_cast<const uint8_t *>(c->d_in.p);
		ddst = static_cast<uint8_t *>(c->d_out.p);
	}
	std::vector<nxgpu_inflate_item> items(n_chunks);
	std::vector<nxgpu_inflate_result> rs(n_chunks);
	for (uint32_t i = 0; i < n_chunks; i++) {
		const uint64_t o = (uint64_t)i * chunk;
		items[i].src = dsrc + chunk_offsets[i];
		items[i].src_len = (uint32_t)(chunk_offsets[i + 1] - chunk_offsets[i]);
		items[i].dst = ddst + o;
		items[i].dst_cap = (uint32_t)(dst_cap - o < chunk ? dst_cap - o : chunk);
		items[i].wrap = NXGPU_WRAP_RAW;
		items[i].hist_len = 0;
	}
	if ((rc = nxgpu_inflate_batch(c, items.data(), n_chunks, rs.data(), NXGPU_MEM_DEVICE))) return rc;
	// every segment but the last fills its chunk exactly and ends on the joiner; the last one carries BFINAL
	uint64_t total = 0;
	uint32_t crc = 0, adler = 1;
	for (uint32_t i = 0; i < n_chunks; i++) {
		const bool last = i + 1 == n_chunks;
		// an inner segment ends on the joiner without a final block: reported as E_DATA + "clean end" (bit 2)
		const bool ok = last ? (rs[i].rc == 0 && (rs[i].flags & 1))
				     : (rs[i].rc == NXGPU_E_DATA && (rs[i].flags & 5) == 4 && rs[i].out_len == chunk);
		if (!ok) {
			set_error("segment %u: rc %d, %u bytes, final %u (a stream whose chunks were primed cannot be split)", i, rs[i].rc, rs[i].out_len, rs[i].flags & 1);
			return rs[i].rc == NXGPU_E_BUF ? NXGPU_E_BUF : NXGPU_E_DATA;
		}
		total += rs[i].out_len;
	}
	// checksums of the whole output in one pass (the segments are the ranges of one job, combined on the device)
	{
		nxgpu_cksum_item whole = { ddst, total, 0, 1 };
		uint32_t ck[2];
		if ((rc = checksum_device(c, &whole, 1, 3))) return rc;
		NXGPU_CUDA_OK(cudaMemcpyAsync(ck, c->d_cks.p, 8, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
		crc = ck[0]; adler = ck[1];
	}
	// trailer (the check lib/nx_inflate.c:763-848 does on the host)
	if (trailer) {
		uint8_t t[8];
		if (mem == NXGPU_MEM_HOST) memcpy(t, static_cast<const uint8_t *>(src) + chunk_offsets[n_chunks], trailer);
		else NXGPU_CUDA_OK(cudaMemcpy(t, dsrc + chunk_offsets[n_chunks], trailer, cudaMemcpyDeviceToHost));
		if (wrap == NXGPU_WRAP_GZIP) {
			const uint32_t tc = t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
			const uint32_t ts = t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
			if (tc != crc || ts != (uint32_t)total) { set_error("gzip trailer mismatch"); return NXGPU_E_DATA; }
		} else {
			const uint32_t ta = (uint32_t)t[0] << 24 | (uint32_t)t[1] << 16 | (uint32_t)t[2] << 8 | t[3];
			if (ta != adler) { set_error("zlib trailer mismatch"); return NXGPU_E_DATA; }
		}
	}
	if (mem == NXGPU_MEM_HOST && total) {
		NXGPU_CUDA_OK(cudaMemcpyAsync(dst, ddst, total, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	}
	res->out_len = total;
	res->crc32 = crc;
	res->adler32 = adler;
	res->n_chunks = n_chunks;
	res->n_tokens = 0;
	return 0;
}

/* A file of concatenated gzip members (`cat a.gz b.gz`, bgzip, the output of pigz -i ...) inflated as ONE batch —
 * the member-indexed reader of SURVEY.md §8f rank 2 (the reference's nx_gzread, lib/nx_gzlib.c:220-263, stops after
 * the first member and pulls 10 bytes per read()).  Members carry no index, so they are discovered on the device:
 * (1) every offset that looks like a member header becomes a candidate, (2) a dry run decodes all candidates at
 * once — no output, only "where does it end, how many bytes does it make" —, (3) the true members are the chain from
 * offset 0 (each must start where the previous one ended; ISIZE must match), (4) their output offsets are the prefix
 * sums and nxgpu_inflate_batch inflates and CRC-checks them all. */
int nxgpu_gunzip_concat(nxgpu_ctx *c, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			uint64_t *out_len, uint32_t *n_members, int mem)
{
	if (!c || !src || (!dst && dst_cap) || !out_len) return NXGPU_E_ARG;
	NXGPU_LOCK(c);
	*out_len = 0;
	if (n_members) *n_members = 0;
	if (src_len < 18) { set_error("shorter than one gzip member"); return NXGPU_E_DATA; }
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	int rc;
	const uint8_t *dsrc = static_cast<const uint8_t *>(src);
	uint8_t *ddst = static_cast<uint8_t *>(dst);
	if (mem == NXGPU_MEM_HOST) {
		if ((rc = c->d_in.reserve(src_len + 64))) return rc;
		NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_in.p, src, src_len, cudaMemcpyHostToDevice, c->stream));
		dsrc = static_cast<const uint8_t *>(c->d_in.p);
	}
	// (1) candidates
	const uint32_t max_cand = 1u << 22;
	if ((rc = c->d_cat.reserve((size_t)max_cand * 8))) return rc;
	if ((rc = c->d_misc.reserve(64))) return rc;
	uint64_t *d_cand = static_cast<uint64_t *>(c->d_cat.p);
	uint32_t *d_count = static_cast<uint32_t *>(c->d_misc.p) + 8;
	NXGPU_CUDA_OK(launch_gzip_candidates(dsrc, src_len, d_cand, max_cand, d_count, c->stream));
	c->launches++;
	uint32_t n_cand = 0;
	NXGPU_CUDA_OK(cudaMemcpyAsync(&n_cand, d_count, 4, cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (n_cand == 0) { set_error("no gzip member header found"); return NXGPU_E_DATA; }
	if (n_cand > max_cand) { set_error("more than %u candidate member headers", max_cand); return NXGPU_E_MEM; }
	std::vector<uint64_t> cand(n_cand);
	NXGPU_CUDA_OK(cudaMemcpy(cand.data(), d_cand, (size_t)n_cand * 8, cudaMemcpyDeviceToHost));
	std::sort(cand.begin(), cand.end());
	// (2) dry run of every candidate
	if ((rc = c->h_jobs.reserve((size_t)n_cand * sizeof(InflateJob)))) return rc;
	if ((rc = c->d_ijobs.reserve((size_t)n_cand * sizeof(InflateJob)))) return rc;
	if ((rc = c->d_iouts.reserve((size_t)n_cand * sizeof(InflateOut)))) return rc;
	InflateJob *jh = static_cast<InflateJob *>(c->h_jobs.p);
	memset(jh, 0, (size_t)n_cand * sizeof(InflateJob));
	for (uint32_t i = 0; i < n_cand; i++) {
		const uint64_t rest = src_len - cand[i];
		jh[i].src = dsrc + cand[i];
		jh[i].src_len = (uint32_t)(rest < 0xfffffff0ull ? rest : 0xfffffff0ull);
		jh[i].wrap = NXGPU_WRAP_GZIP | kWrapDry;
		jh[i].dst = nullptr;
		jh[i].dst_cap = 0xffffffffu;
	}
	// a few candidates in a long buffer (a .gz file with one big member): each is counted by many warps
	std::vector<std::pair<size_t, InflateJob>> par;
	inflate_par_select(jh, n_cand, par, true);
	NXGPU_CUDA_OK(cudaMemcpyAsync(c->d_ijobs.p, jh, (size_t)n_cand * sizeof(InflateJob), cudaMemcpyHostToDevice, c->stream));
	if (!par.empty())
		NXGPU_CUDA_OK(cudaEventRecord(c->ev_main, c->stream));
	timer_begin(c, 1);
	NXGPU_CUDA_OK(launch_inflate(static_cast<const InflateJob *>(c->d_ijobs.p), static_cast<InflateOut *>(c->d_iouts.p), n_cand,
				     static_cast<uint32_t *>(c->d_misc.p), c->stream));
	timer_end(c, 1);
	c->launches++;
	if (!par.empty() && (rc = inflate_parallel(c, par, static_cast<InflateOut *>(c->d_iouts.p), true))) return rc;
	std::vector<InflateOut> dry(n_cand);
	NXGPU_CUDA_OK(cudaMemcpyAsync(dry.data(), c->d_iouts.p, (size_t)n_cand * sizeof(InflateOut), cudaMemcpyDeviceToHost, c->stream));
	NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	// (3) the chain of true members from offset 0
	std::vector<nxgpu_inflate_item> items;
	uint64_t p = 0, total = 0;
	while (p < src_len) {
		const auto it = std::lower_bound(cand.begin(), cand.end(), p);
		if (it == cand.end() || *it != p) {
			bool zeros = mem == NXGPU_MEM_HOST;                 // gzip tolerates zero padding behind the last member
			for (uint64_t q = p; zeros && q < src_len; q++) zeros = static_cast<const uint8_t *>(src)[q] == 0;
			if (zeros && !items.empty()) break;
			set_error("no gzip member starts at offset %llu", (unsigned long long)p);
			return NXGPU_E_DATA;
		}
		const InflateOut &o = dry[it - cand.begin()];
		if (o.rc != 0 || o.in_used == 0 || o.trailer_isize != o.out_len) {
			set_error("member at offset %llu does not decode (rc %d)", (unsigned long long)p, o.rc);
			return NXGPU_E_DATA;
		}
		nxgpu_inflate_item m;
		m.src = dsrc + p; m.src_len = o.in_used;
		m.dst = reinterpret_cast<void *>(total);                 // offset for now, pointer once the buffer is known
		m.dst_cap = o.out_len; m.wrap = NXGPU_WRAP_GZIP; m.hist_len = 0;
		items.push_back(m);
		total += o.out_len;
		p += o.in_used;
	}
	*out_len = total;
	if (n_members) *n_members = (uint32_t)items.size();
	if (total > dst_cap) { set_error("output needs %llu bytes, capacity %llu", (unsigned long long)total, (unsigned long long)dst_cap); return NXGPU_E_BUF; }
	// (4) the real thing
	if (mem == NXGPU_MEM_HOST) {
		if ((rc = c->d_out.reserve(total + 64))) return rc;
		ddst = static_cast<uint8_t *>(c->d_out.p);
	}
	for (nxgpu_inflate_item &m : items)
		m.dst = ddst + reinterpret_cast<uintptr_t>(m.dst);
	std::vector<nxgpu_inflate_result> rs(items.size());
	if ((rc = nxgpu_inflate_batch(c, items.data(), items.size(), rs.data(), NXGPU_MEM_DEVICE))) return rc;
	for (size_t i = 0; i < rs.size(); i++)
		if (rs[i].rc != 0 || !(rs[i].flags & 2)) {
			set_error("member %zu: rc %d, trailer %s", i, rs[i].rc, (rs[i].flags & 2) ? "ok" : "mismatch");
			return NXGPU_E_DATA;
		}
	if (mem == NXGPU_MEM_HOST && total) {
		NXGPU_CUDA_OK(cudaMemcpyAsync(dst, ddst, total, cudaMemcpyDeviceToHost, c->stream));
		NXGPU_CUDA_OK(cudaStreamSynchronize(c->stream));
	}
	return 0;
}

/* ------------------------------ makedata ------------------------------- */

// Same draw order as reference samples/makedata.c:35-70 (srand48/lrand48).  A draw of dist == 0
// copies a byte onto itself there, i.e. keeps the fresh (zero) malloc page, hence the memset.
uint64_t nxgpu_makedata(int seed, int log2size, const void *seedfile, uint64_t seedfile_len, void *out_, uint64_t out_cap)
{
	uint8_t *out = static_cast<uint8_t *>(out_);
	uint64_t bufsz = 1ull << log2size;
	srand48(seed);
	const long a = lrand48() % 2;
	const long b = lrand48() % (long)(bufsz / 10);
	bufsz += (uint64_t)a * (uint64_t)b;
	if (bufsz > out_cap)
		return 0;
	memset(out, 0, bufsz);
	uint64_t idx = seedfile_len < bufsz / 2 ? seedfile_len : bufsz / 2;
	memcpy(out, seedfile, idx);
	const uint64_t len_max = (uint64_t)(lrand48() % 240) + 10;
	const uint64_t dist_max = (uint64_t)(lrand48() % (1L << 16)) + 1;
	while (idx < bufsz) {
		uint64_t dist = (uint64_t)lrand48() % (idx > dist_max ? dist_max : idx);
		uint64_t len = (uint64_t)lrand48() % len_max + 16;
		if (dist > idx)
			dist = idx;
		while (len-- > 0 && idx < bufsz) {
			out[idx] = out[idx - dist];
			idx++;
		}
	}
	return idx;
}

// The bytes [from, to) of the same stream without holding the stream: the generator only ever looks back
// 65536 bytes (dist_max, samples/makedata.c:47), so a 128 KiB ring plus the seed file reproduce any range.
// Used to cut a 16 GiB stream (BASELINE.json configs[3]) into per-GPU shards, each rank generating its own.
uint64_t nxgpu_makedata_range(int seed, int log2size, const void *seedfile, uint64_t seedfile_len, uint64_t from, uint64_t to, void *out_)
{
	uint8_t *out = static_cast<uint8_t *>(out_);
	uint64_t bufsz = 1ull << log2size;
	srand48(seed);
	const long a = lrand48() % 2;
	const long b = lrand48() % (long)(bufsz / 10);
	bufsz += (uint64_t)a * (uint64_t)b;
	if (to > bufsz)
		to = bufsz;
	if (from >= to)
		return 0;
	constexpr uint64_t kRing = 1ull << 17;
	std::vector<uint8_t> ring(kRing, 0);
	const uint8_t *sf = static_cast<const uint8_t *>(seedfile);
	uint64_t idx = seedfile_len < bufsz / 2 ? seedfile_len : bufsz / 2;
	for (uint64_t i = idx > kRing ? idx - kRing : 0; i < idx; i++)
		ring[i & (kRing - 1)] = sf[i];
	if (from < idx)
		memcpy(out, sf + from, (idx < to ? idx : to) - from);
	const uint64_t len_max = (uint64_t)(lrand48() % 240) + 10;
	const uint64_t dist_max = (uint64_t)(lrand48() % (1L << 16)) + 1;
	while (idx < to) {
		uint64_t dist = (uint64_t)lrand48() % (idx > dist_max ? dist_max : idx);
		uint64_t len = (uint64_t)lrand48() % len_max + 16;
		if (dist > idx)
			dist = idx;
		const uint64_t n = len < bufsz - idx ? len : bufsz - idx, i0 = idx;
		uint8_t *r = ring.data();
		if (dist == 0) {
			// dist 0 copies the byte onto itself: the zero the buffer was cleared to
			for (uint64_t k = 0; k < n; k++)
				r[(i0 + k) & (kRing - 1)] = 0;
		} else {
			for (uint64_t k = 0; k < n; k++)
				r[(i0 + k) & (kRing - 1)] = r[(i0 + k - dist) & (kRing - 1)];
		}
		idx += n;
		const uint64_t lo = i0 > from ? i0 : from, hi = idx < to ? idx : to;
		for (uint64_t i = lo; i < hi; i++)
			out[i - from] = r[i & (kRing - 1)];
	}
	return to - from;
}

} // extern "C"
