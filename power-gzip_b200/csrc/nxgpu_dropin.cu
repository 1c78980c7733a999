// nxgpu_dropin.cu — the six boundary symbols libnxz's host code links against
// (reference lib/nx_zlib.h:625-629, inc_nx/nxu.h:71, lib/crc32_ppc.c:30), served by the GPU.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <mutex>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {
int run_job_impl(nxgpu_ctx *ctx, uint8_t *crb_cpb);   // nxgpu_job.cu
}

namespace {
std::mutex g_mu;
nxgpu_ctx *g_ctx[16];

// one shared context per device, created on first use (handles are shared by up to
// 10 000 streams/threads, lib/nx_zlib.c:531-551, so the context is too)
nxgpu_ctx *ctx_for(int dev)
{
	if (dev < 0 || dev >= 16)
		dev = 0;
	if (!g_ctx[dev]) {
		nxgpu_ctx *c = nullptr;
		if (nxgpu_open(dev, &c) != 0)
			return nullptr;
		g_ctx[dev] = c;
	}
	return g_ctx[dev];
}
} // namespace

extern "C" {

uint64_t tb_freq = 0;

int nx_function_begin(int function, int pri, nx_devp_t h)
{
	if (function != 2 /* NX_FUNC_COMP_GZIP */ || !h) {
		errno = EINVAL;
		return -1;
	}
	std::lock_guard<std::mutex> lk(g_mu);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		errno = ENODEV;
		return -1;
	}
	const int dev = pri < 0 ? 0 : pri % ndev;
	nxgpu_ctx *c = ctx_for(dev);
	if (!c) {
		errno = ENODEV;
		return -1;
	}
	nxgpu_dev_prefix *d = reinterpret_cast<nxgpu_dev_prefix *>(h);
	d->paste_addr = c;            // non-NULL == usable
	d->fd = dev;
	d->function = function;
	return 0;
}

int nx_function_end(nx_devp_t h)
{
	if (!h)
		return -1;
	nxgpu_dev_prefix *d = reinterpret_cast<nxgpu_dev_prefix *>(h);
	d->paste_addr = nullptr;      // the shared context lives until process exit
	return 0;
}

uint64_t nx_wait_ticks(uint64_t ticks, uint64_t accumulated_ticks, int do_sleep)
{
	// 512 MHz timebase ticks -> nanoseconds
	struct timespec ts;
	const uint64_t ns = ticks * 1000ull / 512ull;
	(void)do_sleep;
	ts.tv_sec = (time_t)(ns / 1000000000ull);
	ts.tv_nsec = (long)(ns % 1000000000ull);
	nanosleep(&ts, nullptr);
	return accumulated_ticks + ticks;
}

int nxu_run_job(nx_gzip_crb_cpb_t *c, nx_devp_t h)
{
	if (!c || !h)
		return -EAGAIN;
	nxgpu_dev_prefix *d = reinterpret_cast<nxgpu_dev_prefix *>(h);
	if (!d->paste_addr)
		return -EAGAIN;
	std::lock_guard<std::mutex> lk(g_mu);   // one job at a time per process for now (coalescing)
	nxgpu_ctx *ctx = ctx_for(d->fd);
	if (!ctx)
		return -EAGAIN;
	return nxgpu::run_job_impl(ctx, reinterpret_cast<uint8_t *>(c));
}

unsigned int __crc32_vpmsum(unsigned int crc, const void *p, unsigned long len)
{
	// raw register update (no inversion): crc32(seed) = ~raw(~seed)  =>  raw(r) = ~crc32(~r)
	std::lock_guard<std::mutex> lk(g_mu);
	nxgpu_ctx *ctx = ctx_for(0);
	uint32_t out = 0;
	if (!ctx || nxgpu_crc32(ctx, ~crc, p, len, NXGPU_MEM_HOST, &out) != 0) {
		fprintf(stderr, "libnxgpu: __crc32_vpmsum: no usable GPU (%s)\n", nxgpu_last_error());
		abort();                  // no CPU fallback: fail loudly
	}
	return ~out;
}

} // extern "C"
