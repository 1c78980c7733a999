// ctx.cuh — internal definitions shared by the host-side translation units (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <utility>
#include <vector>
#include "common.cuh"
#include "../../include/nxgpu.h"

namespace nxgpu {

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes)
	{
		if (bytes <= cap)
			return 0;
		if (p)
			cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 4096;
		if (cudaMalloc(&p, want) != cudaSuccess) {
			cudaGetLastError();
			want = bytes;
			if (cudaMalloc(&p, want) != cudaSuccess) {
				set_error("cudaMalloc(%zu) failed", want);
				p = nullptr;
				return NXGPU_E_MEM;
			}
		}
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
	void *p = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes)
	{
		if (bytes <= cap)
			return 0;
		if (p)
			cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 4096;
		if (cudaMallocHost(&p, want) != cudaSuccess) {
			set_error("cudaMallocHost(%zu) failed", want);
			cudaGetLastError();
			return NXGPU_E_MEM;
		}
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// one many-warp inflate in flight (inflate_par.cuh): its own stream and scratch, so that several of them overlap
struct ParSlot {
	cudaStream_t st = nullptr;
	cudaEvent_t ev = nullptr;
	DevBuf d1, d2;
	PinBuf h;
};
constexpr int kParSlots = 8;

struct KernelTimer {
	std::vector<cudaEvent_t> ev;      // start/stop pairs
	size_t used = 0;
	double ms_total = 0;
	uint64_t launches = 0;
};

} // namespace nxgpu

using namespace nxgpu;
#define NXGPU_LOCK(c) std::lock_guard<std::recursive_mutex> nxgpu_lock_((c)->mu)

struct nxgpu_ctx {
	// a context owns one stream and one set of staging buffers: the batch calls serialise on this
	// (recursive: nxgpu_inflate_stream -> nxgpu_inflate_batch); contexts are independent of each other
	std::recursive_mutex mu;
	int dev = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr;      // uploads of host-pointer streams, overlapped with compute
	cudaEvent_t ev_main = nullptr, ev_copy = nullptr;
	const uint32_t *ready_flags = nullptr;   // set while a host-pointer stream is being uploaded slice by slice
	uint32_t jobs_per_flag = 0;
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	uint64_t launches = 0;
	DevBuf d_jobs, d_outs, d_tok, d_slots, d_ranges, d_parts, d_rs, d_seeds, d_cks, d_in, d_out, d_offsets, d_misc, d_dst_ptrs, d_dht, d_lz, d_ctr, d_flags, d_ijobs, d_iouts, d_cat, d_catdesc, d_chain, d_par1, d_par2;
	PinBuf h_jobs, h_outs, h_misc, h_stage, h_ones, h_cat, h_par;
	ParSlot par[kParSlots];
	KernelTimer timers[3];           // 0 deflate, 1 inflate, 2 checksum
	bool timing = true;
};


namespace nxgpu {
// nxgpu_api.cu: device-resident cores of the batch calls (pointers are device pointers)
int deflate_device(nxgpu_ctx *c, DeflateJob *jobs_h, size_t n, int level, bool want_cksum, const StreamOut *so = nullptr);
int checksum_device(nxgpu_ctx *c, const nxgpu_cksum_item *items, size_t n, int which);
// nxgpu_deflate_stream in two halves: enqueue (no host synchronisation, totals stay on the device) and collect
struct StreamEnq { size_t n = 0; uint64_t *d_off = nullptr; uint32_t *d_cks = nullptr; uint8_t *ddst = nullptr; bool zero_copy = false; int wrap = 0; };
int deflate_stream_enqueue(nxgpu_ctx *c, const void *src, uint64_t src_len, void *dst, uint64_t dst_cap,
			   int level, int wrap, uint32_t chunk, int mem, StreamEnq *e);
int deflate_stream_collect(nxgpu_ctx *c, const StreamEnq &e, void *dst, uint64_t dst_cap, int mem, uint64_t *chunk_offsets, nxgpu_stream_result *res);
// one stream decoded by many warps (inflate_par.cuh): inflate_par_select() marks the descriptors of a launch that take the
// parallel path (kWrapSkip in jobs[], the originals in `picked`), inflate_parallel() runs them on c->stream behind that
// launch, results in the launch's own result slots
void inflate_par_select(InflateJob *jobs, size_t n, std::vector<std::pair<size_t, InflateJob>> &picked, bool dry_too = false);
int inflate_parallel(nxgpu_ctx *c, const std::vector<std::pair<size_t, InflateJob>> &picked, InflateOut *d_outs, bool inputs_marked = false);
void timer_begin(nxgpu_ctx *c, int fam);
void timer_end(nxgpu_ctx *c, int fam);
// nxgpu_job.cu: NX job descriptors, one at a time or coalesced
int run_job_impl(nxgpu_ctx *c, uint8_t *crb_cpb);
void run_jobs_batch(nxgpu_ctx *c, uint8_t *const *crbs, int *rcs, size_t n);
}
