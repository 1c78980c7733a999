// nxgpu_team.cu — ONE gzip / zlib / raw member deflated by the GPUs of one box (SURVEY.md §8e,
// BASELINE.json configs[3]): one process (or thread) per GPU, each owning a contiguous chunk range.
//
//   rank r:  deflate its range as raw deflate, every chunk ending on the joiner, BFINAL only on the
//            last rank (the same joiner the reference puts between jobs, lib/nx_deflate.c:220-243)
//        ->  publish {compressed size, crc32, adler32, length} FROM THE GPU into a control block that
//            every rank has mapped (pinned shared memory)
//        ->  a one-warp kernel waits for the sizes of the ranks in front, exclusive-scans them
//        ->  a copy kernel writes the range at its global offset straight into the destination:
//              NXGPU_MEM_HOST    the destination lives in the shared segment, every GPU writes its part
//                                over its OWN PCIe link (nothing funnels through GPU 0)
//              NXGPU_MEM_DEVICE  rank 0's device buffer, mapped into the peers by CUDA IPC: P2P stores
//                                over NVLink / NVSwitch
//   No host round trip between the deflate kernel and the copy: the chain is enqueued on one stream.
//   One synchronisation at the end; rank 0 folds the checksums (crc32_combine, lib/nx_crc.c:374, as
//   lib/nx_deflate.c:1562-1577 folds the checksums of consecutive jobs) and writes header + trailer.
//
// The rendezvous is a named POSIX shared-memory segment (a file under /tmp when /dev/shm is too small);
// there is no dependency on MPI / NCCL / torch here.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <string>
#include "ctx.cuh"

namespace {

constexpr int kMaxRanks = 16;
constexpr uint64_t kMagic = 0x6e786770757465ull;   // "nxgpute"
constexpr uint64_t kCtlBytes = 8192;

struct Slot {                         // written by rank r's GPU (publish_kernel), read by the GPUs behind it and by rank 0's host
	volatile uint64_t seq;
	uint64_t size;                    // compressed bytes of the range (no header / trailer)
	uint64_t ulen;                    // uncompressed bytes
	uint32_t crc, adler;
	volatile uint64_t overflow;       // the range did not fit the destination
	uint64_t pad[3];
};
static_assert(sizeof(Slot) == 64, "slot");

struct TeamCtl {
	volatile uint64_t magic;
	uint32_t nranks, dst_mem;
	uint64_t dst_cap, seg_bytes;
	volatile uint32_t attached[kMaxRanks];
	volatile uint64_t done[kMaxRanks];         // host barrier: rank r finished call `seq`
	volatile uint64_t result_seq;              // rank 0 published the result of call `seq`
	uint64_t total; uint32_t crc, adler; uint64_t ulen; int64_t rc;
	// NXGPU_MEM_DEVICE: rank 0's destination buffer
	volatile uint32_t handle_ready;
	int owner_pid, owner_dev;
	uint64_t owner_ptr;
	cudaIpcMemHandle_t handle;
	Slot slot[kMaxRanks];
};
static_assert(sizeof(TeamCtl) <= kCtlBytes, "control block");

double now_s()
{
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

__global__ void publish_kernel(Slot *slot, const uint64_t *d_off, uint32_t n_chunks, const uint32_t *d_cks, uint64_t ulen, uint64_t seq)
{
	slot->size = d_off[n_chunks];             // raw stream, base 0: the end of the last chunk is the size
	slot->ulen = ulen;
	slot->crc = d_cks[0];
	slot->adler = d_cks[1];
	slot->overflow = 0;
	__threadfence_system();
	slot->seq = seq;
}

// one warp: lane r < rank polls rank r's slot; place[0] = global offset of this rank's range, place[1] = its size
__global__ void scan_kernel(const Slot *slots, uint32_t rank, uint64_t base, uint64_t seq, const uint64_t *d_off, uint32_t n_chunks,
			    uint64_t *place, long long timeout_cycles)
{
	const uint32_t lane = threadIdx.x;
	uint64_t v = 0;
	bool ok = true;
	if (lane < rank) {
		const long long t0 = clock64();
		while (slots[lane].seq != seq) {
			__nanosleep(2000);
			if (clock64() - t0 > timeout_cycles) { ok = false; break; }
		}
		__threadfence_system();
		v = slots[lane].size;
	}
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	ok = __all_sync(0xffffffffu, ok);
	if (lane == 0) {
		place[0] = base + v;
		place[1] = d_off[n_chunks];
		place[2] = ok ? 0 : 1;
	}
}

// dst[place[0] .. +place[1]) = src[0 .. place[1]): 16-byte stores to an arbitrarily aligned destination (device, peer or host)
__global__ void __launch_bounds__(256)
place_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t dst_cap, const uint64_t *__restrict__ place, Slot *slot)
{
	const uint64_t off = place[0], len = place[1];
	if (place[2] || off + len > dst_cap) {
		if (blockIdx.x == 0 && threadIdx.x == 0) { slot->overflow = 1; __threadfence_system(); }
		return;
	}
	uint8_t *d = dst + off;
	const uint64_t head = min(len, (uint64_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15));
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
	if (tid < head)
		d[tid] = src[tid];
	const uint64_t nvec = (len - head) >> 4;
	const uint32_t sh = (uint32_t)(head & 3) * 8;
	const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src) + (head >> 2);     // src is 16-byte aligned
	uint4 *d128 = reinterpret_cast<uint4 *>(d + head);
	for (uint64_t v = tid; v < nvec; v += nth) {
		const uint32_t *p = s32 + 4 * v;
		const uint32_t a = p[0], b = p[1], c = p[2], e = p[3], f = sh ? p[4] : 0;
		d128[v] = sh ? make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh))
			     : make_uint4(a, b, c, e);
	}
	const uint64_t done = head + nvec * 16;
	if (tid < len - done)
		d[done + tid] = src[done + tid];
}

__global__ void poke_bytes_kernel(uint8_t *dst, uint64_t v, int n)
{
	for (int k = 0; k < n; k++)
		dst[k] = (uint8_t)(v >> (8 * k));
}

} // namespace

struct nxgpu_team {
	nxgpu_ctx *c = nullptr;
	int rank = 0, nranks = 1;
	std::string name, path;
	bool is_file = false;
	int fd = -1;
	uint8_t *base = nullptr;              // mapped segment (host)
	uint8_t *base_dev = nullptr;          // the same through the GPU's address space
	uint64_t seg_bytes = 0;
	TeamCtl *ctl = nullptr;
	uint8_t *dst_host = nullptr;          // NXGPU_MEM_HOST destination (inside the segment)
	uint8_t *dst_dev = nullptr;           // where this rank's GPU writes the member
	bool ipc_opened = false, own_dst = false, registered = false;
	uint64_t seq = 0;
	void *d_place = nullptr;
	DevBuf d_local;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	float last_ms = 0;
};

extern "C" {

void nxgpu_team_close(nxgpu_team *t)
{
	if (!t)
		return;
	if (t->c)
		cudaSetDevice(t->c->dev);
	if (t->ev0) cudaEventDestroy(t->ev0);
	if (t->ev1) cudaEventDestroy(t->ev1);
	if (t->d_place) cudaFree(t->d_place);
	t->d_local.release();
	if (t->ipc_opened && t->dst_dev) cudaIpcCloseMemHandle(t->dst_dev);
	if (t->own_dst && t->dst_dev) cudaFree(t->dst_dev);
	if (t->registered && t->base) cudaHostUnregister(t->base);
	if (t->base) munmap(t->base, t->seg_bytes);
	if (t->fd >= 0) close(t->fd);
	if (t->rank == 0 && !t->path.empty()) {
		if (t->is_file) unlink(t->path.c_str()); else shm_unlink(t->path.c_str());
	}
	delete t;
}

int nxgpu_team_open(nxgpu_ctx *c, const char *name, int rank, int nranks, uint64_t dst_cap, int dst_mem, nxgpu_team **out)
{
	if (!c || !name || !out || rank < 0 || nranks < 1 || nranks > kMaxRanks || rank >= nranks || dst_cap < 64) return NXGPU_E_ARG;
	if (dst_mem != NXGPU_MEM_HOST && dst_mem != NXGPU_MEM_DEVICE) return NXGPU_E_ARG;
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	nxgpu_team *t = new nxgpu_team();
	struct Guard { nxgpu_team *t; ~Guard() { if (t) nxgpu_team_close(t); } } guard{ t };
	t->c = c; t->rank = rank; t->nranks = nranks; t->name = name;
	const uint64_t seg = kCtlBytes + (dst_mem == NXGPU_MEM_HOST ? ((dst_cap + 4095) & ~4095ull) : 0);
	const std::string shm_name = std::string("/") + name, file_name = std::string("/tmp/") + name + ".nxgpu-team";
	if (rank == 0) {
		shm_unlink(shm_name.c_str());
		unlink(file_name.c_str());
		int fd = shm_open(shm_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
		if (fd >= 0 && posix_fallocate(fd, 0, (off_t)seg) != 0) {      // /dev/shm too small: fall back to a file (page cache)
			close(fd); shm_unlink(shm_name.c_str()); fd = -1;
		}
		if (fd >= 0) {
			t->path = shm_name;
		} else {
			fd = open(file_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
			if (fd < 0 || posix_fallocate(fd, 0, (off_t)seg) != 0) { set_error("team: cannot create the shared segment (%llu bytes): %s", (unsigned long long)seg, strerror(errno)); if (fd >= 0) close(fd); return NXGPU_E_MEM; }
			t->path = file_name; t->is_file = true;
		}
		t->fd = fd;
	} else {
		const double t0 = now_s();
		for (;;) {
			int fd = shm_open(shm_name.c_str(), O_RDWR, 0600);
			bool file = false;
			if (fd < 0) { fd = open(file_name.c_str(), O_RDWR); file = fd >= 0; }
			if (fd >= 0) {
				struct stat st;
				if (fstat(fd, &st) == 0 && (uint64_t)st.st_size >= seg) { t->fd = fd; t->is_file = file; break; }
				close(fd);
			}
			if (now_s() - t0 > 120) { set_error("team: rank 0 never created segment %s", name); return NXGPU_E_NODEV; }
			usleep(2000);
		}
	}
	void *m = mmap(nullptr, seg, PROT_READ | PROT_WRITE, MAP_SHARED, t->fd, 0);
	if (m == MAP_FAILED) { set_error("team: mmap failed: %s", strerror(errno)); return NXGPU_E_MEM; }
	t->base = static_cast<uint8_t *>(m); t->seg_bytes = seg;
	t->ctl = reinterpret_cast<TeamCtl *>(t->base);
	if (rank == 0) {
		memset(t->base, 0, kCtlBytes);
		t->ctl->nranks = (uint32_t)nranks; t->ctl->dst_mem = (uint32_t)dst_mem; t->ctl->dst_cap = dst_cap; t->ctl->seg_bytes = seg;
		__sync_synchronize();
		t->ctl->magic = kMagic;
	} else {
		const double t0 = now_s();
		while (t->ctl->magic != kMagic) {
			if (now_s() - t0 > 120) { set_error("team: segment %s never initialised", name); return NXGPU_E_NODEV; }
			usleep(1000);
		}
		__sync_synchronize();
		if (t->ctl->nranks != (uint32_t)nranks || t->ctl->dst_mem != (uint32_t)dst_mem || t->ctl->dst_cap != dst_cap) { set_error("team: ranks disagree on the team parameters"); return NXGPU_E_ARG; }
	}
	// the segment is pinned and mapped into this GPU's address space: kernels read the control block and (host mode) write the member
	NXGPU_CUDA_OK(cudaHostRegister(t->base, seg, cudaHostRegisterMapped | cudaHostRegisterPortable));
	t->registered = true;
	void *dp = nullptr;
	NXGPU_CUDA_OK(cudaHostGetDevicePointer(&dp, t->base, 0));
	t->base_dev = static_cast<uint8_t *>(dp);
	if (dst_mem == NXGPU_MEM_HOST) {
		t->dst_host = t->base + kCtlBytes;
		t->dst_dev = t->base_dev + kCtlBytes;
	} else if (rank == 0) {
		void *d = nullptr;
		NXGPU_CUDA_OK(cudaMalloc(&d, dst_cap));
		t->dst_dev = static_cast<uint8_t *>(d); t->own_dst = true;
		NXGPU_CUDA_OK(cudaIpcGetMemHandle(const_cast<cudaIpcMemHandle_t *>(&t->ctl->handle), d));
		t->ctl->owner_pid = (int)getpid(); t->ctl->owner_dev = c->dev; t->ctl->owner_ptr = (uint64_t)(uintptr_t)d;
		__sync_synchronize();
		t->ctl->handle_ready = 1;
	} else {
		const double t0 = now_s();
		while (!t->ctl->handle_ready) {
			if (now_s() - t0 > 120) { set_error("team: rank 0 never exported its buffer"); return NXGPU_E_NODEV; }
			usleep(1000);
		}
		__sync_synchronize();
		if (t->ctl->owner_pid == (int)getpid()) {
			// threads of one process: the pointer is valid as it is once peer access is on
			cudaError_t e = t->ctl->owner_dev == c->dev ? cudaSuccess : cudaDeviceEnablePeerAccess(t->ctl->owner_dev, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("team: no peer access to device %d: %s", t->ctl->owner_dev, cudaGetErrorString(e)); return NXGPU_E_NODEV; }
			cudaGetLastError();
			t->dst_dev = reinterpret_cast<uint8_t *>((uintptr_t)t->ctl->owner_ptr);
		} else {
			void *d = nullptr;
			cudaIpcMemHandle_t h;
			memcpy(&h, const_cast<cudaIpcMemHandle_t *>(&t->ctl->handle), sizeof(h));
			NXGPU_CUDA_OK(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
			t->dst_dev = static_cast<uint8_t *>(d); t->ipc_opened = true;
		}
	}
	NXGPU_CUDA_OK(cudaMalloc(&t->d_place, 64));
	NXGPU_CUDA_OK(cudaEventCreate(&t->ev0));
	NXGPU_CUDA_OK(cudaEventCreate(&t->ev1));
	t->ctl->attached[rank] = (uint32_t)getpid();
	__sync_synchronize();
	// everyone attached before anyone computes (rank 0 may unlink the name afterwards; the mapping stays)
	const double t0 = now_s();
	for (int r = 0; r < nranks; r++)
		while (!t->ctl->attached[r]) {
			if (now_s() - t0 > 120) { set_error("team: rank %d never attached", r); return NXGPU_E_NODEV; }
			usleep(1000);
		}
	guard.t = nullptr;
	*out = t;
	return 0;
}

void *nxgpu_team_dst(nxgpu_team *t)
{
	if (!t) return nullptr;
	if (t->ctl->dst_mem == NXGPU_MEM_HOST) return t->dst_host;     // valid in every rank
	return t->rank == 0 ? t->dst_dev : nullptr;                    // rank 0's device buffer
}



int nxgpu_team_deflate(nxgpu_team *t, const void *src, uint64_t src_len, int level, int wrap, uint32_t chunk, int src_mem, nxgpu_team_result *res)
{
	if (!t || !res || (!src && src_len)) return NXGPU_E_ARG;
	if (wrap != NXGPU_WRAP_RAW && wrap != NXGPU_WRAP_ZLIB && wrap != NXGPU_WRAP_GZIP) return NXGPU_E_ARG;
	nxgpu_ctx *c = t->c;
	NXGPU_LOCK(c);
	NXGPU_CUDA_OK(cudaSetDevice(c->dev));
	TeamCtl *ctl = t->ctl;
	const uint64_t seq = ++t->seq;
	const bool last = t->rank == t->nranks - 1;
	if (chunk == 0) chunk = 262144;
	int rc;
	// ---- this rank's range: raw deflate into a device buffer of its own ----
	const uint64_t local_cap = nxgpu_deflate_stream_bound(src_len, chunk);
	uint8_t *local = nullptr;
	uint8_t dummy_host[16];
	void *enq_dst;
	if (src_mem == NXGPU_MEM_HOST) {
		enq_dst = dummy_host;                      // host-pointer flavour: the stream stays in the context's device buffer (e.ddst)
	} else {
		if ((rc = t->d_local.reserve(local_cap + 64))) return rc;
		enq_dst = t->d_local.p;
	}
	NXGPU_CUDA_OK(cudaEventRecord(t->ev0, c->stream));
	StreamEnq e;
	if ((rc = deflate_stream_enqueue(c, src, src_len, enq_dst, local_cap, level, last ? NXGPU_WRAP_RAW : NXGPU_WRAP_RAW_CONT, chunk, src_mem, &e))) return rc;
	local = e.ddst;
	// ---- publish, scan, place: all on the stream, no host in between ----
	uint64_t hdr = 0; int hdr_len = 0;
	if (wrap == NXGPU_WRAP_GZIP) hdr_len = 10;
	else if (wrap == NXGPU_WRAP_ZLIB) hdr_len = 2;
	Slot *slots_dev = reinterpret_cast<Slot *>(t->base_dev + offsetof(TeamCtl, slot));
	uint64_t *place = static_cast<uint64_t *>(t->d_place);
	publish_kernel<<<1, 1, 0, c->stream>>>(slots_dev + t->rank, e.d_off, (uint32_t)e.n, e.d_cks, src_len, seq);
	int clock_khz = 1900000;
	cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, c->dev);
	scan_kernel<<<1, 32, 0, c->stream>>>(slots_dev, (uint32_t)t->rank, (uint64_t)hdr_len, seq, e.d_off, (uint32_t)e.n, place, (long long)clock_khz * 1000 * 30);
	place_kernel<<<kNumSMs * 4, 256, 0, c->stream>>>(local, t->dst_dev, ctl->dst_cap, place, slots_dev + t->rank);
	if (t->rank == 0 && hdr_len) {
		if (wrap == NXGPU_WRAP_GZIP) {
			// 1f 8b 08 00 mtime=0 xfl=0 os=3: the blank header of lib/nx_deflate.c:473-489
			hdr = 0x1full | 0x8bull << 8 | 0x08ull << 16;
			poke_bytes_kernel<<<1, 1, 0, c->stream>>>(t->dst_dev, hdr, 8);
			poke_bytes_kernel<<<1, 1, 0, c->stream>>>(t->dst_dev + 8, 0x0300ull, 2);
		} else {
			const int lv = level <= 0 ? 6 : level;
			const uint32_t flevel = lv < 2 ? 0 : lv < 6 ? 1 : lv == 6 ? 2 : 3;
			uint32_t h = (0x78u << 8) | (flevel << 6);
			h += 31 - (h % 31);
			poke_bytes_kernel<<<1, 1, 0, c->stream>>>(t->dst_dev, (h >> 8) | ((h & 0xff) << 8), 2);
		}
	}
	c->launches += 3;
	NXGPU_CUDA_OK(cudaGetLastError());
	NXGPU_CUDA_OK(cudaEventRecord(t->ev1, c->stream));
	// ---- the one synchronisation; per-chunk status of this rank ----
	nxgpu_stream_result lr;
	uint8_t dummy2[16];
	rc = deflate_stream_collect(c, e, src_mem == NXGPU_MEM_HOST ? static_cast<void *>(dummy2) : enq_dst, local_cap, NXGPU_MEM_DEVICE, nullptr, &lr);
	cudaEventElapsedTime(&t->last_ms, t->ev0, t->ev1);
	uint64_t my_place[3] = { 0, 0, 0 };
	if (!rc) NXGPU_CUDA_OK(cudaMemcpy(my_place, place, sizeof(my_place), cudaMemcpyDeviceToHost));
	if (rc) ctl->slot[t->rank].overflow = 2;
	__sync_synchronize();
	ctl->done[t->rank] = seq;
	// ---- host barrier, then rank 0 finishes the member ----
	{
		const double t0 = now_s();
		for (int r = 0; r < t->nranks; r++)
			while (ctl->done[r] < seq) {
				if (now_s() - t0 > 120) { set_error("team: rank %d did not finish call %llu", r, (unsigned long long)seq); return NXGPU_E_NODEV; }
				sched_yield();
			}
		__sync_synchronize();
	}
	if (t->rank == 0) {
		uint64_t total = hdr_len, ulen = 0;
		uint32_t crc = 0, adler = 1;
		int64_t trc = 0;
		for (int r = 0; r < t->nranks; r++) {
			const Slot &s = ctl->slot[r];
			if (s.overflow == 2 || s.seq != seq) trc = NXGPU_E_DATA;
			else if (s.overflow) trc = NXGPU_E_BUF;
			total += s.size;
			crc = host_crc32_combine(crc, s.crc, s.ulen);
			adler = nxgpu_adler32_combine(adler, s.adler, s.ulen);
			ulen += s.ulen;
		}
		uint8_t tr[8]; int tr_len = 0;
		if (wrap == NXGPU_WRAP_GZIP) {
			for (int k = 0; k < 4; k++) { tr[k] = (uint8_t)(crc >> (8 * k)); tr[4 + k] = (uint8_t)((uint32_t)ulen >> (8 * k)); }
			tr_len = 8;
		} else if (wrap == NXGPU_WRAP_ZLIB) {
			for (int k = 0; k < 4; k++) tr[k] = (uint8_t)(adler >> (24 - 8 * k));
			tr_len = 4;
		}
		if (!trc && total + tr_len > ctl->dst_cap) trc = NXGPU_E_BUF;
		if (!trc && tr_len) {
			if (ctl->dst_mem == NXGPU_MEM_HOST) memcpy(t->dst_host + total, tr, tr_len);
			else if (cudaMemcpy(t->dst_dev + total, tr, tr_len, cudaMemcpyHostToDevice) != cudaSuccess) trc = NXGPU_E_NODEV;
		}
		ctl->total = total + tr_len; ctl->crc = crc; ctl->adler = adler; ctl->ulen = ulen; ctl->rc = trc;
		__sync_synchronize();
		ctl->result_seq = seq;
	} else {
		const double t0 = now_s();
		while (ctl->result_seq < seq) {
			if (now_s() - t0 > 120) { set_error("team: rank 0 never published the result"); return NXGPU_E_NODEV; }
			sched_yield();
		}
		__sync_synchronize();
	}
	res->out_len = ctl->total; res->crc32 = ctl->crc; res->adler32 = ctl->adler; res->src_len = ctl->ulen;
	res->my_offset = my_place[0]; res->my_size = my_place[1];
	res->device_ms = t->last_ms;
	if (rc) return rc;
	if (ctl->rc) { set_error("team deflate: a rank failed or the destination (%llu bytes) is too small", (unsigned long long)ctl->dst_cap); return (int)ctl->rc; }
	return 0;
}

} // extern "C"
