// nxgpu_gz.cu — a gz* READER on top of the batched inflate.
//
// The reference's gz layer pulls at most 10 bytes per read() and stops after the first member
// (lib/nx_gzlib.c:220-263 __gzread, :278-325); samples/gunzip_nx.c walks members one job at a time.  Here a gz
// file is read whole, its members (one or many: `cat a.gz b.gz`, bgzip, pigz -i) are discovered and inflated on
// the device as ONE batch (nxgpu_gunzip_concat: candidate headers, dry decoding run, chain from offset 0, CRC-32 and
// ISIZE of every member verified), and gzread() serves bytes from the inflated image.  Same calling convention as
// zlib's gzopen / gzdopen / gzread / gzeof / gzclose, so lib/nx_gzlib.c can forward its read side here
// (INTEGRATION.md shows the four lines).  The file and its inflated image must fit in host memory.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "ctx.cuh"

struct nxgpu_gzfile {
	nxgpu_ctx *c = nullptr;
	bool own_ctx = false;
	int fd = -1;
	uint8_t *out = nullptr;      // inflated image (pinned)
	uint64_t out_len = 0, pos = 0;
	uint32_t members = 0;
	bool loaded = false;
	int err = 0;
};

namespace {

int load_all(nxgpu_gzfile *g)
{
	g->loaded = true;
	struct stat st;
	if (fstat(g->fd, &st) != 0) return g->err = NXGPU_E_ARG;
	uint64_t cap = st.st_size > 0 ? (uint64_t)st.st_size : (1u << 20), n = 0;
	void *in = nullptr;
	if (cudaMallocHost(&in, cap + 64) != cudaSuccess) { cudaGetLastError(); set_error("gz: cannot allocate %llu bytes", (unsigned long long)cap); return g->err = NXGPU_E_MEM; }
	for (;;) {
		if (n == cap) {                              // a pipe or a growing file
			void *bigger = nullptr;
			if (cudaMallocHost(&bigger, 2 * cap + 64) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(in); return g->err = NXGPU_E_MEM; }
			memcpy(bigger, in, n);
			cudaFreeHost(in);
			in = bigger; cap *= 2;
		}
		const ssize_t k = read(g->fd, static_cast<uint8_t *>(in) + n, cap - n);
		if (k < 0) { if (errno == EINTR) continue; cudaFreeHost(in); return g->err = NXGPU_E_ARG; }
		if (k == 0) break;
		n += (uint64_t)k;
	}
	if (n == 0) { cudaFreeHost(in); return 0; }          // an empty file reads as an empty stream, like zlib
	// sizes first (candidates + dry run), then the real batch into a buffer of exactly that size
	uint64_t need = 0;
	int rc = nxgpu_gunzip_concat(g->c, in, n, nullptr, 0, &need, &g->members, NXGPU_MEM_HOST);
	if (rc != 0 && rc != NXGPU_E_BUF) { cudaFreeHost(in); return g->err = rc; }
	if (need) {
		void *o = nullptr;
		if (cudaMallocHost(&o, need) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(in); return g->err = NXGPU_E_MEM; }
		g->out = static_cast<uint8_t *>(o);
		rc = nxgpu_gunzip_concat(g->c, in, n, g->out, need, &g->out_len, &g->members, NXGPU_MEM_HOST);
		if (rc) { cudaFreeHost(in); return g->err = rc; }
	}
	cudaFreeHost(in);
	return 0;
}

} // namespace

extern "C" {

nxgpu_gzfile *nxgpu_gzdopen(nxgpu_ctx *ctx, int fd, const char *mode)
{
	if (fd < 0 || !mode || mode[0] != 'r') { set_error("nxgpu_gzdopen: only reading is bound to the batched inflate (mode \"%s\")", mode ? mode : ""); return nullptr; }
	nxgpu_gzfile *g = new nxgpu_gzfile();
	g->fd = fd;
	if (ctx) {
		g->c = ctx;
	} else {
		if (nxgpu_open(-1, &g->c) != 0) { delete g; return nullptr; }
		g->own_ctx = true;
	}
	return g;
}

nxgpu_gzfile *nxgpu_gzopen(nxgpu_ctx *ctx, const char *path, const char *mode)
{
	if (!path) return nullptr;
	const int fd = open(path, O_RDONLY);
	if (fd < 0) { set_error("nxgpu_gzopen: %s: %s", path, strerror(errno)); return nullptr; }
	nxgpu_gzfile *g = nxgpu_gzdopen(ctx, fd, mode);
	if (!g) close(fd);
	return g;
}

// like zlib's gzread: the number of uncompressed bytes delivered, 0 at the end of the file, -1 on error
int nxgpu_gzread(nxgpu_gzfile *g, void *buf, unsigned len)
{
	if (!g || (!buf && len)) return -1;
	if (!g->loaded && load_all(g) != 0) return -1;
	if (g->err) return -1;
	const uint64_t left = g->out_len - g->pos;
	const unsigned n = left < len ? (unsigned)left : len;
	if (n > 0x7fffffffu) return -1;
	if (n) memcpy(buf, g->out + g->pos, n);
	g->pos += n;
	return (int)n;
}

int nxgpu_gzeof(nxgpu_gzfile *g) { return g && g->loaded && !g->err && g->pos == g->out_len; }
uint32_t nxgpu_gzmembers(nxgpu_gzfile *g) { return g ? g->members : 0; }

int nxgpu_gzclose(nxgpu_gzfile *g)
{
	if (!g) return NXGPU_E_ARG;
	const int rc = g->err;
	if (g->out) cudaFreeHost(g->out);
	if (g->fd >= 0) close(g->fd);
	if (g->own_ctx) nxgpu_close(g->c);
	delete g;
	return rc;
}

} // extern "C"
