/*
 * cpu_baseline.c — the reference's software path timed on host cores, driven like
 * samples/compdecomp_th.c:134-340,349-473: T pthreads released by a barrier, each
 * owning a contiguous share of the input cut into `piece` byte pieces, one
 * compress2()/uncompress() per piece, throughput = uncompressed bytes / wall time.
 * The arithmetic is system zlib — what lib/sw_zlib.c:283-327 dlopens.  When
 * oracle/_ref/libnxz_ref.so exists bench.py points `lib` at it (kind "reference",
 * NX_GZIP_TYPE_SELECTOR=1 routes through sw_zlib.c); otherwise libz.so.1 ("port").
 * TEST/BENCH INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*compress2_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long, int);
typedef int (*uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);
typedef unsigned long (*bound_fn)(unsigned long);
typedef unsigned long (*cksum_fn)(unsigned long, const unsigned char *, unsigned int);

typedef struct {
	const uint8_t *src; uint64_t len; uint32_t piece; int level; int mode;   /* 0 deflate 1 inflate 2 crc32 3 adler32 */
	compress2_fn c2; uncompress_fn un; bound_fn bound; cksum_fn ck;
	pthread_barrier_t *bar;
	uint64_t out_bytes; int rc;
	/* inflate mode: pre-compressed pieces */
	uint8_t **cbuf; unsigned long *clen; uint64_t npieces;
} work_t;

static void *worker(void *arg)
{
	work_t *w = arg;
	unsigned long cap = w->bound ? w->bound(w->piece) : w->piece + 1024;
	uint8_t *tmp = malloc(cap > w->piece ? cap : w->piece);
	pthread_barrier_wait(w->bar);
	if (w->mode == 0) {
		for (uint64_t o = 0; o < w->len; o += w->piece) {
			unsigned long n = w->len - o < w->piece ? w->len - o : w->piece, ol = cap;
			if (w->c2(tmp, &ol, w->src + o, n, w->level)) w->rc = -1;
			w->out_bytes += ol;
		}
	} else if (w->mode == 1) {
		for (uint64_t i = 0; i < w->npieces; i++) {
			unsigned long ol = w->piece;
			if (w->un(tmp, &ol, w->cbuf[i], w->clen[i])) w->rc = -1;
			w->out_bytes += ol;
		}
	} else {
		unsigned long c = w->mode == 2 ? 0 : 1;
		for (uint64_t o = 0; o < w->len; o += 1u << 30) {
			uint64_t n = w->len - o < (1u << 30) ? w->len - o : (1u << 30);
			c = w->ck(c, w->src + o, (unsigned int)n);
		}
		w->out_bytes = c;
	}
	pthread_barrier_wait(w->bar);
	free(tmp);
	return NULL;
}

/* returns seconds; *out_bytes = compressed bytes (deflate) / output bytes (inflate) */
double oracle_cpu_baseline(const char *lib, const uint8_t *src, uint64_t len, uint32_t piece,
			   int level, int mode, int threads, uint64_t *out_bytes)
{
	void *h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
	if (!h) return -1.0;
	compress2_fn c2 = (compress2_fn)dlsym(h, "compress2");
	uncompress_fn un = (uncompress_fn)dlsym(h, "uncompress");
	bound_fn bound = (bound_fn)dlsym(h, "compressBound");
	cksum_fn ck = (cksum_fn)dlsym(h, mode == 3 ? "adler32" : "crc32");
	if (!c2 || !un || !ck) return -1.0;
	if (threads < 1) threads = 1;
	pthread_barrier_t bar;
	pthread_barrier_init(&bar, NULL, threads + 1);
	work_t *w = calloc(threads, sizeof(*w));
	pthread_t *t = calloc(threads, sizeof(*t));
	uint64_t pieces = (len + piece - 1) / piece, per = (pieces + threads - 1) / threads;
	for (int i = 0; i < threads; i++) {
		uint64_t p0 = (uint64_t)i * per, p1 = p0 + per > pieces ? pieces : p0 + per;
		if (p0 > pieces) p0 = p1 = pieces;
		w[i].src = src + p0 * piece;
		w[i].len = (p1 * piece > len ? len : p1 * piece) - p0 * piece;
		w[i].piece = piece; w[i].level = level; w[i].mode = mode;
		w[i].c2 = c2; w[i].un = un; w[i].bound = bound; w[i].ck = ck; w[i].bar = &bar;
		if (mode == 1) {
			w[i].npieces = p1 - p0;
			w[i].cbuf = calloc(w[i].npieces + 1, sizeof(uint8_t *));
			w[i].clen = calloc(w[i].npieces + 1, sizeof(unsigned long));
			for (uint64_t k = 0; k < w[i].npieces; k++) {
				uint64_t o = k * piece;
				unsigned long n = w[i].len - o < piece ? w[i].len - o : piece;
				unsigned long cl = bound ? bound(n) : n + 1024;
				w[i].cbuf[k] = malloc(cl);
				c2(w[i].cbuf[k], &cl, w[i].src + o, n, level);
				w[i].clen[k] = cl;
			}
		}
		pthread_create(&t[i], NULL, worker, &w[i]);
	}
	struct timespec a, b;
	pthread_barrier_wait(&bar);
	clock_gettime(CLOCK_MONOTONIC, &a);
	pthread_barrier_wait(&bar);
	clock_gettime(CLOCK_MONOTONIC, &b);
	uint64_t tot = 0; int rc = 0;
	for (int i = 0; i < threads; i++) {
		pthread_join(t[i], NULL);
		tot += w[i].out_bytes; rc |= w[i].rc;
		if (mode == 1) {
			for (uint64_t k = 0; k < w[i].npieces; k++) free(w[i].cbuf[k]);
			free(w[i].cbuf); free(w[i].clen);
		}
	}
	if (out_bytes) *out_bytes = tot;
	free(w); free(t);
	pthread_barrier_destroy(&bar);
	dlclose(h);
	if (rc) return -2.0;
	return (double)(b.tv_sec - a.tv_sec) + (double)(b.tv_nsec - a.tv_nsec) * 1e-9;
}
