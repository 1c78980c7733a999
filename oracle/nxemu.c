/*
 * nxemu.c — software NX-GZIP engine: nxu_run_job() (reference lib/gzip_vas.c:281) restated on
 * the CPU for the x86 build of the UNMODIFIED reference host code (oracle/_ref/libnxz_ref.so).
 * TEST INFRASTRUCTURE ONLY (see oracle.h): it pins down what the host code above the boundary
 * expects in the CSB / CPB after a job, so that the same expectations can be checked against
 * the GPU engine.
 *
 * Follows: descriptor layout inc_nx/nxu.h:286-616 (offsets via include/nxgpu.h, asserted by
 * layout_check.c); how compress outputs are consumed lib/nx_deflate.c:919-1078,1267-1282;
 * decompress resume protocol lib/nx_inflate.c:1372-1609; wrap lib/nx_zlib.c:1398-1440;
 * checksum byte order lib/nx_deflate.c:1572-1577, lib/nx_inflate.c:809-817; lzcount format
 * lib/nx_dht.c:187-199; minimal protocols selftest/gzfht_test.c:100-141,335-405.
 *
 * The compressor is deliberately simple (greedy, one hash probe): the NX contract is "one
 * deflate block per job, coded with the table the caller named", not a particular parse.
 */
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define NXGPU_NO_DROPIN_DECLS
#include "../include/nxgpu.h"
#include "oracle.h"

static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static uint64_t be64(const uint8_t *p) { return (uint64_t)be32(p) << 32 | be32(p + 4); }
static void put_be32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static uint32_t le32(const uint8_t *p) { return (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]; }
static void put_le32(uint8_t *p, uint32_t v) { p[3] = v >> 24; p[2] = v >> 16; p[1] = v >> 8; p[0] = v; }

typedef struct { uint8_t *p; uint32_t len; } seg;
typedef struct { seg s[260]; int n; uint64_t total; } seglist;

static int dde_segments(const uint8_t *dde, seglist *out)
{
	uint32_t count = (be32(dde) >> 8) & 0xff, bc = be32(dde + 4);
	uint64_t addr = be64(dde + 8);
	out->n = 0; out->total = 0;
	if (count == 0) {
		if (bc) { out->s[0].p = (uint8_t *)(uintptr_t)addr; out->s[0].len = bc; out->n = 1; }
		out->total = bc;
		return 0;
	}
	const uint8_t *list = (const uint8_t *)(uintptr_t)addr;
	uint64_t left = bc;
	for (uint32_t i = 0; i < count && left; i++) {
		const uint8_t *d = list + 16 * i;
		if ((be32(d) >> 8) & 0xff) return -1;
		uint64_t l = be32(d + 4);
		uint32_t use = (uint32_t)(l < left ? l : left);
		if (use) { out->s[out->n].p = (uint8_t *)(uintptr_t)be64(d + 8); out->s[out->n].len = use; out->n++; }
		left -= use; out->total += use;
	}
	return 0;
}
static void gather(const seglist *l, uint8_t *dst) { for (int i = 0; i < l->n; i++) { memcpy(dst, l->s[i].p, l->s[i].len); dst += l->s[i].len; } }
static void scatter(const seglist *l, const uint8_t *src, uint64_t n)
{
	for (int i = 0; i < l->n && n; i++) {
		uint32_t k = (uint32_t)(l->s[i].len < n ? l->s[i].len : n);
		memcpy(l->s[i].p, src, k); src += k; n -= k;
	}
}
static void complete(uint8_t *crb, uint32_t cc, uint32_t ce_ms3b, uint32_t tpbc)
{
	put_be32(crb + NXGPU_CRB_CSB + 4, tpbc);
	put_be32(crb + NXGPU_CRB_CSB, 0x80000000u | (cc & 0xff) << 8 | ((ce_ms3b & 7) << 5));
}

/* ---- dynamic header (from HLIT) -> code lengths ---- */
typedef struct { const uint8_t *p; uint32_t nbits, bp; } hbits;
static int hget(hbits *b, unsigned n)
{
	if (b->bp + n > b->nbits) return -1;
	uint32_t v = 0;
	for (unsigned i = 0; i < n; i++, b->bp++) v |= (uint32_t)((b->p[b->bp >> 3] >> (b->bp & 7)) & 1) << i;
	return (int)v;
}
static int dht_to_lengths(const uint8_t *bits, uint32_t nbits, uint8_t *ll, uint8_t *dl)
{
	static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	hbits b = { bits, nbits, 0 };
	int v = hget(&b, 14);
	if (v < 0) return -1;
	int hlit = (v & 31) + 257, hdist = ((v >> 5) & 31) + 1, hclen = (v >> 10) + 4;
	if (hlit > 286 || hdist > 30) return -1;
	uint8_t cl[19] = { 0 }, all[320];
	for (int i = 0; i < hclen; i++) { if ((v = hget(&b, 3)) < 0) return -1; cl[order[i]] = (uint8_t)v; }
	uint16_t count[8] = { 0 }, first[8] = { 0 }, code[19], c = 0;
	for (int i = 0; i < 19; i++) count[cl[i]]++;
	count[0] = 0;
	for (int l = 1; l <= 7; l++) { c = (uint16_t)((c + count[l - 1]) << 1); first[l] = c; }
	for (int i = 0; i < 19; i++) code[i] = cl[i] ? first[cl[i]]++ : 0;
	int n = 0;
	while (n < hlit + hdist) {
		int sym = -1; uint32_t acc = 0;
		for (int l = 1; l <= 7 && sym < 0; l++) {
			if ((v = hget(&b, 1)) < 0) return -1;
			acc = (acc << 1) | (uint32_t)v;
			for (int i = 0; i < 19; i++) if (cl[i] == l && code[i] == acc) { sym = i; break; }
		}
		if (sym < 0) return -1;
		if (sym < 16) { all[n++] = (uint8_t)sym; continue; }
		int rep, val = 0;
		if (sym == 16) { if (n == 0 || (v = hget(&b, 2)) < 0) return -1; val = all[n - 1]; rep = 3 + v; }
		else if (sym == 17) { if ((v = hget(&b, 3)) < 0) return -1; rep = 3 + v; }
		else { if ((v = hget(&b, 7)) < 0) return -1; rep = 11 + v; }
		if (n + rep > hlit + hdist) return -1;
		while (rep--) all[n++] = (uint8_t)val;
	}
	memset(ll, 0, 288); memset(dl, 0, 32);
	memcpy(ll, all, hlit); memcpy(dl, all + hlit, hdist);
	return 0;
}

/* ---- one deflate block with a given code ---- */
typedef struct { uint8_t *p; uint64_t bp, cap_bits; int ovf; } bitwr;
static void wput(bitwr *w, uint32_t v, unsigned n)
{
	for (unsigned i = 0; i < n; i++, w->bp++) {
		if (w->bp >= w->cap_bits) { w->ovf = 1; return; }
		if ((w->bp & 7) == 0) w->p[w->bp >> 3] = 0;
		w->p[w->bp >> 3] |= (uint8_t)(((v >> i) & 1) << (w->bp & 7));
	}
}
static void canon(const uint8_t *len, int n, uint16_t *code)
{
	uint16_t cnt[16] = { 0 }, next[16] = { 0 }, c = 0;
	for (int i = 0; i < n; i++) cnt[len[i]]++;
	cnt[0] = 0;
	for (int l = 1; l <= 15; l++) { c = (uint16_t)((c + cnt[l - 1]) << 1); next[l] = c; }
	for (int i = 0; i < n; i++) {
		uint16_t v = len[i] ? next[len[i]]++ : 0, r = 0;
		for (int k = 0; k < len[i]; k++) r |= (uint16_t)(((v >> k) & 1) << (len[i] - 1 - k));
		code[i] = r;                     /* bit-reversed: emitted LSB first */
	}
}
static void len_sym(unsigned len, unsigned *sym, unsigned *nx, unsigned *x)
{
	static const uint16_t base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
	static const uint8_t extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
	int s = 28;
	while (base[s] > len) s--;
	if (len == 258) s = 28;
	*sym = 257 + s; *nx = extra[s]; *x = len - base[s];
}
static void dist_sym(unsigned dist, unsigned *sym, unsigned *nx, unsigned *x)
{
	static const uint16_t base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
	static const uint8_t extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
	int s = 29;
	while (base[s] > dist) s--;
	*sym = s; *nx = extra[s]; *x = dist - base[s];
}

/* returns 0, 13 (no room) or 66 (symbol without a code) */
static int compress_block(const uint8_t *buf, uint32_t hist, uint32_t n, int fixed, const uint8_t *dht, uint32_t dhtlen,
			  uint8_t *out, uint64_t cap, uint32_t *tpbc, uint32_t *tebc, uint32_t *lz)
{
	uint8_t ll[288], dl[32];
	uint16_t lc[288], dc[32];
	if (fixed) {
		for (int i = 0; i < 288; i++) ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
		for (int i = 0; i < 32; i++) dl[i] = 5;
	} else if (dht_to_lengths(dht, dhtlen, ll, dl)) {
		return 68;
	}
	canon(ll, 288, lc); canon(dl, 32, dc);
	bitwr w = { out, 0, cap * 8, 0 };
	wput(&w, fixed ? 2 : 4, 3);                     /* BFINAL=0, BTYPE=01 / 10 */
	if (!fixed)
		for (uint32_t i = 0; i < dhtlen; i++) wput(&w, (dht[i >> 3] >> (i & 7)) & 1, 1);
	memset(lz, 0, 316 * 4);
	int32_t *head = malloc(sizeof(int32_t) << 15);
	for (int i = 0; i < 1 << 15; i++) head[i] = -1;
	const uint32_t end = hist + n;
	int bad = 0;
#define H3(p) (((uint32_t)buf[p] << 10 ^ (uint32_t)buf[(p) + 1] << 5 ^ buf[(p) + 2]) & 0x7fff)
	for (uint32_t p = 0; p + 2 < hist; p++) head[H3(p)] = (int32_t)p;
	for (uint32_t p = hist; p < end && !bad;) {
		unsigned len = 0, dist = 0;
		if (p + 2 < end) {
			int32_t q = head[H3(p)];
			if (q >= 0 && p - (uint32_t)q <= 32768) {
				while (len < 258 && p + len < end && buf[q + len] == buf[p + len]) len++;
				dist = p - (uint32_t)q;
			}
		}
		if (len >= 4) {
			unsigned s, nx, x, ds, dnx, dx;
			len_sym(len, &s, &nx, &x); dist_sym(dist, &ds, &dnx, &dx);
			if (!ll[s] || !dl[ds]) { bad = 1; break; }
			wput(&w, lc[s], ll[s]); wput(&w, x, nx); wput(&w, dc[ds], dl[ds]); wput(&w, dx, dnx);
			lz[s]++; lz[286 + ds]++;
			for (unsigned k = 0; k < len; k++, p++) if (p + 2 < end) head[H3(p)] = (int32_t)p;
		} else {
			if (!ll[buf[p]]) { bad = 1; break; }
			wput(&w, lc[buf[p]], ll[buf[p]]); lz[buf[p]]++;
			if (p + 2 < end) head[H3(p)] = (int32_t)p;
			p++;
		}
	}
#undef H3
	free(head);
	if (bad || !ll[256]) return 66;
	wput(&w, lc[256], ll[256]); lz[256] = 1;
	if (w.ovf) return 13;
	*tpbc = (uint32_t)((w.bp + 7) >> 3);
	*tebc = (uint32_t)(w.bp & 7);
	return 0;
}

int oracle_nxemu_run_job(void *crb_cpb)
{
	uint8_t *crb = crb_cpb, *cpb = crb + NXGPU_CPB;
	const uint32_t fc = be32(crb + NXGPU_CRB_FC) & 0xff;
	seglist *src = malloc(sizeof(seglist)), *dst = malloc(sizeof(seglist));
	int ret = 0;
	uint8_t *in = NULL, *out = NULL;
	if (dde_segments(crb + NXGPU_CRB_SRC_DDE, src) || dde_segments(crb + NXGPU_CRB_DST_DDE, dst)) {
		complete(crb, 30, 2, 0);                 /* ERR_NX_INVALID_DDE, inc_nx/nxu.h:843 */
		goto done;
	}
	const uint32_t w8 = be32(cpb + 8), w12 = be32(cpb + 12);
	in = malloc(src->total + 64);
	gather(src, in);
	const uint32_t adler_seed = be32(cpb + 0), crc_seed = le32(cpb + 4);

	if (fc == 0x1e) {                                /* wrap */
		if (dst->total < src->total) { complete(crb, 13, 0, 0); goto done; }
		scatter(dst, in, src->total);
		put_be32(cpb + 384, oracle_adler32(1, in, src->total));
		put_le32(cpb + 388, oracle_crc32(0, in, src->total));
		put_be32(cpb + 400, (uint32_t)src->total);
		complete(crb, 0, 0, (uint32_t)src->total);
	} else if ((fc & 0x10) == 0) {                   /* compress */
		const int resume = fc & 0x08, use_dht = fc & 0x02, count = fc & 0x04;
		const uint32_t hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
		if (hist > src->total) { complete(crb, 3, 2, 0); goto done; }
		const uint32_t n = (uint32_t)(src->total - hist);
		uint32_t tpbc = 0, tebc = 0, lz[316];
		out = malloc(dst->total + 64);
		/* only the last 32 KiB of history can be referenced */
		const uint32_t skip = hist > 32768 ? hist - 32768 : 0;
		int rc = compress_block(in + skip, hist - skip, n, !use_dht, cpb + 16, w12 & 0xfff, out, dst->total, &tpbc, &tebc, lz);
		if (rc) { complete(crb, rc, rc == 13 ? 0 : 2, 0); goto done; }
		scatter(dst, out, tpbc);
		put_be32(cpb + 384, oracle_adler32(adler_seed, in + hist, n));
		put_le32(cpb + 388, oracle_crc32(crc_seed, in + hist, n));
		put_be32(cpb + 392, (tebc & 7) << 16);
		if (count) {
			for (int i = 0; i < 316; i++) put_be32(cpb + 400 + 4 * i, lz[i] > 0xffffff ? 0xffffff : lz[i]);
			put_be32(cpb + 1664, (uint32_t)src->total);
		} else {
			put_be32(cpb + 400, (uint32_t)src->total);
		}
		complete(crb, tpbc > src->total ? 64 : 0, 0, tpbc);
	} else {                                         /* decompress */
		const int resume = fc & 0x04;
		const uint32_t hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
		if (hist > src->total) { complete(crb, 3, 2, 0); goto done; }
		out = malloc(hist + dst->total + 64);
		memcpy(out, in, hist);
		oracle_inflate_job j;
		memset(&j, 0, sizeof(j));
		j.src = in + hist; j.src_len = src->total - hist;
		j.start_bit = resume ? (8 - (w8 & 7)) & 7 : 0;
		j.in_sfbt = resume ? (w12 >> 16) & 0xf : 0;
		j.in_rembytecnt = w12 & 0xffff;
		j.in_dht = cpb + 16; j.in_dhtlen = w12 & 0xfff;
		j.dst = out + hist; j.dst_cap = dst->total; j.hist_len = hist;
		j.single_block = (fc & 0x02) != 0;
		oracle_inflate_run(&j);
		if (j.err) { complete(crb, (uint32_t)j.err, j.err == 13 ? 0 : 2, 0); goto done; }
		scatter(dst, out + hist, j.out_len);
		put_be32(cpb + 384, oracle_adler32(adler_seed, out + hist, j.out_len));
		put_le32(cpb + 388, oracle_crc32(crc_seed, out + hist, j.out_len));
		if (j.out_subc > 0xffff) abort();          /* SUBC is a 16-bit field and never wraps (Table 5-3 bounds) */
		put_be32(cpb + 392, j.out_subc);
		const int in_dyn = (j.out_sfbt & 0xe) == 0xc;
		put_be32(cpb + 396, (j.out_sfbt & 0xf) << 16 | (in_dyn ? (j.out_dhtlen & 0xfff) : (j.out_rembytecnt & 0xffff)));
		if (in_dyn) memcpy(cpb + 400, j.out_dht, 288);
		put_be32(cpb + 688, (uint32_t)(hist + j.src_read));
		/* Always CC=3 + "partial completion" (manual Table 5-3 allows it for every row; the one CC=0 row -
		 * final block, no byte of source behind it - is a case the host code does not finish a stream on:
		 * lib/nx_inflate.c:1426-1436 takes no is_final from it) */
		complete(crb, 3, 4 | 1, (uint32_t)j.out_len);
	}
done:
	if (getenv("NXEMU_TRACE"))
		fprintf(stderr, "nxemu: fc %02x src %llu dst %llu in(w8 %08x w12 %08x) -> csb %08x tpbc %u out(w392 %08x w396 %08x) crc %08x adler %08x\n", fc,
			(unsigned long long)src->total, (unsigned long long)dst->total, be32(cpb + 8), be32(cpb + 12), be32(crb + NXGPU_CRB_CSB),
			be32(crb + NXGPU_CRB_CSB + 4), be32(cpb + 392), be32(cpb + 396), le32(cpb + 388), be32(cpb + 384));
	free(in); free(out); free(src); free(dst);
	return ret;
}
