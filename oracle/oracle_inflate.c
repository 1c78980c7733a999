/*
 * oracle_inflate.c — CPU restatement of the NX decompress function (SURVEY.md §8a
 * row a8): raw-deflate decoding that can start and stop at any symbol boundary
 * and reports SFBT / SUBC / rembytecnt / DHT exactly as the reference's host code
 * consumes them (lib/nx_inflate.c:1372-1609; field semantics inc_nx/nxu.h:296-540;
 * manual Table 5-3).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Whole-member wrappers follow the header/trailer rules the reference's host side
 * implements in lib/nx_inflate.c:329-730 (gzip/zlib header FSM) and :763-848
 * (trailer check).  The decode arithmetic itself (RFC 1951) lives, for the
 * reference's software path, in system zlib (lib/sw_zlib.c:80-276); this file is
 * checked against zlib 1.3 in tests/test_oracle.py.
 */
#include <string.h>
#include "oracle.h"

typedef struct {
	const uint8_t *p;
	uint64_t nbits;     /* total bits available */
	uint64_t bp;        /* next bit */
} bitrd;

/* returns -1 on underflow (nothing consumed) */
static int64_t br_get(bitrd *b, unsigned n)
{
	if (b->bp + n > b->nbits)
		return -1;
	uint64_t v = 0;
	for (unsigned i = 0; i < n; i++, b->bp++)
		v |= (uint64_t)((b->p[b->bp >> 3] >> (b->bp & 7)) & 1) << i;
	return (int64_t)v;
}

typedef struct {
	uint16_t count[16];
	uint16_t symbol[288];
} hufftab;

/* 0 ok, 1 incomplete, -1 over-subscribed */
static int huff_build(hufftab *h, const uint8_t *len, int n)
{
	uint16_t offs[16];
	int left = 1;
	memset(h->count, 0, sizeof(h->count));
	for (int i = 0; i < n; i++)
		h->count[len[i]]++;
	if (h->count[0] == n)
		return 1;
	for (int l = 1; l <= 15; l++) {
		left <<= 1;
		left -= h->count[l];
		if (left < 0)
			return -1;
	}
	offs[1] = 0;
	for (int l = 1; l < 15; l++)
		offs[l + 1] = offs[l] + h->count[l];
	for (int i = 0; i < n; i++)
		if (len[i])
			h->symbol[offs[len[i]]++] = (uint16_t)i;
	return left ? 1 : 0;
}

/* -1 underflow, -2 no such code, else symbol */
static int huff_decode(bitrd *b, const hufftab *h)
{
	int code = 0, first = 0, index = 0;
	uint64_t save = b->bp;
	for (int l = 1; l <= 15; l++) {
		int64_t bit = br_get(b, 1);
		if (bit < 0) {
			b->bp = save;
			return -1;
		}
		code |= (int)bit;
		int cnt = h->count[l];
		if (code - cnt < first)
			return h->symbol[index + (code - first)];
		index += cnt;
		first += cnt;
		first <<= 1;
		code <<= 1;
	}
	return -2;
}

static const uint16_t len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
	35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
	3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129,
	193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
	7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
static const uint8_t clen_order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

static void fixed_tables(hufftab *ll, hufftab *d)
{
	uint8_t l[288];
	int i = 0;
	for (; i < 144; i++) l[i] = 8;
	for (; i < 256; i++) l[i] = 9;
	for (; i < 280; i++) l[i] = 7;
	for (; i < 288; i++) l[i] = 8;
	huff_build(ll, l, 288);
	for (i = 0; i < 30; i++) l[i] = 5;
	huff_build(d, l, 30);
}

/* Parse the dynamic header starting at HLIT.  0 ok, -1 underflow, >0 NX error code */
static int dyn_tables(bitrd *b, hufftab *ll, hufftab *d)
{
	uint8_t lens[320], cl[19];
	hufftab ch;
	int64_t v;
	if ((v = br_get(b, 14)) < 0)
		return -1;
	int hlit = (int)(v & 31) + 257, hdist = (int)((v >> 5) & 31) + 1, hclen = (int)(v >> 10) + 4;
	if (hlit > 286 || hdist > 30)
		return 68;
	memset(cl, 0, sizeof(cl));
	for (int i = 0; i < hclen; i++) {
		if ((v = br_get(b, 3)) < 0)
			return -1;
		cl[clen_order[i]] = (uint8_t)v;
	}
	if (huff_build(&ch, cl, 19) < 0)
		return 68;
	int n = 0;
	while (n < hlit + hdist) {
		int sym = huff_decode(b, &ch);
		if (sym == -1)
			return -1;
		if (sym < 0)
			return 68;
		if (sym < 16) {
			lens[n++] = (uint8_t)sym;
			continue;
		}
		int rep, val = 0;
		if (sym == 16) {
			if (n == 0)
				return 68;
			val = lens[n - 1];
			if ((v = br_get(b, 2)) < 0) return -1;
			rep = 3 + (int)v;
		} else if (sym == 17) {
			if ((v = br_get(b, 3)) < 0) return -1;
			rep = 3 + (int)v;
		} else {
			if ((v = br_get(b, 7)) < 0) return -1;
			rep = 11 + (int)v;
		}
		if (n + rep > hlit + hdist)
			return 68;
		while (rep--)
			lens[n++] = (uint8_t)val;
	}
	if (lens[256] == 0)
		return 68;
	if (huff_build(ll, lens, hlit) < 0)
		return 68;
	if (huff_build(d, lens + hlit, hdist) < 0)
		return 68;
	return 0;
}

static void copy_bits(uint8_t *dst, const uint8_t *src, uint64_t from, uint64_t nbits)
{
	memset(dst, 0, 288);
	for (uint64_t i = 0; i < nbits && i < 288 * 8; i++) {
		uint64_t s = from + i;
		if ((src[s >> 3] >> (s & 7)) & 1)
			dst[i >> 3] |= (uint8_t)(1u << (i & 7));
	}
}

int oracle_inflate_run(oracle_inflate_job *j)
{
	bitrd b = { j->src, (uint64_t)j->src_len * 8, j->start_bit };
	hufftab ll, dd;
	size_t out = 0;
	unsigned bfinal = 0, kind = 0;    /* kind: 0 header next, 1 stored, 2 fixed, 3 dynamic */
	unsigned rem = 0;
	uint64_t dht_from = 0, dht_bits = 0;
	const uint8_t *dht_src = j->src;
	int rc;

	j->out_len = 0; j->bits_used = b.bp; j->out_sfbt = 0; j->out_subc = 0; j->out_rembytecnt = 0;
	j->out_dhtlen = 0; j->final_seen = 0; j->err = 0; j->src_read = j->src_len;

	if (b.bp > b.nbits) { j->err = 3; return 3; }

	switch (j->in_sfbt & 0xe) {
	case ORA_SFBT_LIT: kind = 1; bfinal = j->in_sfbt & 1; rem = j->in_rembytecnt; break;
	case ORA_SFBT_FHT: kind = 2; bfinal = j->in_sfbt & 1; fixed_tables(&ll, &dd); break;
	case ORA_SFBT_DHT: {
		bitrd t = { j->in_dht, j->in_dhtlen, 0 };
		kind = 3; bfinal = j->in_sfbt & 1;
		rc = dyn_tables(&t, &ll, &dd);
		if (rc) { j->err = 68; return 68; }
		dht_src = j->in_dht; dht_from = 0; dht_bits = j->in_dhtlen;
		break;
	}
	default: kind = 0; break;
	}

	for (;;) {
		if (kind == 0) {
			uint64_t blk = b.bp;
			int64_t v = br_get(&b, 3);
			if (v < 0) {
				/* manual Table 5-3: 1110/1111 by the first header bit if we have it */
				unsigned f = 0;
				if (b.nbits > blk)
					f = (b.p[blk >> 3] >> (blk & 7)) & 1;
				j->out_sfbt = ORA_SFBT_HDR | f;
				j->out_subc = (unsigned)(b.nbits - blk);
				break;
			}
			bfinal = (unsigned)v & 1;
			unsigned btype = (unsigned)v >> 1;
			if (btype == 0) {
				b.bp = (b.bp + 7) & ~7ull;
				if (b.bp > b.nbits) b.bp = b.nbits;
				v = br_get(&b, 32);
				if (v < 0) {
					b.bp = blk;
					j->out_sfbt = ORA_SFBT_HDR | bfinal;
					j->out_subc = (unsigned)(b.nbits - blk);
					break;
				}
				if (((v ^ (v >> 16)) & 0xffff) != 0xffff) { j->err = 68; break; }
				rem = (unsigned)v & 0xffff;
				kind = 1;
			} else if (btype == 1) {
				fixed_tables(&ll, &dd);
				kind = 2;
			} else if (btype == 2) {
				uint64_t from = b.bp;
				rc = dyn_tables(&b, &ll, &dd);
				if (rc < 0) {
					b.bp = blk;
					j->out_sfbt = ORA_SFBT_HDR | bfinal;
					j->out_subc = (unsigned)(b.nbits - blk);
					break;
				}
				if (rc) { j->err = rc; break; }
				dht_src = j->src; dht_from = from; dht_bits = b.bp - from;
				kind = 3;
			} else {
				j->err = 68;
				break;
			}
		}
		if (kind == 1) {
			/* byte aligned here */
			size_t avail = (size_t)((b.nbits - b.bp) >> 3);
			size_t n = rem < avail ? rem : avail;
			if (n > j->dst_cap - out) { j->err = 13; break; }
			memcpy(j->dst + out, b.p + (b.bp >> 3), n);
			out += n; b.bp += (uint64_t)n * 8; rem -= (unsigned)n;
			if (rem) {
				j->out_sfbt = ORA_SFBT_LIT | bfinal;
				j->out_rembytecnt = rem;
				j->out_subc = (unsigned)(b.nbits - b.bp);
				break;
			}
		} else {
			int stop = 0;
			for (;;) {
				uint64_t sym_at = b.bp;
				int sym = huff_decode(&b, &ll);
				if (sym == -1) { stop = 1; }
				else if (sym < 0) { j->err = 66; break; }
				else if (sym < 256) {
					if (out >= j->dst_cap) { j->err = 13; break; }
					j->dst[out++] = (uint8_t)sym;
					continue;
				} else if (sym == 256) {
					break;
				} else {
					int64_t e;
					sym -= 257;
					if (sym >= 29) { j->err = 66; break; }
					if ((e = br_get(&b, len_extra[sym])) < 0) stop = 1;
					else {
						unsigned len = len_base[sym] + (unsigned)e;
						int ds = huff_decode(&b, &dd);
						if (ds == -1) stop = 1;
						else if (ds < 0 || ds >= 30) { j->err = 66; break; }
						else if ((e = br_get(&b, dist_extra[ds])) < 0) stop = 1;
						else {
							size_t dist = dist_base[ds] + (size_t)e;
							if (dist > out + j->hist_len) { j->err = 67; break; }
							if (len > j->dst_cap - out) { j->err = 13; break; }
							for (unsigned k = 0; k < len; k++, out++)
								j->dst[out] = *(j->dst + out - dist);
							continue;
						}
					}
				}
				if (stop) {
					b.bp = sym_at;
					break;
				}
			}
			if (j->err)
				break;
			if (stop) {
				j->out_sfbt = (kind == 2 ? ORA_SFBT_FHT : ORA_SFBT_DHT) | bfinal;
				j->out_subc = (unsigned)(b.nbits - b.bp);
				if (kind == 3) {
					copy_bits(j->out_dht, dht_src, dht_from, dht_bits);
					j->out_dhtlen = (unsigned)dht_bits;
				}
				break;
			}
		}
		/* end of block */
		kind = 0;
		if (bfinal || j->single_block || b.bp == b.nbits) {
			/* The engine stops by itself here, with source possibly left.  The manual (§2.4): "SPBC
			 * indicates the number of compressed source bytes READ by the accelerator, SUBC the number of
			 * source bits that the accelerator discarded because they were past the stream end" - a gzip
			 * trailer gives SUBC 64..71, a zlib trailer 32..39 (inc_nx/nxu.h:454-465).  Model: the engine
			 * has read at most 8 bytes behind the byte that holds the last processed bit; the host
			 * computes the stream end as spbc - histlen - subc/8 (lib/nx_inflate.c:1452-1472). */
			uint64_t end_byte = (b.bp + 7) >> 3, rd = end_byte + ORA_READ_AHEAD;
			if (rd > j->src_len) rd = j->src_len;
			j->src_read = (size_t)rd;
			j->out_subc = (unsigned)(rd * 8 - b.bp);
			j->final_seen = bfinal ? 1 : 0;
			/* Table 5-3: 0000 final EOB; 1110 a block with BFINAL=0 ended (single-block suspend, or the
			 * source ended exactly on the block boundary) */
			j->out_sfbt = bfinal ? ORA_SFBT_FINAL_EOB : ORA_SFBT_HDR;
			break;
		}
	}
	j->out_len = out;
	j->bits_used = b.bp;
	return j->err;
}

int oracle_inflate_member(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
			  int wrap, size_t *out_len, size_t *in_used, uint32_t *crc, uint32_t *adler)
{
	size_t pos = 0;
	if (wrap == 3) {
		if (src_len >= 2 && src[0] == 0x1f && src[1] == 0x8b) wrap = 2;
		else if (src_len >= 2 && (src[0] & 0x0f) == 8 && ((src[0] << 8 | src[1]) % 31) == 0) wrap = 1;
		else wrap = 0;
	}
	if (wrap == 2) {
		if (src_len < 18 || src[0] != 0x1f || src[1] != 0x8b || src[2] != 8) return -3;
		unsigned flg = src[3];
		pos = 10;
		if (flg & 4) {
			if (pos + 2 > src_len) return -3;
			pos += 2 + (src[pos] | src[pos + 1] << 8);
		}
		if (flg & 8) { while (pos < src_len && src[pos]) pos++; pos++; }
		if (flg & 16) { while (pos < src_len && src[pos]) pos++; pos++; }
		if (flg & 2) pos += 2;
		if (pos > src_len) return -3;
	} else if (wrap == 1) {
		if (src_len < 6 || (src[0] & 0x0f) != 8 || ((src[0] << 8 | src[1]) % 31) || (src[1] & 0x20)) return -3;
		pos = 2;
	}
	oracle_inflate_job j;
	memset(&j, 0, sizeof(j));
	j.src = src + pos; j.src_len = src_len - pos; j.dst = dst; j.dst_cap = dst_cap;
	int err = oracle_inflate_run(&j);
	if (out_len) *out_len = j.out_len;
	if (err == 13) return -5;
	if (err || !j.final_seen) return -3;
	pos += (size_t)((j.bits_used + 7) >> 3);
	uint32_t c = oracle_crc32(0, dst, j.out_len), a = oracle_adler32(1, dst, j.out_len);
	if (crc) *crc = c;
	if (adler) *adler = a;
	if (wrap == 2) {
		if (pos + 8 > src_len) return -3;
		uint32_t tc = src[pos] | src[pos + 1] << 8 | src[pos + 2] << 16 | (uint32_t)src[pos + 3] << 24;
		uint32_t ts = src[pos + 4] | src[pos + 5] << 8 | src[pos + 6] << 16 | (uint32_t)src[pos + 7] << 24;
		if (tc != c || ts != (uint32_t)j.out_len) return -3;
		pos += 8;
	} else if (wrap == 1) {
		if (pos + 4 > src_len) return -3;
		uint32_t ta = (uint32_t)src[pos] << 24 | src[pos + 1] << 16 | src[pos + 2] << 8 | src[pos + 3];
		if (ta != a) return -3;
		pos += 4;
	}
	if (in_used) *in_used = pos;
	return 0;
}
