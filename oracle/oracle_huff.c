/*
 * oracle_huff.c — CPU restatement of the reference's dynamic-Huffman-table
 * generation.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows lib/nx_dhtgen.c: sort symbols by count and merge with two queues
 * (huffman_tree, :418-571), retry with flattened counts until no code is longer
 * than the limit (huffmanize, :576-595; count scaling length_limit, :295), then
 * cost the RFC 1951 §3.2.7 header with run-length symbols 16/17/18
 * (encode_lengths, :709-915).  Unlike the reference, which hard-codes the
 * code-length code (:610-653), the header here uses a real Huffman code for the
 * 19 code-length symbols, because that is what the GPU kernel emits.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct { uint32_t f; int s; } leaf_t;

static int leaf_cmp(const void *a, const void *b)
{
	const leaf_t *x = a, *y = b;
	if (x->f != y->f)
		return x->f < y->f ? -1 : 1;
	return x->s - y->s;
}

void oracle_huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len)
{
	uint32_t f[320];
	for (int i = 0; i < n; i++)
		f[i] = freq[i];
	for (;;) {
		leaf_t leaves[320];
		uint64_t w[640];
		int parent[640], depth[640];
		int nl = 0, tot, q1 = 0, q2, mx = 0;
		for (int i = 0; i < n; i++) {
			len[i] = 0;
			if (f[i]) { leaves[nl].f = f[i]; leaves[nl].s = i; nl++; }
		}
		if (nl == 0)
			return;
		if (nl == 1) { len[leaves[0].s] = 1; return; }
		qsort(leaves, nl, sizeof(leaf_t), leaf_cmp);
		for (int i = 0; i < nl; i++)
			w[i] = leaves[i].f;
		tot = q2 = nl;
		while ((nl - q1) + (tot - q2) > 1) {
			int a, b;
			if (q1 < nl && (q2 >= tot || w[q1] <= w[q2])) a = q1++; else a = q2++;
			if (q1 < nl && (q2 >= tot || w[q1] <= w[q2])) b = q1++; else b = q2++;
			w[tot] = w[a] + w[b];
			parent[a] = parent[b] = tot;
			tot++;
		}
		depth[tot - 1] = 0;
		for (int i = tot - 2; i >= 0; i--)
			depth[i] = depth[parent[i]] + 1;
		for (int i = 0; i < nl; i++) {
			len[leaves[i].s] = (uint8_t)depth[i];
			if (depth[i] > mx) mx = depth[i];
		}
		if (mx <= maxbits)
			return;
		for (int i = 0; i < n; i++)
			if (f[i]) f[i] = (f[i] + 1) / 2;
	}
}

static const uint8_t lext[29] = { 0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
static const uint8_t dext[30] = { 0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13 };

uint64_t oracle_dynblock_bits(const uint32_t *ll, const uint32_t *d)
{
	static const int ord[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	uint8_t lll[286], dl[30], seq[320], cl[19];
	uint32_t cf[19] = { 0 };
	int hlit = 286, hdist = 30, hclen = 19, n = 0, extra = 0;
	uint64_t bits;
	oracle_huff_lengths(ll, 286, 15, lll);
	oracle_huff_lengths(d, 30, 15, dl);
	while (hlit > 257 && lll[hlit - 1] == 0) hlit--;
	while (hdist > 1 && dl[hdist - 1] == 0) hdist--;
	for (int i = 0; i < hlit; i++) seq[n++] = lll[i];
	for (int i = 0; i < hdist; i++) seq[n++] = dl[i];
	for (int i = 0; i < n;) {
		int k = i, run, v = seq[i];
		while (k < n && seq[k] == v) k++;
		run = k - i;
		if (v == 0) {
			while (run >= 11) { int r = run > 138 ? 138 : run; cf[18]++; extra += 7; run -= r; }
			if (run >= 3) { cf[17]++; extra += 3; run = 0; }
			cf[0] += run;
		} else {
			cf[v]++; run--;
			while (run >= 3) { int r = run > 6 ? 6 : run; cf[16]++; extra += 2; run -= r; }
			cf[v] += run;
		}
		i = k;
	}
	oracle_huff_lengths(cf, 19, 7, cl);
	while (hclen > 4 && cl[ord[hclen - 1]] == 0) hclen--;
	bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)hclen + extra;
	for (int i = 0; i < 19; i++) bits += (uint64_t)cf[i] * cl[i];
	for (int i = 0; i < 286; i++) bits += (uint64_t)ll[i] * (lll[i] + (i > 256 ? lext[i - 257] : 0));
	for (int i = 0; i < 30; i++) bits += (uint64_t)d[i] * (dl[i] + dext[i]);
	return bits;
}
