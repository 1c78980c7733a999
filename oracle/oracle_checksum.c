/*
 * oracle_checksum.c — CPU restatement of the reference's checksum arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates, in independent code:
 *   crc32        lib/nx_crc.c:215-243 (byte-table loop of generic_crc32)
 *   crc32 raw    lib/crc32_ppc.c:22-28 + contract of __crc32_vpmsum (:30)
 *   combine      lib/nx_crc.c:347-424  (GF(2) "append len2 zero bytes" operator)
 *   adler32      lib/nx_adler32.c:81-147 (BASE 65521, NMAX 5552 deferred modulo)
 *   combine      lib/nx_adler32.c:154-177
 */
#include "oracle.h"

#define CRC_POLY 0xEDB88320u   /* reflected CRC-32, lib/nx_crc.c:386 */
#define ADLER_BASE 65521u      /* lib/nx_adler32.c:12 */
#define ADLER_NMAX 5552u       /* lib/nx_adler32.c:13 */

static uint32_t crc_tab[256];
static int crc_tab_ready;

static void crc_tab_init(void)
{
	for (uint32_t n = 0; n < 256; n++) {
		uint32_t c = n;
		for (int k = 0; k < 8; k++)
			c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
		crc_tab[n] = c;
	}
	crc_tab_ready = 1;
}

uint32_t oracle_crc32_raw(uint32_t crc, const uint8_t *buf, size_t len)
{
	if (!crc_tab_ready)
		crc_tab_init();
	while (len--)
		crc = crc_tab[(crc ^ *buf++) & 0xff] ^ (crc >> 8);
	return crc;
}

uint32_t oracle_crc32(uint32_t crc, const uint8_t *buf, size_t len)
{
	if (buf == NULL)
		return 0;          /* lib/nx_crc.c:218 */
	return ~oracle_crc32_raw(~crc, buf, len);
}

/* a(x)*b(x) mod P(x), reflected bit order (bit 31 = x^0) */
static uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
	uint32_t m = 1u << 31, p = 0;
	for (;;) {
		if (a & m) {
			p ^= b;
			if ((a & (m - 1)) == 0)
				break;
		}
		m >>= 1;
		b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
	}
	return p;
}

/* x^(8*n) mod P */
static uint32_t gf2_x8n(uint64_t n)
{
	uint32_t p = 1u << 31;          /* x^0 */
	uint32_t sq = 1u << 23;         /* x^8 */
	while (n) {
		if (n & 1)
			p = gf2_mulmod(sq, p);
		sq = gf2_mulmod(sq, sq);
		n >>= 1;
	}
	return p;
}

uint32_t oracle_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
	if (len2 == 0)
		return crc1;       /* lib/nx_crc.c:382-383 */
	return gf2_mulmod(gf2_x8n(len2), crc1) ^ crc2;
}

uint32_t oracle_adler32(uint32_t adler, const uint8_t *buf, size_t len)
{
	uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
	if (buf == NULL)
		return 1;          /* lib/nx_adler32.c:102-103 */
	while (len) {
		size_t n = len < ADLER_NMAX ? len : ADLER_NMAX;
		len -= n;
		while (n--) {
			a += *buf++;
			b += a;
		}
		a %= ADLER_BASE;
		b %= ADLER_BASE;
	}
	return a | (b << 16);
}

uint32_t oracle_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
	uint64_t rem = len2 % ADLER_BASE;
	uint64_t s1 = a1 & 0xffff, s2;
	s2 = (rem * s1) % ADLER_BASE;
	s1 += (a2 & 0xffff) + ADLER_BASE - 1;
	s2 += ((a1 >> 16) & 0xffff) + ((a2 >> 16) & 0xffff) + ADLER_BASE - rem;
	s1 %= ADLER_BASE;
	s2 %= ADLER_BASE;
	return (uint32_t)(s1 | (s2 << 16));
}
