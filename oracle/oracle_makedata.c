/*
 * oracle_makedata.c — restatement of the reference's synthetic-text generator
 * (samples/makedata.c:35-70).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Same lrand48 draw order, so `-s 1 -b 26 < alice29.txt` gives crc32 ece3d95e
 * (BASELINE.md §2).  A draw of dist == 0 makes the reference copy a byte onto
 * itself, i.e. keep whatever malloc returned; for the large mmap-backed buffers
 * it uses that is a zero byte, which is what this restatement writes.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

uint64_t oracle_makedata(int seed, int log2size, const uint8_t *seedfile, uint64_t seedfile_len,
			 uint8_t *out, uint64_t out_cap)
{
	uint64_t bufsz = 1ull << log2size, idx, len_max, dist_max;
	srand48(seed);
	/* samples/makedata.c:38 — both draws are always made */
	{
		long a = lrand48() % 2;
		long b = lrand48() % (long)(bufsz / 10);
		bufsz += (uint64_t)a * (uint64_t)b;
	}
	if (bufsz > out_cap)
		return 0;
	memset(out, 0, bufsz);
	idx = seedfile_len < bufsz / 2 ? seedfile_len : bufsz / 2;   /* :45 */
	memcpy(out, seedfile, idx);
	len_max = (uint64_t)(lrand48() % 240) + 10;                  /* :51 */
	dist_max = (uint64_t)(lrand48() % (1L << 16)) + 1;           /* :52 */
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
