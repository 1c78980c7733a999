/*
 * oracle.h — CPU restatement of the NX-GZIP engine path (SURVEY.md §8a rows a5,
 * a8-a11).  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is linked, imported
 * or executed by the product (power-gzip_b200/, include/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker.
 *
 * Parity status: crc32/adler32/combine are pinned against the reference's own
 * known-answer tests (test/test_crc32.c:38-180, test/test_adler32.c:38-179) and
 * against the reference's nx_crc.c/nx_adler32.c compiled in place
 * (oracle/_ref/libnxz_ref.so).  Inflate is pinned against system zlib 1.3 — the
 * library the reference's software path dlopens (lib/sw_zlib.c:283-327) — and
 * against the reference's 611-byte scp stream (test/test_buf_error.c:107).
 * Deflate *bytes* are unpinned by design: the reference never
 * pins compressed bytes, only round trips.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/nx_crc.c:215-243 (generic_crc32), :437-453 (exported crc32) */
uint32_t oracle_crc32(uint32_t crc, const uint8_t *buf, size_t len);
/* lib/crc32_ppc.c:30 contract of __crc32_vpmsum: no pre/post inversion */
uint32_t oracle_crc32_raw(uint32_t crc, const uint8_t *buf, size_t len);
/* lib/nx_crc.c:374-424 (crc32_combine_) */
uint32_t oracle_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
/* lib/nx_adler32.c:81-147 */
uint32_t oracle_adler32(uint32_t adler, const uint8_t *buf, size_t len);
/* lib/nx_adler32.c:154-177 */
uint32_t oracle_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

/* Raw-deflate decoder restating the NX decompress function (inc_nx/nxu.h:812-815,
 * CPB comments :296-540, manual §5.2.5.5 Table 5-3). */
enum {
	ORA_SFBT_FINAL_EOB = 0x0,
	ORA_SFBT_LIT = 0x8, ORA_SFBT_FHT = 0xa, ORA_SFBT_DHT = 0xc, ORA_SFBT_HDR = 0xe
};
#define ORA_READ_AHEAD 8
typedef struct {
	/* in */
	const uint8_t *src; size_t src_len;     /* compressed bytes (no history)        */
	unsigned start_bit;                     /* bits of src[0] already consumed (0-7) */
	unsigned in_sfbt;                       /* 0 (= fresh, block header next) or 8..15 */
	unsigned in_rembytecnt;                 /* for sfbt 100x                        */
	const uint8_t *in_dht; unsigned in_dhtlen; /* for sfbt 110x: header bits from HLIT */
	uint8_t *dst; size_t dst_cap;           /* dst[-hist_len .. -1] is the window    */
	size_t hist_len;
	int single_block;                       /* stop after one block (FC 0x12/0x16)   */
	/* out */
	size_t out_len;                         /* tpbc                                  */
	size_t src_read;                        /* source bytes the engine read (spbc - history): all of it when
	                                           the source ran out, else at most ORA_READ_AHEAD bytes behind
	                                           the last processed bit                                     */
	uint64_t bits_used;                     /* from bit 0 of src[0]                  */
	unsigned out_sfbt, out_subc, out_rembytecnt;
	uint8_t out_dht[288]; unsigned out_dhtlen;
	int final_seen;
	int err;                                /* 0 ok, 13 target full, 66/67/68 data   */
} oracle_inflate_job;
int oracle_inflate_run(oracle_inflate_job *j);

/* whole member: wrap 0 raw / 1 zlib / 2 gzip / 3 auto.  Returns 0, -3 data error,
 * -5 output too small.  *in_used counts header+trailer too. */
int oracle_inflate_member(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
			  int wrap, size_t *out_len, size_t *in_used, uint32_t *crc, uint32_t *adler);

/* Length-limited Huffman code lengths (<= maxbits) for n symbol counts, in the
 * spirit of lib/nx_dhtgen.c:418-595; used to sanity-check GPU-built tables
 * (Kraft equality, optimal cost within a bound). */
void oracle_huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len);
/* cost in bits of one dynamic block given lit/len (286) and dist (30) counts,
 * including the RFC 1951 §3.2.7 header built like lib/nx_dhtgen.c:709-915 */
uint64_t oracle_dynblock_bits(const uint32_t *ll, const uint32_t *d);

/* samples/makedata.c:35-70 restated: same bytes as `makedata -s seed -b log2`. */
uint64_t oracle_makedata(int seed, int log2size, const uint8_t *seedfile, uint64_t seedfile_len,
			 uint8_t *out, uint64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif
