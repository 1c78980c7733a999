// nxgpu_job.cu — interprets one NX job descriptor (nx_gzip_crb_cpb_t, reference
// inc_nx/nxu.h:286-616) and runs it on the GPU: the body of nxu_run_job (lib/gzip_vas.c:281).
//
// The host code above this boundary (lib/nx_deflate.c, lib/nx_inflate.c, lib/nx_zlib.c) fills the
// descriptor with big-endian fields and HOST addresses, calls nxu_run_job and reads the completion
// status block (CSB) and the output half of the parameter block (CPB) afterwards.  This file only
// moves bytes and fields: every codec and checksum step is a kernel launch (deflate.cu, inflate.cu,
// checksum.cu).  Field offsets are the NXGPU_* constants of include/nxgpu.h, asserted against the
// reference's own headers by oracle/layout_check.c.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ctx.cuh"

namespace nxgpu {
namespace {

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline uint64_t be64(const uint8_t *p) { return (uint64_t)be32(p) << 32 | be32(p + 4); }
inline void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]; }
inline void put_le32(uint8_t *p, uint32_t v) { p[3] = (uint8_t)(v >> 24); p[2] = (uint8_t)(v >> 16); p[1] = (uint8_t)(v >> 8); p[0] = (uint8_t)v; }

struct Seg { uint8_t *p; uint32_t len; };

// A data descriptor element is direct (count 0: address + byte count) or points to a list of
// direct ones; of an indirect list only the first `ddebc` bytes count (inc_nx/nxu.h:155-170,
// lib/nx_deflate.c:1248-1251).
bool dde_segments(const uint8_t *dde, std::vector<Seg> &out, uint64_t &total)
{
	const uint32_t count = (be32(dde) >> 8) & 0xff;
	const uint32_t bc = be32(dde + 4);
	const uint64_t addr = be64(dde + 8);
	total = 0;
	if (count == 0) {
		if (bc)
			out.push_back({ reinterpret_cast<uint8_t *>(addr), bc });
		total = bc;
		return true;
	}
	const uint8_t *list = reinterpret_cast<const uint8_t *>(addr);
	uint64_t left = bc;
	for (uint32_t i = 0; i < count && left; i++) {
		const uint8_t *d = list + 16 * i;
		if ((be32(d) >> 8) & 0xff)
			return false;                                   // only one level of indirection
		const uint64_t l = be32(d + 4);
		const uint32_t use = (uint32_t)(l < left ? l : left);
		if (use)
			out.push_back({ reinterpret_cast<uint8_t *>(be64(d + 8)), use });
		left -= use;
		total += use;
	}
	return true;
}

// LSB-first bit reader over a small host buffer (the dynamic header the caller supplies)
struct HostBits {
	const uint8_t *p; uint32_t nbits, bp;
	int get(uint32_t n)
	{
		if (bp + n > nbits) return -1;
		uint32_t v = 0;
		for (uint32_t i = 0; i < n; i++, bp++)
			v |= (uint32_t)((p[bp >> 3] >> (bp & 7)) & 1) << i;
		return (int)v;
	}
};

// The caller's dynamic Huffman table (cpb.in_dht, written by lib/nx_dhtgen.c:709-915 or copied
// from lib/nx_dht_builtin.c) is the RFC 1951 §3.2.7 block header from HLIT on.  The deflate kernel
// wants the 286 + 30 code lengths next to the raw bits; this only re-reads the header, it builds
// nothing.
bool dht_to_lengths(const uint8_t *bits, uint32_t nbits, uint8_t *lens)
{
	static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	HostBits b = { bits, nbits, 0 };
	int v = b.get(14);
	if (v < 0) return false;
	const int hlit = (v & 31) + 257, hdist = ((v >> 5) & 31) + 1, hclen = (v >> 10) + 4;
	if (hlit > 286 || hdist > 30) return false;
	uint8_t cl[19] = { 0 };
	for (int i = 0; i < hclen; i++) {
		if ((v = b.get(3)) < 0) return false;
		cl[order[i]] = (uint8_t)v;
	}
	// canonical code of the code-length alphabet
	uint16_t count[8] = { 0 }, code[19];
	for (int i = 0; i < 19; i++) count[cl[i]]++;
	count[0] = 0;
	uint16_t c = 0;
	uint16_t first[8] = { 0 };
	for (int l = 1; l <= 7; l++) { c = (uint16_t)((c + count[l - 1]) << 1); first[l] = c; }
	for (int i = 0; i < 19; i++) code[i] = cl[i] ? first[cl[i]]++ : 0;
	uint8_t all[320];
	int n = 0;
	while (n < hlit + hdist) {
		// decode one code-length symbol bit by bit (MSB of the code first)
		int sym = -1;
		uint32_t acc = 0;
		for (int l = 1; l <= 7 && sym < 0; l++) {
			if ((v = b.get(1)) < 0) return false;
			acc = (acc << 1) | (uint32_t)v;
			for (int i = 0; i < 19; i++)
				if (cl[i] == l && code[i] == acc) { sym = i; break; }
		}
		if (sym < 0) return false;
		if (sym < 16) { all[n++] = (uint8_t)sym; continue; }
		int rep, val = 0;
		if (sym == 16) { if (n == 0 || (v = b.get(2)) < 0) return false; val = all[n - 1]; rep = 3 + v; }
		else if (sym == 17) { if ((v = b.get(3)) < 0) return false; rep = 3 + v; }
		else { if ((v = b.get(7)) < 0) return false; rep = 11 + v; }
		if (n + rep > hlit + hdist) return false;
		while (rep--) all[n++] = (uint8_t)val;
	}
	memset(lens, 0, 316);
	memcpy(lens, all, hlit);
	memcpy(lens + 286, all + hlit, hdist);
	return true;
}

void complete(uint8_t *c, uint32_t cc, uint32_t ce_ms3b, uint32_t tpbc)
{
	uint8_t *csb = c + NXGPU_CRB_CSB;
	put_be32(csb + 4, tpbc);
	// V (bit 0), CC (bits 16:23), CE (bits 24:31, only its three most significant bits are defined)
	put_be32(csb, 0x80000000u | (cc & 0xff) << 8 | ((ce_ms3b & 7) << 5));
}

bool gather(const std::vector<Seg> &segs, uint8_t *dst)
{
	for (const Seg &s : segs) {
		memcpy(dst, s.p, s.len);
		dst += s.len;
	}
	return true;
}
void scatter(const std::vector<Seg> &segs, const uint8_t *src, uint64_t n)
{
	for (const Seg &s : segs) {
		if (!n) break;
		const uint32_t l = (uint32_t)(s.len < n ? s.len : n);
		memcpy(s.p, src, l);
		src += l;
		n -= l;
	}
}

inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

int job_level()
{
	static int lv = [] { const char *e = getenv("NXGPU_JOB_LEVEL"); int v = e ? atoi(e) : 6; return v < 1 ? 1 : v > 9 ? 9 : v; }();
	return lv;
}

constexpr uint32_t CE_PARTIAL = 0x4, CE_TERMINATE = 0x2, CE_TPBC_VALID = 0x1;

} // namespace

static int run_job_inner(nxgpu_ctx *c, uint8_t *crb);

// returns 0 (CSB/CPB filled) or -EAGAIN when the device cannot be used
int run_job_impl(nxgpu_ctx *c, uint8_t *crb)
{
	const int rc = run_job_inner(c, crb);
	static const bool trace = getenv("NXGPU_TRACE") != nullptr;
	if (trace) {
		const uint8_t *cpb = crb + NXGPU_CPB;
		fprintf(stderr, "nxgpu job: fc %02x src bc %u dst bc %u in(w8 %08x w12 %08x) -> rc %d csb %08x tpbc %u out(w392 %08x w396 %08x)\n",
			be32(crb) & 0xff, be32(crb + NXGPU_CRB_SRC_DDE + 4), be32(crb + NXGPU_CRB_DST_DDE + 4), be32(cpb + 8), be32(cpb + 12), rc,
			be32(crb + NXGPU_CRB_CSB), be32(crb + NXGPU_CRB_CSB + 4), be32(cpb + 392), be32(cpb + 396));
	}
	return rc;
}

static int run_job_inner(nxgpu_ctx *c, uint8_t *crb)
{
	uint8_t *cpb = crb + NXGPU_CPB;
	const uint32_t fc = be32(crb + NXGPU_CRB_FC) & 0xff;
	std::vector<Seg> src, dst;
	uint64_t src_total = 0, dst_total = 0;
	if (!dde_segments(crb + NXGPU_CRB_SRC_DDE, src, src_total) || !dde_segments(crb + NXGPU_CRB_DST_DDE, dst, dst_total)) {
		complete(crb, 9 /* ERR_NX_BAD_DDE */, CE_TERMINATE, 0);
		return 0;
	}
	if (src_total > 0xffffff00ull || dst_total > 0xffffff00ull) {
		complete(crb, 3, CE_TERMINATE, 0);
		return 0;
	}
	if (cudaSetDevice(c->dev) != cudaSuccess)
		return -EAGAIN;
	const bool is_compress = (fc & 0x10) == 0;
	const bool is_wrap = fc == 0x1e;
	const uint32_t w8 = be32(cpb + 8), w12 = be32(cpb + 12);

	// ---- source into pinned staging (history first, exactly as the DDE list has it) ----
	if (c->h_stage.reserve(src_total + 64)) return -EAGAIN;
	uint8_t *hs = static_cast<uint8_t *>(c->h_stage.p);
	gather(src, hs);

	if (is_wrap) {
		// GZIP_FC_WRAP (inc_nx/nxu.h:816; caller lib/nx_zlib.c:1398): copy + fresh crc32/adler32
		if (dst_total < src_total) { complete(crb, 13, 0, 0); return 0; }
		if (c->d_in.reserve(src_total + 16)) return -EAGAIN;
		if (cudaMemcpyAsync(c->d_in.p, hs, src_total, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
		nxgpu_cksum_item it = { c->d_in.p, src_total, 0, 1 };
		if (checksum_device(c, &it, 1, 3)) return -EAGAIN;
		if (c->h_outs.reserve(src_total + 64)) return -EAGAIN;
		uint8_t *ho = static_cast<uint8_t *>(c->h_outs.p);
		if (cudaMemcpyAsync(ho, c->d_cks.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (cudaMemcpyAsync(ho + 16, c->d_in.p, src_total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -EAGAIN;
		scatter(dst, ho + 16, src_total);
		const uint32_t *ck = reinterpret_cast<const uint32_t *>(ho);
		put_be32(cpb + NXGPU_CPB_OUT_ADLER - NXGPU_CPB, ck[1]);
		put_le32(cpb + NXGPU_CPB_OUT_CRC - NXGPU_CPB, ck[0]);
		put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP - NXGPU_CPB, (uint32_t)src_total);
		complete(crb, 0, 0, (uint32_t)src_total);
		return 0;
	}

	// running checksums: in_adler is a plain big-endian field, in_crc holds the CRC with its
	// bytes the other way round (lib/nx_deflate.c:1572-1577 and lib/nx_inflate.c:809-817 rely on it)
	const uint32_t adler_seed = be32(cpb + NXGPU_CPB_IN_ADLER - NXGPU_CPB);
	const uint32_t crc_seed = le32(cpb + NXGPU_CPB_IN_CRC - NXGPU_CPB);

	if (is_compress) {
		const bool resume = (fc & 0x08) != 0, use_dht = (fc & 0x02) != 0, count = (fc & 0x04) != 0;
		const uint32_t hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
		if (hist > src_total) { complete(crb, 3, CE_TERMINATE, 0); return 0; }   // history length error
		const uint32_t n_new = (uint32_t)(src_total - hist);
		if (c->d_in.reserve(src_total + 32)) return -EAGAIN;
		if (src_total && cudaMemcpyAsync(c->d_in.p, hs, src_total, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
		DeflateJob job;
		memset(&job, 0, sizeof(job));
		job.src = static_cast<const uint8_t *>(c->d_in.p) + hist;
		job.src_len = n_new;
		job.hist_len = hist > 32768 ? 32768 : hist;
		job.flags = NXGPU_F_NO_JOINER | (use_dht ? 0 : NXGPU_F_FIXED);
		if (use_dht) {
			const uint32_t dhtlen = w12 & 0xfff;
			uint8_t blob[320 + 288];
			memset(blob, 0, sizeof(blob));
			if (dhtlen < 42 || dhtlen > 288 * 8 || !dht_to_lengths(cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, dhtlen, blob)) {
				complete(crb, 68 /* invalid DHT */, CE_TERMINATE, 0);
				return 0;
			}
			memcpy(blob + 320, cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, (dhtlen + 7) / 8);
			if (c->d_dht.reserve(1024)) return -EAGAIN;
			if (cudaMemcpyAsync(c->d_dht.p, blob, sizeof(blob), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
			job.dht = static_cast<const uint8_t *>(c->d_dht.p);
			job.dht_bits = dhtlen;
		}
		if (count) {
			if (c->d_lz.reserve(316 * 4)) return -EAGAIN;
			job.lzcount = static_cast<uint32_t *>(c->d_lz.p);
		}
		if (deflate_device(c, &job, 1, job_level(), false)) return -EAGAIN;
		nxgpu_cksum_item it = { job.src, n_new, crc_seed, adler_seed };
		if (checksum_device(c, &it, 1, 3)) return -EAGAIN;
		if (c->h_outs.reserve(sizeof(DeflateOut) + 64 + 316 * 4 + 2 * (size_t)n_new + 2048)) return -EAGAIN;
		uint8_t *ho = static_cast<uint8_t *>(c->h_outs.p);
		DeflateOut *o = reinterpret_cast<DeflateOut *>(ho);
		uint32_t *ck = reinterpret_cast<uint32_t *>(ho + sizeof(DeflateOut));
		uint32_t *lz = reinterpret_cast<uint32_t *>(ho + sizeof(DeflateOut) + 64);
		uint8_t *data = ho + sizeof(DeflateOut) + 64 + 316 * 4;
		if (cudaMemcpyAsync(o, c->d_outs.p, sizeof(DeflateOut), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (cudaMemcpyAsync(ck, c->d_cks.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (count && cudaMemcpyAsync(lz, c->d_lz.p, 316 * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -EAGAIN;
		if (o->rc == 66) { complete(crb, 66, CE_TERMINATE, 0); return 0; }            // a needed symbol has no code
		if (o->rc != 0) { complete(crb, 13, 0, 0); return 0; }
		if (o->out_len > dst_total) { complete(crb, 13, 0, 0); return 0; }           // ERR_NX_TARGET_SPACE: caller halves the input
		if (o->out_len && cudaMemcpyAsync(data, job.out, o->out_len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
		if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -EAGAIN;
		scatter(dst, data, o->out_len);
		put_be32(cpb + NXGPU_CPB_OUT_ADLER - NXGPU_CPB, ck[1]);
		put_le32(cpb + NXGPU_CPB_OUT_CRC - NXGPU_CPB, ck[0]);
		put_be32(cpb + NXGPU_CPB_OUT_TEBC - NXGPU_CPB, (o->tebc & 7) << 16);
		if (count) {
			// 286 + 30 symbol counts, big-endian like every other field (lib/nx_dht.c:187-199 detects the
			// byte order by looking at the end-of-block count, which is always 1)
			uint8_t *p = cpb + NXGPU_CPB_OUT_LZCOUNT - NXGPU_CPB;
			for (int i = 0; i < 316; i++)
				put_be32(p + 4 * i, lz[i] > 0xffffff ? 0xffffff : lz[i]);
			put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP_WITH_COUNT - NXGPU_CPB, (uint32_t)src_total);
		} else {
			put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP - NXGPU_CPB, (uint32_t)src_total);
		}
		// manual Table 6-8: the target came out larger than the source
		complete(crb, o->out_len > src_total ? 64 : 0, 0, o->out_len);
		return 0;
	}

	// ---- decompress (inc_nx/nxu.h:812-815): raw deflate from a bit offset, with a preloaded window ----
	const bool resume = (fc & 0x04) != 0;
	const uint32_t hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
	if (hist >= src_total && !(hist == 0 && src_total == 0)) {
		if (hist > src_total) { complete(crb, 3, CE_TERMINATE, 0); return 0; }
	}
	const uint32_t comp_len = (uint32_t)(src_total - hist);
	const uint32_t in_subc = resume ? (w8 & 7) : 0;
	const uint32_t sfbt = resume ? (w12 >> 16) & 0xf : 0;
	const size_t out_off = align16(hist);
	if (c->d_in.reserve(comp_len + 32)) return -EAGAIN;
	if (c->d_out.reserve(out_off + dst_total + 32)) return -EAGAIN;
	if (c->d_dht.reserve(1024)) return -EAGAIN;
	if (c->d_jobs.reserve(sizeof(InflateJob))) return -EAGAIN;
	if (c->d_outs.reserve(sizeof(InflateOut))) return -EAGAIN;
	if (c->d_misc.reserve(64)) return -EAGAIN;
	uint8_t *d_out = static_cast<uint8_t *>(c->d_out.p);
	if (hist && cudaMemcpyAsync(d_out + out_off - hist, hs, hist, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
	if (comp_len && cudaMemcpyAsync(c->d_in.p, hs + hist, comp_len, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
	InflateJob job;
	memset(&job, 0, sizeof(job));
	job.src = static_cast<const uint8_t *>(c->d_in.p);
	job.src_len = comp_len;
	job.wrap = kWrapJob;
	job.dst = d_out + out_off;
	job.dst_cap = (uint32_t)dst_total;
	job.hist_len = hist;
	job.start_bit = (8 - in_subc) & 7;
	job.sfbt = sfbt;
	job.rembytecnt = w12 & 0xffff;
	job.out_dht = static_cast<uint8_t *>(c->d_dht.p) + 512;
	if ((sfbt & 0xe) == 0xc) {
		job.dht_bits = w12 & 0xfff;
		if (cudaMemcpyAsync(c->d_dht.p, cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, 288, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
		job.dht = static_cast<const uint8_t *>(c->d_dht.p);
	}
	if (c->h_jobs.reserve(sizeof(InflateJob))) return -EAGAIN;
	memcpy(c->h_jobs.p, &job, sizeof(job));
	if (cudaMemcpyAsync(c->d_jobs.p, c->h_jobs.p, sizeof(job), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return -EAGAIN;
	timer_begin(c, 1);
	if (launch_inflate(static_cast<const InflateJob *>(c->d_jobs.p), static_cast<InflateOut *>(c->d_outs.p), 1,
			   static_cast<uint32_t *>(c->d_misc.p), c->stream) != cudaSuccess) return -EAGAIN;
	timer_end(c, 1);
	if (c->h_outs.reserve(sizeof(InflateOut) + 64 + 288 + dst_total + 64)) return -EAGAIN;
	uint8_t *ho = static_cast<uint8_t *>(c->h_outs.p);
	InflateOut *o = reinterpret_cast<InflateOut *>(ho);
	uint32_t *ck = reinterpret_cast<uint32_t *>(ho + sizeof(InflateOut));
	uint8_t *odht = ho + sizeof(InflateOut) + 64;
	uint8_t *data = odht + 288;
	if (cudaMemcpyAsync(o, c->d_outs.p, sizeof(InflateOut), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
	if (cudaMemcpyAsync(odht, job.out_dht, 288, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
	if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -EAGAIN;
	if (o->rc != 0) {
		// 13: target full, the caller retries with less input; 66/67/68: bad code / distance / table
		complete(crb, (uint32_t)o->rc, o->rc == 13 ? 0 : CE_TERMINATE, 0);
		return 0;
	}
	nxgpu_cksum_item it = { job.dst, o->out_len, crc_seed, adler_seed };
	if (checksum_device(c, &it, 1, 3)) return -EAGAIN;
	if (cudaMemcpyAsync(ck, c->d_cks.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
	if (o->out_len && cudaMemcpyAsync(data, job.dst, o->out_len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -EAGAIN;
	if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -EAGAIN;
	scatter(dst, data, o->out_len);
	put_be32(cpb + NXGPU_CPB_OUT_ADLER - NXGPU_CPB, ck[1]);
	put_le32(cpb + NXGPU_CPB_OUT_CRC - NXGPU_CPB, ck[0]);
	put_be32(cpb + NXGPU_CPB_OUT_TEBC - NXGPU_CPB, o->subc & 0xffff);                   // out_subc: low half of this word
	const bool in_dyn = (o->sfbt & 0xe) == 0xc;
	put_be32(cpb + NXGPU_CPB_OUT_SFBT - NXGPU_CPB, (o->sfbt & 0xf) << 16 | (in_dyn ? (o->dhtlen & 0xfff) : (o->rembytecnt & 0xffff)));
	if (in_dyn)
		memcpy(cpb + NXGPU_CPB_OUT_DHT - NXGPU_CPB, odht, 288);
	put_be32(cpb + NXGPU_CPB_OUT_SPBC_DECOMP - NXGPU_CPB, (uint32_t)src_total);
	// CC=3 with CE "partial completion" is the normal way a decompress job ends (lib/nx_inflate.c:1372-1390)
	complete(crb, 3, CE_PARTIAL | CE_TPBC_VALID, o->out_len);
	return 0;
}

} // namespace nxgpu
