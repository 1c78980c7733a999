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

// One descriptor of a batch, from parsing to completion.
namespace {
struct JobState {
	enum Kind { DONE, WRAP, COMP, DECOMP };
	uint8_t *crb = nullptr, *cpb = nullptr;
	Kind kind = DONE;
	uint32_t fc = 0;
	std::vector<Seg> src, dst;
	uint64_t src_total = 0, dst_total = 0;
	size_t in_off = 0;            // this job's source bytes in h_stage / d_in (history first)
	size_t out_off = 0;           // decompress: first target byte in d_out (history sits right in front)
	size_t host_off = 0;          // where the job's output bytes land in h_outs
	uint32_t hist = 0, n_new = 0; // compress: new bytes; decompress: compressed bytes
	uint32_t crc_seed = 0, adler_seed = 1;
	bool use_dht = false, count = false;
	size_t idx = 0;               // index in the DeflateJob / InflateJob array
	size_t ck = 0;                // index in the checksum batch
	uint32_t out_len = 0;
	bool ok = false;              // output is to be returned
	size_t np = 1;                // compress: pieces the descriptor is cut into (one CTA each)
	uint64_t out_bits = 0;        // compress: valid bits of the block
	size_t cat_off = 0;           // compress, np > 1: where the joined block sits in d_cat
};
// A large compress descriptor is cut into pieces that separate CTAs compress with the SAME table (fixed or
// the caller's DHT, so no table has to be agreed on): piece 0 carries the block header, the last one the
// end-of-block, each is primed with the 32 KiB in front of it, and bitconcat_kernel joins the bit strings
// into the single block the descriptor asks for.  One descriptor then runs at 16 SMs' speed instead of one.
// The piece size follows the load: the pieces of one batch should fill the GPU without drowning in per-piece
// fixed cost — total compress bytes / 148 SMs, rounded up to a power of two within [8 KiB, 64 KiB].  A lone 64 KiB
// descriptor is cut into eight pieces (0.48 -> 0.23 ms), a batch of many MiB keeps 64 KiB pieces.
uint32_t piece_bytes(uint64_t batch_bytes)
{
	static const uint32_t forced = [] { const char *e = getenv("NXGPU_JOB_PIECE"); return e ? (uint32_t)strtoul(e, nullptr, 0) : 0u; }();   // developer switch
	if (forced)
		return forced;
	uint32_t p = 8192;
	while (p < 65536 && (uint64_t)p * kNumSMs < batch_bytes)
		p *= 2;
	return p;
}
} // namespace

// Runs n descriptors as ONE batch: one upload of all sources, one deflate launch for every compress
// job, one inflate launch for every decompress job, one checksum launch pair for everything, two
// stream synchronisations in total.  rcs[i] = 0 (CSB/CPB filled) or -EAGAIN (device unusable).
// This is what makes many concurrent small z_streams (test/test_multithread_stress.c, LD_PRELOAD
// users) fast without touching the zlib surface: nxu_run_job coalesces whatever is pending
// (nxgpu_dropin.cu) and hands it over here.
void run_jobs_batch(nxgpu_ctx *c, uint8_t *const *crbs, int *rcs, size_t n)
{
	NXGPU_LOCK(c);
	std::vector<JobState> js(n);
	auto fail_all = [&]() { for (size_t i = 0; i < n; i++) if (js[i].kind != JobState::DONE) rcs[i] = -EAGAIN; };
	for (size_t i = 0; i < n; i++) rcs[i] = 0;
	if (cudaSetDevice(c->dev) != cudaSuccess) {
		for (size_t i = 0; i < n; i++) rcs[i] = -EAGAIN;
		return;
	}
	// ---- parse ----
	size_t in_total = 0, out_total = 0, nd = 0, ni = 0;
	uint64_t comp_bytes = 0;
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		j.crb = crbs[i];
		j.cpb = j.crb + NXGPU_CPB;
		j.fc = be32(j.crb + NXGPU_CRB_FC) & 0xff;
		if (!dde_segments(j.crb + NXGPU_CRB_SRC_DDE, j.src, j.src_total) || !dde_segments(j.crb + NXGPU_CRB_DST_DDE, j.dst, j.dst_total)) {
			complete(j.crb, 30 /* ERR_NX_INVALID_DDE, inc_nx/nxu.h:843 */, CE_TERMINATE, 0);
			continue;
		}
		if (j.src_total > 0xffffff00ull || j.dst_total > 0xffffff00ull) {
			complete(j.crb, 3, CE_TERMINATE, 0);
			continue;
		}
		const uint32_t w8 = be32(j.cpb + 8);
		// running checksums: in_adler is a plain big-endian field, in_crc holds the CRC with its
		// bytes the other way round (lib/nx_deflate.c:1572-1577 and lib/nx_inflate.c:809-817 rely on it)
		j.adler_seed = be32(j.cpb + NXGPU_CPB_IN_ADLER - NXGPU_CPB);
		j.crc_seed = le32(j.cpb + NXGPU_CPB_IN_CRC - NXGPU_CPB);
		if (j.fc == 0x1e) {
			// GZIP_FC_WRAP (inc_nx/nxu.h:816; caller lib/nx_zlib.c:1398): copy + fresh crc32/adler32
			if (j.dst_total < j.src_total) { complete(j.crb, 13, 0, 0); continue; }
			j.kind = JobState::WRAP;
			j.crc_seed = 0; j.adler_seed = 1;
		} else if ((j.fc & 0x10) == 0) {
			const bool resume = (j.fc & 0x08) != 0;
			j.use_dht = (j.fc & 0x02) != 0;
			j.count = (j.fc & 0x04) != 0;
			j.hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
			if (j.hist > j.src_total) { complete(j.crb, 3, CE_TERMINATE, 0); continue; }   // history length error
			j.n_new = (uint32_t)(j.src_total - j.hist);
			j.kind = JobState::COMP;
			comp_bytes += j.n_new;
		} else {
			const bool resume = (j.fc & 0x04) != 0;
			j.hist = resume ? ((w8 >> 20) & 0xfff) * 16 : 0;
			if (j.hist > j.src_total) { complete(j.crb, 3, CE_TERMINATE, 0); continue; }
			j.n_new = (uint32_t)(j.src_total - j.hist);
			j.kind = JobState::DECOMP;
			j.idx = ni++;
			out_total += align16(j.hist);
			j.out_off = out_total;
			out_total += align16(j.dst_total + 32);
		}
		j.in_off = in_total;
		in_total += align16(j.src_total + 48);
	}
	const uint32_t kPiece = piece_bytes(comp_bytes);
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		if (j.kind != JobState::COMP)
			continue;
		j.np = j.n_new >= 2 * kPiece ? (j.n_new + kPiece - 1) / kPiece : 1;
		j.idx = nd;
		nd += j.np;
	}
	if (in_total == 0 && nd == 0 && ni == 0)
		return;

	// ---- stage every source in pinned memory (history first, exactly as the DDE list has it), one upload ----
	if (c->h_stage.reserve(in_total + 64) || c->d_in.reserve(in_total + 64) || c->d_dht.reserve(n * 1024) ||
	    c->h_jobs.reserve(n * 1024 + ni * sizeof(InflateJob)) || c->d_lz.reserve((nd + 1) * 316 * 4) ||
	    c->d_out.reserve(out_total + 64) || c->d_ijobs.reserve((ni + 1) * sizeof(InflateJob)) ||
	    c->d_iouts.reserve((ni + 1) * sizeof(InflateOut)) || c->d_misc.reserve(64)) {
		fail_all();
		return;
	}
	uint8_t *hs = static_cast<uint8_t *>(c->h_stage.p);
	uint8_t *d_in = static_cast<uint8_t *>(c->d_in.p);
	uint8_t *d_out = static_cast<uint8_t *>(c->d_out.p);
	uint8_t *d_dht = static_cast<uint8_t *>(c->d_dht.p);
	uint8_t *h_dht = static_cast<uint8_t *>(c->h_jobs.p);                      // n x 1024: caller tables, as the kernels want them
	InflateJob *ij = reinterpret_cast<InflateJob *>(h_dht + n * 1024);
	for (size_t i = 0; i < n; i++)
		if (js[i].kind != JobState::DONE)
			gather(js[i].src, hs + js[i].in_off);
	if (in_total && cudaMemcpyAsync(d_in, hs, in_total, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail_all(); return; }

	// ---- device job arrays ----
	std::vector<DeflateJob> dj(nd);
	bool any_dht = false;
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		if (j.kind == JobState::COMP) {
			uint32_t dhtlen = 0;
			bool dht_ok = true;
			if (j.use_dht) {
				dhtlen = be32(j.cpb + 12) & 0xfff;
				uint8_t *blob = h_dht + i * 1024;
				memset(blob, 0, 320 + 288);
				dht_ok = dhtlen >= 42 && dhtlen <= 288 * 8 && dht_to_lengths(j.cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, dhtlen, blob);
				if (dht_ok) {
					memcpy(blob + 320, j.cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, (dhtlen + 7) / 8);
					any_dht = true;
				}
			}
			for (size_t p = 0; p < j.np; p++) {
				DeflateJob &job = dj[j.idx + p];
				memset(&job, 0, sizeof(job));
				const uint32_t off = (uint32_t)p * kPiece;
				job.src = d_in + j.in_off + j.hist + off;
				job.src_len = j.np == 1 ? j.n_new : (j.n_new - off < kPiece ? j.n_new - off : kPiece);
				const uint64_t before = (uint64_t)j.hist + off;
				job.hist_len = before > 32768 ? 32768 : (uint32_t)before;
				job.flags = NXGPU_F_NO_JOINER | (j.use_dht ? 0 : NXGPU_F_FIXED) | (p > 0 ? NXGPU_F_NO_HEADER : 0) | (p + 1 < j.np ? NXGPU_F_NO_EOB : 0);
				if (!dht_ok) {
					job.src_len = 0; job.flags |= NXGPU_F_FIXED;       // keeps its slot in the launch, result ignored
					continue;
				}
				if (j.use_dht) {
					job.dht = d_dht + i * 1024;
					job.dht_bits = dhtlen;
				}
				if (j.count)
					job.lzcount = static_cast<uint32_t *>(c->d_lz.p) + (j.idx + p) * 316;
			}
			if (!dht_ok) {
				complete(j.crb, 68 /* invalid DHT */, CE_TERMINATE, 0);
				j.kind = JobState::DONE;
			}
		} else if (j.kind == JobState::DECOMP) {
			const uint32_t w8 = be32(j.cpb + 8), w12 = be32(j.cpb + 12);
			const bool resume = (j.fc & 0x04) != 0;
			const uint32_t in_subc = resume ? (w8 & 7) : 0;
			const uint32_t sfbt = resume ? (w12 >> 16) & 0xf : 0;
			InflateJob &job = ij[j.idx];
			memset(&job, 0, sizeof(job));
			job.src = d_in + j.in_off + j.hist;
			job.src_len = j.n_new;
			job.wrap = kWrapJob;
			job.dst = d_out + j.out_off;
			job.dst_cap = (uint32_t)j.dst_total;
			job.hist_len = j.hist;
			job.start_bit = (8 - in_subc) & 7;
			job.sfbt = sfbt;
			job.rembytecnt = w12 & 0xffff;
			job.single_block = (j.fc & 0x02) != 0;      // GZIP_FC_DECOMPRESS[_RESUME]_SINGLE_BLK_N_SUSPEND
			job.out_dht = d_dht + i * 1024 + 640;
			if ((sfbt & 0xe) == 0xc) {
				job.dht_bits = w12 & 0xfff;
				memcpy(h_dht + i * 1024, j.cpb + NXGPU_CPB_IN_DHT - NXGPU_CPB, 288);
				job.dht = d_dht + i * 1024;
				any_dht = true;
			}
			// the window: the history bytes go right in front of the target (device-to-device, they are already up)
			if (j.hist && cudaMemcpyAsync(d_out + j.out_off - j.hist, d_in + j.in_off, j.hist, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { fail_all(); return; }
		}
	}
	if (any_dht && cudaMemcpyAsync(d_dht, h_dht, n * 1024, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail_all(); return; }

	// ---- launches ----
	if (nd && deflate_device(c, dj.data(), nd, job_level(), false)) { fail_all(); return; }
	if (ni) {
		std::vector<std::pair<size_t, InflateJob>> par;
		inflate_par_select(ij, ni, par);
		if (cudaMemcpyAsync(c->d_ijobs.p, ij, ni * sizeof(InflateJob), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail_all(); return; }
		if (!par.empty() && cudaEventRecord(c->ev_main, c->stream) != cudaSuccess) { fail_all(); return; }   // sources and histories are up
		timer_begin(c, 1);
		if (launch_inflate(static_cast<const InflateJob *>(c->d_ijobs.p), static_cast<InflateOut *>(c->d_iouts.p), (uint32_t)ni,
				   static_cast<uint32_t *>(c->d_misc.p), c->stream) != cudaSuccess) { fail_all(); return; }
		timer_end(c, 1);
		c->launches++;
		// (ij[] may still be on its way up: the skip marks stay)
		if (!par.empty() && inflate_parallel(c, par, static_cast<InflateOut *>(c->d_iouts.p), true)) { fail_all(); return; }
	}
	// ---- results of the codec kernels (sizes, states) ----
	const size_t res_bytes = align16(nd * sizeof(DeflateOut)) + align16(ni * sizeof(InflateOut)) + align16(nd * 316 * 4) + n * 288 + n * 8 + 64;
	size_t data_total = 0;
	for (size_t i = 0; i < n; i++) {
		js[i].host_off = res_bytes + data_total;
		if (js[i].kind == JobState::COMP) data_total += align16(2 * (size_t)js[i].n_new + 2048);
		else if (js[i].kind == JobState::DECOMP) data_total += align16(js[i].dst_total + 64);
		else if (js[i].kind == JobState::WRAP) data_total += align16(js[i].src_total + 64);
	}
	if (c->h_outs.reserve(res_bytes + data_total + 64)) { fail_all(); return; }
	uint8_t *ho = static_cast<uint8_t *>(c->h_outs.p);
	DeflateOut *dout = reinterpret_cast<DeflateOut *>(ho);
	InflateOut *iout = reinterpret_cast<InflateOut *>(ho + align16(nd * sizeof(DeflateOut)));
	uint32_t *lz = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(iout) + align16(ni * sizeof(InflateOut)));
	uint8_t *odht = reinterpret_cast<uint8_t *>(lz) + align16(nd * 316 * 4);
	uint32_t *ck = reinterpret_cast<uint32_t *>(odht + n * 288);
	bool any_count = false, any_decomp_dht = false;
	for (size_t i = 0; i < n; i++) { any_count |= js[i].kind == JobState::COMP && js[i].count; any_decomp_dht |= js[i].kind == JobState::DECOMP; }
	if (nd && cudaMemcpyAsync(dout, c->d_outs.p, nd * sizeof(DeflateOut), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
	if (ni && cudaMemcpyAsync(iout, c->d_iouts.p, ni * sizeof(InflateOut), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
	if (any_count && cudaMemcpyAsync(lz, c->d_lz.p, nd * 316 * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
	if (cudaStreamSynchronize(c->stream) != cudaSuccess) { fail_all(); return; }

	// ---- per-job verdict; pieces of cut descriptors are joined on the device ----
	std::vector<nxgpu_cksum_item> items;
	std::vector<BitPiece> pieces;
	std::vector<BitGroup> groups;
	size_t cat_total = 0;
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		if (j.kind == JobState::WRAP) {
			j.ok = true; j.out_len = (uint32_t)j.src_total;
		} else if (j.kind == JobState::COMP) {
			bool missing = false, failed = false;
			uint64_t bits = 0;
			for (size_t p = 0; p < j.np; p++) {
				const DeflateOut &o = dout[j.idx + p];
				if (o.rc == 66) missing = true;
				else if (o.rc != 0) failed = true;
				else if (o.out_len) bits += (uint64_t)(o.out_len - 1) * 8 + (o.tebc ? o.tebc : 8);
			}
			if (missing) { complete(j.crb, 66, CE_TERMINATE, 0); continue; }                    // a needed symbol has no code
			if (failed || (bits + 7) / 8 > j.dst_total) { complete(j.crb, 13, 0, 0); continue; }  // ERR_NX_TARGET_SPACE: caller halves the input
			j.ok = true; j.out_bits = bits; j.out_len = (uint32_t)((bits + 7) / 8);
			if (j.np > 1) {
				j.cat_off = cat_total;
				cat_total += align16(j.out_len + 32);
				BitGroup g = { nullptr, bits, (uint32_t)pieces.size(), (uint32_t)j.np };
				uint64_t at = 0;
				for (size_t p = 0; p < j.np; p++) {
					const DeflateOut &o = dout[j.idx + p];
					const uint64_t nb = o.out_len ? (uint64_t)(o.out_len - 1) * 8 + (o.tebc ? o.tebc : 8) : 0;
					pieces.push_back({ dj[j.idx + p].out, nb, at });
					at += nb;
				}
				groups.push_back(g);
			}
		} else if (j.kind == JobState::DECOMP) {
			const InflateOut &o = iout[j.idx];
			if (o.rc != 0) {
				// 13: target full, the caller retries with less input; 66/67/68: bad code / distance / table
				complete(j.crb, (uint32_t)o.rc, o.rc == 13 ? 0 : CE_TERMINATE, 0);
				continue;
			}
			j.ok = true; j.out_len = o.out_len;
		}
	}
	if (!groups.empty()) {
		const size_t pb = pieces.size() * sizeof(BitPiece), gb = groups.size() * sizeof(BitGroup);
		if (c->d_cat.reserve(cat_total + 64) || c->d_catdesc.reserve(pb + gb + 64) || c->h_cat.reserve(pb + gb + 64)) { fail_all(); return; }
		size_t gi = 0;
		for (size_t i = 0; i < n; i++)
			if (js[i].ok && js[i].kind == JobState::COMP && js[i].np > 1)
				groups[gi++].dst = static_cast<uint8_t *>(c->d_cat.p) + js[i].cat_off;
		uint8_t *hc = static_cast<uint8_t *>(c->h_cat.p);
		memcpy(hc, pieces.data(), pb);
		memcpy(hc + align16(pb), groups.data(), gb);
		if (cudaMemcpyAsync(c->d_catdesc.p, hc, align16(pb) + gb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail_all(); return; }
		const uint8_t *dd = static_cast<const uint8_t *>(c->d_catdesc.p);
		if (launch_bitconcat(reinterpret_cast<const BitPiece *>(dd), reinterpret_cast<const BitGroup *>(dd + align16(pb)),
				     (uint32_t)groups.size(), c->stream) != cudaSuccess) { fail_all(); return; }
		c->launches++;
	}
	// ---- checksums and output bytes of the good ones ----
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		if (!j.ok)
			continue;
		j.ck = items.size();
		if (j.kind == JobState::WRAP) {
			// the copy goes through the engine like every other function code: up with the batch, back from the device
			items.push_back({ d_in + j.in_off, j.src_total, 0, 1 });
			if (j.src_total && cudaMemcpyAsync(ho + j.host_off, d_in + j.in_off, j.src_total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
		} else if (j.kind == JobState::COMP) {
			items.push_back({ dj[j.idx].src, j.n_new, j.crc_seed, j.adler_seed });
			const uint8_t *from = j.np > 1 ? static_cast<const uint8_t *>(c->d_cat.p) + j.cat_off : dj[j.idx].out;
			if (j.out_len && cudaMemcpyAsync(ho + j.host_off, from, j.out_len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
		} else {
			const InflateOut &o = iout[j.idx];
			items.push_back({ ij[j.idx].dst, o.out_len, j.crc_seed, j.adler_seed });
			if (o.out_len && cudaMemcpyAsync(ho + j.host_off, ij[j.idx].dst, o.out_len, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
			if ((o.sfbt & 0xe) == 0xc &&
			    cudaMemcpyAsync(odht + i * 288, ij[j.idx].out_dht, 288, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { fail_all(); return; }
		}
	}
	(void)any_decomp_dht;
	if (items.empty())
		return;
	if (checksum_device(c, items.data(), items.size(), 3)) { for (size_t i = 0; i < n; i++) if (js[i].ok) rcs[i] = -EAGAIN; return; }
	if (cudaMemcpyAsync(ck, c->d_cks.p, items.size() * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
	    cudaStreamSynchronize(c->stream) != cudaSuccess) { for (size_t i = 0; i < n; i++) if (js[i].ok) rcs[i] = -EAGAIN; return; }
	const size_t nck = items.size();

	// ---- hand the bytes and the parameter-block outputs back ----
	for (size_t i = 0; i < n; i++) {
		JobState &j = js[i];
		if (!j.ok)
			continue;
		uint8_t *cpb = j.cpb;
		put_be32(cpb + NXGPU_CPB_OUT_ADLER - NXGPU_CPB, ck[nck + j.ck]);
		put_le32(cpb + NXGPU_CPB_OUT_CRC - NXGPU_CPB, ck[j.ck]);
		if (j.kind == JobState::WRAP) {
			scatter(j.dst, ho + j.host_off, j.src_total);
			put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP - NXGPU_CPB, (uint32_t)j.src_total);
			complete(j.crb, 0, 0, (uint32_t)j.src_total);
		} else if (j.kind == JobState::COMP) {
			scatter(j.dst, ho + j.host_off, j.out_len);
			put_be32(cpb + NXGPU_CPB_OUT_TEBC - NXGPU_CPB, (uint32_t)(j.out_bits & 7) << 16);
			if (j.count) {
				// 286 + 30 symbol counts, big-endian like every other field (lib/nx_dht.c:187-199 detects the
				// byte order by looking at the end-of-block count, which is always 1); pieces add up
				uint8_t *p = cpb + NXGPU_CPB_OUT_LZCOUNT - NXGPU_CPB;
				for (int k = 0; k < 316; k++) {
					uint64_t v = 0;
					for (size_t q = 0; q < j.np; q++)
						v += lz[(j.idx + q) * 316 + k];
					if (k == 256)
						v = 1;
					put_be32(p + 4 * k, v > 0xffffff ? 0xffffff : (uint32_t)v);
				}
				put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP_WITH_COUNT - NXGPU_CPB, (uint32_t)j.src_total);
			} else {
				put_be32(cpb + NXGPU_CPB_OUT_SPBC_COMP - NXGPU_CPB, (uint32_t)j.src_total);
			}
			// manual Table 6-8: the target came out larger than the source
			complete(j.crb, j.out_len > j.src_total ? 64 : 0, 0, j.out_len);
		} else {
			const InflateOut &o = iout[j.idx];
			scatter(j.dst, ho + j.host_off, o.out_len);
			put_be32(cpb + NXGPU_CPB_OUT_TEBC - NXGPU_CPB, o.subc > 0xffff ? 0xffff : o.subc);  // out_subc: low half of this word; never wraps
			const bool in_dyn = (o.sfbt & 0xe) == 0xc;
			put_be32(cpb + NXGPU_CPB_OUT_SFBT - NXGPU_CPB, (o.sfbt & 0xf) << 16 | (in_dyn ? (o.dhtlen & 0xfff) : (o.rembytecnt & 0xffff)));
			if (in_dyn)
				memcpy(cpb + NXGPU_CPB_OUT_DHT - NXGPU_CPB, odht + i * 288, 288);
			// source bytes the engine read, history included (inc_nx/nxu.h:454-465, lib/nx_inflate.c:1452-1472)
			put_be32(cpb + NXGPU_CPB_OUT_SPBC_DECOMP - NXGPU_CPB, j.hist + o.in_used);
			// CC=3 with CE "partial completion" is the normal way a decompress job ends (lib/nx_inflate.c:1372-1390)
			complete(j.crb, 3, CE_PARTIAL | CE_TPBC_VALID, o.out_len);
		}
	}
}

// returns 0 (CSB/CPB filled) or -EAGAIN when the device cannot be used
int run_job_impl(nxgpu_ctx *c, uint8_t *crb)
{
	int rc = 0;
	run_jobs_batch(c, &crb, &rc, 1);
	static const bool trace = getenv("NXGPU_TRACE") != nullptr;
	if (trace) {
		const uint8_t *cpb = crb + NXGPU_CPB;
		fprintf(stderr, "nxgpu job: fc %02x src bc %u dst bc %u in(w8 %08x w12 %08x) -> rc %d csb %08x tpbc %u out(w392 %08x w396 %08x)\n",
			be32(crb) & 0xff, be32(crb + NXGPU_CRB_SRC_DDE + 4), be32(crb + NXGPU_CRB_DST_DDE + 4), be32(cpb + 8), be32(cpb + 12), rc,
			be32(crb + NXGPU_CRB_CSB), be32(crb + NXGPU_CRB_CSB + 4), be32(cpb + 392), be32(cpb + 396));
	}
	return rc;
}

} // namespace nxgpu
