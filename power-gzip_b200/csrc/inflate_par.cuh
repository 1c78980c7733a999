// inflate_par.cuh — ONE deflate stream decoded by many warps (included by inflate.cu, inside its namespace).
//
// A decompress descriptor (inc_nx/nxu.h:812-815) or a foreign gzip/zlib member is a single bit stream: one warp pair walks
// it at ~150 MB/s, below one host core, while 147 SMs idle.  Deflate blocks are independent Huffman-wise, and what they
// need from their predecessors is only the 32 KiB window, so:
//
//   1. blockfind     every bit offset of the source is tested for "a dynamic block header starts here" (BTYPE=10, HLIT and
//                    HDIST in range, a complete code-length code, code lengths that decode to complete literal/length
//                    and distance codes).  The result is a bit map and a sorted list of candidates.  False positives
//                    only cost work; a real block start that the test refuses (the lenient codes an NX job may carry)
//                    only costs parallelism — the map merely has to be a fixed function of the source.
//   2. speculation   one warp pair per candidate (a walker and a copier, see DuoQueue) decodes from its candidate to the
//                    first block boundary that is a candidate again (stored and fixed blocks are walked through).  The
//                    window in front of it is unknown: the copier keeps the last 32 Ki symbols as 16-bit values in
//                    shared memory, initialised with markers "window byte s", so a match that reaches in front of the
//                    piece copies markers.  Nothing is
//                    written but the ring, the piece's length and where it ended.  In the same launch piece 0 — from
//                    the descriptor's own start state (container header, resumed block, history) — is decoded for real.
//   3. link          one thread follows piece 0's end through the candidate list: piece k starts where piece k-1
//                    ended; output offsets are the running sum.  The chain stops at the first piece that did not end
//                    on a candidate (final block, source end, error, target full): that piece is decoded to the end
//                    of the job by the serial rule.
//   4. windows       the 32 KiB in front of every chained piece: a marker is chased through the rings of the
//                    predecessors until it names a byte (usually zero or one hop), all pieces and slots in parallel.
//   5. real decode   the serial decoder (inflate_one + duo_copier, window in shared memory) runs every chained piece with
//                    its true window, a warp pair per piece, writing the target; the last one reports the completion state (SFBT, SUBC, DHT, errors)
//                    exactly as the serial engine would, because it is the serial engine.
//   6. finish        offsets are folded into the last piece's result; pieces that did not end where speculation said
//                    they would (cannot happen) send the whole descriptor through the serial path.
//
// Critical path: two block decodes instead of all of them.  A dry run (kWrapDry: where does the member end, how long is
// its output — nxgpu_gunzip_concat) stops after step 3: the chain is the answer (inflate_dry_finish_kernel).

constexpr uint32_t kSpecLinked = 0, kSpecFinal = 1, kSpecSrcEnd = 2, kSpecError = 3, kSpecTooLong = 4;
constexpr uint32_t kRingSyms = 32768;

// ---- 1. block-start candidates ----
// stage 1: thread per aligned source word, 32 bit offsets each: 3 header bits, HLIT / HDIST ranges, Kraft sum of the code-length code
__global__ void __launch_bounds__(256)
blockfind_filter_kernel(const uint8_t *__restrict__ src, uint32_t src_len, uint64_t first_bit, uint64_t *__restrict__ surv, uint32_t surv_cap,
			uint32_t *__restrict__ counts)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(src);
	const uint32_t *base32 = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
	const uint32_t skip = (uint32_t)(a & 3);
	const uint64_t end = (uint64_t)skip + src_len;
	const uint64_t total_bits = (uint64_t)src_len * 8;
	const uint32_t n_words = (uint32_t)((end + 3) >> 2);
	// Kraft sum of four 3-bit code lengths at once (unit 1/128)
	__shared__ uint16_t kraft4[4096];
	for (uint32_t v = threadIdx.x; v < 4096; v += blockDim.x) {
		uint32_t sum = 0;
		for (int f = 0; f < 4; f++) {
			const uint32_t l = (v >> (3 * f)) & 7;
			if (l) sum += 128u >> l;
		}
		kraft4[v] = (uint16_t)sum;
	}
	__syncthreads();
	for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
		uint32_t w[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
			w[k] = BitReader::load_word(base32, skip, end, wi + k);
#pragma unroll
		for (int o = 0; o < 32; o++) {
			const uint32_t k = o >> 5 ? 1 : 0, sh = o & 31;
			const uint32_t h = __funnelshift_r(w[k], w[k + 1], sh);
			// BFINAL any, BTYPE = 2 (bits 1..2 = 0, 1), HLIT <= 29, HDIST <= 29
			if (((h >> 1) & 3) != 2 || ((h >> 3) & 31) > 29 || ((h >> 8) & 31) > 29)
				continue;
			const uint32_t hclen = ((h >> 13) & 15) + 4;
			// 19 x 3 bits from bit o + 17
			const uint32_t o2 = o + 17;
			const uint32_t k2 = o2 >> 5, s2 = o2 & 31;
			const uint32_t c0 = __funnelshift_r(w[k2], w[k2 + 1], s2), c1 = __funnelshift_r(w[k2 + 1], w[k2 + 2], s2);
			const uint64_t cl = ((uint64_t)c0 | ((uint64_t)c1 << 32)) & ((1ull << (3 * hclen)) - 1);
			const uint32_t sum = kraft4[(uint32_t)cl & 4095] + kraft4[(uint32_t)(cl >> 12) & 4095] + kraft4[(uint32_t)(cl >> 24) & 4095] +
					     kraft4[(uint32_t)(cl >> 36) & 4095] + kraft4[(uint32_t)(cl >> 48) & 4095];
			if (sum != 128)
				continue;
			const int64_t p = (int64_t)wi * 32 + o - (int64_t)skip * 8;
			if (p <= (int64_t)first_bit || (uint64_t)p + 17 + 3 * hclen + 16 > total_bits)
				continue;
			const uint32_t idx = atomicAdd(&counts[0], 1u);
			if (idx < surv_cap)
				surv[idx] = (uint64_t)p;
		}
	}
}

// plain LSB-first reader over global memory (one thread)
struct GlobalBits {
	const uint32_t *base32; uint32_t skip; uint64_t end;
	uint64_t buf; uint32_t cnt; uint32_t w;
	uint64_t consumed;                 // bits, counted from base32
	__device__ void init(const uint8_t *src, uint32_t src_len, uint64_t bit)
	{
		const uintptr_t a = reinterpret_cast<uintptr_t>(src);
		base32 = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		skip = (uint32_t)(a & 3);
		end = (uint64_t)skip + src_len;
		const uint64_t abs = bit + skip * 8;
		w = (uint32_t)(abs >> 5);
		buf = BitReader::load_word(base32, skip, end, w++) >> (abs & 31);
		cnt = 32 - (uint32_t)(abs & 31);
		consumed = abs;
	}
	__device__ uint32_t get(uint32_t n)          // n <= 16
	{
		if (cnt < n) {
			buf |= (uint64_t)BitReader::load_word(base32, skip, end, w++) << cnt;
			cnt += 32;
		}
		const uint32_t v = (uint32_t)buf & ((1u << n) - 1);
		buf >>= n; cnt -= n; consumed += n;
		return v;
	}
	__device__ bool overrun() const { return consumed > end * 8; }
};

// stage 2: thread per survivor: the code lengths themselves.  Accepts exactly what build_table(strict) accepts.
__global__ void __launch_bounds__(128)
blockfind_verify_kernel(const uint8_t *__restrict__ src, uint32_t src_len, const uint64_t *__restrict__ surv, uint32_t surv_cap,
			uint32_t *__restrict__ map, uint64_t *__restrict__ cand, uint32_t cand_cap, uint32_t *__restrict__ counts)
{
	const uint32_t n = min(counts[0], surv_cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint64_t p = surv[i];
		GlobalBits br;
		br.init(src, src_len, p);
		const uint32_t v = br.get(14 + 3) >> 3;
		const int hlit = (int)(v & 31) + 257, hdist = (int)((v >> 5) & 31) + 1, hclen = (int)(v >> 10) + 4;
		uint8_t cl[19];
		for (int k = 0; k < 19; k++) cl[k] = 0;
		for (int k = 0; k < hclen; k++) cl[k_clorder[k]] = (uint8_t)br.get(3);
		// canonical decode of the code-length code (complete: stage 1 checked the Kraft sum)
		uint8_t ccount[8], csorted[19], coffs[9];
		for (int l = 0; l < 8; l++) ccount[l] = 0;
		for (int k = 0; k < 19; k++) ccount[cl[k]]++;
		ccount[0] = 0;
		coffs[1] = 0;
		for (int l = 1; l < 8; l++) coffs[l + 1] = coffs[l] + ccount[l];
		for (int k = 0; k < 19; k++) if (cl[k]) csorted[coffs[cl[k]]++] = (uint8_t)k;
		uint32_t lcnt[16], dcnt[16];                 // codes per length
		for (int l = 0; l < 16; l++) lcnt[l] = dcnt[l] = 0;
		int nsym = 0, prev = 0;
		uint32_t kraft_l = 0, kraft_d = 0;           // running Kraft sums, unit 2^-15
		bool ok = true, eob = false;
		while (ok && nsym < hlit + hdist) {
			int code = 0, first = 0, index = 0, sym = -1;
			for (int l = 1; l <= 7; l++) {
				code |= (int)br.get(1);
				const int c = ccount[l];
				if (code - c < first) { sym = csorted[index + (code - first)]; break; }
				index += c; first += c; first <<= 1; code <<= 1;
			}
			if (sym < 0) { ok = false; break; }
			int rep = 1, val = sym;
			if (sym == 16) { if (nsym == 0) { ok = false; break; } val = prev; rep = 3 + (int)br.get(2); }
			else if (sym == 17) { val = 0; rep = 3 + (int)br.get(3); }
			else if (sym == 18) { val = 0; rep = 11 + (int)br.get(7); }
			if (nsym + rep > hlit + hdist) { ok = false; break; }
			for (int r = 0; r < rep; r++, nsym++) {
				if (nsym < hlit) { lcnt[val]++; if (nsym == 256 && val) eob = true; if (val) kraft_l += 32768u >> val; }
				else { dcnt[val]++; if (val) kraft_d += 32768u >> val; }
			}
			if (kraft_l > 32768u || kraft_d > 32768u) { ok = false; break; }     // over-subscribed already: most look-alikes end here
			prev = val;
		}
		if (!ok || !eob || br.overrun())
			continue;
		int left = 1, maxl = 0;
		for (int l = 1; l <= 15; l++) { left = (left << 1) - (int)lcnt[l]; if (left < 0) break; if (lcnt[l]) maxl = l; }
		if (left < 0 || (left > 0 && maxl > 1))
			continue;
		left = 1; maxl = 0;
		for (int l = 1; l <= 15; l++) { left = (left << 1) - (int)dcnt[l]; if (left < 0) break; if (dcnt[l]) maxl = l; }
		if (left < 0 || (left > 0 && maxl > 1))
			continue;
		const uint32_t idx = atomicAdd(&counts[1], 1u);
		if (idx < cand_cap) {
			cand[idx] = p;
			atomicOr(&map[p >> 5], 1u << (p & 31));
		}
	}
}

// ---- 2. speculative decode of one candidate: a walker warp and a copier warp (see DuoQueue in inflate.cu) ----
// walker: headers, tables, symbols; hands token batches and stored runs to the copier; decides how the piece ends
__device__ void spec_walker(const uint8_t *src, uint32_t src_len, uint64_t start_bit, const uint32_t *__restrict__ map, bool strict,
			    WarpTables &T, DuoQueue &Q, SpecOut &O)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t total_bits = (uint64_t)src_len * 8;
	uint32_t tsa = (uint32_t)__cvta_generic_to_shared(&T);
	asm volatile("" : "+r"(tsa));
	BitReader br;
	br.setup(src, src_len, T.in);
	if (lane == 0) {
		br.seek((uint32_t)(start_bit >> 3));
		br.drop((uint32_t)(start_bit & 7));
	}
	uint32_t status = kSpecError, head = 0;
	for (;;) {
		if (ld_vol(&Q.cerr)) { status = kSpecTooLong; break; }
		// ---- block header (lane 0) ----
		uint32_t btype = 0, stored_len = 0, stored_at = 0, fin = 0, bad = 0;
		int hlit = 0, hdist = 0;
		if (lane == 0) {
			const uint32_t h = br.get(3);
			fin = h & 1;
			btype = h >> 1;
			if (btype == 0) {
				br.align_byte();
				const uint32_t v = br.peek32();
				br.drop(32);
				if (((v ^ (v >> 16)) & 0xffff) != 0xffff) bad = 1;
				stored_len = v & 0xffff;
				stored_at = br.byte_pos();
			} else if (btype == 2) {
				if (parse_dyn_header(br, T.lens, hlit, hdist, strict)) bad = 1;
			} else if (btype == 3) {
				bad = 1;
			}
			if (br.overrun()) bad = 2;
		}
		bad = __shfl_sync(0xffffffffu, bad, 0);
		btype = __shfl_sync(0xffffffffu, btype, 0);
		fin = __shfl_sync(0xffffffffu, fin, 0);
		if (bad) { status = bad == 2 ? kSpecSrcEnd : kSpecError; break; }
		if (btype == 0) {
			stored_len = __shfl_sync(0xffffffffu, stored_len, 0);
			stored_at = __shfl_sync(0xffffffffu, stored_at, 0);
			if ((uint64_t)stored_at + stored_len > src_len) { status = kSpecSrcEnd; break; }
			if (stored_len)
				duo_push(Q, head, lane, kDuoStored, stored_at, stored_len, 0);
			if (lane == 0)
				br.seek(stored_at + stored_len);
			__syncwarp();
		} else {
			if (btype == 1) {
				for (int i = lane; i < 288; i += 32)
					T.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
				T.lens[288 + lane] = 5;
				hlit = 288; hdist = 30;
			} else {
				hlit = __shfl_sync(0xffffffffu, hlit, 0);
				hdist = __shfl_sync(0xffffffffu, hdist, 0);
			}
			__syncwarp();
			bool ok = build_table(T.lens, hlit, T.lit, kLitBits, T.lit_count, T.lit_sorted, false, strict && btype == 2);
			ok = build_table(T.lens + hlit, hdist, T.dist, kDistBits, T.dist_count, T.dist_sorted, true, strict && btype == 2) && ok;
			if (!ok) { status = kSpecError; break; }
			bool block_done = false;
			uint32_t stop = 0;               // 0 go on, else the status to leave with
			while (!block_done && !stop) {
				{
					const uint32_t wp = __shfl_sync(0xffffffffu, br.wpos, 0);
					uint32_t st = __shfl_sync(0xffffffffu, br.staged, 0);
					if (st - wp < kInAhead && (uint64_t)st * 4 < br.end) {
						const uint32_t w0 = BitReader::load_word(br.base32, br.skip, br.end, st + lane);
						const uint32_t w1 = BitReader::load_word(br.base32, br.skip, br.end, st + 32 + lane);
						T.in[(st + lane) % kInWords] = w0;
						T.in[(st + 32 + lane) % kInWords] = w1;
						br.staged = st + 64;
						__syncwarp();
					}
				}
				uint32_t wst;
				uint64_t sym_at = 0;
				const uint32_t qn = walk_batch(br, T, tsa, lane, wst, sym_at);
				if (wst == kWalkSrcEnd) stop = kSpecSrcEnd;
				else if (wst == kWalkBadCode) stop = kSpecError;
				else if (wst == kWalkEob) block_done = true;
				if (stop)
					break;
				if (ld_vol(&Q.cerr)) { stop = kSpecTooLong; break; }
				if (qn)
					duo_push(Q, head, lane, kDuoTokens, qn, 0, lane < qn ? T.q[lane] : 0);
			}
			if (stop) { status = stop; break; }
		}
		// ---- a block ended ----
		if (fin) { status = kSpecFinal; break; }
		uint64_t at = 0;
		if (lane == 0)
			at = br.bits_used();
		at = __shfl_sync(0xffffffffu, at, 0);
		if (at >= total_bits) { status = kSpecSrcEnd; break; }
		if ((map[at >> 5] >> (at & 31)) & 1) { status = kSpecLinked; break; }
	}
	if (lane == 0)
		O.end_bit = br.bits_used();
	duo_push(Q, head, lane, kDuoEnd, status, 0, 0);
}

// copier: the ring of the last 32 Ki symbols (bytes, or markers "window byte s" for what lies in front of the piece), the
// piece's length, how far back it reached; dumps the ring if the piece ended on a candidate
__device__ void spec_copier(const uint8_t *src, uint32_t out_cap, DuoQueue &Q, uint16_t *win, uint16_t *ring_out, SpecOut &O)
{
	const uint32_t lane = threadIdx.x & 31;
	{
		uint32_t *w32 = reinterpret_cast<uint32_t *>(win);
		for (uint32_t i = lane; i < kRingSyms / 2; i += 32)
			w32[i] = (0x8000u | (2 * i)) | ((0x8000u | (2 * i + 1)) << 16);
	}
	__syncwarp();
	uint32_t out = 0, max_back = 0, tail = 0, status = kSpecError;
	bool too_long = false;
	for (;;) {
		while (ld_vol(&Q.head) == tail)
			__nanosleep(40);
		__threadfence_block();
		const DuoQueue::Cmd &c = Q.cmd[tail & 1];
		const uint32_t kind = ld_vol(&c.kind), a = ld_vol(&c.a), b = ld_vol(&c.b);
		const uint32_t t = kind == kDuoTokens && lane < a ? ld_vol(&c.tok[lane]) : 0;
		__syncwarp();
		tail++;
		if (lane == 0)
			st_vol(&Q.tail, tail);
		if (kind == kDuoEnd) { status = a; break; }
		if (too_long)
			continue;
		if (kind == kDuoStored) {
			if (b > out_cap - out) {
				too_long = true;
			} else {
				for (uint32_t i = lane; i < b; i += 32)
					win[(out + i) & (kRingSyms - 1)] = src[a + i];
				out += b;
				__syncwarp();
			}
		} else {
			const bool is_m = lane < a && tok_is_match(t);
			const uint32_t mylen = lane < a ? (is_m ? tok_len(t) : 1) : 0;
			uint32_t incl = mylen;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= (uint32_t)o)
					incl += y;
			}
			const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
			const uint32_t my_out = out + incl - mylen;
			if (total > out_cap - out) {
				too_long = true;
			} else {
				if (is_m && tok_dist(t) > my_out)
					max_back = max(max_back, tok_dist(t) - my_out);
				ring_fill<uint16_t>(win, out, lane, t, is_m, mylen, incl, total);
				out += total;
			}
		}
		if (too_long && lane == 0)
			st_vol(&Q.cerr, 1);
	}
	if (too_long)
		status = kSpecTooLong;
	for (int o = 16; o; o >>= 1)
		max_back = max(max_back, __shfl_xor_sync(0xffffffffu, max_back, o));
	if (lane == 0) {
		O.out_len = out;
		O.status = status;
		O.max_back = max_back;
		O.pad_ = 0;
	}
	if (status == kSpecLinked) {
		__syncwarp();
		const uint4 *s4 = reinterpret_cast<const uint4 *>(win);
		uint4 *d4 = reinterpret_cast<uint4 *>(ring_out);
		for (uint32_t i = lane; i < kRingSyms * 2 / 16; i += 32)
			d4[i] = s4[i];
	}
}

// piece 0 (the descriptor itself, stopping at the first candidate boundary) and every candidate: two warps per CTA
__global__ void __launch_bounds__(64)
inflate_spec_kernel(const ParPlan P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	if (blockIdx.x == 0) {
		InflateJob J = P.job;
		J.stop_map = P.map;
		J.map_bit0 = 0;
		if (J.wrap & kWrapDry) {
			// counting only (member discovery): nothing to copy, one warp
			if (threadIdx.x < 32)
				inflate_one<false>(J, *P.head_out, *reinterpret_cast<WarpTables *>(smem_raw), nullptr);
		} else {
			inflate_duo(J, *P.head_out, smem_raw);
		}
		return;
	}
	WarpTables &T = *reinterpret_cast<WarpTables *>(smem_raw);
	DuoQueue &Q = *reinterpret_cast<DuoQueue *>(smem_raw + kDuoQueueAt);
	uint16_t *win = reinterpret_cast<uint16_t *>(smem_raw + kDuoWinAt);
	if (threadIdx.x < 8)
		reinterpret_cast<uint32_t *>(&Q)[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t u = blockIdx.x - 1;
	if (threadIdx.x < 32)
		spec_walker(P.job.src, P.job.src_len, P.cands[u], P.map, (P.job.wrap & 0xff) != kWrapJob, T, Q, P.spec[u]);
	else
		spec_copier(P.job.src, P.job.dst_cap, Q, win, P.rings + (size_t)u * kRingSyms, P.spec[u]);
}

// ---- 3. link: piece k starts where piece k-1 ended (one thread) ----
__global__ void inflate_link_kernel(const ParPlan P)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	const InflateOut &H = *P.head_out;
	uint32_t n = 0;
	if (H.rc == 0 && (H.flags & kInflateMapStop)) {
		uint64_t e = (uint64_t)H.end_bit_lo | ((uint64_t)H.end_bit_hi << 32);
		uint64_t off = H.out_len;
		uint32_t idx = 0;
		const uint32_t wrap = (P.job.wrap & 0xff) == kWrapJob ? kWrapJob : ((H.flags >> 8) & 0xff) | kWrapNoHeader;
		while (n < P.n_cand) {
			// the candidate that starts at bit e: the chain only moves forward, gallop from the last hit
			uint32_t lo = idx, hi = idx, step = 1;
			while (hi < P.n_cand && P.cands[hi] < e) { lo = hi + 1; hi += step; step *= 2; }
			if (hi > P.n_cand) hi = P.n_cand;
			while (lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if (P.cands[mid] < e) lo = mid + 1; else hi = mid;
			}
			if (lo >= P.n_cand || P.cands[lo] != e)
				break;                              // (cannot happen: the map and the list hold the same bits)
			idx = lo + 1;
			const SpecOut &S = P.spec[lo];
			const uint64_t have = (uint64_t)P.job.hist_len + off;      // bytes in front of the piece
			const bool last = S.status != kSpecLinked || S.max_back > have || off + S.out_len > P.job.dst_cap;
			ChainMeta &M = P.meta[n];
			M.bit = e; M.exp_end = S.end_bit; M.unit = lo; M.out_off = (uint32_t)off; M.exp_len = S.out_len; M.last = last;
			InflateJob J = P.job;
			J.src = P.job.src + (e >> 3);
			J.src_len = P.job.src_len - (uint32_t)(e >> 3);
			J.start_bit = (uint32_t)(e & 7);
			J.wrap = wrap;
			J.dst = P.job.dst + off;
			J.dst_cap = P.job.dst_cap - (uint32_t)off;
			J.hist_len = have > kWinBytes ? kWinBytes : (uint32_t)have;
			J.hist_ptr = P.hists + (size_t)n * kWinBytes;
			J.dht = nullptr; J.dht_bits = 0; J.sfbt = 0; J.rembytecnt = 0; J.single_block = 0;
			J.stop_map = last ? nullptr : P.map;
			J.map_bit0 = (e >> 3) * 8;
			P.cjobs[n] = J;
			n++;
			if (last)
				break;
			off += S.out_len;
			e = S.end_bit;
		}
		if (n && !P.meta[n - 1].last) {
			// ran out of candidates with the chain still open (cannot happen): the last piece takes the rest of the job
			P.meta[n - 1].last = 1;
			P.cjobs[n - 1].stop_map = nullptr;
		}
	}
	*P.n_chain = n;
}

// ---- 4. the 32 KiB in front of every chained piece ----
__global__ void __launch_bounds__(1024)
inflate_windows_kernel(const ParPlan P)
{
	const uint32_t n = *P.n_chain;
	for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
		const int64_t S = P.meta[c].out_off;
		uint8_t *hist = P.hists + (size_t)c * kWinBytes;
		// (gridDim.y CTAs share a piece: the chase is a chain of dependent loads, more threads hide it)
		for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < kWinBytes; i += blockDim.x * gridDim.y) {
			int64_t a = S - (int64_t)kWinBytes + i;       // output position, counted from the descriptor's first target byte
			int cc = (int)c - 1;
			uint32_t byte = 0;
			for (;;) {
				if (cc < 0) {
					// piece 0 and the descriptor's history are real bytes
					if (a >= -(int64_t)P.job.hist_len)
						byte = P.job.hist_ptr && a < 0 ? P.job.hist_ptr[(int64_t)kWinBytes + a] : P.job.dst[a];
					break;
				}
				const ChainMeta &M = P.meta[cc];
				const uint32_t sym = P.rings[(size_t)M.unit * kRingSyms + ((uint32_t)(a - (int64_t)M.out_off) & (kRingSyms - 1))];
				if (sym < 0x8000) { byte = sym; break; }
				a = (int64_t)M.out_off - (int64_t)kRingSyms + (sym & 0x7fff);
				cc--;
			}
			hist[i] = (uint8_t)byte;
		}
	}
}

// ---- 5. the real decode of the chained pieces: inflate_one with the window in shared memory ----
__global__ void __launch_bounds__(64)
inflate_chain_kernel(const ParPlan P, uint32_t *next_job)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	__shared__ uint32_t s_job;
	const uint32_t n = *P.n_chain;
	for (;;) {
		if (threadIdx.x == 0)
			s_job = atomicAdd(next_job, 1u);
		__syncthreads();
		const uint32_t j = s_job;
		__syncthreads();
		if (j >= n)
			break;
		// the longest-running piece first: the last one runs to the end of the job
		const uint32_t k = n - 1 - j;
		const InflateJob J = P.cjobs[k];
		inflate_duo(J, P.couts[k], smem_raw);
	}
}

// ---- 6. one result for the caller ----
__global__ void inflate_finish_kernel(const ParPlan P)
{
	const uint32_t lane = threadIdx.x;          // one warp
	const uint32_t n = *P.n_chain;
	bool good = true;
	for (uint32_t k = lane; k + 1 < n; k += 32) {
		const InflateOut &o = P.couts[k];
		const ChainMeta &M = P.meta[k];
		const uint64_t end = (uint64_t)o.end_bit_lo | ((uint64_t)o.end_bit_hi << 32);
		if (o.rc != 0 || !(o.flags & kInflateMapStop) || o.out_len != M.exp_len || end != M.exp_end)
			good = false;
	}
	good = __all_sync(0xffffffffu, good);
	if (lane != 0)
		return;
	InflateJob R = P.job;
	R.wrap |= kWrapSkip;
	if (n == 0) {
		InflateOut O = *P.head_out;
		if (O.flags & kInflateMapStop) {
			// stopped on a candidate that the list does not hold (cannot happen): serial
			O.rc = kInflateRetry;
			R.wrap &= ~kWrapSkip;
		}
		*P.final_out = O;
		*P.retry_job = R;
		return;
	}
	const ChainMeta &L = P.meta[n - 1];
	InflateOut O = P.couts[n - 1];
	if (O.flags & kInflateMapStop)
		good = false;
	if (!good) {
		O.rc = kInflateRetry;
		R.wrap &= ~kWrapSkip;
	} else {
		O.out_len += L.out_off;
		const bool job = (P.job.wrap & 0xff) == kWrapJob;
		if (job || (O.rc == 0 && (O.flags & 1)))
			O.in_used += (uint32_t)(L.bit >> 3);
	}
	*P.final_out = O;
	*P.retry_job = R;
}

// ---- dry run (kWrapDry: where does the member end, how long is its output): pieces 1.. were only ever counted, so the
// chain itself is the answer when its last piece ran into the final block ----
__global__ void inflate_dry_finish_kernel(const ParPlan P)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	const uint32_t n = *P.n_chain;
	InflateJob R = P.job;
	R.wrap |= kWrapSkip;
	InflateOut O = *P.head_out;
	bool good = true;
	if (n == 0) {
		good = !(O.flags & kInflateMapStop);
	} else {
		const ChainMeta &L = P.meta[n - 1];
		const SpecOut &S = P.spec[L.unit];
		const uint32_t wrap = (O.flags >> 8) & 0xff;
		const uint64_t p = (S.end_bit + 7) >> 3;
		const uint64_t tr = wrap == NXGPU_WRAP_GZIP ? 8 : wrap == NXGPU_WRAP_ZLIB ? 4 : 0;
		good = S.status == kSpecFinal && S.max_back <= (uint64_t)P.job.hist_len + L.out_off &&
		       (uint64_t)L.out_off + S.out_len <= P.job.dst_cap && p + tr <= P.job.src_len;
		if (good) {
			const uint8_t *s = P.job.src + p;
			O.rc = 0;
			O.out_len = L.out_off + S.out_len;
			O.in_used = (uint32_t)(p + tr);
			O.flags = 1 | (wrap << 8);
			O.trailer_crc = 0; O.trailer_isize = 0;
			if (wrap == NXGPU_WRAP_GZIP) {
				O.trailer_crc = s[0] | (uint32_t)s[1] << 8 | (uint32_t)s[2] << 16 | (uint32_t)s[3] << 24;
				O.trailer_isize = s[4] | (uint32_t)s[5] << 8 | (uint32_t)s[6] << 16 | (uint32_t)s[7] << 24;
			} else if (wrap == NXGPU_WRAP_ZLIB) {
				O.trailer_crc = (uint32_t)s[0] << 24 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 8 | s[3];
			}
		}
	}
	if (!good) {
		O.rc = kInflateRetry;
		R.wrap &= ~kWrapSkip;
	}
	*P.final_out = O;
	*P.retry_job = R;
}
