#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

Workload (configs[1]): 1 GiB of makedata.c-style synthetic text (seed 1, `-b 30`, seeded with
alice29.txt) deflated at level 6 with dynamic Huffman in 256 KiB chunks primed with the previous
32 KiB, stitched into one gzip member with crc32; one "step" = one pass over the 1 GiB that is
already resident in HBM.  `value` = uncompressed GB/s (CUDA events on the engine's stream, max over
ranks); `e2e` = the same metric through the host-pointer C-ABI call (pinned host buffer in,
compressed host buffer out, copies inside the timed region).  Level 1 deflate and the batched
inflate of 64 KiB gzip members (configs[2]) are measured in the same run and reported under
`extra`.

N > 1 (torchrun, one rank per GPU, NCCL): every rank deflates its own 1 GiB slice of an N GiB
stream (weak scaling), sizes are all-gathered and exclusive-scanned, the compressed pieces are
sent to rank 0 at their scanned offsets over NCCL P2P and the per-rank CRCs are folded with
crc32_combine (configs[3]).

`--impl reference` times the reference's own software path (lib/sw_zlib.c -> system zlib through
oracle/_ref/libnxz_ref.so when it was built, else libz directly) on the host cores.
"""
import argparse
import ctypes as C
import gzip
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
GIB = 1 << 30
CHUNK = 262144


def load_pkg():
    spec = importlib.util.spec_from_file_location(
        "power_gzip_b200", os.path.join(ROOT, "power-gzip_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "power-gzip_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["power_gzip_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def ncu_traffic_per_input_byte():
    """DRAM bytes (read + write) per uncompressed input byte of deflate_kernel, from the committed
    `ncu --set full` capture summarised in profiles/ (None if absent)."""
    p = os.path.join(ROOT, "profiles", "ncu_deflate_summary.json")
    try:
        j = json.load(open(p))
        return (j["dram_bytes_read"] + j["dram_bytes_write"]) / j["input_bytes"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"          # /opt/skills/guides/B200_PROFILING.md


class ClockSampler:
    """nvidia-smi polled every 100 ms in ONE long-lived process (spawning it per sample takes longer than a
    step); only samples that fall inside [mark_start, mark_end] — the timed region — are summarised."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.t0, self.t1 = gpu, [], None, None
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append((time.time(), parts))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        lo, hi = (self.t0 or 0) - 0.05, (self.t1 or time.time()) + 0.15
        rows = [r for t, r in self.rows if lo <= t <= hi] or [r for _, r in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for i, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(len(r) > i and r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


def oracle_lib():
    p = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(p):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(p)
    lib.oracle_cpu_baseline.restype = C.c_double
    lib.oracle_cpu_baseline.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    return lib


def cpu_baseline(data_addr, nbytes, level, threads, mode=0):
    """Reference software path on `threads` host cores over `nbytes` of the workload (bounded sample)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libnxz_ref.so")
    kind = "reference" if os.path.exists(ref) else "port"
    lib_path = ref if kind == "reference" else "libz.so.1"
    os.environ["NX_GZIP_TYPE_SELECTOR"] = "1"       # libnxz: software (sw_zlib.c) path
    os.environ.setdefault("NX_GZIP_LOGFILE", "/tmp/nx.log")
    out = C.c_uint64()
    secs = oracle_lib().oracle_cpu_baseline(lib_path.encode(), data_addr, nbytes, CHUNK, level, mode, threads, C.byref(out))
    if secs <= 0 and kind == "reference":
        kind, lib_path = "port", "libz.so.1"
        secs = oracle_lib().oracle_cpu_baseline(lib_path.encode(), data_addr, nbytes, CHUNK, level, mode, threads, C.byref(out))
    return secs, out.value, kind


def gen_input(pg, log2):
    alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
    lib = pg.load_library()
    cap = (1 << log2) + 16
    hptr = C.c_void_p()
    rc = lib.nxgpu_host_alloc(cap, C.byref(hptr))       # pinned: the e2e leg copies from here
    if rc != 0:
        raise RuntimeError("pinned allocation failed: " + pg.last_error())
    seed_buf = C.create_string_buffer(alice, len(alice))
    n = lib.nxgpu_makedata(1, log2, C.addressof(seed_buf), len(alice), hptr, cap)
    assert n == 1 << log2, n
    return hptr.value, n


def workload_config(n, world):
    """the `config` object both arms print (BASELINE.json configs[1])"""
    return {"workload": "1 GiB makedata text (seed 1) per GPU, deflate level 6, dynamic Huffman, 256 KiB chunks primed with 32 KiB, one gzip member + crc32",
            "bytes_per_gpu": n, "chunk": CHUNK, "level": 6, "l2": "inputs (1 GiB) larger than the 126 MB L2; no flush needed",
            "parallelism": f"chunk-range x{world}" if world > 1 else "1 GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pg = load_pkg()
    threads = os.cpu_count() or 1
    sample = min(GIB, 8 * 1024 * 1024 * threads)     # about 10-30 s of level-6 zlib in total
    sample = max(CHUNK * threads, sample // CHUNK * CHUNK)
    log2 = 30 if sample > (1 << 28) else 28
    alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
    data = pg.makedata(1, log2, alice)
    buf = C.create_string_buffer(data[:sample], sample)
    addr = C.addressof(buf)
    per = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        secs, outb, kind = cpu_baseline(addr, sample, 6, threads)
        if i >= args.warmup:
            per.append(secs)
    t = sum(per) / len(per)
    val = sample / t / 1e9
    line = {"impl": "reference", "metric": "deflate uncompressed GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(GIB, args.gpus),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind,
                             "sample": f"first {sample >> 20} MiB of the workload, compress2(level 6) per 256 KiB piece, zlib {zlib.ZLIB_RUNTIME_VERSION}"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--log2", type=int, default=30, help="log2 of the per-GPU input size (default 1 GiB)")
    ap.add_argument("--skip-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = load_pkg()
    from power_gzip_b200 import multi
    eng = pg.Engine(local)
    lib = eng.lib
    hbm_peak, peak_kind = peaks()

    # ---- inputs: synthetic makedata text, resident in HBM before the timed region ----
    hsrc, n = gen_input(pg, args.log2)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    eng._check(lib.nxgpu_memcpy_h2d(eng.ctx, src.data_ptr(), hsrc, n), "h2d")
    cap = eng.deflate_bound(n, CHUNK)
    dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    hdst = C.c_void_p()
    lib.nxgpu_host_alloc(cap, C.byref(hdst))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def deflate_step(level, gather_to_rank0=True):
        """one pass of the hot path; returns (device ms, StreamResult)"""
        eng.timer_start()
        if world == 1:
            res = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=level, wrap=pg.WRAP_GZIP, chunk=CHUNK)
            return eng.timer_stop(), res, res.out_len
        # rank r owns bytes [r*n, (r+1)*n) of the N*n stream: raw deflate of its slice, joiner unless last
        wrap = pg.WRAP_RAW if rank == world - 1 else pg.WRAP_RAW_CONT
        res = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=level, wrap=wrap, chunk=CHUNK)
        ms = eng.timer_stop()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        # sizes + CRCs all-gathered, exclusive scan, ranges sent to rank 0 at their offsets, CRCs folded
        # (power-gzip_b200/multi.py; same functions the gloo tests cover)
        out, total, crc, _ = multi.stitch_to_rank0(dist, torch, dst, int(res.out_len), int(res.crc32), n,
                                                   eng.crc32_combine, stitched[0])
        if rank == 0:
            stitched[0] = out
            last_crc[0] = crc
        t1.record(); t1.synchronize()
        return ms + t0.elapsed_time(t1), res, total

    stitched, last_crc = [None], [0]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        per = []
        for _ in range(steps):
            per.append(fn())
        barrier()
        return per

    # ---- headline: deflate level 6, device resident ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    eng.kernel_time_reset()
    for _ in range(args.warmup):
        deflate_step(6)
    barrier()
    eng.kernel_time_reset()
    l0 = eng.launch_count()
    sampler.mark_start()
    per = []
    last = None
    for _ in range(args.steps):
        ms, res, total = deflate_step(6)
        per.append(ms); last = (res, total)
    barrier()
    sampler.mark_end()
    launches = eng.launch_count() - l0
    kms, kn = eng.kernel_time("deflate")
    step_ms = sum(per) / len(per)
    if world > 1:
        t = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t[0])
    value = world * n / step_ms / 1e6
    res6, total6 = last
    comp_bytes = res6.out_len
    k_avg_ms = kms / max(kn, 1)
    achieved = (n + comp_bytes) / k_avg_ms / 1e6                          # algorithmic U + C bytes per launch
    extra = {"ratio_level6": n / comp_bytes, "tokens_level6": int(res6.n_tokens), "deflate_kernel_ms": k_avg_ms,
             "deflate_kernel_GBps_uncompressed": n / k_avg_ms / 1e6}

    # ---- end to end through the host-pointer C-ABI call (rank-local) ----
    res_h = pg.StreamResult()

    def e2e_step():
        t = time.perf_counter()
        eng._check(lib.nxgpu_deflate_stream(eng.ctx, hsrc, n, hdst, cap, 6, pg.WRAP_GZIP, CHUNK, None, C.byref(res_h), pg.MEM_HOST), "e2e deflate")
        return (time.perf_counter() - t) * 1e3
    e2e_per = timed(e2e_step, max(2, min(args.steps, 3)), 1)
    e2e_ms = sum(e2e_per) / len(e2e_per)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    sampler.stop()
    e2e = {"value": world * n / e2e_ms / 1e6, "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(res_h.out_len)}

    cpu = None
    if rank == 0:
        # ---- correctness of what was just timed: zlib must decode the stitched stream ----
        check_n = min(n, 64 << 20)
        if world == 1:
            blob = bytes(C.string_at(hdst.value, int(res_h.out_len)))
            d = zlib.decompressobj(31)
            got = d.decompress(blob, check_n)
            want = C.string_at(hsrc, check_n)
            assert got == want, "zlib does not reproduce the input from the GPU stream"
            assert res_h.crc32 == res6.crc32
        extra["verified"] = f"zlib inflate of the e2e output reproduces the first {check_n >> 20} MiB; crc32 {res6.crc32:08x}"
        if world > 1 and stitched[0] is not None:
            # the stitched N-GPU stream is ONE gzip member: inflate all of it with zlib, compare crc32 + length
            blob = stitched[0][:total6].cpu().numpy().tobytes()
            d = zlib.decompressobj(31)
            crc, length, pos = 0, 0, 0
            while pos < len(blob):
                piece = d.decompress(blob[pos:pos + (8 << 20)])
                pos += 8 << 20
                crc = zlib.crc32(piece, crc); length += len(piece)
            tail = d.flush(); crc = zlib.crc32(tail, crc); length += len(tail)
            assert d.eof and length == world * n and crc == last_crc[0], (d.eof, length, world * n, hex(crc), hex(last_crc[0]))
            extra["verified_stitched"] = f"zlib inflates the {world}-GPU stream: {length} bytes, crc32 {crc:08x} == folded crc"

    if not args.skip_extra:
        # ---- level 1 ----
        eng.kernel_time_reset()
        p1 = timed(lambda: deflate_step(1)[0], max(2, min(args.steps, 3)), 1)
        k1, kn1 = eng.kernel_time("deflate")
        r1 = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=1, wrap=pg.WRAP_GZIP, chunk=CHUNK)
        extra["deflate_level1_GBps"] = world * n / (sum(p1) / len(p1)) / 1e6
        extra["ratio_level1"] = n / r1.out_len
        # ---- batched inflate of independent 64 KiB gzip members (configs[2], scaled to the input) ----
        M = 65536
        nm = 100000 if args.log2 >= 30 else n // M          # configs[2]: 100 k independent 64 KiB gzip members
        sample_members = min(n // M, 2048)
        hb = C.string_at(hsrc, sample_members * M)
        blobs = [zlib.compress(hb[i * M:(i + 1) * M], 6, wbits=31) for i in range(sample_members)]
        reps = -(-nm // sample_members)
        packed = b"".join(blobs * reps)
        lens = ([len(b) for b in blobs] * reps)[:nm]
        packed = packed[:sum(lens)]
        comp = torch.empty(len(packed), dtype=torch.uint8, device="cuda")
        pk = C.create_string_buffer(packed, len(packed))
        eng._check(lib.nxgpu_memcpy_h2d(eng.ctx, comp.data_ptr(), C.addressof(pk), len(packed)), "h2d")
        out = torch.empty(nm * M, dtype=torch.uint8, device="cuda")
        items = (pg.InflateItem * nm)()
        o = 0
        for i in range(nm):
            items[i] = pg.InflateItem(comp.data_ptr() + o, lens[i], out.data_ptr() + i * M, M, pg.WRAP_GZIP, 0)
            o += lens[i]
        ires = (pg.InflateResult * nm)()

        def inflate_step():
            eng.timer_start()
            eng._check(lib.nxgpu_inflate_batch(eng.ctx, items, nm, ires, pg.MEM_DEVICE), "inflate")
            return eng.timer_stop()
        eng.kernel_time_reset()
        pi = timed(inflate_step, max(2, min(args.steps, 3)), 1)
        ki, kni = eng.kernel_time("inflate")
        assert all(r.rc == 0 and r.out_len == M for r in ires), "batched inflate failed"
        assert ires[0].crc32 == zlib.crc32(hb[:M])
        extra["inflate_members"] = nm
        extra["inflate_64KiB_members_GBps"] = world * nm * M / (sum(pi) / len(pi)) / 1e6
        extra["inflate_kernel_GBps"] = nm * M / (ki / max(kni, 1)) / 1e6
        extra["inflate_roofline_frac"] = (nm * M + len(packed)) / (ki / max(kni, 1)) / 1e6 / hbm_peak
        # ---- the same members WITHOUT their index: one concatenated multi-member buffer, members discovered on the
        # device (candidate headers, dry decoding run, chain from offset 0), then inflated as one batch ----
        if world == 1:
            ng = min(nm, 32768)
            glen = sum(lens[:ng])
            tot, mem_n = C.c_uint64(), C.c_uint32()

            def gunzip_step():
                eng.timer_start()
                eng._check(lib.nxgpu_gunzip_concat(eng.ctx, comp.data_ptr(), glen, out.data_ptr(), ng * M, C.byref(tot), C.byref(mem_n),
                                                   pg.MEM_DEVICE), "gunzip_concat")
                return eng.timer_stop()
            pgz = timed(gunzip_step, 2, 1)
            assert tot.value == ng * M and mem_n.value == ng
            extra["gunzip_concat_GBps"] = {"value": ng * M / (sum(pgz) / len(pgz)) / 1e6, "members": ng,
                                           "note": "members discovered on the device, no index given"}
        del comp, out
        # ---- the same members end to end through the host-pointer call: pinned host buffers, H2D of the compressed
        # bytes and D2H of the inflated ones inside the timed region (wall clock) ----
        if world == 1:
            nh = min(nm, 16384)
            hin_len = sum(lens[:nh])
            hin, hout = C.c_void_p(), C.c_void_p()
            lib.nxgpu_host_alloc(hin_len, C.byref(hin)); lib.nxgpu_host_alloc(nh * M, C.byref(hout))
            C.memmove(hin, packed[:hin_len], hin_len)
            hitems = (pg.InflateItem * nh)()
            o = 0
            for i in range(nh):
                hitems[i] = pg.InflateItem(hin.value + o, lens[i], hout.value + i * M, M, pg.WRAP_GZIP, 0)
                o += lens[i]
            hres = (pg.InflateResult * nh)()
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                eng._check(lib.nxgpu_inflate_batch(eng.ctx, hitems, nh, hres, pg.MEM_HOST), "inflate e2e")
                d = time.perf_counter() - t0
                best = d if best is None else min(best, d)
            assert all(r.rc == 0 and r.out_len == M for r in hres) and C.string_at(hout.value, M) == hb[:M]
            extra["inflate_e2e_host_GBps"] = {"value": nh * M / best / 1e9, "members": nh, "h2d_bytes": hin_len, "d2h_bytes": nh * M}
            lib.nxgpu_host_free(hin); lib.nxgpu_host_free(hout)
        # ---- ONE big member inflated as the segments of its sync-point index: the workload
        # deflated with independent chunks, then nxgpu_inflate_stream over the whole stream, device resident ----
        if world == 1:
            nchunks = -(-n // CHUNK)
            idx = (C.c_uint64 * (nchunks + 1))()
            ri = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=6, wrap=pg.WRAP_GZIP | pg.STREAM_INDEPENDENT,
                                           chunk=CHUNK, index=idx)
            back = torch.empty(n, dtype=torch.uint8, device="cuda")

            def inflate_stream_step():
                eng.timer_start()
                r = eng.inflate_stream_device(dst.data_ptr(), ri.out_len, back.data_ptr(), n, idx, nchunks, chunk=CHUNK, wrap=pg.WRAP_GZIP)
                assert r.out_len == n and r.crc32 == ri.crc32
                return eng.timer_stop()
            ps = timed(inflate_stream_step, max(2, min(args.steps, 3)), 1)
            extra["ratio_level6_independent_chunks"] = n / ri.out_len
            extra["inflate_stream_one_member_GBps"] = n / (sum(ps) / len(ps)) / 1e6
            assert bool(torch.equal(back[: 1 << 24], src[: 1 << 24]))
            del back
        # ---- crc32 + adler32 (configs[4]): one buffer of 4 KiB .. 1 GiB, and a storm of small buffers ----
        sweep = {}
        for lg in range(12, args.log2 + 1, 2):
            sz = 1 << lg
            best = None
            for _ in range(3):
                eng.timer_start()
                r = eng.checksum_batch([(src.data_ptr(), sz, 0, 1)], mem=pg.MEM_DEVICE)
                ms = eng.timer_stop()
                best = ms if best is None else min(best, ms)
            sweep[str(sz)] = round(sz / best / 1e6, 3)
        assert r[0][0] == res6.crc32, "crc32 of the whole buffer differs from the deflate path's"
        extra["crc32_adler32_GBps_by_size"] = sweep
        small = [(src.data_ptr() + (i * 4099) % (n - 70000), 1 << (12 + i % 5), 0, 1) for i in range(100000)]
        sarr = (pg.CksumItem * len(small))(*[pg.CksumItem(a, l, cs, ads) for a, l, cs, ads in small])
        sres = (pg.CksumResult * len(small))()
        dt = None
        for _ in range(3):                               # the C-ABI call alone (arrays marshalled once), best of 3
            t0 = time.perf_counter()
            eng._check(lib.nxgpu_checksum_batch(eng.ctx, sarr, len(small), sres, pg.MEM_DEVICE), "checksum storm")
            d = time.perf_counter() - t0
            dt = d if dt is None else min(dt, d)
        probe = small[12345]
        assert sres[12345].crc32 == zlib.crc32(C.string_at(hsrc + (probe[0] - src.data_ptr()), probe[1]))
        extra["checksum_storm"] = {"buffers": len(small), "bytes": sum(x[1] for x in small), "buffers_per_s": round(len(small) / dt),
                                   "GBps": round(sum(x[1] for x in small) / dt / 1e9, 2), "note": "4-64 KiB buffers, one batched call, host wall clock"}

    if rank == 0 and world == 1 and not args.skip_extra:
        # ---- many concurrent small z_streams through the UNCHANGED zlib surface (configs[4]): the reference's host
        # code over the GPU engine, test/test_multithread_stress.c pattern; descriptors coalesce inside nxu_run_job ----
        gpu_nxz = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")
        if os.path.exists(gpu_nxz):
            import subprocess
            try:
                env = dict(os.environ, NX_GZIP_LOGFILE="/tmp/nx_bench.log")
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nx_dropin_driver.py"), gpu_nxz, "stress", "64", "2"],
                                   capture_output=True, text=True, timeout=300, env=env)
                st = json.loads(p.stdout.strip().splitlines()[-1])
                extra["zstream_storm"] = {"threads": st["threads"], "calls_per_s": round(st["calls_per_s"]), "MBps": round(st["MBps"], 1),
                                          "descriptors": st.get("jobs"), "gpu_batches": st.get("batches"), "largest_batch": st.get("max_batch"),
                                          "errors": len(st["errors"]),
                                          "note": "compress()/uncompress() of 4 KiB-1 MiB buffers from 64 threads via libnxz host code + nxu_run_job, Python harness"}
            except Exception as e:                      # noqa: BLE001 - an extra, never fatal
                extra["zstream_storm"] = {"error": repr(e)[:200]}
            try:
                # samples/bench_initend.c: deflateInit2/deflateEnd and inflateInit2/inflateEnd pairs over the GPU engine
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nx_dropin_driver.py"), gpu_nxz, "initend", "1000"],
                                   capture_output=True, text=True, timeout=300, env=env)
                ie = json.loads(p.stdout.strip().splitlines()[-1])
                extra["zstream_init_end"] = {k: round(v, 2) for k, v in ie.items() if k.endswith(("_us", "_ms"))}
            except Exception as e:                      # noqa: BLE001
                extra["zstream_init_end"] = {"error": repr(e)[:200]}

    if rank == 0:
        threads = os.cpu_count() or 1
        sample = min(n, max(CHUNK * threads, 4 * 1024 * 1024 * threads))
        secs, outb, kind = cpu_baseline(hsrc, sample, 6, threads)
        cpu = {"value": sample / secs / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
               "sample": f"first {sample >> 20} MiB of the workload, compress2(level 6) per 256 KiB piece on {threads} threads, zlib {zlib.ZLIB_RUNTIME_VERSION}",
               "ratio": sample / max(outb, 1)}
        line = {
            "metric": "deflate uncompressed GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(n, world),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": (ncu_traffic_per_input_byte() * n) if ncu_traffic_per_input_byte() else None,
                         "peak_source": peak_kind, "kernel": "deflate_kernel",
                         "note": "algorithmic bytes = U + C per launch (SURVEY.md §8d); LZ77 search is latency/issue bound, not HBM bound"},
            "cpu_baseline": cpu, "clocks": sampler.summary(), "extra": extra,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
