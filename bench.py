#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

N = 1 (configs[1]): 1 GiB of makedata.c-style synthetic text (seed 1, `-b 30`, seeded with alice29.txt) deflated at
level 6 with dynamic Huffman in 256 KiB chunks primed with the previous 32 KiB, one gzip member with crc32.  One "step"
= one pass over the 1 GiB already resident in HBM.  `value` = uncompressed GB/s (CUDA events on the engine's stream);
`e2e` = the same through the host-pointer C-ABI call nxgpu_deflate_stream (pinned host buffer in, compressed host
buffer out, copies inside the timed region, wall clock over --steps).  Level 1 (the other half of configs[1]) gets its
own value / e2e / roofline / cpu_baseline under `extra.level1`.  configs[2] (100 000 distinct 64 KiB gzip members of a
seed-4 stream, each made by zlib level 6) and configs[4] (crc32/adler32 4 KiB .. 4 GiB; the reference's own
test_multithread_stress over the GPU engine with the same program over the software path beside it) are under `extra`.

N > 1 (configs[3], torchrun, one rank per GPU): a 16 GiB stream (seed 5, `-b 34`) is cut into N contiguous ranges;
nxgpu_team_deflate (C-ABI, csrc/nxgpu_team.cu) deflates them in parallel, exchanges the compressed sizes on the
devices, writes every range at its scanned offset and folds the CRCs into ONE gzip member.  `value`: ranges resident
in HBM, member assembled in rank 0's HBM over NVLink P2P, device time (max over ranks).  `e2e`: ranges in pinned host
memory, member assembled in shared HOST memory (each GPU writes its part over its own PCIe link), wall clock, stitch
included.  torch.distributed (NCCL) is the plumbing: barriers and the max over ranks.  Total work is fixed as N grows
("strong"); N = 1 keeps the 1 GiB workload.

`--impl reference` times the reference's own software path (lib/sw_zlib.c -> system zlib through
oracle/_ref/libnxz_ref.so when it was built, else libz directly) on the host cores; it does not load the product.
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
import tempfile
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
GIB = 1 << 30
CHUNK = 262144
MEMBER = 65536


def load_pkg():
    spec = importlib.util.spec_from_file_location(
        "power_gzip_b200", os.path.join(ROOT, "power-gzip_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "power-gzip_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["power_gzip_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def ncu_traffic_per_input_byte(level=6):
    """DRAM bytes (read + write) per uncompressed input byte of deflate_kernel, from the committed
    `ncu --set full` capture summarised in profiles/ (None if absent)."""
    for name in (f"ncu_deflate_r2_l{level}_summary.json", "ncu_deflate_summary.json") if level == 6 else (f"ncu_deflate_r2_l{level}_summary.json",):
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", name)))
            return (j["dram_bytes_read"] + j["dram_bytes_write"]) / j["input_bytes"]
        except Exception:
            continue
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


# ---------------------------------------------------------------------------------------------------------------
# the CPU arm: oracle/ (the checker) may be executed here and only here
# ---------------------------------------------------------------------------------------------------------------
def oracle_lib():
    p = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(p):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(p)
    lib.oracle_cpu_baseline.restype = C.c_double
    lib.oracle_cpu_baseline.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.oracle_makedata.restype = C.c_uint64
    lib.oracle_makedata.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64]
    return lib


def cpu_baseline(data_addr, nbytes, level, threads, mode=0, piece=CHUNK):
    """Reference software path on `threads` host cores over `nbytes` of the workload (bounded sample).
    mode 0 compress2 per piece, 1 uncompress per piece, 2 crc32, 3 adler32 (oracle/cpu_baseline.c)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libnxz_ref.so")
    kind = "reference" if os.path.exists(ref) else "port"
    lib_path = ref if kind == "reference" else "libz.so.1"
    os.environ["NX_GZIP_TYPE_SELECTOR"] = "1"       # libnxz: software (sw_zlib.c) path
    os.environ.setdefault("NX_GZIP_LOGFILE", "/tmp/nx.log")
    out = C.c_uint64()
    secs = oracle_lib().oracle_cpu_baseline(lib_path.encode(), data_addr, nbytes, piece, level, mode, threads, C.byref(out))
    if secs <= 0 and kind == "reference":
        kind, lib_path = "port", "libz.so.1"
        secs = oracle_lib().oracle_cpu_baseline(lib_path.encode(), data_addr, nbytes, piece, level, mode, threads, C.byref(out))
    return secs, out.value, kind


def alice_bytes():
    return gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())


def workload_config(world, total_bytes, per_gpu):
    if world == 1:
        return {"workload": "1 GiB makedata text (seed 1, -b 30), deflate level 6, dynamic Huffman, 256 KiB chunks primed with 32 KiB, one gzip member + crc32",
                "bytes_per_gpu": per_gpu, "chunk": CHUNK, "level": 6, "l2": "inputs (1 GiB) larger than the 126 MB L2; no flush needed",
                "parallelism": "1 GPU"}
    return {"workload": f"{total_bytes >> 30} GiB makedata text (seed 5, -b 34) deflate level 6 + crc32 combine, chunk ranges over {world} GPUs, stitched into one gzip member",
            "bytes_total": total_bytes, "bytes_per_gpu": per_gpu, "chunk": CHUNK, "level": 6,
            "l2": f"inputs ({per_gpu >> 30} GiB per GPU) larger than the 126 MB L2; no flush needed",
            "parallelism": f"chunk-range x{world}, sizes scanned on the devices, ranges written at their offsets (NVLink P2P / per-GPU PCIe), CRCs folded",
            "note": "N=1 runs the 1 GiB configs[1] workload; N>1 runs configs[3] at a fixed 16 GiB"}


def run_reference(args):
    """The reference's software path on the host cores.  Loads oracle/ only (never the product library)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = min(GIB, 8 * 1024 * 1024 * threads)     # about 10-30 s of level-6 zlib in total
    sample = max(CHUNK * threads, sample // CHUNK * CHUNK)
    world = args.gpus
    seed, log2 = (1, 30 if sample > (1 << 28) else 28) if world == 1 else (5, 30 if sample > (1 << 28) else 28)
    alice = alice_bytes()
    buf = C.create_string_buffer((1 << log2) + 16)
    n = oracle_lib().oracle_makedata(seed, log2, alice, len(alice), buf, len(buf))
    assert n == 1 << log2
    addr = C.addressof(buf)
    per = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        secs, outb, kind = cpu_baseline(addr, sample, 6, threads)
        if i >= args.warmup:
            per.append(secs)
    t = sum(per) / len(per)
    val = sample / t / 1e9
    total = (1 << args.total_log2) if world > 1 else GIB
    line = {"impl": "reference", "metric": "deflate uncompressed GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world, total, total // world),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind,
                             "sample": f"first {sample >> 20} MiB of the workload's generator (seed {seed}), compress2(level 6) per 256 KiB piece on {threads} threads, zlib {zlib.ZLIB_RUNTIME_VERSION}"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--log2", type=int, default=30, help="N=1: log2 of the input size (default 1 GiB)")
    ap.add_argument("--total-log2", type=int, default=34, help="N>1: log2 of the whole stream (default 16 GiB, configs[3])")
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
    eng = pg.Engine(local)
    lib = eng.lib
    hbm_peak, peak_kind = peaks()
    alice = alice_bytes()
    seed_buf = C.create_string_buffer(alice, len(alice))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- inputs: synthetic makedata text in pinned host memory, copied into HBM before the timed region ----
    if world == 1:
        n = 1 << args.log2
        total = n
        lo = 0
        gen_seed, gen_log2 = 1, args.log2
    else:
        total = 1 << args.total_log2
        n_chunks = total // CHUNK
        per_rank = (n_chunks + world - 1) // world * CHUNK
        lo = min(total, rank * per_rank)
        n = min(total, lo + per_rank) - lo
        gen_seed, gen_log2 = 5, args.total_log2
    hptr = C.c_void_p()
    if lib.nxgpu_host_alloc(n + 16, C.byref(hptr)) != 0:
        raise RuntimeError("pinned allocation failed: " + pg.last_error())
    hsrc = hptr.value
    t_gen = time.time()
    if world == 1:
        got = lib.nxgpu_makedata(gen_seed, gen_log2, C.addressof(seed_buf), len(alice), hsrc, n + 16)
    else:
        got = lib.nxgpu_makedata_range(gen_seed, gen_log2, C.addressof(seed_buf), len(alice), lo, lo + n, hsrc)
    assert got == n, (got, n)
    t_gen = time.time() - t_gen
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    eng._check(lib.nxgpu_memcpy_h2d(eng.ctx, src.data_ptr(), hsrc, n), "h2d")
    torch.cuda.synchronize()
    extra = {"input_generation_s": round(t_gen, 1)}

    # ---- the step ----
    team_dev = team_host = None
    if world == 1:
        cap = eng.deflate_bound(n, CHUNK)
        dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
        hdst = C.c_void_p()
        lib.nxgpu_host_alloc(cap, C.byref(hdst))
        res_h = pg.StreamResult()

        def deflate_step(level):
            """one pass of the hot path, device resident; returns (device ms, compressed bytes, crc32, tokens)"""
            eng.timer_start()
            res = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=level, wrap=pg.WRAP_GZIP, chunk=CHUNK)
            return eng.timer_stop(), int(res.out_len), int(res.crc32), int(res.n_tokens)

        def e2e_step(level):
            t = time.perf_counter()
            eng._check(lib.nxgpu_deflate_stream(eng.ctx, hsrc, n, hdst, cap, level, pg.WRAP_GZIP, CHUNK, None, C.byref(res_h), pg.MEM_HOST), "e2e deflate")
            return (time.perf_counter() - t) * 1e3, int(res_h.out_len), int(res_h.crc32)
    else:
        # the member is about total/5.5 bytes on this text; a third of the input leaves slack (NXGPU_E_BUF otherwise)
        dst_cap = total // 3 + (1 << 20)
        name = f"nxgpu-bench-{os.environ.get('MASTER_PORT', '0')}-{os.getppid()}"
        team_dev = pg.Team(eng, name + "-d", rank, world, dst_cap, pg.MEM_DEVICE)
        team_host = pg.Team(eng, name + "-h", rank, world, dst_cap, pg.MEM_HOST)

        def deflate_step(level):
            r = team_dev.deflate(src.data_ptr(), n, level=level, wrap=pg.WRAP_GZIP, chunk=CHUNK, src_mem=pg.MEM_DEVICE)
            return float(r.device_ms), int(r.out_len), int(r.crc32), 0

        def e2e_step(level):
            t = time.perf_counter()
            r = team_host.deflate(hsrc, n, level=level, wrap=pg.WRAP_GZIP, chunk=CHUNK, src_mem=pg.MEM_HOST)
            e2e_last[0] = r
            return (time.perf_counter() - t) * 1e3, int(r.out_len), int(r.crc32)
    e2e_last = [None]

    def measure(level, steps, warmup, sampler=None):
        """returns dict(step_ms, e2e_ms, comp_bytes, crc, kernel ms per launch, launches)"""
        for _ in range(warmup):
            deflate_step(level)
        barrier()
        eng.kernel_time_reset()
        l0 = eng.launch_count()
        if sampler:
            sampler.mark_start()
        per = []
        out = None
        for _ in range(steps):
            ms, comp, crc, ntok = deflate_step(level)
            per.append(ms); out = (comp, crc, ntok)
            if world > 1:
                dist.barrier()          # a collective step ends when the slowest rank is done
        barrier()
        if sampler:
            sampler.mark_end()
        launches = eng.launch_count() - l0
        kms, kn = eng.kernel_time("deflate")
        step_ms = max_over_ranks(sum(per) / len(per))
        # end to end: host buffers, copies (and for N>1 the stitch into host memory) inside the timed region
        e2e_step(level)
        barrier()
        eper = []
        for _ in range(steps):
            ms, ecomp, ecrc = e2e_step(level)
            eper.append(ms)
            if world > 1:
                dist.barrier()
        barrier()
        e2e_ms = max_over_ranks(sum(eper) / len(eper))
        return {"step_ms": step_ms, "e2e_ms": e2e_ms, "comp": out[0], "crc": out[1], "ntok": out[2], "e2e_comp": ecomp, "e2e_crc": ecrc,
                "k_ms": kms / max(kn, 1), "launches": int(launches)}

    sampler = ClockSampler(local)
    sampler.start()
    m6 = measure(6, args.steps, args.warmup, sampler)
    sampler.stop()
    value = total / m6["step_ms"] / 1e6
    comp_bytes = m6["comp"]
    # roofline of the dominant kernel: algorithmic U + C bytes of THIS rank's launch over its average duration
    my_comp = comp_bytes if world == 1 else int(e2e_last[0].my_size)
    achieved = (n + my_comp) / m6["k_ms"] / 1e6
    e2e = {"value": total / m6["e2e_ms"] / 1e6, "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": my_comp,
           "ms_per_step": m6["e2e_ms"], "steps": args.steps}
    extra.update({"ratio_level6": total / comp_bytes, "tokens_level6": m6["ntok"], "deflate_kernel_ms": m6["k_ms"],
                  "deflate_kernel_GBps_uncompressed": n / m6["k_ms"] / 1e6})
    traffic6 = ncu_traffic_per_input_byte(6)

    # ---- correctness of what was just timed ----
    if world == 1:
        check_n = min(n, 64 << 20)
        blob = bytes(C.string_at(hdst.value, m6["e2e_comp"]))
        got = zlib.decompressobj(31).decompress(blob, check_n)
        assert got == C.string_at(hsrc, check_n), "zlib does not reproduce the input from the GPU stream"
        assert m6["e2e_crc"] == m6["crc"]
        extra["verified"] = f"zlib inflate of the e2e output reproduces the first {check_n >> 20} MiB; crc32 {m6['crc']:08x}"
    else:
        # every rank checks ITS range of the member that the e2e leg assembled in shared host memory: the range starts byte
        # aligned (the rank in front ended on a joiner) and without history, so zlib decodes it on its own
        r = e2e_last[0]
        check_n = min(n, 64 << 20)
        piece = C.string_at(team_host.dst() + int(r.my_offset), int(r.my_size))
        got = zlib.decompressobj(-15).decompress(piece, check_n)
        ok = got == C.string_at(hsrc, check_n)
        # the folded crc32 in the trailer against zlib's own crc32 of every rank's range, folded the same way
        my_crc = 0
        for o in range(0, n, 1 << 30):
            my_crc = zlib.crc32(C.string_at(hsrc + o, min(1 << 30, n - o)), my_crc)
        meta = torch.tensor([my_crc, n, 1 if ok else 0], dtype=torch.int64, device="cuda")
        allm = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(allm, meta)
        rows = [[int(x) for x in m.tolist()] for m in allm]
        if rank == 0:
            crc = 0
            for c_, n_, ok_ in rows:
                crc = eng.crc32_combine(crc, c_, n_)
                assert ok_, "a rank's range of the stitched member does not inflate to its input"
            member = team_host.dst()
            tail = C.string_at(member + int(r.out_len) - 8, 8)
            assert C.string_at(member, 4) == b"\x1f\x8b\x08\x00"
            assert int.from_bytes(tail[:4], "little") == crc == int(r.crc32), (hex(crc), hex(int(r.crc32)))
            assert int.from_bytes(tail[4:], "little") == total & 0xffffffff
            assert m6["crc"] == crc, "the device-assembled member carries a different crc"
            extra["verified_stitched"] = (f"every rank's range of the {world}-GPU member ({int(r.out_len)} bytes in shared host memory) inflates with zlib to "
                                          f"the first {check_n >> 20} MiB of its input; trailer crc32 {crc:08x} == zlib crc32 of the {total >> 30} GiB folded; ISIZE ok")

    # ---- level 1: the other half of configs[1] ----
    if not args.skip_extra:
        m1 = measure(1, args.steps, 2)
        a1 = (n + (m1["comp"] if world == 1 else int(e2e_last[0].my_size))) / m1["k_ms"] / 1e6
        t1 = ncu_traffic_per_input_byte(1)
        extra["level1"] = {"value": total / m1["step_ms"] / 1e6, "unit": "GB/s", "ms_per_step": m1["step_ms"], "ratio": total / m1["comp"],
                           "e2e": {"value": total / m1["e2e_ms"] / 1e6, "unit": "GB/s", "ms_per_step": m1["e2e_ms"]},
                           "roofline": {"bound": "hbm", "achieved": a1, "peak": hbm_peak, "unit": "GB/s", "frac": a1 / hbm_peak,
                                        "traffic": t1 * n if t1 else None, "kernel": "deflate_kernel"}}

    if world == 1 and not args.skip_extra:
        extras_single_gpu(args, pg, eng, lib, torch, src, hsrc, n, extra, hbm_peak, m6, seed_buf, len(alice))

    cpu = None
    if rank == 0:
        threads = os.cpu_count() or 1
        sample = min(n, max(CHUNK * threads, 4 * 1024 * 1024 * threads))
        secs, outb, kind = cpu_baseline(hsrc, sample, 6, threads)
        cpu = {"value": sample / secs / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
               "sample": f"first {sample >> 20} MiB of rank 0's input, compress2(level 6) per 256 KiB piece on {threads} threads, zlib {zlib.ZLIB_RUNTIME_VERSION}",
               "ratio": sample / max(outb, 1)}
        if "level1" in extra:
            secs1, outb1, kind1 = cpu_baseline(hsrc, sample, 1, threads)
            extra["level1"]["cpu_baseline"] = {"value": sample / secs1 / 1e9, "unit": "GB/s", "cores": threads, "kind": kind1,
                                               "sample": f"first {sample >> 20} MiB, compress2(level 1) per 256 KiB piece", "ratio": sample / max(outb1, 1)}
        line = {
            "metric": "deflate uncompressed GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m6["step_ms"], "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world, total, n),
            "e2e": e2e, "gpu_launches": m6["launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": (traffic6 * n) if traffic6 else None,
                         "peak_source": peak_kind, "kernel": "deflate_kernel",
                         "note": "algorithmic bytes = U + C per launch (SURVEY.md §8d); LZ77 search is latency/issue bound, not HBM bound"},
            "cpu_baseline": cpu, "clocks": sampler.summary(), "extra": extra,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        team_dev.close(); team_host.close()
        dist.barrier()
        dist.destroy_process_group()


def extras_single_gpu(args, pg, eng, lib, torch, src, hsrc, n, extra, hbm_peak, m6, seed_buf, seed_len):
    """configs[2] and configs[4] on one GPU, plus the library-level extras."""
    threads = os.cpu_count() or 1

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        per = [fn() for _ in range(steps)]
        torch.cuda.synchronize()
        return per

    # ---- configs[2]: batched inflate of 100 000 DISTINCT 64 KiB gzip members: member i = bytes [64 KiB * i, 64 KiB * (i+1)) of the
    # makedata stream `-s 4 -b 33`, each compress2'd by zlib level 6 with the gzip wrapper ----
    nm = 100000 if args.log2 >= 30 else max(1024, (n // MEMBER) // 4)
    t0 = time.time()
    raw = C.create_string_buffer(nm * MEMBER)
    got = lib.nxgpu_makedata_range(4, 33, C.addressof(seed_buf), seed_len, 0, nm * MEMBER, C.addressof(raw))
    assert got == nm * MEMBER
    mv = memoryview(raw).cast("B")

    def comp_member(i):
        return zlib.compress(mv[i * MEMBER:(i + 1) * MEMBER], 6, wbits=31)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        blobs = list(ex.map(comp_member, range(nm), chunksize=256))
    lens = [len(b) for b in blobs]
    packed_len = sum(lens)
    hin, hout = C.c_void_p(), C.c_void_p()
    lib.nxgpu_host_alloc(packed_len + 64, C.byref(hin)); lib.nxgpu_host_alloc(nm * MEMBER, C.byref(hout))
    o = 0
    for b in blobs:
        C.memmove(hin.value + o, b, len(b)); o += len(b)
    crc0, crc_last = zlib.crc32(mv[:MEMBER]), zlib.crc32(mv[(nm - 1) * MEMBER: nm * MEMBER])
    extra["inflate_input_prep_s"] = round(time.time() - t0, 1)
    comp = torch.empty(packed_len, dtype=torch.uint8, device="cuda")
    eng._check(lib.nxgpu_memcpy_h2d(eng.ctx, comp.data_ptr(), hin, packed_len), "h2d")
    out = torch.empty(nm * MEMBER, dtype=torch.uint8, device="cuda")
    items = (pg.InflateItem * nm)()
    hitems = (pg.InflateItem * nm)()
    o = 0
    for i in range(nm):
        items[i] = pg.InflateItem(comp.data_ptr() + o, lens[i], out.data_ptr() + i * MEMBER, MEMBER, pg.WRAP_GZIP, 0)
        hitems[i] = pg.InflateItem(hin.value + o, lens[i], hout.value + i * MEMBER, MEMBER, pg.WRAP_GZIP, 0)
        o += lens[i]
    ires = (pg.InflateResult * nm)()

    def inflate_step():
        eng.timer_start()
        eng._check(lib.nxgpu_inflate_batch(eng.ctx, items, nm, ires, pg.MEM_DEVICE), "inflate")
        return eng.timer_stop()
    eng.kernel_time_reset()
    pi = timed(inflate_step, args.steps, 2)
    ki, kni = eng.kernel_time("inflate")
    assert all(r.rc == 0 and r.out_len == MEMBER for r in ires), "batched inflate failed"
    assert ires[0].crc32 == crc0 and ires[nm - 1].crc32 == crc_last
    hres = (pg.InflateResult * nm)()

    def inflate_e2e():
        t = time.perf_counter()
        eng._check(lib.nxgpu_inflate_batch(eng.ctx, hitems, nm, hres, pg.MEM_HOST), "inflate e2e")
        return (time.perf_counter() - t) * 1e3
    pe = timed(inflate_e2e, max(2, min(args.steps, 3)), 1)
    assert all(r.rc == 0 and r.out_len == MEMBER for r in hres) and zlib.crc32(C.string_at(hout.value + (nm - 1) * MEMBER, MEMBER)) == crc_last
    k_ms = ki / max(kni, 1)
    a_inf = (nm * MEMBER + packed_len) / k_ms / 1e6
    inflate = {"members": nm, "member_bytes": MEMBER, "compressed_bytes": packed_len,
               "input": "member i = 64 KiB slice i of makedata -s 4 -b 33, zlib level 6, gzip wrapper (100 000 distinct members)",
               "value": nm * MEMBER / (sum(pi) / len(pi)) / 1e6, "unit": "GB/s", "ms_per_step": sum(pi) / len(pi),
               "e2e": {"value": nm * MEMBER / (sum(pe) / len(pe)) / 1e6, "unit": "GB/s", "h2d_bytes_per_step": packed_len, "d2h_bytes_per_step": nm * MEMBER},
               "roofline": {"bound": "hbm", "achieved": a_inf, "peak": hbm_peak, "unit": "GB/s", "frac": a_inf / hbm_peak, "kernel": "inflate_kernel"}}
    # CPU beside it: uncompress() per member through the reference's software path, bounded sample
    sample_m = min(nm, 512 * threads)
    secs, outb, kind = cpu_baseline(C.addressof(raw), sample_m * MEMBER, 6, threads, mode=1, piece=MEMBER)
    inflate["cpu_baseline"] = {"value": sample_m * MEMBER / secs / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                               "sample": f"first {sample_m} members, uncompress() per member on {threads} threads"}
    extra["inflate_100k_members"] = inflate
    # the same members WITHOUT their index: one concatenated multi-member buffer, members discovered on the device
    ng = min(nm, 32768)
    glen = sum(lens[:ng])
    tot, mem_n = C.c_uint64(), C.c_uint32()

    def gunzip_step():
        eng.timer_start()
        eng._check(lib.nxgpu_gunzip_concat(eng.ctx, comp.data_ptr(), glen, out.data_ptr(), ng * MEMBER, C.byref(tot), C.byref(mem_n), pg.MEM_DEVICE), "gunzip_concat")
        return eng.timer_stop()
    pgz = timed(gunzip_step, 2, 1)
    assert tot.value == ng * MEMBER and mem_n.value == ng
    extra["gunzip_concat_GBps"] = {"value": ng * MEMBER / (sum(pgz) / len(pgz)) / 1e6, "members": ng, "note": "members discovered on the device, no index given"}
    del comp, out, raw, mv
    lib.nxgpu_host_free(hin); lib.nxgpu_host_free(hout)

    # ---- ONE big member inflated as the segments of its sync-point index ----
    cap = eng.deflate_bound(n, CHUNK)
    dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    nchunks = -(-n // CHUNK)
    idx = (C.c_uint64 * (nchunks + 1))()
    ri = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=6, wrap=pg.WRAP_GZIP | pg.STREAM_INDEPENDENT, chunk=CHUNK, index=idx)
    back = torch.empty(n, dtype=torch.uint8, device="cuda")

    def inflate_stream_step():
        eng.timer_start()
        r = eng.inflate_stream_device(dst.data_ptr(), ri.out_len, back.data_ptr(), n, idx, nchunks, chunk=CHUNK, wrap=pg.WRAP_GZIP)
        assert r.out_len == n and r.crc32 == ri.crc32
        return eng.timer_stop()
    ps = timed(inflate_stream_step, max(2, min(args.steps, 3)), 1)
    extra["ratio_level6_independent_chunks"] = n / ri.out_len
    # the ordinary (primed) member of the same input, no index: block starts found on the device, many warps (csrc/inflate_par.cuh)
    rp = eng.deflate_stream_device(src.data_ptr(), n, dst.data_ptr(), cap, level=6, wrap=pg.WRAP_GZIP, chunk=CHUNK)
    if rp.out_len < (1 << 31):
        def inflate_primed_step():
            eng.timer_start()
            r = eng.inflate_batch([pg.InflateItem(dst.data_ptr(), rp.out_len, back.data_ptr(), n, pg.WRAP_GZIP, 0)], mem=pg.MEM_DEVICE)[0]
            ms = eng.timer_stop()
            assert r.rc == 0 and r.out_len == n and r.crc32 == rp.crc32 and (r.flags & 3) == 3
            return ms
        pp = timed(inflate_primed_step, 3, 1)
        extra["inflate_primed_member_no_index_GBps"] = n / (sum(pp) / len(pp)) / 1e6
    extra["inflate_stream_one_member_GBps"] = n / (sum(ps) / len(ps)) / 1e6
    assert bool(torch.equal(back[: 1 << 24], src[: 1 << 24]))
    del back, dst

    # ---- configs[4]: crc32 + adler32, one buffer of 4 KiB .. 4 GiB (prefixes of the seed-5 stream the 16 GiB workload is cut from),
    # device resident and from host memory; CPU beside it ----
    big_log2 = 32 if args.log2 >= 30 else args.log2
    big = 1 << big_log2
    hbig = C.c_void_p()
    lib.nxgpu_host_alloc(big, C.byref(hbig))
    lib.nxgpu_makedata_range(5, 34, C.addressof(seed_buf), seed_len, 0, big, hbig)
    dbig = torch.empty(big, dtype=torch.uint8, device="cuda")
    eng._check(lib.nxgpu_memcpy_h2d(eng.ctx, dbig.data_ptr(), hbig, big), "h2d")
    sweep, sweep_host = {}, {}
    r = None
    for lg in range(12, big_log2 + 1, 2):
        sz = 1 << lg
        best = bh = None
        for _ in range(3):
            eng.timer_start()
            r = eng.checksum_batch([(dbig.data_ptr(), sz, 0, 1)], mem=pg.MEM_DEVICE)
            ms = eng.timer_stop()
            best = ms if best is None else min(best, ms)
            t = time.perf_counter()
            rh = eng.checksum_batch([(hbig.value, sz, 0, 1)], mem=pg.MEM_HOST)
            d = (time.perf_counter() - t) * 1e3
            bh = d if bh is None else min(bh, d)
        assert rh[0] == r[0]
        sweep[str(sz)] = round(sz / best / 1e6, 3)
        sweep_host[str(sz)] = round(sz / bh / 1e6, 3)
    want_crc = 0
    for o in range(0, big, 1 << 30):
        want_crc = zlib.crc32(C.string_at(hbig.value + o, min(1 << 30, big - o)), want_crc)
    assert r[0][0] == want_crc, "crc32 of the 4 GiB buffer differs from zlib's"
    secs_c, _, kind_c = cpu_baseline(hbig.value, min(big, 256 * 1024 * 1024 * min(threads, 8)), 0, threads, mode=2, piece=1 << 26)
    secs_a, _, _ = cpu_baseline(hbig.value, min(big, 256 * 1024 * 1024 * min(threads, 8)), 0, threads, mode=3, piece=1 << 26)
    cs = min(big, 256 * 1024 * 1024 * min(threads, 8))
    extra["crc32_adler32"] = {"GBps_by_size_device": sweep, "GBps_by_size_host_buffer": sweep_host, "largest_bytes": big,
                              "roofline": {"bound": "hbm", "achieved": sweep[str(big)], "peak": hbm_peak, "unit": "GB/s", "frac": sweep[str(big)] / hbm_peak,
                                           "kernel": "checksum_ranges_kernel", "note": "algorithmic bytes = U (both checksums in one pass)"},
                              "cpu_baseline": {"crc32_GBps": cs / secs_c / 1e9, "adler32_GBps": cs / secs_a / 1e9, "cores": threads, "kind": kind_c,
                                               "sample": f"first {cs >> 20} MiB in 64 MiB pieces over {threads} threads"}}
    small = [(dbig.data_ptr() + (i * 4099) % (big - 70000), 1 << (12 + i % 5), 0, 1) for i in range(100000)]
    sarr = (pg.CksumItem * len(small))(*[pg.CksumItem(a, l, cs_, ads) for a, l, cs_, ads in small])
    sres = (pg.CksumResult * len(small))()
    dt = None
    for _ in range(3):                               # the C-ABI call alone (arrays marshalled once), best of 3
        t0 = time.perf_counter()
        eng._check(lib.nxgpu_checksum_batch(eng.ctx, sarr, len(small), sres, pg.MEM_DEVICE), "checksum storm")
        d = time.perf_counter() - t0
        dt = d if dt is None else min(dt, d)
    probe = small[12345]
    assert sres[12345].crc32 == zlib.crc32(C.string_at(hbig.value + (probe[0] - dbig.data_ptr()), probe[1]))
    extra["checksum_storm"] = {"buffers": len(small), "bytes": sum(x[1] for x in small), "buffers_per_s": round(len(small) / dt),
                               "GBps": round(sum(x[1] for x in small) / dt / 1e9, 2), "note": "4-64 KiB buffers, one batched call, host wall clock"}
    del dbig
    lib.nxgpu_host_free(hbig)

    # ---- ONE foreign stream (no index, not our encoder): 64 MiB of the benchmark text as system zlib level 6 writes it.  One warp
    # walks such a stream at ~70 MB/s; the engine finds the block starts, decodes the blocks speculatively and then for real
    # (csrc/inflate_par.cuh).  Device-resident, from host memory, and through the reference's own nx_uncompress() ----
    try:
        ln = min(n, 64 << 20)
        one = C.string_at(hsrc, ln)
        t0 = time.perf_counter()
        zc = zlib.compress(one, 6)
        t_comp = time.perf_counter() - t0
        t0 = time.perf_counter()
        assert zlib.decompress(zc) == one
        t_cpu = time.perf_counter() - t0
        dzc = torch.frombuffer(bytearray(zc), dtype=torch.uint8).cuda()
        dback = torch.empty(ln, dtype=torch.uint8, device="cuda")

        one_crc = zlib.crc32(one)

        def lone_step():
            eng.timer_start()
            r = eng.inflate_batch([pg.InflateItem(dzc.data_ptr(), len(zc), dback.data_ptr(), ln, pg.WRAP_ZLIB, 0)], mem=pg.MEM_DEVICE)[0]
            ms = eng.timer_stop()
            assert r.rc == 0 and r.out_len == ln and r.crc32 == one_crc
            return ms
        pl = timed(lone_step, 5, 2)
        hz, hb = C.c_void_p(), C.c_void_p()
        lib.nxgpu_host_alloc(len(zc) + 64, C.byref(hz)); lib.nxgpu_host_alloc(ln, C.byref(hb))
        C.memmove(hz, zc, len(zc))
        th = None
        for _ in range(3):
            t0 = time.perf_counter()
            r = eng.inflate_batch([pg.InflateItem(hz.value, len(zc), hb.value, ln, pg.WRAP_ZLIB, 0)], mem=pg.MEM_HOST)[0]
            d = time.perf_counter() - t0
            th = d if th is None else min(th, d)
        assert r.rc == 0 and C.string_at(hb.value, 4096) == one[:4096]
        lib.nxgpu_host_free(hz); lib.nxgpu_host_free(hb)
        lone = {"input": f"first {ln >> 20} MiB of the benchmark text, zlib.compress level 6 ({len(zc)} bytes, one stream, no flush points)",
                "device_resident_GBps": ln / (sum(pl) / len(pl)) / 1e6, "device_resident_ms_per_step": [round(x, 2) for x in pl],
                "host_to_host_GBps": ln / th / 1e9,
                "system_zlib_one_core_GBps": ln / t_cpu / 1e9}
        gpu_nxz0 = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")
        if os.path.exists(gpu_nxz0):
            through = {}
            for lg in (20, 22, 24, 26):
                pr = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nx_dropin_driver.py"), gpu_nxz0, "lone", str(lg), "6"],
                                    capture_output=True, text=True, timeout=300, env=dict(os.environ, NX_GZIP_LOGFILE="/tmp/nx_bench.log"))
                j = json.loads(pr.stdout.strip().splitlines()[-1])
                through[f"{1 << (lg - 20)} MiB"] = {"GBps": j["GBps"], "system_zlib_one_core_GBps": j["system_zlib_one_core_GBps"], "ok": j["ok"]}
            lone["through_nx_uncompress"] = {"note": "the reference's nx_uncompress() in NX mode over the engine, alice29-like text, zlib level 6", **through}
        extra["inflate_one_foreign_stream"] = lone
        del dzc, dback
    except Exception as e:                          # noqa: BLE001 - an extra, never fatal
        extra["inflate_one_foreign_stream"] = {"error": repr(e)[:300]}

    # ---- configs[4], second half: many concurrent small z_streams through the UNCHANGED zlib surface.  The reference's own
    # test/test_multithread_stress.c (compress()/uncompress() of 4 KiB - 1 MiB buffers, 64 threads, no Python in the loop) linked
    # against the drop-in library (NX mode: every call is a job on the GPU engine), and the same binary over the reference's
    # software path (sw_zlib.c -> zlib) as the CPU number beside it ----
    binp = os.path.join(ROOT, "oracle", "_ref", "reftests", "test_multithread_stress")
    gpu_nxz = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")
    cpu_nxz = os.path.join(ROOT, "oracle", "_ref", "libnxz_ref.so")

    def storm(libpath, selector, nthreads, seconds):
        with tempfile.TemporaryDirectory() as wd:
            os.symlink(libpath, os.path.join(wd, "libnxz.so.1"))
            env = dict(os.environ, LD_LIBRARY_PATH=wd + os.pathsep + os.path.dirname(libpath) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""),
                       NX_GZIP_TYPE_SELECTOR=str(selector), NX_GZIP_LOGFILE=os.path.join(wd, "nx.log"))
            t0 = time.time()
            p = subprocess.run([binp, str(nthreads), str(seconds), "1"], cwd=wd, env=env, capture_output=True, text=True, timeout=300)
            wall = time.time() - t0
            gb = None
            for ln in p.stdout.splitlines():
                if ln.startswith("Total data:"):
                    gb = float(ln.split()[2])
            return {"rc": p.returncode, "threads": nthreads, "seconds": seconds, "total_GB": gb, "GBps": (gb / seconds) if gb else None, "wall_s": round(wall, 1)}
    if os.path.exists(binp) and os.path.exists(gpu_nxz):
        try:
            st = {"program": "reference test/test_multithread_stress.c, unmodified", "gpu_engine": storm(gpu_nxz, 2, 64, 5)}
            if os.path.exists(cpu_nxz):
                st["cpu_software_path"] = storm(cpu_nxz, 1, threads, 5)
            extra["zstream_storm"] = st
        except Exception as e:                      # noqa: BLE001 - an extra, never fatal
            extra["zstream_storm"] = {"error": repr(e)[:200]}
        try:
            # samples/bench_initend.c: deflateInit2/deflateEnd and inflateInit2/inflateEnd pairs over the GPU engine
            env = dict(os.environ, NX_GZIP_LOGFILE="/tmp/nx_bench.log")
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nx_dropin_driver.py"), gpu_nxz, "initend", "1000"],
                               capture_output=True, text=True, timeout=300, env=env)
            ie = json.loads(p.stdout.strip().splitlines()[-1])
            extra["zstream_init_end"] = {k: round(v, 2) for k, v in ie.items() if k.endswith(("_us", "_ms"))}
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nx_dropin_driver.py"), gpu_nxz, "initend", "10"],
                               capture_output=True, text=True, timeout=300, env=dict(env, NXGPU_PREWARM="1"))
            extra["zstream_init_end"]["first_use_ms_with_NXGPU_PREWARM"] = round(json.loads(p.stdout.strip().splitlines()[-1])["first_use_ms"], 2)
        except Exception as e:                      # noqa: BLE001
            extra["zstream_init_end"] = {"error": repr(e)[:200]}


if __name__ == "__main__":
    main()
