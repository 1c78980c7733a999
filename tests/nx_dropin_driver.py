#!/usr/bin/env python3
"""Drives the reference's UNMODIFIED zlib-compatible surface (libnxz.h:119-192: compress2 /
uncompress / deflate / inflate / crc32 / adler32) in NX mode (NX_GZIP_TYPE_SELECTOR=2) through a
build of its host code whose six boundary symbols come from either engine:

    oracle/_ref/libnxz_ref.so   nxu_run_job = oracle/nxemu.c (CPU, the checker)
    power-gzip_b200/libnxz_gpu.so   nxu_run_job = power-gzip_b200/libnxgpu.so (the product)

Run as a subprocess (the selector is read when the library is loaded).  Prints one JSON line.
Mirrors the round trips of the reference's test/test_deflate.c and test/test_inflate.c: output of
the nx deflate must inflate with system zlib, and streams made by system zlib must inflate through
nx, in one shot and in small pieces.
usage: nx_dropin_driver.py <libnxz .so> [size_log2]
       nx_dropin_driver.py <libnxz .so> stress <threads> <iterations>
       nx_dropin_driver.py <libnxz .so> initend [pairs]
       nx_dropin_driver.py <libnxz .so> lone [size_log2] [level]
The stress mode follows the reference's test/test_multithread_stress.c:26-120: every thread runs
compress()/uncompress() over the same ten buffers (4 KiB .. 1 MiB of the 33-symbol alphabet of
test/test_utils.c:22-28, srand(1)) and checks the round trip; it also reports how many descriptors
the engine coalesced per GPU batch (nxgpu_job_stats) when the engine exports that."""
import ctypes as C
import faulthandler
import gzip
import json
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NX_GZIP_TYPE_SELECTOR"] = "2"          # lib/nx_zlib.c: 2 = NX only, no software fallback
os.environ.setdefault("NX_GZIP_LOGFILE", "/tmp/nx_dropin.log")
if os.environ.get("NX_DRIVER_WATCHDOG"):
    faulthandler.dump_traceback_later(int(os.environ["NX_DRIVER_WATCHDOG"]), exit=True)
lib = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)
STRESS = len(sys.argv) > 2 and sys.argv[2] == "stress"
log2 = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 20


class ZStream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong),
                ("next_out", C.c_void_p), ("avail_out", C.c_uint), ("total_out", C.c_ulong),
                ("msg", C.c_char_p), ("state", C.c_void_p), ("zalloc", C.c_void_p), ("zfree", C.c_void_p),
                ("opaque", C.c_void_p), ("data_type", C.c_int), ("adler", C.c_ulong), ("reserved", C.c_ulong)]


lib.compress2.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong, C.c_int]
lib.uncompress.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong]
lib.nx_uncompress.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong]
lib.compressBound.restype = C.c_ulong
lib.compressBound.argtypes = [C.c_ulong]
lib.crc32.restype = C.c_ulong
lib.crc32.argtypes = [C.c_ulong, C.c_void_p, C.c_uint]
lib.adler32.restype = C.c_ulong
lib.adler32.argtypes = [C.c_ulong, C.c_void_p, C.c_uint]
for f in ("deflateInit2_", "inflateInit2_"):
    getattr(lib, f).restype = C.c_int
lib.deflateInit2_.argtypes = [C.POINTER(ZStream), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
lib.inflateInit2_.argtypes = [C.POINTER(ZStream), C.c_int, C.c_char_p, C.c_int]
for f in ("deflate", "inflate"):
    getattr(lib, f).argtypes = [C.POINTER(ZStream), C.c_int]
for f in ("deflateEnd", "inflateEnd"):
    getattr(lib, f).argtypes = [C.POINTER(ZStream)]
VER = b"1.2.11"
Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH, Z_OK, Z_STREAM_END, Z_BUF_ERROR = 0, 2, 3, 4, 0, 1, -5


def nx_compress2(data, level):
    cap = lib.compressBound(len(data)) + 64
    out = C.create_string_buffer(cap)
    n = C.c_ulong(cap)
    rc = lib.compress2(out, C.byref(n), data, len(data), level)
    assert rc == 0, f"compress2 rc={rc}"
    return out.raw[: n.value]


def nx_uncompress(blob, n_out):
    out = C.create_string_buffer(max(n_out, 1))
    n = C.c_ulong(n_out)
    # nx_uncompress is what uncompress() calls in NX mode (lib/nx_uncompr.c:123-139); uncompress() itself goes through the PLT to
    # uncompress2, which in this interpreter binds to the libz that Python's zlib module has already loaded
    rc = lib.nx_uncompress(out, C.byref(n), blob, len(blob))
    assert rc == 0, f"uncompress rc={rc}"
    return out.raw[: n.value]


def nx_deflate_stream(data, wbits, in_piece, out_piece, flush_every=0):
    """deflate() fed in_piece bytes at a time, drained out_piece bytes at a time"""
    s = ZStream()
    assert lib.deflateInit2_(C.byref(s), 6, 8, wbits, 8, 0, VER, C.sizeof(ZStream)) == 0
    src = C.create_string_buffer(data, len(data))
    obuf = C.create_string_buffer(out_piece)
    out = bytearray()
    pos, k = 0, 0
    while True:
        take = min(in_piece, len(data) - pos)
        s.next_in = C.addressof(src) + pos
        s.avail_in = take
        pos += take
        last = pos == len(data)
        k += 1
        flush = Z_FINISH if last else (Z_FULL_FLUSH if flush_every and k % flush_every == 0 else Z_NO_FLUSH)
        while True:
            s.next_out = C.addressof(obuf)
            s.avail_out = out_piece
            rc = lib.deflate(C.byref(s), flush)
            assert rc in (Z_OK, Z_STREAM_END, Z_BUF_ERROR), f"deflate rc={rc}"
            out += obuf.raw[: out_piece - s.avail_out]
            if rc == Z_STREAM_END:
                break
            if s.avail_in == 0 and s.avail_out != 0 and flush != Z_FINISH:
                break
        if last:
            assert rc == Z_STREAM_END
            break
    lib.deflateEnd(C.byref(s))
    return bytes(out)


def nx_inflate_stream(blob, wbits, in_piece, out_piece, expect_len):
    s = ZStream()
    assert lib.inflateInit2_(C.byref(s), wbits, VER, C.sizeof(ZStream)) == 0
    src = C.create_string_buffer(blob, len(blob))
    obuf = C.create_string_buffer(out_piece)
    out = bytearray()
    pos = 0
    rc = Z_OK
    guard = 0
    while rc != Z_STREAM_END:
        if s.avail_in == 0 and pos < len(blob):
            take = min(in_piece, len(blob) - pos)
            s.next_in = C.addressof(src) + pos
            s.avail_in = take
            pos += take
        s.next_out = C.addressof(obuf)
        s.avail_out = out_piece
        rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
        assert rc in (Z_OK, Z_STREAM_END, Z_BUF_ERROR), f"inflate rc={rc} after {len(out)} bytes"
        out += obuf.raw[: out_piece - s.avail_out]
        guard += 1
        assert guard < 10_000_000 and len(out) <= expect_len, "inflate does not terminate"
        if rc == Z_BUF_ERROR and pos >= len(blob) and s.avail_out != 0:
            raise AssertionError("inflate starved")
    lib.inflateEnd(C.byref(s))
    return bytes(out)


def stress(n_threads, iterations):
    import threading
    import time
    dict33 = b"abcdefghijklmnopqrstuvwxyz,.!?.{}"
    rnd = random.Random(1)
    bufs = [bytes(rnd.choice(dict33) for _ in range(n)) for n in (4096, 4096, 65536, 65536, 131072, 131072, 262144, 262144, 1048576, 1048576)]
    refs = [zlib.compress(b, 6) for b in bufs]
    nx_compress2(bufs[0], 6)                     # opens the device outside the timed region
    errors, done_bytes = [], [0] * n_threads
    start = threading.Barrier(n_threads + 1)

    def worker(t):
        try:
            start.wait()
            for it in range(iterations):
                for k, b in enumerate(bufs):
                    z = nx_compress2(b, 6)
                    if zlib.decompress(z) != b:
                        raise AssertionError(f"thread {t}: zlib cannot decode nx compress of buffer {k}")
                    if nx_uncompress(z, len(b)) != b or nx_uncompress(refs[k], len(b)) != b:
                        raise AssertionError(f"thread {t}: nx uncompress of buffer {k} differs")
                    done_bytes[t] += 3 * len(b)
        except Exception as e:                       # noqa: BLE001 - reported to the parent
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    start.wait()
    t0 = time.time()
    for x in th:
        x.join()
    dt = time.time() - t0
    rep = {"lib": os.path.basename(sys.argv[1]), "threads": n_threads, "iterations": iterations, "errors": errors,
           "seconds": dt, "MBps": sum(done_bytes) / dt / 1e6, "calls_per_s": n_threads * iterations * len(bufs) * 3 / dt}
    if hasattr(lib, "nxgpu_job_stats"):
        b, j, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.nxgpu_job_stats(0, C.byref(b), C.byref(j), C.byref(m))
        rep.update(batches=b.value, jobs=j.value, max_batch=m.value)
    print(json.dumps(rep))


def initend(n):
    """samples/bench_initend.c: cost of deflateInit2/deflateEnd and inflateInit2/inflateEnd pairs.
    The fifos and DHT tables are the reference's host code; the engine's share is nx_function_begin on first use."""
    import time
    if os.environ.get("NXGPU_PREWARM"):
        time.sleep(3.0)                          # the application's own start-up, during which the library opens the device in the background
    t0 = time.perf_counter()
    nx_compress2(b"warm up the device handle", 6)
    first = time.perf_counter() - t0
    out = {"first_use_ms": first * 1e3}
    for name, init, end, args in (("deflate", lib.deflateInit2_, lib.deflateEnd, (6, 8, 31, 8, 0, VER, C.sizeof(ZStream))),
                                  ("inflate", lib.inflateInit2_, lib.inflateEnd, (31, VER, C.sizeof(ZStream)))):
        t0 = time.perf_counter()
        for _ in range(n):
            s = ZStream()
            assert init(C.byref(s), *args) == 0
            end(C.byref(s))
        out[f"{name}_init_end_us"] = (time.perf_counter() - t0) / n * 1e6
    print(json.dumps({"lib": os.path.basename(sys.argv[1]), "pairs": n, **out}))


if STRESS:
    stress(int(sys.argv[3]), int(sys.argv[4]))
    sys.exit(0)
def lone(lg, level):
    """ONE foreign zlib stream through uncompress() (the LD_PRELOAD single-stream case): text-like data compressed by system zlib,
    inflated by the library under test; system zlib's own inflate on one core is timed beside it."""
    import time
    alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
    rnd = random.Random(7)
    parts, n = [], 0
    while n < (1 << lg):
        o, l = rnd.randrange(len(alice) - 4096), rnd.randrange(200, 4096)
        parts.append(alice[o:o + l]); n += l
    data = b"".join(parts)[: 1 << lg]
    comp = zlib.compress(data, level)
    src = C.create_string_buffer(comp, len(comp))
    dst = C.create_string_buffer(len(data))
    lib.nx_uncompress.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong]
    best = 1e9
    for it in range(4):
        dl = C.c_ulong(len(data))
        t0 = time.perf_counter()
        rc = lib.nx_uncompress(dst, C.byref(dl), src, len(comp))     # (uncompress() itself calls uncompress2 through the PLT, which binds to the libz this interpreter already loaded)
        dt = time.perf_counter() - t0
        assert rc == 0 and dl.value == len(data), (rc, dl.value)
        if it:
            best = min(best, dt)
    ok = dst.raw == data
    t0 = time.perf_counter()
    ref = zlib.decompress(comp)
    cpu = time.perf_counter() - t0
    print(json.dumps({"bytes": len(data), "compressed_bytes": len(comp), "level": level, "ok": bool(ok and ref == data),
                      "uncompress_ms": round(best * 1e3, 2), "GBps": round(len(data) / best / 1e9, 3),
                      "system_zlib_one_core_ms": round(cpu * 1e3, 2), "system_zlib_one_core_GBps": round(len(data) / cpu / 1e9, 3)}))


if len(sys.argv) > 2 and sys.argv[2] == "lone":
    lone(int(sys.argv[3]) if len(sys.argv) > 3 else 26, int(sys.argv[4]) if len(sys.argv) > 4 else 6)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "initend":
    initend(int(sys.argv[3]) if len(sys.argv) > 3 else 1000)
    sys.exit(0)

def dictionary_roundtrip(data, dictionary):
    """Preset dictionaries through the zlib surface (test/test_dict.c:40-140): the dictionary travels as job history.
    zlib mode: FDICT set, DICTID = adler32(dict), trailer = adler32(data); raw mode: no header.  Both directions:
    nx deflate -> system zlib inflate, system zlib deflate -> nx inflate (Z_NEED_DICT, then inflateSetDictionary)."""
    lib.deflateSetDictionary.argtypes = [C.POINTER(ZStream), C.c_char_p, C.c_uint]
    lib.inflateSetDictionary.argtypes = [C.POINTER(ZStream), C.c_char_p, C.c_uint]
    Z_NEED_DICT = 2
    sizes = {}
    for wbits in (15, -15):
        s = ZStream()
        assert lib.deflateInit2_(C.byref(s), 6, 8, wbits, 8, 0, VER, C.sizeof(ZStream)) == 0
        assert lib.deflateSetDictionary(C.byref(s), dictionary, len(dictionary)) == 0
        if wbits > 0:
            assert s.adler == zlib.adler32(dictionary), "deflateSetDictionary must leave the dictionary's adler32 in strm.adler"
        src = C.create_string_buffer(data, len(data))
        out = C.create_string_buffer(2 * len(data) + 1024)
        s.next_in, s.avail_in = C.addressof(src), len(data)
        s.next_out, s.avail_out = C.addressof(out), len(out)
        assert lib.deflate(C.byref(s), Z_FINISH) == Z_STREAM_END
        blob = out.raw[: s.total_out]
        lib.deflateEnd(C.byref(s))
        if wbits > 0:
            assert blob[1] & 0x20, "FLG.FDICT not set"
            assert int.from_bytes(blob[2:6], "big") == zlib.adler32(dictionary), "wrong DICTID"
            assert int.from_bytes(blob[-4:], "big") == zlib.adler32(data), "trailer must be the adler32 of the data alone"
        d = zlib.decompressobj(wbits, zdict=dictionary)
        assert d.decompress(blob) == data, f"zlib cannot decode nx deflate with a dictionary (wbits {wbits})"
        sizes[wbits] = len(blob)
        # the other direction
        co = zlib.compressobj(6, zlib.DEFLATED, wbits, zdict=dictionary)
        foreign = co.compress(data) + co.flush()
        s = ZStream()
        assert lib.inflateInit2_(C.byref(s), wbits, VER, C.sizeof(ZStream)) == 0
        fsrc = C.create_string_buffer(foreign, len(foreign))
        back = C.create_string_buffer(len(data) + 64)
        s.next_in, s.avail_in = C.addressof(fsrc), len(foreign)
        s.next_out, s.avail_out = C.addressof(back), len(back)
        if wbits < 0:
            assert lib.inflateSetDictionary(C.byref(s), dictionary, len(dictionary)) == 0
        rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
        if wbits > 0:
            assert rc == Z_NEED_DICT, f"expected Z_NEED_DICT, got {rc}"
            assert s.adler == zlib.adler32(dictionary)
            assert lib.inflateSetDictionary(C.byref(s), dictionary, len(dictionary)) == 0
            rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
        guard = 0
        while rc == Z_OK and guard < 1000:
            rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
            guard += 1
        assert rc == Z_STREAM_END, f"inflate with dictionary ended with {rc}"
        assert back.raw[: s.total_out] == data
        lib.inflateEnd(C.byref(s))
    return sizes[15]


def trailing_data(data):
    """Bytes BEHIND the end of a stream inside the same job (concatenated gzip members, a tar of .gz, PNG chunks, HTTP):
    the engine reports the source bytes it read (SPBC) and the bits behind the final end-of-block it discarded (SUBC,
    16 bits wide; 32..39 / 64..71 for a bare zlib / gzip trailer, inc_nx/nxu.h:454-465) and the host finds the trailer at
    spbc - histlen - subc/8 (lib/nx_inflate.c:1452-1472).  Round 1 counted everything supplied in SUBC, which wrapped
    at 8 KiB.  Done = Z_STREAM_END with total_in == the stream's own length for 0 .. 100 000 appended bytes."""
    rnd = random.Random(11)
    checked = 0
    for wbits, blob in ((15, zlib.compress(data, 6)), (31, gzip.compress(data, 6)), (-15, zlib.compress(data, 6)[2:-4])):
        for extra in (0, 1, 7, 8, 9, 100, 8100, 8191, 8192, 8193, 8300, 9000, 20000, 65536, 70000, 100000):
            buf = blob + rnd.randbytes(extra)
            for hold_back in (0, 300):
                if hold_back >= len(blob):
                    continue
                s = ZStream()
                assert lib.inflateInit2_(C.byref(s), wbits, VER, C.sizeof(ZStream)) == 0
                src = C.create_string_buffer(buf, len(buf))
                out = C.create_string_buffer(len(data) + 64)
                s.next_out, s.avail_out = C.addressof(out), len(out)
                first = len(blob) - hold_back if hold_back else len(buf)
                s.next_in, s.avail_in = C.addressof(src), first
                rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
                guard = 0
                while rc == Z_OK and guard < 100:
                    if s.avail_in == 0 and first < len(buf):
                        s.next_in, s.avail_in = C.addressof(src) + first, len(buf) - first
                        first = len(buf)
                    rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
                    guard += 1
                assert rc == Z_STREAM_END, f"wbits {wbits}, {extra} bytes behind the stream, hold back {hold_back}: rc {rc}, total_in {s.total_in}"
                # a second piece below cache_threshold (8 KiB) is swallowed whole into fifo_in by the host code before
                # any job runs (lib/nx_inflate.c:1197-1205): total_in then counts it, on any engine
                if hold_back == 0 or hold_back + extra >= 8192:
                    assert s.total_in == len(blob), f"wbits {wbits}, {extra} trailing: total_in {s.total_in} != {len(blob)}"
                assert out.raw[: s.total_out] == data
                lib.inflateEnd(C.byref(s))
                checked += 1
    # concatenated gzip members in one buffer, inflateReset between them (what gunzip does with `cat a.gz b.gz`)
    lib.inflateReset.argtypes = [C.POINTER(ZStream)]
    parts = [data[:50000], data[50000:50010], b"", data[60000:]]
    cat = b"".join(gzip.compress(p, 6) for p in parts)
    src = C.create_string_buffer(cat, len(cat))
    out = C.create_string_buffer(len(data) + 64)
    s = ZStream()
    assert lib.inflateInit2_(C.byref(s), 31, VER, C.sizeof(ZStream)) == 0
    s.next_in, s.avail_in = C.addressof(src), len(cat)
    got = []
    for k in range(len(parts)):
        s.next_out, s.avail_out = C.addressof(out), len(out)
        before = s.avail_out
        rc, guard = Z_OK, 0
        while rc == Z_OK and guard < 100:
            rc = lib.inflate(C.byref(s), Z_NO_FLUSH)
            guard += 1
        assert rc == Z_STREAM_END, f"member {k}: rc {rc}"
        got.append(out.raw[: before - s.avail_out])
        assert lib.inflateReset(C.byref(s)) == 0
    assert got == parts and s.avail_in == 0, "concatenated members do not come back one by one"
    lib.inflateEnd(C.byref(s))
    return checked


def zero_input_and_reset(data):
    """test/test_zeroinput.c:39-60: an empty stream finished with every flush mode must be a valid zlib stream of zero
    bytes; test/test_reset.c:69-140: deflateReset / inflateReset give a stream that works again, several times."""
    lib.deflateReset.argtypes = [C.POINTER(ZStream)]
    lib.inflateReset.argtypes = [C.POINTER(ZStream)]
    out = C.create_string_buffer(2 * len(data) + 1024)
    for flush in (Z_NO_FLUSH, 1, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH):
        s = ZStream()
        assert lib.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, VER, C.sizeof(ZStream)) == 0
        s.next_in, s.avail_in = C.addressof(out), 0
        s.next_out, s.avail_out = C.addressof(out), len(out)
        rc = lib.deflate(C.byref(s), flush)
        assert rc == (Z_STREAM_END if flush == Z_FINISH else Z_OK), f"empty deflate, flush {flush}: {rc}"
        if flush != Z_FINISH:
            assert lib.deflate(C.byref(s), Z_FINISH) == Z_STREAM_END
        assert zlib.decompress(out.raw[: s.total_out]) == b"", f"empty stream (flush {flush}) is not a valid zlib stream"
        lib.deflateEnd(C.byref(s))
    # the sequence of test/test_reset.c:80-140: reset between init and the first call is fine, the stream works, and a
    # reset afterwards zeroes the counters (re-using a deflate stream after a reset that follows data trips an assertion
    # in the reference's own host code, lib/nx_deflate.c:843, on any engine — the reference's test never does that)
    back = C.create_string_buffer(len(data) + 64)
    src = C.create_string_buffer(data, len(data))
    s = ZStream()
    assert lib.deflateInit2_(C.byref(s), 6, 8, 31, 8, 0, VER, C.sizeof(ZStream)) == 0
    assert lib.deflateReset(C.byref(s)) == 0
    s.next_in, s.avail_in = C.addressof(src), len(data)
    s.next_out, s.avail_out = C.addressof(out), len(out)
    assert lib.deflate(C.byref(s), Z_FINISH) == Z_STREAM_END
    blob = out.raw[: s.total_out]
    assert gzip.decompress(blob) == data
    assert lib.deflateReset(C.byref(s)) == 0 and s.total_in == 0 and s.total_out == 0
    lib.deflateEnd(C.byref(s))
    d = ZStream()
    assert lib.inflateInit2_(C.byref(d), 31, VER, C.sizeof(ZStream)) == 0
    assert lib.inflateReset(C.byref(d)) == 0
    for k in range(3):                               # an inflate stream is re-used after every reset
        zs = C.create_string_buffer(blob, len(blob))
        d.next_in, d.avail_in = C.addressof(zs), len(blob)
        d.next_out, d.avail_out = C.addressof(back), len(back)
        rc, guard = Z_OK, 0
        while rc == Z_OK and guard < 1000:
            rc = lib.inflate(C.byref(d), Z_NO_FLUSH)
            guard += 1
        assert rc == Z_STREAM_END and back.raw[: d.total_out] == data, f"round {k}: inflate after inflateReset differs ({rc})"
        assert lib.inflateReset(C.byref(d)) == 0 and d.total_out == 0
    lib.inflateEnd(C.byref(d))
    return True


def gz_file_roundtrip(data):
    """The reference's gz* file layer (lib/nx_gzlib.c:130-351: gzopen / gzwrite / gzread / gzclose) over the engine:
    a file it writes must gunzip with Python, a file Python wrote must read back through it."""
    import tempfile
    # the nx_-prefixed entry points, as in the reference's test/test_gz.c:15-60
    gzopen, gzwrite, gzread, gzclose = lib.nx_gzopen, lib.nx_gzwrite, lib.nx_gzread, lib.nx_gzclose
    gzopen.restype = C.c_void_p
    gzopen.argtypes = [C.c_char_p, C.c_char_p]
    gzwrite.argtypes = [C.c_void_p, C.c_char_p, C.c_uint]
    gzread.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    gzclose.argtypes = [C.c_void_p]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "w.gz").encode()
        f = gzopen(path, b"wb6")
        assert f, "gzopen(wb6) failed"
        for o in range(0, len(data), 50000):
            piece = data[o:o + 50000]
            assert gzwrite(f, piece, len(piece)) == len(piece)
        gzclose(f)
        assert gzip.decompress(open(path, "rb").read()) == data, "gunzip of a file written by gzwrite differs"
        size = os.path.getsize(path)
        path2 = os.path.join(td, "r.gz").encode()
        with open(path2, "wb") as fh:
            fh.write(gzip.compress(data, 6))
        f = gzopen(path2, b"rb")
        assert f, "gzopen(rb) failed"
        got = bytearray()
        buf = C.create_string_buffer(65536)
        while len(got) < len(data):
            n = gzread(f, buf, 65536)
            assert n > 0, f"gzread returned {n} after {len(got)} bytes"
            got += buf.raw[:n]
        gzclose(f)
        assert bytes(got) == data, "gzread of a foreign gzip file differs"
    return size


alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
rnd = random.Random(7)
text = (alice * ((1 << log2) // len(alice) + 1))[: 1 << log2]
cases = {
    "alice": alice,
    "text": text,
    "zeros": bytes(300000),
    "random": rnd.randbytes(100000),
    "tiny": b"hello hello hello hello",
}
report = {"lib": os.path.basename(sys.argv[1]), "cases": {}}
for name, data in cases.items():
    r = {}
    print("case", name, len(data), file=sys.stderr, flush=True)
    # libnxz.h compress2 / uncompress (lib/nx_compress.c:82, lib/nx_uncompr.c:91)
    z = nx_compress2(data, 6)
    assert zlib.decompress(z) == data, f"{name}: zlib cannot decode nx compress2 output"
    r["compress2"] = len(z)
    assert nx_uncompress(zlib.compress(data, 6), len(data)) == data, f"{name}: nx uncompress(zlib -6) differs"
    assert nx_uncompress(z, len(data)) == data, f"{name}: nx uncompress(nx compress2) differs"
    # checksums through the exported crc32/adler32 (lib/nx_crc.c:437, lib/nx_adler32.c:182)
    assert lib.crc32(0, data, len(data)) == zlib.crc32(data)
    assert lib.adler32(1, data, len(data)) == zlib.adler32(data)
    # streaming deflate: gzip wrapper, small pieces, full flushes in between
    g = nx_deflate_stream(data, 31, 60000, 50000, flush_every=3)
    assert gzip.decompress(g) == data, f"{name}: gzip cannot decode streamed nx deflate"
    r["deflate_gzip_stream"] = len(g)
    raw = nx_deflate_stream(data, -15, 1 << 20, 1 << 20)
    assert zlib.decompress(raw, -15) == data
    # streaming inflate of foreign streams, cut at awkward places (resume protocol, lib/nx_inflate.c:1447-1609)
    for lvl, wb in ((6, 31), (1, 15), (9, -15), (0, 31)):
        co = zlib.compressobj(lvl, zlib.DEFLATED, wb)
        blob = co.compress(data) + co.flush()
        for in_piece, out_piece in ((1 << 20, 1 << 20), (4099, 70001), (len(blob) // 3 + 1, 1 << 16)):
            got = nx_inflate_stream(blob, wb, in_piece, out_piece, len(data))
            assert got == data, f"{name}: nx inflate(level {lvl}, wbits {wb}, pieces {in_piece}/{out_piece}) differs"
    fx = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    blob = fx.compress(data) + fx.flush()
    assert nx_inflate_stream(blob, -15, 5000, 9000, len(data)) == data
    report["cases"][name] = r
print("zero input, reset", file=sys.stderr, flush=True)
report["zero_input_and_reset"] = zero_input_and_reset(alice[:70000])
print("trailing data", file=sys.stderr, flush=True)
report["trailing_data"] = trailing_data(alice)
print("dictionary", file=sys.stderr, flush=True)
# sizes: one job, and a first job that yields more than the 32 KiB window — the reference's host code hands the
# dictionary to the first decompress job only (lib/nx_inflate.c:1711-1712), whatever engine sits below it
report["dictionary"] = [dictionary_roundtrip(alice[40000:45000], alice[:32768]), dictionary_roundtrip(alice[40000:150000], alice[:20000])]
print("gz file layer", file=sys.stderr, flush=True)
report["gz_file"] = gz_file_roundtrip(alice[:60000])
print(json.dumps(report))
