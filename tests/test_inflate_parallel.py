"""One deflate stream decoded by many warps (csrc/inflate_par.cuh): block-start candidates, speculative decode with window
markers, chain, real decode per piece.  The gate is the serial engine and zlib: every observable of a call — bytes,
lengths, bytes consumed, checksums, and for NX job descriptors CC / SFBT / SUBC / SPBC / rembytecnt / out_dht — must be
what one warp walking the whole stream reports (oracle/nxemu.c for descriptors, system zlib for members)."""
import ctypes as C
import os
import random
import zlib

import pytest

from nxjob import Job

PAR = "NXGPU_INFLATE_PAR_MIN"


def _inputs(pg, alice):
    rnd = random.Random(11)
    text = pg.makedata(1, 22, alice)                       # 4 MiB of the benchmark text
    src = b"".join(open(os.path.join(os.path.dirname(__file__), f), "rb").read() for f in sorted(os.listdir(os.path.dirname(__file__))) if f.endswith(".py")) * 6
    mix = b"".join((alice[:70000], rnd.randbytes(90000), bytes(300000), alice[20000:140000], rnd.randbytes(200), b"ab" * 40000) * 3)
    return {"text": text, "source": src, "mix": mix}


def _streams(pg, alice):
    """(name, stream, data, wrap)"""
    ins = _inputs(pg, alice)
    out = []
    for name, d in ins.items():
        for lvl in (1, 6, 9):
            out.append((f"{name}-zlib{lvl}", zlib.compress(d, lvl), d))
        out.append((f"{name}-gzip6", zlib.compress(d, 6, wbits=31), d))
        out.append((f"{name}-raw6", zlib.compress(d, 6, wbits=-15), d))
    d = ins["text"][:1 << 20]
    fx = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    out.append(("fixed", fx.compress(d) + fx.flush(), d))
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    out.append(("sync-flushes", b"".join(co.compress(d[k:k + 3000]) + co.flush(zlib.Z_SYNC_FLUSH) for k in range(0, len(d), 3000)) + co.flush(), d))
    co = zlib.compressobj(9, zlib.DEFLATED, 15)
    out.append(("full-flushes", b"".join(co.compress(d[k:k + 50000]) + co.flush(zlib.Z_FULL_FLUSH) for k in range(0, len(d), 50000)) + co.flush(), d))
    out.append(("stored", zlib.compress(random.Random(3).randbytes(700000), 0), random.Random(3).randbytes(700000)))
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY)
    out.append(("huffman-only", co.compress(d) + co.flush(), d))
    z = bytes(1 << 24)
    out.append(("zeros", zlib.compress(z, 6), z))
    return out


def _inflate(engine, pg, stream, cap, hist=b""):
    sb = C.create_string_buffer(stream, len(stream))
    ob = (C.c_char * (len(hist) + cap + 16))()
    C.memmove(ob, hist, len(hist))
    r = engine.inflate_batch([pg.InflateItem(C.addressof(sb), len(stream), C.addressof(ob) + len(hist), cap, pg.WRAP_AUTO, len(hist))], mem=pg.MEM_HOST)[0]
    return (r.rc, r.out_len, r.in_used, r.flags, r.crc32, r.adler32), bytes(memoryview(ob)[len(hist): len(hist) + min(r.out_len, cap)])


@pytest.mark.gpu
def test_members_decoded_by_many_warps_match_zlib_and_the_serial_engine(engine, pg, alice, monkeypatch):
    """Foreign members of every block mix (dynamic, fixed, stored, empty stored blocks from flushes, one huge block), three
    wrappers and levels 1/6/9, with trailing bytes behind some: parallel path == serial path == zlib."""
    for name, stream, data in _streams(pg, alice):
        for trailing in (b"", b"\x00" * 9 + b"trailing"):
            s = stream + trailing
            monkeypatch.setenv(PAR, "65536")
            par, out = _inflate(engine, pg, s, len(data))
            monkeypatch.setenv(PAR, "0")
            ser, out_s = _inflate(engine, pg, s, len(data))
            assert par == ser, (name, par, ser)
            assert par[0] == 0 and out == data and par[2] == len(stream), (name, par, len(stream))
            assert par[4] == zlib.crc32(data) and par[5] == zlib.adler32(data)


@pytest.mark.gpu
def test_members_errors_and_limits_are_the_serial_engines(engine, pg, alice, monkeypatch):
    """Bit flips anywhere in a long stream, truncations, a target that is too small: the verdict and the bytes in front of the
    error are those of the serial path."""
    rnd = random.Random(9)
    data = pg.makedata(5, 21, alice)
    stream = zlib.compress(data, 6)
    cases = []
    for _ in range(24):
        b = bytearray(stream)
        pos = rnd.randrange(len(b))
        b[pos] ^= 1 << rnd.randrange(8)
        cases.append((bytes(b), len(data)))
    for cut in (len(stream) // 3, len(stream) - 5, len(stream) - 1):
        cases.append((stream[:cut], len(data)))
    for cap in (len(data) - 1, len(data) // 2, 100000):
        cases.append((stream, cap))
    n_bad = 0
    for s, cap in cases:
        monkeypatch.setenv(PAR, "65536")
        par, out = _inflate(engine, pg, s, cap)
        monkeypatch.setenv(PAR, "0")
        ser, out_s = _inflate(engine, pg, s, cap)
        assert par[0] == ser[0], (par, ser)
        if par[0] == 0:
            assert par == ser and out == out_s
        else:
            n_bad += 1
    assert n_bad >= 6


class _Dev(C.Structure):
    _fields_ = [("i", C.c_int * 8), ("paste_addr", C.c_void_p), ("fd", C.c_int), ("function", C.c_int), ("pad", C.c_char * 256)]


def _drive(run, stream, data, piece, dst_cap, limit=400):
    """feeds one raw deflate stream through decompress descriptors the way lib/nx_inflate.c:1447-1609 does; returns what
    every descriptor reported"""
    seen = []
    byte_pos, out_done = 0, 0
    in_subc = in_sfbt = in_rem = in_dhtlen = 0
    in_dht = b""
    extra = 0
    cur = piece
    for step in range(limit):
        end = min(len(stream), byte_pos + cur + extra)
        hist = data[max(0, out_done - 32768): out_done]
        hist = bytes((-len(hist)) % 16) + hist
        in_dyn = (in_sfbt & 0xe) == 0xc
        j = Job(0x10 if step == 0 else 0x14, [hist, stream[byte_pos:end]] if hist else [stream[byte_pos:end]], dst_cap, histlen_qw=len(hist) // 16,
                subc=in_subc, sfbt=in_sfbt, rem_or_dhtlen=in_dhtlen if in_dyn else in_rem, dht=in_dht,
                crc=zlib.crc32(data[:out_done]), adler=zlib.adler32(data[:out_done]))
        run(j)
        sfbt, low = (j.w396() >> 16) & 0xf, j.w396() & 0xffff
        rec = {"cc": j.cc(), "ce": j.ce(), "tpbc": j.tpbc() if j.cc() != 13 else 0}
        if j.cc() == 13:
            seen.append(rec)
            cur = max(cur // 2, 4096)
            continue
        rec.update(sfbt=sfbt, subc=j.w392(), spbc=j.spbc_decomp(), crc=j.crc(), adler=j.adler(), low=low if (sfbt & 0xe) in (0x8, 0xc) else 0)
        out = j.out()
        rec["out_ok"] = out == data[out_done: out_done + len(out)]
        if (sfbt & 0xe) == 0xc:
            nb = ((low & 0xfff) + 7) // 8
            dht = bytearray(j.get(256 + 400, nb))
            if (low & 7) and nb:
                dht[-1] &= (1 << (low & 7)) - 1
            rec["dht"] = bytes(dht)
        seen.append(rec)
        if j.cc() not in (0, 3):
            break
        out_done += len(out)
        if sfbt == 0:
            break
        consumed = (rec["spbc"] - len(hist)) - (rec["subc"] + 7) // 8
        progress = consumed > 0 or len(out) > 0
        byte_pos += consumed
        in_subc, in_sfbt = rec["subc"] % 8, sfbt
        in_rem = low if (sfbt & 0xe) == 0x8 else 0
        in_dhtlen = low & 0xfff if (sfbt & 0xe) == 0xc else 0
        in_dht = j.get(256 + 400, 288) if (sfbt & 0xe) == 0xc else b""
        extra = extra + cur if not progress else 0
        if byte_pos >= len(stream) and not progress:
            break
    return seen, out_done


@pytest.mark.gpu
def test_descriptors_decoded_by_many_warps_report_what_the_cpu_engine_reports(oracle, pg, alice, monkeypatch):
    """Long raw streams through chains of decompress / resume descriptors (sources that end inside blocks, targets that
    fill up, resumed dynamic / fixed / stored blocks with history): every field of every descriptor against oracle/nxemu.c,
    once with the parallel path forced on and once with it off."""
    oracle.oracle_nxemu_run_job.argtypes = [C.c_void_p]
    oracle.oracle_nxemu_run_job.restype = C.c_int
    lib = pg.load_library()
    dev = _Dev()
    assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0
    ins = _inputs(pg, alice)
    streams = [("text6", zlib.compress(ins["text"], 6, wbits=-15), ins["text"]),
               ("mix1", zlib.compress(ins["mix"], 1, wbits=-15), ins["mix"]),
               ("source9", zlib.compress(ins["source"], 9, wbits=-15) + b"\x01\x02\x03\x04trailer!", ins["source"])]
    fx = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    d = ins["text"][: 1 << 20]
    streams.append(("fixed", fx.compress(d) + fx.flush(), d))
    plans = [(1 << 30, 1 << 26), (300000, 1 << 22), (1 << 20, 1 << 20), (150001, 700000)]

    def cpu(job):
        assert oracle.oracle_nxemu_run_job(job.addr) == 0

    def gpu(job):
        assert lib.nxu_run_job(job.addr, C.byref(dev)) == 0
    try:
        for name, stream, data in streams:
            for piece, cap in plans:
                want, done_w = _drive(cpu, stream, data, piece, cap)
                assert done_w == len(data) and all(r.get("out_ok", True) for r in want), (name, piece, cap)
                for mode in ("65536", "0"):
                    monkeypatch.setenv(PAR, mode)
                    got, done_g = _drive(gpu, stream, data, piece, cap)
                    assert done_g == done_w and len(got) == len(want), (name, piece, cap, mode, len(got), len(want))
                    for k, (g, w) in enumerate(zip(got, want)):
                        assert g == w, (name, piece, cap, mode, k, g, w)
    finally:
        lib.nx_function_end(C.byref(dev))


@pytest.mark.gpu
def test_descriptor_errors_by_many_warps(oracle, pg, alice, monkeypatch):
    """a corrupt long stream: the completion code and the byte counts in front of the error are the CPU engine's"""
    oracle.oracle_nxemu_run_job.argtypes = [C.c_void_p]
    oracle.oracle_nxemu_run_job.restype = C.c_int
    lib = pg.load_library()
    dev = _Dev()
    assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0
    rnd = random.Random(21)
    data = pg.makedata(4, 21, alice)
    stream = zlib.compress(data, 6, wbits=-15)
    monkeypatch.setenv(PAR, "65536")
    try:
        n_err = 0
        for _ in range(16):
            b = bytearray(stream)
            pos = rnd.randrange(len(b))
            b[pos] ^= 1 << rnd.randrange(8)
            jc = Job(0x10, [bytes(b)], len(data) + 64)
            jg = Job(0x10, [bytes(b)], len(data) + 64)
            assert oracle.oracle_nxemu_run_job(jc.addr) == 0
            assert lib.nxu_run_job(jg.addr, C.byref(dev)) == 0
            assert jg.cc() == jc.cc(), (pos, jg.cc(), jc.cc())
            if jc.cc() in (0, 3):
                assert (jg.tpbc(), jg.w392(), jg.w396() >> 16, jg.spbc_decomp(), jg.crc()) == (jc.tpbc(), jc.w392(), jc.w396() >> 16, jc.spbc_decomp(), jc.crc())
                assert jg.out() == jc.out()
            else:
                n_err += 1
        assert n_err >= 1
    finally:
        lib.nx_function_end(C.byref(dev))


@pytest.mark.gpu
def test_gunzip_of_files_with_a_few_big_members(engine, pg, alice, monkeypatch):
    """nxgpu_gunzip_concat on what .gz files usually are — one big member, or a few: the dry run that finds where every
    member ends goes through the many-warp path too (counting only), then the real decode; result == gzip.decompress, with
    header look-alikes in the data, zero padding behind the last member, and the one-warp dry run as the reference."""
    rnd = random.Random(4)
    fake = b"\x1f\x8b\x08\x00" + bytes(14)
    a = pg.makedata(1, 22, alice)
    b = (alice[:200000] + fake * 3 + rnd.randbytes(70000)) * 5
    files = {"one": [a], "two": [a, b], "three-with-stored": [b, rnd.randbytes(300000), a[: 1 << 20]]}
    for name, parts in files.items():
        blob = b"".join(gzip_member(d, 6 if i != 1 else (0 if name == "three-with-stored" else 9)) for i, d in enumerate(parts))
        want = b"".join(parts)
        for pad in (b"", bytes(300)):
            monkeypatch.setenv(PAR, "65536")
            got, members = engine.gunzip(blob + pad, len(want))
            assert members == len(parts) and got == want, (name, members, len(got), len(want))
            monkeypatch.setenv(PAR, "0")
            assert engine.gunzip(blob + pad, len(want)) == (want, len(parts)), name


def gzip_member(d, level):
    import gzip as _gzip
    return _gzip.compress(d, level, mtime=0)


@pytest.mark.gpu
def test_full_size_member_round_trip_through_the_many_warp_decode(engine, pg, alice, monkeypatch):
    """BASELINE.json configs[1] at full size: 1 GiB of the benchmark text -> one gzip member with
    primed 256 KiB chunks (no index kept) -> inflated again as ONE foreign member: 4096+ block-start candidates, speculative
    pieces, chain.  Device-resident end to end; the checks are the size-independent ones: lengths, the trailer, and the
    crc32 / adler32 of the 1 GiB output against the values the deflate side computed from the input."""
    monkeypatch.delenv(PAR, raising=False)
    lib = pg.load_library()
    n = 1 << 30
    dsrc = engine.alloc(n)
    seed = C.create_string_buffer(alice, len(alice))
    host = C.c_void_p()
    lib.nxgpu_host_alloc(n + 64, C.byref(host))
    try:
        assert lib.nxgpu_makedata(1, 30, C.addressof(seed), len(alice), host, n + 16) == n
        engine._check(lib.nxgpu_memcpy_h2d(engine.ctx, dsrc.ptr, host, n), "h2d")
    finally:
        lib.nxgpu_host_free(host)
    cap = engine.deflate_bound(n)
    ddst = engine.alloc(cap)
    res = engine.deflate_stream_device(dsrc.ptr, n, ddst.ptr, cap, level=6, wrap=pg.WRAP_GZIP)
    assert res.crc32 == 0xf50ac623 and res.n_chunks == 4096           # crc32 of makedata -s 1 -b 30 (bench.py verifies the same value with zlib)
    dsrc.free()
    dback = engine.alloc(n)
    r = engine.inflate_batch([pg.InflateItem(ddst.ptr, res.out_len, dback.ptr, n, pg.WRAP_GZIP, 0)], mem=pg.MEM_DEVICE)[0]
    assert r.rc == 0 and r.out_len == n and r.in_used == res.out_len and (r.flags & 3) == 3
    assert r.crc32 == res.crc32 and r.adler32 == res.adler32
    ddst.free(); dback.free()


@pytest.mark.gpu
def test_many_shapes_of_streams_through_the_many_warp_decode(engine, pg, alice, monkeypatch):
    """Streams made with every knob of deflateInit2 that changes the block structure — memLevel 1 (a block every 128 symbols:
    thousands of pieces far shorter than the window, markers chased through many predecessors), memLevel 9, windowBits 9-15,
    the RLE / FILTERED / HUFFMAN_ONLY / FIXED strategies, levels 0-9, Z_PARTIAL/SYNC/FULL flushes at random places — over
    text, binary and mixed inputs of 100 KB - 3 MB: bytes, lengths and checksums against the input, verdict against zlib."""
    monkeypatch.setenv(PAR, "65536")
    rnd = random.Random(2024)
    text = pg.makedata(5, 22, alice)
    bins = b"".join(rnd.randrange(1 << 30).to_bytes(4, "little") * rnd.randrange(1, 6) for _ in range(150000))
    pool = [text, alice * 4, bins, bytes(1 << 20) + alice[:100000] + rnd.randbytes(300000) + text[:500000]]
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    n_par = 0
    for case in range(60):
        src = pool[case % len(pool)]
        o = rnd.randrange(0, len(src) // 2)
        d = src[o: o + rnd.randrange(100000, min(len(src) - o, 3000000))]
        level = rnd.choice([0, 1, 2, 4, 6, 6, 9])
        wbits = rnd.choice([9, 11, 13, 15, 15, 15])
        mem = rnd.choice([1, 1, 4, 8, 9])
        strat = rnd.choice(strategies)
        wrap = rnd.choice([wbits, -wbits, wbits + 16])
        co = zlib.compressobj(level, zlib.DEFLATED, wrap, mem, strat)
        parts, p = [], 0
        while p < len(d):
            step = rnd.randrange(20000, 400000)
            parts.append(co.compress(d[p:p + step])); p += step
            if rnd.random() < 0.3:
                parts.append(co.flush(rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
        parts.append(co.flush())
        z = b"".join(parts)
        got, out = _inflate(engine, pg, z, len(d))
        assert got[0] == 0 and out == d and got[2] == len(z), (case, level, wbits, mem, strat, wrap, got, len(z), len(d))
        assert got[4] == zlib.crc32(d) and got[5] == zlib.adler32(d)
        n_par += len(z) >= 65536
    assert n_par >= 30


@pytest.mark.gpu
def test_a_batch_of_long_streams_decodes_side_by_side(engine, pg, alice, monkeypatch):
    """twelve long members in ONE batch: more many-warp decodes than there are slots (each has its own stream and scratch),
    running beside each other and beside the launch that takes the short members"""
    monkeypatch.setenv(PAR, "65536")
    rnd = random.Random(12)
    text = pg.makedata(4, 22, alice)
    members = []
    for i in range(12):
        o = rnd.randrange(0, len(text) // 2)
        d = text[o: o + rnd.randrange(300000, 1500000)]
        members.append((zlib.compress(d, rnd.choice([1, 6, 9]), wbits=rnd.choice([15, 31, -15])), d))
    for i in range(6):                                                     # short ones: the ordinary launch
        d = alice[i * 1000: i * 1000 + 20000]
        members.append((zlib.compress(d, 6), d))
    rnd.shuffle(members)
    blob = b"".join(m for m, _ in members)
    src = C.create_string_buffer(blob, len(blob))
    outb = (C.c_char * (sum(len(d) for _, d in members) + 64))()
    items, so, do = [], 0, 0
    for m, d in members:
        items.append(pg.InflateItem(C.addressof(src) + so, len(m), C.addressof(outb) + do, len(d), pg.WRAP_AUTO, 0))
        so += len(m); do += len(d)
    for rounds in range(2):
        res = engine.inflate_batch(items, mem=pg.MEM_HOST)
        do = 0
        for (m, d), r in zip(members, res):
            assert r.rc == 0 and r.out_len == len(d) and r.in_used == len(m) and r.crc32 == zlib.crc32(d), (len(m), len(d), r.rc, r.out_len)
            assert bytes(memoryview(outb)[do: do + len(d)]) == d
            do += len(d)
