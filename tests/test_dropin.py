"""The drop-in boundary: nxu_run_job and friends.

* The reference's UNMODIFIED host code (oracle/_ref/libnxz_*.so, compiled in place from
  /root/reference) is driven through its zlib-compatible surface in NX mode, once over the CPU
  engine of oracle/nxemu.c (pins what the host code expects from a job) and once over the GPU
  engine (libnxgpu.so).  Mirrors test/test_deflate.c, test/test_inflate.c of the reference.
* Single job descriptors are run through both engines and compared field by field: decompress
  jobs are bit-exact (output bytes, tpbc, SFBT, SUBC, rembytecnt, DHT, checksums); compress jobs
  must decode to the source and agree on spbc / checksums / completion code.
"""
import ctypes as C
import json
import os
import random
import subprocess
import sys
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "nx_dropin_driver.py")
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "libnxz_ref.so")
REF_GPU = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")


def _drive(lib, log2):
    p = subprocess.run([sys.executable, DRIVER, lib, str(log2)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(REF_CPU), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_host_code_over_cpu_engine():
    rep = _drive(REF_CPU, 18)
    assert set(rep["cases"]) == {"alice", "text", "zeros", "random", "tiny"}
    assert rep["gz_file"] > 0                     # gzopen / gzwrite / gzread / gzclose (lib/nx_gzlib.c) round trip
    assert all(x > 0 for x in rep["dictionary"])  # deflate/inflateSetDictionary both ways (test/test_dict.c)
    assert rep["zero_input_and_reset"]            # test/test_zeroinput.c, test/test_reset.c


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="power-gzip_b200/libnxz_gpu.so not built (needs /root/reference at build time)")
def test_reference_host_code_over_gpu_engine():
    rep = _drive(REF_GPU, 20)
    # the GPU engine's jobs compress for real: the reference's compress2 over it lands near zlib
    assert rep["cases"]["alice"]["compress2"] < 70000, rep
    assert 0 < rep["gz_file"] < 40000, rep        # the gz* file layer over the GPU engine; 60 000 bytes of text
    assert all(x > 0 for x in rep["dictionary"]) and rep["zero_input_and_reset"], rep


def _stress(lib, threads, iterations):
    p = subprocess.run([sys.executable, DRIVER, lib, "stress", str(threads), str(iterations)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(REF_CPU), reason="oracle/_ref not built (needs /root/reference)")
def test_multithread_stress_over_cpu_engine():
    # reference test/test_multithread_stress.c: concurrent compress()/uncompress() of ten buffers per thread
    rep = _stress(REF_CPU, 4, 1)
    assert rep["errors"] == [], rep


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="power-gzip_b200/libnxz_gpu.so not built (needs /root/reference at build time)")
def test_multithread_stress_is_coalesced_on_the_gpu():
    # descriptors from concurrent z_streams share GPU launches, results stay per stream
    rep = _stress(REF_GPU, 32, 2)
    assert rep["errors"] == [], rep
    assert rep["jobs"] >= 32 * 2 * 10 and rep["max_batch"] > 1, rep
    one = _stress(REF_GPU, 1, 2)
    assert one["errors"] == [] and one["max_batch"] == 1, one


# ---------------------------------------------------------------------------------------------
# single descriptors
# ---------------------------------------------------------------------------------------------
from nxjob import Job  # noqa: E402  (tests/nxjob.py: the 2048-byte descriptor)


@pytest.fixture(scope="module")
def engines(pg, oracle):
    lib = pg.load_library()

    class Dev(C.Structure):
        _fields_ = [("i", C.c_int * 8), ("paste_addr", C.c_void_p), ("fd", C.c_int), ("function", C.c_int), ("pad", C.c_char * 256)]
    dev = Dev()
    assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0
    oracle.oracle_nxemu_run_job.argtypes = [C.c_void_p]
    oracle.oracle_nxemu_run_job.restype = C.c_int

    def gpu(job):
        rc = lib.nxu_run_job(job.addr, C.byref(dev))
        assert rc == 0 and job.valid() == 1
        return job

    def cpu(job):
        assert oracle.oracle_nxemu_run_job(job.addr) == 0 and job.valid() == 1
        return job
    yield gpu, cpu
    lib.nx_function_end(C.byref(dev))


def _streams(data):
    yield "dyn6", zlib.compress(data, 6)[2:-4]
    yield "dyn1", zlib.compress(data, 1)[2:-4]
    yield "stored", zlib.compress(data, 0)[2:-4]
    fx = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    yield "fixed", fx.compress(data) + fx.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = b"".join(co.compress(data[i:i + 30000]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), 30000))
    yield "multiblock", parts + co.flush()


@pytest.mark.gpu
def test_decompress_jobs_match_the_oracle_field_by_field(engines, alice):
    """Cut raw deflate streams at arbitrary byte positions and chain jobs exactly like
    lib/nx_inflate.c:1447-1609 does (FC 0x10 first, then 0x14 with history + SFBT/SUBC/DHT fed
    back); after every job the GPU's CSB/CPB must equal the CPU engine's, and so must the bytes."""
    gpu, cpu = engines
    rnd = random.Random(3)
    data = alice[:90000] + bytes(5000) + rnd.randbytes(3000) + alice[:20000]
    for name, stream in _streams(data):
        for piece in (len(stream), 7001, 257):
            outs = {"gpu": bytearray(), "cpu": bytearray()}
            state = {k: dict(pos=0, subc=0, sfbt=0, rem=0, dht=b"", crc=0, adler=1, first=True) for k in outs}
            for step in range(100000):
                done = 0
                jobs = {}
                for k, run in (("gpu", gpu), ("cpu", cpu)):
                    st, produced = state[k], outs[k]
                    if st.get("final"):
                        done += 1
                        continue
                    hist = bytes(produced[-32768:])
                    hist = bytes((-len(hist)) % 16) + hist            # whole quadwords (lib/nx_deflate.c:853)
                    chunk = stream[st["pos"]: st["pos"] + piece + st.get("extra", 0)]
                    fc = 0x10 if st["first"] else 0x14
                    j = Job(fc, [hist, chunk] if hist else [chunk], 200000, histlen_qw=len(hist) // 16, subc=st["subc"] % 8,
                            sfbt=st["sfbt"], rem_or_dhtlen=st["rem"], dht=st["dht"], crc=st["crc"], adler=st["adler"],
                            split_dst=(step % 2 == 1))
                    run(j)
                    jobs[k] = j
                    assert j.cc() == 3 and j.ce() & 4, (name, piece, step, k, j.cc(), j.ce())
                    produced += j.out()
                    sfbt, subc = (j.w396() >> 16) & 0xf, j.w392() & 0xffff
                    used = len(chunk) - (subc + 7) // 8 if sfbt else len(chunk) - subc // 8
                    st.update(pos=st["pos"] + used, subc=subc, sfbt=sfbt, rem=j.w396() & 0xffff, first=False,
                              dht=j.get(256 + 400, 288) if (sfbt & 0xe) == 0xc else b"", crc=j.crc(), adler=j.adler())
                    # no progress (e.g. a block header longer than the piece): give more source, lib/nx_inflate.c:1222-1240
                    st["extra"] = st.get("extra", 0) + piece if used == 0 and j.tpbc() == 0 else 0
                    if sfbt == 0:
                        st["final"] = True
                if done == 2:
                    break
                g, c = jobs["gpu"], jobs["cpu"]
                assert g.tpbc() == c.tpbc() and g.out() == c.out(), (name, piece, step)
                assert (g.w392() & 0xffff, g.w396(), g.spbc_decomp()) == (c.w392() & 0xffff, c.w396(), c.spbc_decomp()), (name, piece, step)
                assert (g.crc(), g.adler()) == (c.crc(), c.adler()), (name, piece, step)
                if (g.w396() >> 16) & 0xe == 0xc:
                    nbits = g.w396() & 0xfff
                    assert g.get(256 + 400, (nbits + 7) // 8) == c.get(256 + 400, (nbits + 7) // 8), (name, piece, step)
            assert bytes(outs["gpu"]) == data and bytes(outs["cpu"]) == data, (name, piece)
            assert state["gpu"]["crc"] == zlib.crc32(data) and state["gpu"]["adler"] == zlib.adler32(data)


@pytest.mark.gpu
def test_decompress_job_errors_and_target_space(engines, alice):
    gpu, cpu = engines
    stream = zlib.compress(alice, 6)[2:-4]
    for run in (gpu, cpu):
        j = run(Job(0x10, [stream], 1000))                      # target too small: ERR_NX_TARGET_SPACE
        assert j.cc() == 13
        bad = bytearray(stream); bad[len(bad) // 2] ^= 0x55
        j = run(Job(0x10, [bytes(bad)], 200000))
        assert j.cc() in (66, 67, 68, 3)
        j = run(Job(0x10, [b"\x07"], 100))                      # reserved block type
        assert j.cc() in (66, 68)


def _decode_one_block(fc, job, dht_bits=b"", dht_len=0):
    """the job's target holds one deflate block without BFINAL; close the stream and inflate it"""
    out = bytearray(job.out())
    tebc = (job.w392() >> 16) & 7
    bits = (len(out) - 1) * 8 + (tebc or 8) if out else 0
    # append an empty final stored block behind the last valid bit
    stream = int.from_bytes(bytes(out), "little") & ((1 << bits) - 1)
    stream |= 1 << bits                      # BFINAL=1, BTYPE=00
    bits += 3
    bits = (bits + 7) & ~7
    stream |= 0xffff0000 << bits
    bits += 32
    return stream.to_bytes(bits // 8, "little")


@pytest.mark.gpu
def test_compress_jobs(engines, oracle, alice):
    """FHT / DHT / COUNT / RESUME function codes (inc_nx/nxu.h:803-811): the block must decode to the
    new source bytes (history as dictionary), spbc counts history, checksums continue the seeds, and
    the COUNT variants report a histogram that matches the emitted block."""
    gpu, cpu = engines
    data = alice[:120000]
    hist = alice[100000:100000 + 32768]
    # a dynamic table the way the reference makes one: from the LZ counts of a COUNT job
    for run in (gpu, cpu):
        j = run(Job(0x04, [data], 300000))                      # COMPRESS_FHT_COUNT
        assert j.cc() == 0 and j.spbc_comp(True) == len(data)
        lz = [int.from_bytes(j.get(256 + 400 + 4 * i, 4), "big") for i in range(316)]
        assert lz[256] == 1 and sum(lz[:256]) + sum(lz[257:286]) > 0
        assert sum(lz[257:286]) == sum(lz[286:])                # every length has a distance
        got = zlib.decompress(_decode_one_block(0x04, j), -15)
        assert got == data
        assert j.crc() == zlib.crc32(data) and j.adler() == zlib.adler32(data)
        # resume with history: spbc includes it, checksums continue
        hq = hist
        j2 = run(Job(0x08, [hq, data], 300000, histlen_qw=len(hq) // 16, crc=zlib.crc32(b"abc"), adler=zlib.adler32(b"abc")))
        assert j2.cc() == 0 and j2.spbc_comp(False) == len(hq) + len(data)
        d = zlib.decompressobj(-15, zdict=hq)
        assert d.decompress(_decode_one_block(0x08, j2)) == data
        assert j2.crc() == zlib.crc32(data, zlib.crc32(b"abc")) and j2.adler() == zlib.adler32(data, zlib.adler32(b"abc"))
    # DHT: take a real dynamic header from zlib, hand it over as cpb.in_dht (bits from HLIT on)
    z = zlib.compress(alice, 6)[2:-4]
    hdr = int.from_bytes(z[:400], "little")
    assert hdr & 7 in (4, 5)                                     # BTYPE=10
    lens = C.create_string_buffer(320)
    # find the header length by letting the oracle parse it
    job = (C.c_uint8 * 1)()
    del job
    from_hlit = hdr >> 3
    # generous: 288 bytes of bits; the engines read only what the header needs, dhtlen must be exact, so search it
    blob = from_hlit.to_bytes(400, "little")[:288]
    ok = None
    for nbits in range(60, 288 * 8):
        j = cpu(Job(0x02, [b"aaaa"], 1000, rem_or_dhtlen=nbits, dht=blob))
        if j.cc() in (0, 64):
            ok = nbits
            break
    assert ok, "no dynamic header length accepted"
    text = alice[:60000]
    for run in (gpu, cpu):
        j = run(Job(0x02, [text], 300000, rem_or_dhtlen=ok, dht=blob))
        if j.cc() == 66:
            continue                                             # zlib's table lacks a symbol the parse needs: legal outcome
        assert j.cc() == 0
        assert zlib.decompress(_decode_one_block(0x02, j), -15) == text
    # wrap
    for run in (gpu, cpu):
        j = run(Job(0x1e, [alice[:5000], alice[5000:60000]], 60000, split_dst=True))
        assert j.cc() == 0 and j.out() == alice[:60000]
        assert j.crc() == zlib.crc32(alice[:60000]) and j.adler() == zlib.adler32(alice[:60000])
        j = run(Job(0x1e, [alice[:5000]], 100))
        assert j.cc() == 13


@pytest.mark.gpu
def test_mixed_descriptors_from_many_threads_coalesce_and_stay_exact(engines, pg, alice):
    """descriptors submitted concurrently (compress FHT/COUNT, decompress fresh and
    resumed with history, wrap) are run as shared GPU batches; every one must come back exactly as if it
    had been alone: decompress/wrap bit-exact with the CPU engine, compress decoding to its source."""
    import threading
    gpu, cpu = engines
    lib = pg.load_library()
    rnd = random.Random(11)

    def make(i):
        k = i % 4
        data = (alice[(i * 997) % 50000:][: 3000 + 2311 * (i % 13)]) + rnd.randbytes(50 * (i % 3))
        if k == 0:                                   # fresh decompress of a whole raw stream
            return "dec", data, lambda: Job(0x10, [zlib.compress(data, 6)[2:-4]], len(data) + 100, split_dst=(i % 8 == 0))
        if k == 1:                                   # resumed decompress: second half, first half as history
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            a = co.compress(data[: len(data) // 2]) + co.flush(zlib.Z_FULL_FLUSH)
            b = co.compress(data[len(data) // 2:]) + co.flush()
            hist = data[: len(data) // 2][-32768:]
            hist = bytes((-len(hist)) % 16) + hist
            return "dec2", data[len(data) // 2:], lambda: Job(0x14, [hist, b], len(data) + 100, histlen_qw=len(hist) // 16)
        if k == 2:                                   # wrap: copy + checksums
            return "wrap", data, lambda: Job(0x1e, [data], len(data))
        return "comp", data, lambda: Job(0x04 if i % 8 == 3 else 0x00, [data], 2 * len(data) + 1000)

    specs = [make(i) for i in range(96)]
    jobs = [mk() for _, _, mk in specs]
    refs = [cpu(mk()) for _, _, mk in specs]
    b0, j0, m0 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.nxgpu_job_stats(0, C.byref(b0), C.byref(j0), C.byref(m0))
    start = threading.Barrier(len(jobs))
    errs = []

    def worker(j):
        try:
            start.wait()
            gpu(j)
        except Exception as e:                       # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    b1, j1, m1 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.nxgpu_job_stats(0, C.byref(b1), C.byref(j1), C.byref(m1))
    assert j1.value - j0.value == len(jobs)
    assert b1.value - b0.value < len(jobs), "no descriptor shared a batch"
    for (kind, data, _), g, c in zip(specs, jobs, refs):
        assert g.cc() == c.cc(), (kind, g.cc(), c.cc())
        if kind in ("dec", "dec2"):
            assert g.out() == c.out() == data, kind
            assert (g.tpbc(), g.w392() & 0xffff, g.w396(), g.spbc_decomp()) == (c.tpbc(), c.w392() & 0xffff, c.w396(), c.spbc_decomp()), kind
            assert (g.crc(), g.adler()) == (c.crc(), c.adler()), kind
        elif kind == "wrap":
            assert g.out() == data and (g.crc(), g.adler()) == (zlib.crc32(data), zlib.adler32(data))
        else:
            assert zlib.decompress(_decode_one_block(0, g), -15) == data
            assert (g.crc(), g.adler()) == (zlib.crc32(data), zlib.adler32(data))


@pytest.mark.gpu
def test_dhtgen_against_the_reference_tables(engines, pg, alice):
    """nxgpu_dhtgen stands in for lib/nx_dhtgen.c:945 dhtgen().  For the histograms of
    tests/golden/dhtgen_vectors.json (tables there were made by the reference's own code) the GPU's table must
    be a complete code of at most 15 bits covering every counted symbol and must not cost more bits than the
    reference's; the single and the batch call agree; and a table made from a COUNT job's lzcounts drives a
    DHT compress job on both engines whose block decodes to the source."""
    from dht_util import parse_dht, block_cost, kraft
    gpu, cpu = engines
    vec = json.load(open(os.path.join(ROOT, "tests", "golden", "dhtgen_vectors.json")))["vectors"]
    with pg.Engine(0) as eng:
        batch = eng.dhtgen_batch([v["counts"] for v in vec])
        for v, (bdht, bbits) in zip(vec, batch):
            c = v["counts"]
            dht, bits = eng.dhtgen(c[:286], c[286:])
            assert (dht, bits) == (bdht, bbits), v["name"]
            ll, dd = parse_dht(dht, bits)
            assert max(ll) <= 15 and max(dd) <= 15
            assert abs(kraft(ll) - 1.0) < 1e-9 and abs(kraft(dd) - 1.0) < 1e-9, v["name"]
            rl, rd = parse_dht(bytes.fromhex(v["ref_dht_hex"]), v["ref_bits"])
            assert block_cost(c, ll, dd, bits) <= block_cost(c, rl, rd, v["ref_bits"]), v["name"]
        # sparse histograms: symbols without a count get no code; single-symbol trees are completed
        for counts in ([0] * 97 + [5] + [0] * 158 + [1] + [0] * 29 + [0] * 30,               # only 'a' and EOB, no distances
                       [3] * 256 + [1] + [7] + [0] * 28 + [9] + [0] * 29):                     # one length, one distance
            dht, bits = eng.dhtgen(counts[:286], counts[286:])
            ll, dd = parse_dht(dht, bits)
            assert all(ll[s] for s in range(286) if counts[s]) and all(dd[s] for s in range(30) if counts[286 + s])
            assert kraft(ll) <= 1.0 + 1e-9 and kraft(dd) <= 1.0 + 1e-9
        # the reference's flow (lib/nx_dht.c:568-676): COUNT job -> lzcounts -> dhtgen -> DHT job
        text = alice[20000:90000]
        j = gpu(Job(0x04, [text], 300000))
        lz = [max(1, int.from_bytes(j.get(256 + 400 + 4 * i, 4), "big")) for i in range(316)]
        dht, bits = eng.dhtgen(lz[:286], lz[286:])
        for run in (gpu, cpu):
            j = run(Job(0x02, [text], 300000, rem_or_dhtlen=bits, dht=dht))
            assert j.cc() == 0, j.cc()
            assert zlib.decompress(_decode_one_block(0x02, j), -15) == text


@pytest.mark.gpu
def test_large_compress_descriptors_are_cut_into_pieces(engines, pg, alice):
    """A large compress descriptor is compressed by several CTAs (pieces of 8-64 KiB, same table, each
    primed with the 32 KiB in front of it) and the bit strings are joined into the ONE block the descriptor asks
    for: it must decode to the source for FHT / DHT / COUNT / RESUME, with spbc, checksums, lzcounts and the
    completion code as for a single piece; a table lacking a needed symbol still gives CC=66."""
    gpu, cpu = engines
    data = pg.makedata(4, 20, alice)[:1000003] + alice[:70001]            # 1,070,004 bytes: 17 pieces, ragged tail
    hist = alice[100000:100000 + 32768]
    for fc in (0x00, 0x04):                                                   # FHT, FHT + COUNT
        j = gpu(Job(fc, [data], 2 * len(data)))
        assert j.cc() == 0 and j.spbc_comp(fc == 0x04) == len(data)
        assert zlib.decompress(_decode_one_block(fc, j), -15) == data
        assert j.crc() == zlib.crc32(data) and j.adler() == zlib.adler32(data)
        if fc == 0x04:
            lz = [int.from_bytes(j.get(256 + 400 + 4 * i, 4), "big") for i in range(316)]
            assert lz[256] == 1 and sum(lz[257:286]) == sum(lz[286:316]) > 0      # one distance per length
            # the histogram is the histogram of the emitted block: a table made from it costs what the block costs
            with pg.Engine(0) as eng:
                dht, bits = eng.dhtgen([max(1, x) for x in lz[:286]], [max(1, x) for x in lz[286:]])
            jd = gpu(Job(0x02, [data], 2 * len(data), rem_or_dhtlen=bits, dht=dht))
            assert jd.cc() == 0 and zlib.decompress(_decode_one_block(0x02, jd), -15) == data
            assert jd.tpbc() < j.tpbc()                                       # dynamic beats fixed on this text
            jc = cpu(Job(0x02, [data[:300000]], 600000, rem_or_dhtlen=bits, dht=dht))
            assert jc.cc() == 0                                               # the CPU engine accepts the same table
    # resume with history in front: pieces further in use the source itself as their window
    j2 = gpu(Job(0x08, [hist, data], 2 * len(data), histlen_qw=len(hist) // 16, crc=zlib.crc32(b"abc"), adler=zlib.adler32(b"abc")))
    assert j2.cc() == 0 and j2.spbc_comp(False) == len(hist) + len(data)
    assert zlib.decompressobj(-15, zdict=hist).decompress(_decode_one_block(0x08, j2)) == data
    assert j2.crc() == zlib.crc32(data, zlib.crc32(b"abc"))
    # target too small: CC=13, nothing written past the target
    j3 = gpu(Job(0x00, [data], 50000))
    assert j3.cc() == 13
    # a table without codes for bytes that occur: CC=66 like a single piece
    with pg.Engine(0) as eng:
        dht, bits = eng.dhtgen([1 if 97 <= s <= 122 or s >= 256 else 0 for s in range(286)], [1] * 30)
    j4 = gpu(Job(0x02, [data], 2 * len(data), rem_or_dhtlen=bits, dht=dht))
    assert j4.cc() == 66
    # through the reference's own host code: deflate() of a multi-MiB buffer issues such descriptors
    rep = _drive(REF_GPU, 22) if os.path.exists(REF_GPU) else None
    assert rep is None or rep["cases"]["text"]["compress2"] > 0


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="power-gzip_b200/libnxz_gpu.so not built (needs /root/reference at build time)")
def test_ld_preload_puts_an_unmodified_program_on_the_gpu():
    """reference README.md:9-18: LD_PRELOAD libnxz under an unmodified zlib user — here CPython's own zlib module."""
    code = (
        "import zlib, gzip, ctypes\n"
        f"d = gzip.decompress(open({os.path.join(ROOT, 'tests', 'golden', 'alice29.txt.gz')!r}, 'rb').read())\n"
        "z = zlib.compress(d, 6)\n"
        "assert zlib.decompress(z) == d\n"
        "co = zlib.compressobj(6, zlib.DEFLATED, 31)\n"
        "g = b''.join(co.compress(d[i:i + 40000]) for i in range(0, len(d), 40000)) + co.flush()\n"
        "assert zlib.decompressobj(31).decompress(g) == d\n"
        "lib = ctypes.CDLL(None)\n"
        "b, j, m = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()\n"
        "lib.nxgpu_job_stats(0, ctypes.byref(b), ctypes.byref(j), ctypes.byref(m))\n"
        "print(len(z), j.value)\n")
    env = dict(os.environ, LD_PRELOAD=REF_GPU, NX_GZIP_TYPE_SELECTOR="2", NX_GZIP_LOGFILE="/tmp/nx_preload.log")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-3000:]
    size, descriptors = (int(x) for x in p.stdout.split()[-2:])
    assert descriptors >= 4 and size < 80000, (size, descriptors)
    # the same program without the preload inflates what the GPU wrote (checked inside), and system zlib agrees on the data


# ---------------------------------------------------------------------------------------------
# north star: "deflate output must be a valid stream that both system zlib and libnxz's own inflate decode"
# ---------------------------------------------------------------------------------------------
_NX_INFLATE_CHILD = r"""
import ctypes as C, os, sys, json
os.environ["NX_GZIP_TYPE_SELECTOR"] = "2"
os.environ.setdefault("NX_GZIP_LOGFILE", "/tmp/nx_dropin.log")
lib = C.CDLL(sys.argv[1])
class ZStream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong), ("next_out", C.c_void_p),
                ("avail_out", C.c_uint), ("total_out", C.c_ulong), ("msg", C.c_char_p), ("state", C.c_void_p),
                ("zalloc", C.c_void_p), ("zfree", C.c_void_p), ("opaque", C.c_void_p), ("data_type", C.c_int),
                ("adler", C.c_ulong), ("reserved", C.c_ulong)]
lib.nx_inflateInit2_.argtypes = [C.POINTER(ZStream), C.c_int, C.c_char_p, C.c_int]
lib.nx_inflate.argtypes = [C.POINTER(ZStream), C.c_int]
lib.nx_inflateEnd.argtypes = [C.POINTER(ZStream)]
rep = []
for spec in json.load(open(sys.argv[2])):
    blob = open(spec["blob"], "rb").read()
    s = ZStream()
    assert lib.nx_inflateInit2_(C.byref(s), spec["wbits"], b"1.2.11", C.sizeof(ZStream)) == 0
    src = C.create_string_buffer(blob, len(blob)); out = C.create_string_buffer(spec["n"] + 64)
    got = bytearray(); pos = 0; rc = 0; guard = 0
    while rc == 0 and guard < 100000:
        if s.avail_in == 0 and pos < len(blob):
            take = min(spec["piece"], len(blob) - pos)
            s.next_in, s.avail_in = C.addressof(src) + pos, take
            pos += take
        s.next_out, s.avail_out = C.addressof(out), len(out)
        rc = lib.nx_inflate(C.byref(s), 0)
        got += out.raw[: len(out) - s.avail_out]
        guard += 1
    lib.nx_inflateEnd(C.byref(s))
    open(spec["blob"] + ".out", "wb").write(bytes(got))
    rep.append({"rc": rc, "total_in": s.total_in, "n": len(got)})
print(json.dumps(rep))
"""


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_CPU), reason="oracle/_ref not built (needs /root/reference)")
def test_batch_deflate_output_decodes_with_libnxz_own_inflate(engine, pg, alice, tmp_path):
    """Row N1: what nxgpu_deflate_stream writes (primed and independent chunks, raw / zlib / gzip wrappers, levels 1 and 6)
    goes through nx_inflate() of the reference's own host code (lib/nx_inflate.c:277 over the CPU engine, NX mode) —
    in one piece and in 50 000-byte pieces — and must come back bit-exact with Z_STREAM_END and total_in == its length."""
    data = pg.makedata(4, 21, alice)[: (1 << 21) - 12345] + alice[:7777]
    specs = []
    for level in (1, 6):
        for wrap, wbits in ((pg.WRAP_RAW, -15), (pg.WRAP_ZLIB, 15), (pg.WRAP_GZIP, 31)):
            for flag in (0, pg.STREAM_INDEPENDENT):
                blob = engine.compress(data, level=level, wrap=wrap | flag, chunk=65536 if flag else 0)
                for piece in (len(blob), 50000):
                    path = tmp_path / f"l{level}_w{wrap}_{flag}_{piece}.z"
                    path.write_bytes(blob)
                    specs.append({"blob": str(path), "wbits": wbits, "n": len(data), "piece": piece, "len": len(blob)})
    (tmp_path / "specs.json").write_text(json.dumps(specs))
    p = subprocess.run([sys.executable, "-c", _NX_INFLATE_CHILD, REF_CPU, str(tmp_path / "specs.json")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    rep = json.loads(p.stdout.strip().splitlines()[-1])
    assert len(rep) == len(specs) == 24
    for spec, r in zip(specs, rep):
        assert r["rc"] == 1 and r["total_in"] == spec["len"] and r["n"] == len(data), (spec, r)
        assert open(spec["blob"] + ".out", "rb").read() == data, spec
