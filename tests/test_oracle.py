"""CPU tests: the oracle (oracle/*.c) against the reference's golden vectors and against
zlib 1.3, the library the reference's software path dlopens (lib/sw_zlib.c:283-327)."""
import ctypes
import gzip
import hashlib
import json
import os
import random
import zlib

import pytest

from conftest import GOLDEN, oracle_inflate


def test_alice_fixture_is_the_reference_file(alice, vectors):
    # sha256 pinned by the reference itself: oct/snappy-alice29.source:3
    assert hashlib.sha256(alice).hexdigest().startswith("7467306e")
    assert len(alice) == vectors["alice29"]["len"] == 152089


@pytest.mark.parametrize("name", ["crc32", "adler32"])
def test_reference_known_answer_vectors(oracle, name):
    # test/test_crc32.c:38-180 (141 cases) and test/test_adler32.c:38-179 (140 cases)
    kats = json.load(open(os.path.join(GOLDEN, f"kat_{name}.json")))
    assert len(kats) == (141 if name == "crc32" else 140)
    fn = oracle.oracle_crc32 if name == "crc32" else oracle.oracle_adler32
    for k in kats:
        buf = None if k["null"] else bytes.fromhex(k["data_hex"])
        assert fn(k["seed"], buf, k["len"]) == k["expect"], k


def test_checksums_match_zlib_and_reference_pins(oracle, alice, vectors):
    assert oracle.oracle_crc32(0, alice, len(alice)) == vectors["alice29"]["crc32"] == 0x66007dba
    assert oracle.oracle_adler32(1, alice, len(alice)) == vectors["alice29"]["adler32"] == 0xc39d8c10
    rnd = random.Random(7)
    for n in (0, 1, 15, 16, 17, 5551, 5552, 5553, 70000):
        d = rnd.randbytes(n)
        seed = rnd.getrandbits(32)
        assert oracle.oracle_crc32(seed, d, n) == zlib.crc32(d, seed)
        assert oracle.oracle_adler32(seed, d, n) == zlib.adler32(d, seed)
        # raw update convention of __crc32_vpmsum (lib/crc32_ppc.c:30)
        assert oracle.oracle_crc32_raw(seed ^ 0xffffffff, d, n) ^ 0xffffffff == zlib.crc32(d, seed)


def test_combine_matches_concatenation(oracle, alice):
    rnd = random.Random(3)
    for _ in range(50):
        a = rnd.randrange(0, 5000)
        b = rnd.randrange(0, 100000)
        x, y = alice[:a], alice[a:a + b]
        assert oracle.oracle_crc32_combine(zlib.crc32(x), zlib.crc32(y), len(y)) == zlib.crc32(x + y)
        assert oracle.oracle_adler32_combine(zlib.adler32(x), zlib.adler32(y), len(y)) == zlib.adler32(x + y)
    # huge lengths only exercise the operator powers
    c1, c2 = 0x12345678, 0x9abcdef0
    big = (1 << 34) + 12345
    step = oracle.oracle_crc32_combine(oracle.oracle_crc32_combine(c1, 0, 1 << 34), c2, 12345)
    assert oracle.oracle_crc32_combine(c1, c2, big) == step


def test_product_host_combine_matches_oracle(pg, oracle):
    # nxgpu_crc32_combine / nxgpu_adler32_combine are host arithmetic inside the product library
    lib = pg.load_library()
    rnd = random.Random(11)
    for _ in range(200):
        c1, c2 = rnd.getrandbits(32), rnd.getrandbits(32)
        n = rnd.choice([0, 1, 7, 4096, 262144, (1 << 32) + 5, rnd.getrandbits(40)])
        assert lib.nxgpu_crc32_combine(c1, c2, n) == oracle.oracle_crc32_combine(c1, c2, n)
        a1 = rnd.randrange(65521) | rnd.randrange(65521) << 16
        a2 = rnd.randrange(65521) | rnd.randrange(65521) << 16
        assert lib.nxgpu_adler32_combine(a1, a2, n) == oracle.oracle_adler32_combine(a1, a2, n)


def _streams(data):
    yield "gzip6", gzip.compress(data, 6, mtime=0), 2
    yield "zlib1", zlib.compress(data, 1), 1
    yield "zlib9", zlib.compress(data, 9), 1
    yield "stored", zlib.compress(data, 0), 1
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    yield "rawfixed", c.compress(data) + c.flush(), 0
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = []
    for i in range(0, len(data), 20000):
        parts.append(c.compress(data[i:i + 20000]) + c.flush(zlib.Z_SYNC_FLUSH))
    yield "rawsync", b"".join(parts) + c.flush(), 0


def test_inflate_oracle_matches_zlib(oracle, alice):
    rnd = random.Random(5)
    for data in (b"", b"a", alice, bytes(70000), rnd.randbytes(40000), alice[:1000] * 40):
        for name, blob, wrap in _streams(data):
            for w in (wrap, 3) if wrap else (0,):
                rc, out, used, crc, adler = oracle_inflate(oracle, blob, len(data) + 8, w)
                assert rc == 0, (name, len(data))
                assert out == data
                assert used == len(blob)
                assert crc == zlib.crc32(data) and adler == zlib.adler32(data)


def test_inflate_oracle_reference_scp_stream(oracle):
    # the 611-byte stream of test/test_buf_error.c:107: must decode exactly like zlib
    blob = bytes.fromhex(json.load(open(os.path.join(GOLDEN, "scp_stream.json")))["zlib_stream_hex"])
    d = zlib.decompressobj()
    want = d.decompress(blob)
    rc, out, used, crc, adler = oracle_inflate(oracle, blob, len(want) + 64, 1)
    # the capture is a sync-flushed prefix of a longer stream: no final block, so rc is a data error
    # in one-shot mode, but every byte zlib produced must already be there
    assert out[:len(want)] == want and len(out) == len(want)
    assert rc in (0, -3)


def test_inflate_oracle_rejects_corruption(oracle, alice):
    blob = bytearray(zlib.compress(alice, 6))
    rc, *_ = oracle_inflate(oracle, bytes(blob[:-1]), len(alice), 1)
    assert rc == -3                       # truncated trailer
    blob[-1] ^= 1
    rc, *_ = oracle_inflate(oracle, bytes(blob), len(alice), 1)
    assert rc == -3                       # adler mismatch
    rc, *_ = oracle_inflate(oracle, zlib.compress(alice, 6), 1000, 1)
    assert rc == -5                       # target too small (the NX CC=13 case)


def test_makedata_matches_reference_pins(pg, oracle, alice, vectors):
    # samples/makedata.c:35-70; pins were produced by the reference binary (tests/golden/make_fixtures.py)
    for key, v in vectors["makedata"].items():
        seed, b = int(key.split("_")[0][1:]), int(key.split("_")[1][1:])
        if b > 24:
            continue
        cap = (1 << b) + (1 << b) // 10 + 16
        buf = ctypes.create_string_buffer(cap)
        n = oracle.oracle_makedata(seed, b, alice, len(alice), buf, cap)
        assert n == v["len"]
        assert zlib.crc32(buf.raw[:n]) == v["crc32"]
        assert pg.makedata(seed, b, alice) == buf.raw[:n]      # product generator == oracle generator


def test_huffman_oracle_is_valid_and_near_optimal(oracle):
    rnd = random.Random(9)
    u32 = ctypes.c_uint32
    for trial in range(30):
        n = rnd.choice([19, 30, 286])
        maxbits = 7 if n == 19 else 15
        if trial % 3 == 0:
            freq = [min(int(1.6 ** i), 1 << 24) for i in range(n)]   # Fibonacci-like: forces the length limit
        else:
            freq = [rnd.choice([0, 0, 1, 2, 50, 1000, rnd.randrange(1 << 18)]) for _ in range(n)]
        if sum(1 for f in freq if f) < 2:
            freq[0], freq[1] = 1, 1
        lens = (ctypes.c_uint8 * n)()
        oracle.oracle_huff_lengths((u32 * n)(*freq), n, maxbits, lens)
        assert all((l > 0) == (f > 0) for l, f in zip(lens, freq))
        assert max(lens) <= maxbits
        assert sum(2.0 ** -l for l in lens if l) <= 1.0 + 1e-12     # Kraft


def test_dynblock_cost_brackets_zlib(oracle, alice):
    # a single dynamic block over zlib's own token statistics can not be much worse than zlib's output
    ll = (ctypes.c_uint32 * 286)()
    d = (ctypes.c_uint32 * 30)()
    for b in alice:
        ll[b] += 1
    ll[256] = 1
    bits = oracle.oracle_dynblock_bits(ll, d)
    # order-0 entropy coding of alice29 is about 4.6 bits/byte
    assert 4.4 * len(alice) < bits < 4.8 * len(alice)


def test_reference_dhtgen_vectors_are_valid_tables(oracle):
    """tests/golden/dhtgen_vectors.json holds tables made by the reference's own lib/nx_dhtgen.c
    (oracle/_ref/dhtgen_test): each must parse as a complete, length-limited code covering every
    counted symbol, and the oracle's Huffman restatement must not cost more than the reference."""
    import ctypes
    from dht_util import parse_dht, block_cost, kraft
    vec = json.load(open(os.path.join(GOLDEN, "dhtgen_vectors.json")))["vectors"]
    assert len(vec) >= 6
    u32 = ctypes.c_uint32
    for v in vec:
        c = v["counts"]
        ll, dd = parse_dht(bytes.fromhex(v["ref_dht_hex"]), v["ref_bits"])
        assert max(ll) <= 15 and max(dd) <= 15
        assert abs(kraft(ll) - 1.0) < 1e-9 and abs(kraft(dd) - 1.0) < 1e-9, v["name"]
        ref_cost = block_cost(c, ll, dd, v["ref_bits"])
        ol = (ctypes.c_uint8 * 286)(); od = (ctypes.c_uint8 * 30)()
        oracle.oracle_huff_lengths((u32 * 286)(*c[:286]), 286, 15, ol)
        oracle.oracle_huff_lengths((u32 * 30)(*c[286:]), 30, 15, od)
        sym_bits_ref = ref_cost - 3 - v["ref_bits"]
        sym_bits_oracle = block_cost(c, list(ol), list(od), 0) - 3
        assert sym_bits_oracle <= sym_bits_ref, (v["name"], sym_bits_oracle, sym_bits_ref)
