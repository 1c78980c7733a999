"""Both engines against the INDEPENDENT per-function-code descriptor dumps of tests/golden/job_vectors.json.gz
(SURVEY.md §8c).  The vectors come from tests/golden/make_job_vectors.py — an RFC 1951 walker plus the NX-GZIP manual's
Table 5-3 / §2.4 semantics, sharing no code with oracle/ or csrc/ — so a convention that the CPU engine and the GPU
engine merely share (round 1: SUBC counting everything behind the final end-of-block, wrapping at 16 bits) shows up
here instead of cancelling out.

Covered: FC 0x10 / 0x14 (decompress, resume) and 0x12 / 0x16 (single block and suspend) over stored, fixed, dynamic and
multi-block streams; sources that end inside a block header, inside a stored / fixed / dynamic block, exactly on a
block boundary, and behind the final end-of-block with 0 .. 70 000 trailing bytes; resume with history, in_subc,
in_sfbt, in_rembytecnt and in_dht fed back the way lib/nx_inflate.c:1480-1609 does; the no-forward-progress case
(manual §5.2.5.6).  Every job checks CC, CE, TPBC, the target bytes, SFBT, SUBC, rembytecnt / dhtlen, out_dht, SPBC,
out_crc and out_adler."""
import base64
import ctypes as C
import gzip
import json
import os
import zlib

import pytest

from nxjob import Job

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vectors():
    doc = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "job_vectors.json.gz")).read())
    doc["data"] = zlib.decompress(base64.b64decode(doc["data_zlib_b64"]))
    doc["streams"] = {k: base64.b64decode(v) for k, v in doc["streams"].items()}
    return doc


def _check(run, doc):
    data = doc["data"]
    n = 0
    for v in doc["jobs"]:
        tail = bytes((i * 131 + 7) & 0xff for i in range(v["trailing"]))
        full = doc["streams"][v["stream"]] + tail
        hist = data[v["out_done"] - v["hist"]: v["out_done"]]
        hist = bytes((-len(hist)) % 16) + hist                     # whole quadwords (lib/nx_deflate.c:853)
        chunk = full[v["src_from"]: v["src_to"]]
        in_dyn = (v["in_sfbt"] & 0xe) == 0xc
        j = Job(v["fc"], [hist, chunk] if hist else [chunk], len(data) + 64, histlen_qw=len(hist) // 16, subc=v["in_subc"],
                sfbt=v["in_sfbt"], rem_or_dhtlen=v["in_dhtlen"] if in_dyn else v["in_rem"], dht=base64.b64decode(v["in_dht"]),
                crc=zlib.crc32(data[: v["out_done"]]), adler=zlib.adler32(data[: v["out_done"]]), split_dst=(n % 3 == 1))
        run(j)
        e = v["exp"]
        where = (v["stream"], v["plan"], v["step"], v["trailing"])
        assert j.valid() == 1 and j.cc() == e["cc"] and j.ce() & 4, (where, j.cc(), j.ce())
        assert j.tpbc() == e["tpbc"], (where, j.tpbc(), e["tpbc"])
        assert j.out() == data[v["out_done"]: v["out_done"] + e["tpbc"]], where
        sfbt, low = (j.w396() >> 16) & 0xf, j.w396() & 0xffff
        assert sfbt == e["sfbt"], (where, sfbt, e["sfbt"])
        assert j.w392() == e["subc"], (where, "subc", j.w392(), e["subc"])
        assert j.spbc_decomp() == e["spbc"], (where, "spbc", j.spbc_decomp(), e["spbc"])
        if (sfbt & 0xe) == 0x8:
            assert low == e["rem"], (where, "rembytecnt", low, e["rem"])
        if (sfbt & 0xe) == 0xc:
            assert (low & 0xfff) == e["dhtlen"], (where, "dhtlen", low, e["dhtlen"])
            want = base64.b64decode(e["dht"])
            got = j.get(256 + 400, len(want))
            if e["dhtlen"] & 7:
                mask = (1 << (e["dhtlen"] & 7)) - 1
                got, want = got[:-1] + bytes([got[-1] & mask]), want[:-1] + bytes([want[-1] & mask])
            assert got == want, (where, "out_dht")
        assert (j.crc(), j.adler()) == (e["crc"], e["adler"]), (where, "checksums")
        n += 1
    return n


def test_cpu_engine_matches_the_manual_derived_vectors(oracle, vectors):
    oracle.oracle_nxemu_run_job.argtypes = [C.c_void_p]
    oracle.oracle_nxemu_run_job.restype = C.c_int

    def run(job):
        assert oracle.oracle_nxemu_run_job(job.addr) == 0
    assert _check(run, vectors) == len(vectors["jobs"]) > 300


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["solo", "batch"])
def test_gpu_engine_matches_the_manual_derived_vectors(pg, vectors, kernel, monkeypatch):
    # a single descriptor runs on the one-warp-per-stream kernel (window in shared memory); NXGPU_INFLATE_SOLO_MAX=0 sends the
    # same descriptors through the batch kernel (window read from L1/L2)
    if kernel == "batch":
        monkeypatch.setenv("NXGPU_INFLATE_SOLO_MAX", "0")
    lib = pg.load_library()

    class Dev(C.Structure):
        _fields_ = [("i", C.c_int * 8), ("paste_addr", C.c_void_p), ("fd", C.c_int), ("function", C.c_int), ("pad", C.c_char * 256)]
    dev = Dev()
    assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0

    def run(job):
        assert lib.nxu_run_job(job.addr, C.byref(dev)) == 0
    try:
        assert _check(run, vectors) == len(vectors["jobs"])
    finally:
        lib.nx_function_end(C.byref(dev))


def test_vectors_keep_subc_inside_its_field(vectors):
    # the contract the vectors encode: SUBC never needs more than 16 bits (Table 5-3: at most 2285), and behind a final
    # end-of-block it is the 0..7 padding bits plus at most 8 bytes of read-ahead, whatever follows the stream
    for v in vectors["jobs"]:
        e = v["exp"]
        assert e["subc"] <= 2285
        if e["sfbt"] == 0:
            assert e["subc"] <= 71
            if v["trailing"] == 8:
                assert 64 <= e["subc"] <= 71          # a gzip trailer (inc_nx/nxu.h:454-465)
            if v["trailing"] == 4:
                assert 32 <= e["subc"] <= 39          # a zlib trailer
