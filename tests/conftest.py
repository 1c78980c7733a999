import ctypes
import gzip
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_package():
    """The package directory is named power-gzip_b200 (not an identifier), so load it by path."""
    if "power_gzip_b200" in sys.modules:
        return sys.modules["power_gzip_b200"]
    spec = importlib.util.spec_from_file_location(
        "power_gzip_b200", os.path.join(ROOT, "power-gzip_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "power-gzip_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["power_gzip_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pg():
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    """liboracle.so — the CPU checker (tests only)."""
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    lib = ctypes.CDLL(path)
    u32, u64, vp, sz = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t
    lib.oracle_crc32.restype = u32
    lib.oracle_crc32.argtypes = [u32, ctypes.c_char_p, sz]
    lib.oracle_crc32_raw.restype = u32
    lib.oracle_crc32_raw.argtypes = [u32, ctypes.c_char_p, sz]
    lib.oracle_crc32_combine.restype = u32
    lib.oracle_crc32_combine.argtypes = [u32, u32, u64]
    lib.oracle_adler32.restype = u32
    lib.oracle_adler32.argtypes = [u32, ctypes.c_char_p, sz]
    lib.oracle_adler32_combine.restype = u32
    lib.oracle_adler32_combine.argtypes = [u32, u32, u64]
    lib.oracle_inflate_member.restype = ctypes.c_int
    lib.oracle_inflate_member.argtypes = [ctypes.c_char_p, sz, ctypes.c_char_p, sz, ctypes.c_int,
                                          ctypes.POINTER(sz), ctypes.POINTER(sz), ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.oracle_huff_lengths.argtypes = [ctypes.POINTER(u32), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint8)]
    lib.oracle_dynblock_bits.restype = u64
    lib.oracle_dynblock_bits.argtypes = [ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.oracle_makedata.restype = u64
    lib.oracle_makedata.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, u64, ctypes.c_char_p, u64]
    return lib


def oracle_inflate(lib, blob: bytes, cap: int, wrap: int = 3):
    out = ctypes.create_string_buffer(max(cap, 1))
    ol, used = ctypes.c_size_t(), ctypes.c_size_t()
    crc, adler = ctypes.c_uint32(), ctypes.c_uint32()
    rc = lib.oracle_inflate_member(blob, len(blob), out, cap, wrap, ctypes.byref(ol), ctypes.byref(used),
                                   ctypes.byref(crc), ctypes.byref(adler))
    return rc, out.raw[:ol.value], used.value, crc.value, adler.value


@pytest.fixture(scope="session")
def alice():
    return gzip.decompress(open(os.path.join(GOLDEN, "alice29.txt.gz"), "rb").read())


@pytest.fixture(scope="session")
def vectors():
    return json.load(open(os.path.join(GOLDEN, "ref_vectors.json")))


@pytest.fixture(scope="session")
def engine(pg):
    eng = pg.Engine(0)
    yield eng
    eng.close()
