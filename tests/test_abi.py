"""CPU tests: the C-ABI shared library loads and exports every symbol include/nxgpu.h declares.
No compute entry point is called here (there is no GPU in the build container)."""
import os
import re
import subprocess

from conftest import ROOT


def test_library_exports_every_declared_symbol(pg):
    lib = pg.load_library()
    header = open(os.path.join(ROOT, "include", "nxgpu.h")).read()
    declared = set(re.findall(r"\b(nxgpu_[a-z0-9_]+)\s*\(", header))
    declared |= {"nx_function_begin", "nx_function_end", "nx_wait_ticks", "nxu_run_job", "__crc32_vpmsum", "tb_freq"}
    assert declared <= set(pg.EXPORTS) | {"nxgpu_dev_prefix"}, declared - set(pg.EXPORTS)
    for name in pg.EXPORTS:
        assert hasattr(lib, name), name


def test_library_is_built_for_sm_100a_only():
    so = os.path.join(ROOT, "power-gzip_b200", "libnxgpu.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_engine_fails_loudly_without_gpu(pg):
    import torch
    if torch.cuda.is_available():
        return
    try:
        pg.Engine(0)
    except pg.NxGpuError as e:
        assert e.rc == pg.E_NODEV
    else:
        raise AssertionError("Engine() must not succeed without a GPU: there is no CPU fallback")


def test_product_does_not_touch_the_oracle():
    # the product tree must never import, link or execute anything under oracle/
    bad = []
    for base in ("power-gzip_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"liboracle|#include\s*[<\"][^>\"]*oracle|import\s+oracle|from\s+oracle|dlopen\([^)]*oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    so = os.path.join(ROOT, "power-gzip_b200", "libnxgpu.so")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "libz" not in needed


def test_drop_in_library_keeps_the_reference_abi():
    """The reference guards its ABI with test/test_abi (abidiff against test/libnxz.abi; abidiff is not in this image).
    Same intent with readelf: the drop-in library (reference host code over libnxgpu.so, linked with the reference's own
    version script lib/Versions) must define every function symbol of libnxz.abi under the same version node, and nothing
    of the engine's internals may leak into its dynamic symbol table."""
    import json
    import pytest
    so = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")
    if not os.path.exists(so):
        pytest.skip("power-gzip_b200/libnxz_gpu.so not built (needs /root/reference at build time)")
    want = {(n, v) for n, v in json.load(open(os.path.join(ROOT, "tests", "golden", "libnxz_abi_symbols.json")))["functions"]}
    out = subprocess.run(["readelf", "--dyn-syms", "-W", so], capture_output=True, text=True).stdout
    have = set()
    for line in out.splitlines():
        f = line.split()
        if len(f) >= 8 and f[3] == "FUNC" and f[6] != "UND" and f[4] in ("GLOBAL", "WEAK"):
            name, _, ver = f[7].partition("@@")
            have.add((name, ver))
    missing = want - have
    assert not missing, sorted(missing)[:10]
    extra = {n for n, _ in have - want}
    # the six boundary symbols are imported (UND), never re-exported; nothing else may appear
    assert not extra, sorted(extra)[:10]
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libnxgpu.so" in needed and "libnxz.so.1" in needed


def test_crc32_vpmsum_small_calls_stay_on_the_calling_thread(pg):
    """Below NXGPU_CRC_MIN_BYTES (64 KiB) the boundary's CRC is a table loop on the caller, as the reference keeps one below
    its own break-even (lib/nx_crc.c:247-255): no device is needed, so this runs on the build container too."""
    import ctypes as C
    import random
    import zlib
    lib = pg.load_library()
    rnd = random.Random(4)
    for n in (0, 16, 32, 48, 4096, 65520):
        data = rnd.randbytes(n)
        for seed in (0, 0xffffffff, 0x12345678):
            buf = C.create_string_buffer(data, max(n, 1))
            got = lib.__crc32_vpmsum(seed, C.addressof(buf), n)
            # raw update: no pre/post inversion (lib/crc32_ppc.c:33-67 inverts around the call)
            assert got == (zlib.crc32(data, seed ^ 0xffffffff) ^ 0xffffffff), (n, hex(seed))
