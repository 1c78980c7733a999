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
