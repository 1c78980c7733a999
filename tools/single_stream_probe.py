#!/usr/bin/env python3
"""Developer probe: ONE stream on an otherwise idle GPU (the LD_PRELOAD single-z_stream case): inflate of one
zlib member on one warp, deflate of one 1 MiB job on one CTA."""
import gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
data = pg.makedata(1, 24, alice)
eng = pg.Engine(0)
blob = zlib.compress(data, 6)
dc = eng.alloc(len(blob)); dc.upload(blob)
do = eng.alloc(len(data))
for it in range(2):
    eng.kernel_time_reset()
    r = eng.inflate_batch([pg.InflateItem(dc.ptr, len(blob), do.ptr, len(data), pg.WRAP_ZLIB, 0)], mem=pg.MEM_DEVICE)[0]
    kms, _ = eng.kernel_time("inflate")
print(f"one 16 MiB zlib member on one warp: {kms:.1f} ms = {len(data)/kms/1e3:.1f} MB/s, rc {r.rc}")
one = data[: 1 << 20]
ds = eng.alloc(len(one)); ds.upload(one)
dd = eng.alloc(2 << 20)
for level in (1, 6):
    for it in range(2):
        eng.kernel_time_reset()
        res = eng.deflate_stream_device(ds.ptr, len(one), dd.ptr, 2 << 20, level=level, wrap=pg.WRAP_RAW, chunk=1 << 20)
        kms, _ = eng.kernel_time("deflate")
    print(f"one 1 MiB deflate job on one CTA, level {level}: {kms:.2f} ms = {len(one)/kms/1e3:.1f} MB/s")
# one compress descriptor (nxu_run_job, host buffers, wall clock): cut into 64 KiB pieces above 128 KiB
import ctypes as C
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nxjob import Job
lib = pg.load_library()


class Dev(C.Structure):
    _fields_ = [("i", C.c_int * 8), ("paste_addr", C.c_void_p), ("fd", C.c_int), ("function", C.c_int), ("pad", C.c_char * 256)]


dev = Dev()
assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0
for size in (65536, 131072, 1 << 20):
    best = 1e9
    for it in range(5):
        j = Job(0x00, [data[:size]], 2 * size)
        t0 = time.perf_counter()
        assert lib.nxu_run_job(j.addr, C.byref(dev)) == 0
        best = min(best, time.perf_counter() - t0)
    print(f"one FHT compress descriptor of {size >> 10} KiB through nxu_run_job (host to host): {best * 1e3:.2f} ms = {size / best / 1e6:.0f} MB/s, cc {j.cc()}, {j.tpbc()} bytes out")
