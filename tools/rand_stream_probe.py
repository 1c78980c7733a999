import ctypes as C, os, sys, time, zlib, random
sys.path.insert(0, "/root/repo/tests")
ROOT="/root/repo"
import importlib.util
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
from nxjob import Job
lib = pg.load_library()
class Dev(C.Structure):
    _fields_ = [("i", C.c_int * 8), ("paste_addr", C.c_void_p), ("fd", C.c_int), ("function", C.c_int), ("pad", C.c_char * 256)]
dev = Dev(); assert lib.nx_function_begin(2, -1, C.byref(dev)) == 0
data = random.Random(1).randbytes(1 << 20)
jc = Job(0x00, [data], 2 << 20)          # FHT compress
assert lib.nxu_run_job(jc.addr, C.byref(dev)) == 0
comp = jc.out(); print("compressed", len(comp), "cc", jc.cc())
for mode in ("0", "262144", "0", "262144"):
    os.environ["NXGPU_INFLATE_PAR_MIN"] = mode
    j = Job(0x10, [comp], 2 << 20)
    t0 = time.perf_counter(); assert lib.nxu_run_job(j.addr, C.byref(dev)) == 0; dt = time.perf_counter() - t0
    print("par_min", mode, "ms %.1f" % (dt * 1e3), "cc", j.cc(), "tpbc", j.tpbc(), "ok", j.out() == data)
