"""The reference's OWN test programs (test/Makefile.am:9-76) as the acceptance gate.

oracle/Makefile compiles /root/reference/test/*.c unmodified into oracle/_ref/reftests/ (build products,
they travel to the GPU box).  They are linked against soname libnxz.so.1; here LD_LIBRARY_PATH points that
name at

  * oracle/_ref/libnxz_ref.so         the reference's host code over the CPU engine oracle/nxemu.c   (not gpu)
  * power-gzip_b200/libnxz_gpu.so     the same host code over the B200 engine (the product)           (gpu)

both in NX mode (NX_GZIP_TYPE_SELECTOR=2, the ".nx" rows of the reference's selector matrix,
test/gen_test.sh:6-9) so that every deflate()/inflate()/compress()/uncompress() goes through nxu_run_job.
Exit status 0 = pass, 77 = skip (test/test.h:18).

Not run: test_reset / test_reset2 (they assert the AUTO-mode `switchable` flag, which needs a POWER device
tree: lib/nx_zlib.c:1222-1226 returns before the stream map exists when no NX engine is enumerated),
test_pid_reuse and serial-tests/ (PowerVM VAS-window kernel behaviour), test_abi (needs abidiff; the symbol
versions are checked in tests/test_abi.py instead).
"""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "reftests")
CPU_LIB = os.path.join(ROOT, "oracle", "_ref", "libnxz_ref.so")
GPU_LIB = os.path.join(ROOT, "power-gzip_b200", "libnxz_gpu.so")

TESTS = ["test_crc32", "test_adler32", "test_buf_error", "test_resetKeep", "test_inflatesyncpoint", "test_zeroinput",
         "test_gz", "test_dict", "test_stress", "test_deflate", "test_inflate", "test_multithread_stress"]

needs_bins = pytest.mark.skipif(not os.path.isdir(BIN), reason="oracle/_ref/reftests not built (needs /root/reference at build time)")


def _env(lib, workdir):
    link = os.path.join(workdir, "libnxz.so.1")
    if not os.path.exists(link):
        os.symlink(lib, link)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = workdir + os.pathsep + os.path.dirname(lib) + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env["NX_GZIP_TYPE_SELECTOR"] = "2"
    env["NX_GZIP_LOGFILE"] = os.path.join(workdir, "nx.log")
    return env


def _report(name, rc, out):
    tail = "\n".join(out.decode(errors="replace").splitlines()[-15:])
    if rc == 77:
        pytest.skip(f"{name} skipped itself")
    assert rc == 0, f"{name} exit {rc}\n{tail}"


@pytest.fixture(scope="module")
def cpu_runs():
    """All programs at once over the CPU engine (they are independent processes): the slowest one bounds the wall time."""
    if not os.path.isdir(BIN):
        yield {}
        return
    with tempfile.TemporaryDirectory() as wd:
        env = _env(CPU_LIB, wd)
        # test_multithread_stress <threads> <seconds per iteration> <iterations>: the checker is a scalar C engine on a few cores
        procs = {t: subprocess.Popen([os.path.join(BIN, t)] + (["4", "2", "1"] if t == "test_multithread_stress" else []),
                                     cwd=wd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                 for t in TESTS}
        yield procs
        for p in procs.values():
            if p.poll() is None:
                p.kill()


@needs_bins
@pytest.mark.parametrize("name", TESTS)
def test_reference_program_over_cpu_engine(name, cpu_runs):
    p = cpu_runs[name]
    out, _ = p.communicate(timeout=900)
    _report(name, p.returncode, out)


@needs_bins
@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPU_LIB), reason="power-gzip_b200/libnxz_gpu.so not built")
@pytest.mark.parametrize("name", TESTS)
def test_reference_program_over_gpu_engine(name):
    with tempfile.TemporaryDirectory() as wd:
        env = _env(GPU_LIB, wd)
        args = ["64", "3", "2"] if name == "test_multithread_stress" else []     # 64 threads, 2 x 3 s (the default is 20 threads per core for 60 s)
        p = subprocess.run([os.path.join(BIN, name)] + args, cwd=wd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
        # the product really ran: the drop-in library maps libnxgpu.so, and nothing of the oracle
        _report(name, p.returncode, p.stdout)


@needs_bins
@pytest.mark.gpu
def test_gpu_drop_in_links_no_oracle():
    out = subprocess.run(["ldd", GPU_LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "libnxgpu.so" in out and "oracle" not in out and "libnxz_ref" not in out, out
