"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports exactly what include/pi3b200.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pi3b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pi3b_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from pi3_slam_b200.build import build_lib
    path = build_lib()
    lib = ctypes.CDLL(str(path))
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pi3b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (pi3b_[a-z0-9_]+)", out)))
    assert exported == names, "exported pi3b_* symbols and the header disagree"


def test_ctypes_signature_table_matches_header():
    from pi3_slam_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} params, ctypes table {len(argtypes)}"


def test_version_and_error_string_without_gpu():
    from pi3_slam_b200 import _lib
    lib = _lib.load()
    assert lib.pi3b_version() == 100
    # argument validation happens before any CUDA call, so it works on a CPU-only box
    rc = lib.pi3b_gemm(None, 0, None, 0, None, 0, 1, 1, 1, None, None, None, 0, None, 0, 0, 0, 0, 0, 0)
    assert rc == -1 and b"null pointer" in lib.pi3b_last_error()
    # the query pre-scale of the fused QKV epilogue must be positive (1 = none)
    fake = ctypes.c_void_p(0x1000)
    rc = lib.pi3b_gemm_qkv(None, fake, 1024, fake, 1024, 128, 16, 1024, None, fake, 3072, fake, 1e-5, None, 0, 1, 0, 1, 0.0)
    assert rc == -1 and b"q_scale" in lib.pi3b_last_error()


def test_kernels_are_blackwell_native():
    """SASS evidence: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    from pi3_slam_b200.build import LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass and "STTM" in sass
    assert "HMMA.16816" not in sass      # no legacy mma.sync tensor path
    # CTA pairs (tcgen05.mma.cta_group::2 with 2-CTA TMA loads and multicast commits) and packed fp32x2 arithmetic
    assert "UTCHMMA.2CTA" in sass and "UTMALDG.2D.2CTA" in sass and "UTCBAR.2CTA.MULTICAST" in sass
    assert "FFMA2" in sass and "FADD2" in sass
    arch = subprocess.run(["cuobjdump", "-lelf", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in arch


def test_missing_library_fails_loudly():
    """No silent fallback: if libpi3b200.so cannot be found the package raises instead of computing on the CPU."""
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['PI3B_LIB'] = '/nonexistent/libpi3b200.so';"
            "from pi3_slam_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.Pi3bError as e:\n    print('RAISED', e)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in out.stdout and "no CPU or PyTorch fallback" in out.stdout, out.stdout + out.stderr
