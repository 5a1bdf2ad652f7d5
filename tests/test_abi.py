"""CPU tests of the drop-in boundary: libssfe.so builds for sm_100a, loads, and exports every
symbol include/ssfe.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from speechsplit_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ssfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssfe_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from speechsplit_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libssfe.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_geometry_helpers(lib):
    for L, T in ((48000, 188), (48128, 189), (32768, 129), (960000, 3751), (20011, 79)):
        assert lib.ssfe_num_frames(L) == T
        assert lib.ssfe_fixed_length(L) == L + (1 if L % 256 == 0 else 0)
    so = np.array([0, 48000, 48000 + 32768], np.int64)
    fix = np.zeros(3, np.int64)
    fr = np.zeros(3, np.int64)
    p = ctypes.POINTER(ctypes.c_int64)
    assert lib.ssfe_plan_offsets(so.ctypes.data_as(p), 2, fix.ctypes.data_as(p), fr.ctypes.data_as(p)) == 0
    assert list(fix) == [0, 48000, 48000 + 32769] and list(fr) == [0, 188, 188 + 129]


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speechsplit_b200 import FrontEnd
    with pytest.raises(RuntimeError):
        FrontEnd(0)
    # and straight through the C ABI: a negative status plus a message, not a crash
    from speechsplit_b200 import _lib
    from speechsplit_b200.melbasis import reference_mel_basis
    cfg = _lib.Config()
    cfg.sample_rate, cfg.n_fft, cfg.hop, cfg.n_mels = 16000, 1024, 256, 80
    mel = reference_mel_basis()
    cfg.mel_basis = mel.ctypes.data_as(_lib.c_f32p)
    h = ctypes.c_void_p()
    rc = lib.ssfe_create(ctypes.byref(h), 0, ctypes.byref(cfg))
    assert rc == _lib.SSFE_ERR_CUDA and b"no CPU fallback" in lib.ssfe_last_error(None)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no product file may import, include, load or run it."""
    pkg = os.path.join(ROOT, "speechsplit_b200")
    bad_py = re.compile(r"^\s*(from|import)\s+oracle\b|importlib.*oracle|oracle[/.]_build|librapt_ref", re.M)
    bad_c = re.compile(r"#\s*include[^\n]*oracle|dlopen[^\n]*oracle|librapt_ref")
    for dp, _, files in os.walk(pkg):
        for f in files:
            txt = None
            if f.endswith(".py"):
                txt, pat = open(os.path.join(dp, f)).read(), bad_py
            elif f.endswith((".cu", ".cuh", ".cpp", ".h")):
                txt, pat = open(os.path.join(dp, f)).read(), bad_c
            if txt is not None:
                assert not pat.search(txt), "%s reaches into the oracle" % f


def test_sass_has_tma_bulk_copy(lib):
    """The fused STFT kernel stages frames with TMA bulk copies (SASS UBLKCP) on sm_100a."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    from speechsplit_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "arch = sm_100a" in sass
    body = sass[sass.index("stft_mel_kernel"):]
    assert "UBLKCP" in body and "SYNCS.ARRIVE.TRANS64" in body


def test_fft_core_host_emulation(tmp_path):
    """The warp-level 1024-point FFT of the fused STFT kernel (csrc/fft_core.cuh: radix-32 DIF, twiddle,
    transpose, radix-32 DIF, two real frames per complex transform, spectrum split) compiled for the host
    with the lanes executed one after another, against a direct O(N^2) DFT in double: the index math and
    the fp32 error (<= 3e-6 at |X| ~ 16) are checked without a GPU."""
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "fft_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "speechsplit_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "fft_emu.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    err = float(r.stdout.split("=")[1].split()[0])
    assert err <= 3e-6, r.stdout
