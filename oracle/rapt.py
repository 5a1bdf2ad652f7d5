"""ctypes binding of oracle/rapt_ref.c (TEST INFRASTRUCTURE; PARITY UNPINNED, see that file).

``rapt(x, fs, hopsize, min, max, voice_bias=0.0, otype=2)`` mirrors the call at
/root/reference/make_spect_f0.py:64 (``pysptk.sptk.rapt``): float32 in, float32
``ceil(len(x)/hopsize)`` out, unvoiced = -1e10 for otype 2.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librapt_ref.so")
_lib = None
CMAX = 20


class _Debug(ctypes.Structure):
    _fields_ = [("max_frames", ctypes.c_int),
                ("ncands", ctypes.c_void_p), ("locs", ctypes.c_void_p), ("pvals", ctypes.c_void_p),
                ("mpvals", ctypes.c_void_p), ("stat", ctypes.c_void_p), ("rms_ratio", ctypes.c_void_p),
                ("f0cand", ctypes.c_void_p), ("prept", ctypes.c_void_p), ("dpvals", ctypes.c_void_p),
                ("ds", ctypes.c_void_p), ("ds_cap", ctypes.c_int),
                ("n_frames", ctypes.c_int), ("n_forced", ctypes.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "rapt_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.rapt_ref.restype = ctypes.c_int
        _lib.rapt_ref.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return _lib


def n_frames_out(length, hopsize):
    return -(-int(length) // int(hopsize))


def rapt(x, fs=16000, hopsize=256, min=60, max=240, voice_bias=0.0, otype=2, debug=False):
    lib = _load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 1
    n_out = n_frames_out(x.shape[0], hopsize)
    out = np.empty(n_out, dtype=np.float32)
    dbg = None
    arrays = {}
    if debug:
        mf = n_out + 8
        arrays = dict(ncands=np.zeros(mf, np.int32), locs=np.zeros((mf, CMAX), np.int32),
                      pvals=np.zeros((mf, CMAX), np.float32), mpvals=np.zeros((mf, CMAX), np.float32),
                      stat=np.zeros(mf, np.float32), rms_ratio=np.zeros(mf, np.float32),
                      f0cand=np.zeros((mf, CMAX), np.float32), prept=np.zeros((mf, CMAX), np.int32),
                      dpvals=np.zeros((mf, CMAX), np.float32),
                      ds=np.zeros(x.shape[0] // 8 + 512, np.float32))
        dbg = _Debug()
        dbg.max_frames = mf
        for k, v in arrays.items():
            setattr(dbg, k, v.ctypes.data)
        dbg.ds_cap = arrays["ds"].shape[0]
    rc = lib.rapt_ref(x.ctypes.data, x.shape[0], float(fs), int(hopsize), float(min), float(max),
                      float(voice_bias), int(otype), out.ctypes.data, n_out,
                      ctypes.byref(dbg) if dbg is not None else None)
    if rc == 2:
        raise ValueError("input range too small for analysis by get_f0")
    if rc != 0:
        raise RuntimeError("rapt_ref failed (%d)" % rc)
    if debug:
        arrays["n_frames"] = dbg.n_frames
        arrays["n_forced"] = dbg.n_forced
        return out, arrays
    return out
