"""ctypes binding of libssfe.so (include/ssfe.h).  There is no CPU fallback: if the CUDA
library is missing or no GPU is present, everything here fails loudly."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libssfe.so")

SSFE_OK = 0
SSFE_ERR_INVALID = -1
SSFE_ERR_CUDA = -2
SSFE_ERR_TOO_SHORT = -3
SSFE_ERR_RANGE = -4
SSFE_ERR_NOMEM = -5
SSFE_ERR_GENDER = -6
F32, F64, I16 = 0, 1, 2

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p


class Config(ctypes.Structure):
    _fields_ = [("sample_rate", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("hop", ctypes.c_int32),
                ("n_mels", ctypes.c_int32),
                ("b", ctypes.c_double * 6), ("a", ctypes.c_double * 6), ("zi", ctypes.c_double * 5),
                ("mel_basis", c_f32p), ("min_level", ctypes.c_double), ("ref_db", ctypes.c_double),
                ("wav_scale", ctypes.c_double), ("dither_scale", ctypes.c_double),
                ("filtfilt_mode", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Batch(ctypes.Structure):
    _fields_ = [("n_utts", ctypes.c_int32), ("sample_offsets", c_i64p), ("f0_lo", c_f32p), ("f0_hi", c_f32p),
                ("spk_seed", c_u32p), ("dither_skip", c_u64p)]


class Outputs(ctypes.Structure):
    _fields_ = [("mel", vp), ("f0_norm", vp), ("f0_raw", vp), ("onehot", vp), ("bins", vp), ("wav", vp),
                ("wav64", vp)]


# every symbol include/ssfe.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ssfe_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(Config)]),
    "ssfe_destroy": (None, [vp]),
    "ssfe_last_error": (ctypes.c_char_p, [vp]),
    "ssfe_set_stream": (ctypes.c_int, [vp, vp]),
    "ssfe_synchronize": (ctypes.c_int, [vp]),
    "ssfe_version": (ctypes.c_char_p, []),
    "ssfe_launch_count": (ctypes.c_int64, [vp]),
    "ssfe_enable_timing": (ctypes.c_int, [vp, ctypes.c_int]),
    "ssfe_stage_ms": (ctypes.c_int, [vp, c_f32p, ctypes.c_int]),
    "ssfe_fixed_length": (ctypes.c_int64, [ctypes.c_int64]),
    "ssfe_num_frames": (ctypes.c_int64, [ctypes.c_int64]),
    "ssfe_plan_offsets": (ctypes.c_int, [c_i64p, ctypes.c_int, c_i64p, c_i64p]),
    "ssfe_filtfilt": (ctypes.c_int, [vp, vp, ctypes.c_int, c_i64p, ctypes.c_int, vp]),
    "ssfe_rand": (ctypes.c_int, [vp, c_u32p, c_u64p, c_i64p, ctypes.c_int, vp]),
    "ssfe_interp_lnr": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, vp]),
    "ssfe_filt_cascade": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, vp, vp, vp]),
    "ssfe_filt_cascade_taps": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "ssfe_filt_cascade_powers": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "ssfe_mt_charpoly_terms": (ctypes.c_int, [vp, ctypes.c_int]),
    "ssfe_mt_jump_poly": (ctypes.c_int, [ctypes.c_uint64, vp]),
    "ssfe_mt_jump_taps": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int]),
    "ssfe_stft_mag": (ctypes.c_int, [vp, vp, c_i64p, ctypes.c_int, vp]),
    "ssfe_stft_mel_db": (ctypes.c_int, [vp, vp, c_i64p, ctypes.c_int, vp]),
    "ssfe_rapt": (ctypes.c_int, [vp, vp, c_i64p, ctypes.c_int, c_f32p, c_f32p, vp]),
    "ssfe_rapt_dump": (ctypes.c_int64, [vp, ctypes.c_int64, vp, vp, vp, vp, vp, vp]),
    "ssfe_f0_normalize": (ctypes.c_int, [vp, vp, c_i64p, ctypes.c_int, vp, vp]),
    "ssfe_speaker_normalization": (ctypes.c_int, [vp, vp, ctypes.c_int, vp, ctypes.c_double, ctypes.c_double,
                                                  ctypes.c_int64, vp]),
    "ssfe_quantize_f0": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, vp, vp, ctypes.c_int]),
    "ssfe_extract": (ctypes.c_int, [vp, ctypes.POINTER(Batch), vp, ctypes.c_int, ctypes.POINTER(Outputs)]),
    "ssfe_extract_host": (ctypes.c_int, [vp, ctypes.POINTER(Batch), vp, ctypes.c_int, vp, vp, vp]),
    "ssfe_collate": (ctypes.c_int, [vp, vp, vp, c_i64p, ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int,
                                    vp, vp, vp, vp]),
}

_lib = None


def load():
    """Load libssfe.so.  Raises if the library has not been built (python -m speechsplit_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libssfe.so is missing (%s): build it with `python -m speechsplit_b200.build`; "
                               "there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
