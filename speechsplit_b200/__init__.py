"""speechsplit_b200 - B200-native (sm_100a) drop-in for SpeechSplit's feature front end
(make_spect_f0.py + utils.quantize_f0_numpy).  See DESIGN.md / INTEGRATION.md."""
from .frontend import FrontEnd, FrontEndConfig, GENDER_RANGE, SsfeError, butter_highpass, default_frontend  # noqa: F401

__all__ = ["FrontEnd", "FrontEndConfig", "GENDER_RANGE", "SsfeError", "butter_highpass", "default_frontend"]
