"""Process-wide GPU engine used by the reference-shaped seams (one context per device, created on first use)."""
import os

_ENGINES = {}


def get_engine(device=None):
    from .engine import Engine
    if device is None:
        device = int(os.environ.get("SVIM_ASM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]
