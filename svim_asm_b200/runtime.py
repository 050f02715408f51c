"""Process-wide GPU engines used by the reference-shaped seams (one context per device and role, created on first use)."""
import os

_ENGINES = {}


def get_engine(device=None, role="main"):
    """role "loader": the context (stream, pinned staging slots) of the background FASTA load, kept for the life of the
    process like the main one (setting up 64 MB of pinned memory again for every run would be paid inside the ingest
    that runs next to it)."""
    from .engine import Engine
    if device is None:
        device = int(os.environ.get("SVIM_ASM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if (device, role) not in _ENGINES:
        _ENGINES[(device, role)] = Engine(device)
    return _ENGINES[(device, role)]
