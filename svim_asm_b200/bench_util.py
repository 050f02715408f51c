"""Pinned host buffers for the end-to-end legs of bench.py (torch is only the allocator here)."""
import numpy as np


def pin(array):
    """(copy of `array` in pinned host memory as a numpy view, the torch tensor that owns it)."""
    import torch
    flat = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
    t = torch.empty(max(flat.shape[0], 1), dtype=torch.uint8, pin_memory=True)
    view = t.numpy()[:flat.shape[0]]
    view[:] = flat
    return view.view(array.dtype).reshape(array.shape), t


def pinned_host(host):
    """Move the flat arrays of a HostBatch into pinned memory (in place)."""
    keep = []
    for name in ("hdr", "cigar", "seg", "sa_count", "seq4", "seq_off", "contig_lengths"):
        arr, t = pin(getattr(host, name))
        setattr(host, name, arr)
        keep.append(t)
    host._pinned = keep
    return host
