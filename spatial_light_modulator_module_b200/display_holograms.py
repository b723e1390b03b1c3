"""Display-side finalisation of the reference (``display_holograms.py:253-266``,
``move_traps.py:64-68,135-140``, ``show_hologram.py:7-13``): wavefront-correction mask add and
8-bit conversion on the device.  The tkinter window / REPL around them is out of scope."""
from __future__ import annotations

import os

import numpy as np
from PIL import Image as im

from . import _ffi
from .wavefront_correction import _util_engine


def mask_hologram(path, mask_arr, ct2pi):
    """reference: display_holograms.py:253-266 -> PIL "L" image."""
    base, ext = os.path.splitext(path)
    eng = _util_engine()
    mask = np.asarray(mask_arr, dtype=np.float64)
    if ext == ".npy":
        hologram_arr_2pi = np.load(path)
        if hologram_arr_2pi.shape != mask.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {hologram_arr_2pi.shape} {mask.shape}")
        out = eng.quantize(hologram_arr_2pi.astype(np.float64), mask, ct2pi, _ffi.QUANT_PIL_FLOAT)
    else:
        hologram_arr = np.array(im.open(path).convert("L"))
        if hologram_arr.shape != mask.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {hologram_arr.shape} {mask.shape}")
        out = eng.quantize_grey(hologram_arr, mask, ct2pi)
    return im.fromarray(eng.to_host(out))


def hologram_to_grey(hologram, mask=None, ct2pi=256):
    """((hologram [+ mask]) % 2pi * ct2pi / 2pi).astype(uint8) -- the array part of
    move_traps.display_hologram (move_traps.py:135-138) and show_hologram (show_hologram.py:9-11,
    whose default ct2pi is 255)."""
    eng = _util_engine()
    return eng.to_host(eng.quantize(np.asarray(hologram, dtype=np.float64),
                                    None if mask is None else np.asarray(mask, dtype=np.float64),
                                    ct2pi, _ffi.QUANT_FLOOR))


def preview_to_grey(expected):
    """PIL fromarray(float64).convert("L") (generate_hologram_sequence.py:29) as a uint8 array."""
    eng = _util_engine()
    return eng.to_host(eng.quantize(np.asarray(expected, dtype=np.float64), None, 256, _ffi.QUANT_PREVIEW))
