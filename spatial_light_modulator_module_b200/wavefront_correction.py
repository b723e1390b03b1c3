"""The two functions of the reference's ``wavefront_correction.py`` that sit on the hologram
path (the camera-bound calibration itself is out of scope)."""
from __future__ import annotations

import os

import numpy as np

from . import _ffi, constants as c
from .engine import get_engine

_UTIL_SHAPE = (64, 64)      # elementwise kernels do not depend on the context's plane shape


def _util_engine():
    return get_engine(_UTIL_SHAPE, "fp64")


def deflect_2pi(angle: tuple) -> np.ndarray:
    """Return hologram for deflecting light (reference: wavefront_correction.py:440-449)."""
    eng = _util_engine()
    return eng.to_host(eng.deflect_phase(angle, c.px_distance, c.wavelength, c.u, (c.slm_height, c.slm_width)))


def convert_2pi_hologram_to_int_hologram(hologram: np.ndarray, ct2pi: int) -> np.ndarray:
    """np.round(hologram*ct2pi/2pi).astype(uint8) (reference: wavefront_correction.py:458-459)."""
    eng = _util_engine()
    return eng.to_host(eng.quantize(np.asarray(hologram, dtype=np.float64), None, ct2pi, _ffi.QUANT_ROUND_WRAP))


def convert_2pi_holograms_to_int_holograms(sample: list, ct2pi: int) -> list:
    """reference: wavefront_correction.py:452-455."""
    return [convert_2pi_hologram_to_int_hologram(hologram, ct2pi) for hologram in sample]


def originalize_name(name: str) -> str:
    """reference: wavefront_correction.py:325-337."""
    if not os.path.exists(name):
        return name
    base, ext = os.path.splitext(name)
    i = 1
    while True:
        new_name = f"{base}_{i}{ext}"
        if not os.path.exists(new_name):
            return new_name
        i += 1
