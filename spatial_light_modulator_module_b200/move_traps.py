"""The array part of the reference's interactive trap steering (``move_traps.py``): the hologram of a
single trap and its 8-bit frame.  The keyboard thread and the tkinter window stay out of scope."""
from __future__ import annotations

import numpy as np

from .display_holograms import hologram_to_grey
from .wavefront_correction import _util_engine


def update_hologram(black_image, coords, which):
    """np.angle(ifft2(one-hot image)) for the trap ``coords[which]`` (reference: move_traps.py:64-68),
    evaluated in closed form on the device: 2*pi*(row*i/H + col*j/W) wrapped into (-pi, pi]."""
    h, w = np.asarray(black_image).shape
    row, col = int(coords[which][0]), int(coords[which][1])
    if not (0 <= row < h and 0 <= col < w):
        raise IndexError(f"index {row if not 0 <= row < h else col} is out of bounds")
    eng = _util_engine()
    return eng.to_host(eng.single_trap_phase(row, col, (h, w)))


def display_hologram_array(hologram, mask, mask_flag, ct2pi):
    """The uint8 frame move_traps.display_hologram shows (reference: move_traps.py:135-138)."""
    return hologram_to_grey(hologram, mask if mask_flag else None, ct2pi)
