"""Import the UNMODIFIED reference modules from /root/reference/src (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so this
loader is used solely by ``oracle/make_golden.py`` (fixture generation) and by the
CPU tests that are skipped when the directory is absent.

The reference imports plotting / GUI / camera packages that are not installed
(matplotlib, imageio, tkinter, screeninfo, skimage, pylablib, keyboard, cv2 is present).
They are irrelevant to the numeric path, so inert stand-ins are registered in
``sys.modules`` before the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_SRC = "/root/reference/src"

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_agg",
    "matplotlib.ticker", "matplotlib.colors", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "tkinter", "PIL.ImageTk", "screeninfo", "skimage", "skimage.restoration",
    "skimage.restoration.inpaint", "skimage.transform", "pylablib", "pylablib.devices",
    "pylablib.devices.uc480", "imageio", "keyboard",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "algorithms.py"))


def _install_stubs() -> None:
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=f"stub:{name}")


def load(name: str) -> types.ModuleType:
    """Return reference module ``name`` (e.g. 'algorithms', 'generate_hologram')."""
    if not available():
        raise RuntimeError("reference sources are not present on this machine")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    return importlib.import_module(name)
