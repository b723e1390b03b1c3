"""Drop-in for the hologram-path functions of the reference's ``generate_hologram.py``:
target preparation (PIL, byte-exact with the reference), the algorithm call, analytic
deflect / lens holograms and the expected-outcome preview on the device."""
from __future__ import annotations

import os

import numpy as np
import PIL.ImageOps
from PIL import Image as im

from . import constants as c
from . import wavefront_correction as wfc
from .algorithms import gerchberg_saxton, gradient_descent
from .engine import get_engine


def main(args):
    """reference: generate_hologram.py:13-21."""
    if args.img_name is None:
        hologram = np.zeros((c.slm_height, c.slm_width))
    else:
        hologram = make_hologram(args)
    hologram = transform_hologram(hologram, args)
    if args.preview:
        show_expected_outcome(hologram, args)
    save_hologram(hologram, args)


def expected_outcome(hologram, norm=255, precision="fp64"):
    """Numeric part of show_expected_outcome (generate_hologram.py:24-29):
    |fft2(exp(1j*hologram))|^2 / max * norm."""
    hologram = np.asarray(hologram, dtype=np.float64)
    eng = get_engine(hologram.shape, precision)
    return eng.to_host(eng.expected_outcome(hologram, norm))


def show_expected_outcome(hologram, args):
    """reference: generate_hologram.py:24-34."""
    normed = expected_outcome(hologram, find_out_norm(args))
    im.fromarray(normed).resize((c.slm_height, c.slm_height)).show()


def find_out_norm(args):
    """reference: generate_hologram.py:37-42."""
    if args.img_name is None:
        return 255
    img = im.open(f"images/{args.img_name}").convert("L")
    return np.amax(np.array(img))


def pad_to_square(img):
    """Pads the image with black pixels to make it square (reference: generate_hologram.py:45-67)."""
    width, height = img.size
    if width == height:
        return img
    new_size = max(width, height)
    new_img = im.new("L", (new_size, new_size), 0)
    new_img.paste(img, ((new_size - width) // 2, (new_size - height) // 2))
    return new_img


def make_hologram(args):
    """reference: generate_hologram.py:70-79."""
    algorithm = gerchberg_saxton if args.algorithm == "gerchberg_saxton" else gradient_descent
    target = prepare_target(args.img_name, args)
    hologram, _, _ = algorithm(target, args)
    return hologram


def transform_hologram(hologram, args):
    """reference: generate_hologram.py:82-87."""
    if args.deflect is not None:
        hologram = deflect_hologram(hologram, args.deflect)
    if args.lens:
        hologram = add_lens(hologram, args.lens)
    return hologram


def prepare_target(img_name, args):
    """reference: generate_hologram.py:102-110 (host-side PIL pipeline, unchanged)."""
    target_img = im.open(f"images/{img_name}").convert("L")
    if args.invert:
        target_img = PIL.ImageOps.invert(target_img)
    target_img = pad_to_square(target_img)
    if args.quarterize:
        target_img = quarter(target_img)
    resized = target_img.resize((int(c.slm_width), int(c.slm_height)))
    return np.array(resized)


def save_hologram(hologram, args):
    """reference: generate_hologram.py:113-122 (the .npy part of save_hologram_and_gif)."""
    img_name = os.path.basename(args.img_name).split(".")[0] if args.img_name else "analytical"
    dest_dir = args.destination_directory
    if not os.path.exists(dest_dir):
        os.makedirs(dest_dir)
    name = wfc.originalize_name(f"{dest_dir}/{make_hologram_name(args, img_name)}.npy")
    np.save(name, hologram)
    return name


def make_hologram_name(args, img_name):
    """reference: generate_hologram.py:131-154 (including the runs of spaces its f-string embeds)."""
    alg_params = ""
    transforms = ""
    img_transforms = ""
    if args.deflect:
        transforms += f"_deflect_x{args.deflect[0]}_y{args.deflect[1]}"
    if args.lens:
        transforms += f"_lens{args.lens}"
    if args.algorithm == "gradient_descent":
        alg_params += f"_lr{args.learning_rate}_mr{args.white_attention}_unsettle{args.unsettle}"
    if args.quarterize:
        img_transforms += "_quarter"
    if args.invert:
        img_transforms += "_inverted"
    if args.img_name is None:
        return f"{img_name}{transforms}"
    pad = " " * 8
    return f"{img_name}{img_transforms}{pad}_{args.algorithm}{pad}{alg_params}{pad}_loops{args.max_loops}{pad}{transforms}"


def quarter(image):
    """reference: generate_hologram.py:166-175."""
    w, h = image.size
    resized = image.resize((w // 2, h // 2))
    ground = im.new("L", (w, h))
    ground.paste(resized)
    return ground


def deflect_hologram(hologram, angle: tuple):
    """(hologram + deflect_2pi(angle)) % 2pi on the device (reference: generate_hologram.py:178-182)."""
    hologram = np.asarray(hologram, dtype=np.float64)
    if hologram.shape[-2:] != (c.slm_height, c.slm_width):
        raise ValueError(f"operands could not be broadcast together with shapes {hologram.shape} "
                         f"({c.slm_height},{c.slm_width})")
    eng = wfc._util_engine()
    deflect = eng.deflect_phase(angle, c.px_distance, c.wavelength, c.u, (c.slm_height, c.slm_width))
    return eng.to_host(eng.add_mod2pi(hologram, deflect))


def add_lens(hologram, focal_len: float, uint8_quirk: bool = True):
    """(hologram + lens(f, shape)) % 2pi (reference: generate_hologram.py:185-186)."""
    hologram = np.asarray(hologram, dtype=np.float64)
    eng = wfc._util_engine()
    ln = eng.lens_phase(focal_len, c.px_distance, c.wavelength, hologram.shape[-2:], uint8_quirk)
    return eng.to_host(eng.add_mod2pi(hologram, ln))


def lens(focal_length, shape, uint8_quirk: bool = True):
    """Thin-lens phase (reference: generate_hologram.py:189-203).  With ``uint8_quirk`` (default)
    the result is the reference's uint8 array holding the truncated phase 0..6; pass False for the
    float64 phase the reference presumably intended."""
    eng = wfc._util_engine()
    out = eng.to_host(eng.lens_phase(focal_length, c.px_distance, c.wavelength, tuple(shape), uint8_quirk))
    return out.astype(np.uint8) if uint8_quirk else out
