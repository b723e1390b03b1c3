"""Parameters of the optical setup -- values of the reference's src/constants.py:5-10."""

slm_width = 1024  # pixels
slm_height = 768  # pixels
wavelength = 5.32e-7  # metres
px_distance = 3.6e-5  # SLM pixel pitch, metres
first_diff_max = wavelength / px_distance
u = first_diff_max / 4  # deflection unit: quarter of the first diffraction maximum
