"""B200-native hologram-synthesis engine: a drop-in for the Gerchberg-Saxton / gradient-descent
path of pranislav/Spatial_Light_Modulator_Module (src/algorithms.py and the drivers around it).

Sub-modules are named after the reference files they replace::

    from spatial_light_modulator_module_b200 import algorithms, generate_hologram
    hologram, expected, errors = algorithms.gerchberg_saxton(target_uint8, args)

All array work runs in hand-written sm_100a kernels (csrc/, built into lib/libslmholo.so by
``python -m spatial_light_modulator_module_b200.build``) behind the C ABI of include/slm_holo.h.
There is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["algorithms", "compare_error_evolution_algorithms", "constants", "display_holograms", "engine", "generate_hologram",
           "generate_hologram_sequence", "host_logic", "move_traps", "slab", "synthetic", "wavefront_correction"]
