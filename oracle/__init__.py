"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the GS/GD hologram path of pranislav/Spatial_Light_Modulator_Module.
Nothing in the shipped package may import from here: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker (or as the timed CPU baseline), never as
part of the product path.

Parity pinning: the reference ships no tests or golden vectors,
so ``numpy_port`` is pinned against *outputs of the reference itself*, run in the
build container through ``reference_loader`` and committed as ``tests/golden/*.npz``
by ``make_golden.py`` (numpy/scipy/Pillow versions are recorded in every fixture).
"""
