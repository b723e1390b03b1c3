"""Developer tool: device timeline of the pipelined GD Fourier-plane pass (CGM_GD_PIPE) -- needs the -DSLM_TRACE build
(python -c "from spatial_light_modulator_module_b200 import build; build.build(defines=['-DSLM_TRACE'],
out='.../lib/libslmholo_trace.so', lengths=[1024])").  Prints the median duration of each phase of the two groups."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SLM_HOLO_LIB"] = os.path.join(ROOT, "spatial_light_modulator_module_b200", "lib", "libslmholo_trace.so")
import torch
from spatial_light_modulator_module_b200 import _ffi, host_logic as hl, synthetic
from spatial_light_modulator_module_b200.engine import Engine

shape, batch, loops = (1024, 1024), 32, 6
eng = Engine(shape, "fp32", batch)
lib = _ffi.load()
dev = torch.device("cuda", 0)
targets = torch.from_numpy(np.stack([synthetic.noise_target(shape, seed=i) for i in range(batch)])).to(dev)
x = torch.from_numpy(np.exp(2j * np.pi * np.random.default_rng(0).random((batch,) + shape)).astype(np.complex64)).to(dev)
during, _ = hl.learning_rate_schedule(0.005, 0, loops)
norms = np.full(batch, 255.0)
run = lambda: eng.gd(targets, x, during, loops, want_expected=False, norms=norms)
run(); run()
lib.slm_trace_arm.argtypes = [C.c_int]
lib.slm_trace_arm(7)
run()
buf = np.zeros(148 * 64 * 16, dtype=np.uint64)
lib.slm_trace_read.argtypes = [C.c_void_p]
lib.slm_trace_read(buf.ctypes.data)
tr = buf.reshape(148, 64, 16).astype(np.int64)
valid = (tr[:, :, 3] > 0) & (tr[:, :, 9] > 0)
print("tiles per CTA:", valid.sum(axis=1).min(), valid.sum(axis=1).max())
print(f"kernel span {(tr[:, :, 9].max() - tr[:, 0, 0].min()) / 1e3:.1f} us")
names = {(0, 1): "group 0: wait for the tile (full)", (1, 2): "group 0: load + forward transform", (2, 3): "group 0: max + write back + arrive",
         (4, 5): "group 1: wait for group 0 (fwd)", (5, 6): "group 1: load + wait for the plane's max", (6, 7): "group 1: gradient step + sums",
         (7, 8): "group 1: inverse transform", (8, 9): "group 1: write tile + fence + arrive"}
for (a, b), nm in names.items():
    d = (tr[:, :, b] - tr[:, :, a])[valid]
    print(f"  {nm:44s} median {np.median(d) / 1e3:6.2f} us   p90 {np.percentile(d, 90) / 1e3:6.2f}")
for ev, nm in ((0, "group 0"), (4, "group 1")):
    per = (tr[:, 1:, ev] - tr[:, :-1, ev])[valid[:, 1:] & valid[:, :-1]]
    print(f"  {nm}: tile period median {np.median(per) / 1e3:6.2f} us")

# one CTA's timeline, a few consecutive tiles (us from the kernel's start)
t0 = tr[:, 0, 0].min()
for cta in (5, 140):
    print(f"CTA {cta}: tile | g0: ask  got  fwd-done arrive | g1: ask  fwd  ready  grad  inv  done | seq: done-seen taken store-read(slot free)")
    for k in range(8, 16):
        r = (tr[cta, k] - t0) / 1e3
        print(f"   {k:3d} | " + " ".join(f"{r[i]:7.2f}" for i in (0, 1, 2, 3)) + " | " + " ".join(f"{r[i]:7.2f}" for i in (4, 5, 6, 7, 8, 9)) +
              " | " + " ".join(f"{r[i]:7.2f}" for i in (10, 11, 12)))
