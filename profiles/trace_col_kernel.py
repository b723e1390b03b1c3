"""Developer tool: device timeline of the warp-specialised column kernel (needs the -DSLM_TRACE build,
lib/libslmholo_trace.so).  Prints, per phase, the median duration over all CTAs and tiles."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SLM_HOLO_LIB"] = os.path.join(ROOT, "spatial_light_modulator_module_b200", "lib", "libslmholo_trace.so")
import torch
from spatial_light_modulator_module_b200 import _ffi, host_logic as hl, synthetic
from spatial_light_modulator_module_b200.engine import Engine

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 5      # 5 = GD_POST, 4 = STATS_KEEP, 0 = GS
alg = "gs" if mode == 0 else "gd"
shape, batch, loops = (1024, 1024), 32, 6
eng = Engine(shape, "fp32", batch)
lib = _ffi.load()
dev = torch.device("cuda", 0)
targets = torch.from_numpy(np.stack([synthetic.noise_target(shape, seed=i) for i in range(batch)])).to(dev)
x = torch.from_numpy(np.exp(2j * np.pi * np.random.default_rng(0).random((batch,) + shape)).astype(np.complex64)).to(dev)
during, _ = hl.learning_rate_schedule(0.005, 0, loops)
norms = np.full(batch, 255.0)
run = (lambda: eng.gd(targets, x, during, loops, want_expected=False, norms=norms)) if alg == "gd" else \
      (lambda: eng.gs(targets, loops, want_expected=False, norms=norms))
run(); run()
lib.slm_trace_arm.argtypes = [C.c_int]
lib.slm_trace_arm(mode)
run()
buf = np.zeros(148 * 64 * 16, dtype=np.uint64)
lib.slm_trace_read.argtypes = [C.c_void_p]
lib.slm_trace_read(buf.ctypes.data)
tr = buf.reshape(148, 64, 16).astype(np.int64)
valid = tr[:, :, 6] > 0
ntiles = valid.sum(axis=1)
print("tiles per CTA:", ntiles.min(), ntiles.max())
names = {(0, 1): "wait full", (1, 2): "load regs + forward FFT", (2, 3): "pointwise", (3, 4): "reduce/publish",
         (4, 5): "inverse FFT", (5, 6): "write tile + fence + arrive"}
t0 = tr[:, 0, 0].min()
span = (tr[:, :, 6].max() - t0) / 1e3
print(f"kernel span {span:.1f} us")
out = {}
for (a, b), nm in names.items():
    d = (tr[:, :, b] - tr[:, :, a])[valid]
    out[nm] = float(np.median(d)) / 1e3
    print(f"  compute warp 0: {nm:32s} median {np.median(d)/1e3:7.2f} us   p90 {np.percentile(d,90)/1e3:7.2f}")
per_tile = (tr[:, 1:, 0] - tr[:, :-1, 0])[valid[:, 1:]]
print(f"  tile period (start to start)        median {np.median(per_tile)/1e3:7.2f} us")
pv = tr[:, :, 12] > 0
for (a, b), nm in {(8, 9): "producer: wait store drained", (9, 10): "producer: TMA issue + grey copy", (11, 12): "producer: wait done",
                   (12, 13): "producer: issue store"}.items():
    m = (tr[:, :, a] > 0) & (tr[:, :, b] > 0)
    d = (tr[:, :, b] - tr[:, :, a])[m]
    if d.size:
        print(f"  {nm:36s} median {np.median(d)/1e3:7.2f} us   p90 {np.percentile(d,90)/1e3:7.2f}")
# how early does the tile land relative to when compute asks for it?  full-arrive stamp (10) of tile k vs compute stamp 0 of tile k
m = (tr[:, :, 10] > 0) & valid
lead = (tr[:, :, 0] - tr[:, :, 10])[m]
print(f"  grey copy done before compute asks   median {np.median(lead)/1e3:7.2f} us   p10 {np.percentile(lead,10)/1e3:7.2f}")
