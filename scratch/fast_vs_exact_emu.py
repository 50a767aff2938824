"""Does the fast chain arithmetic (fused multiply-adds, reciprocal / rsqrt) cost parity beyond what the FFT rounding
already costs?  Emulated kernels (same float arithmetic as the GPU build) against the golden vector of the reference
header, stereo 0.8x: RMS error per 2-block window, exact chain vs fast chain, both with the float FFT."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
from signalsmith_stretch_b200 import BatchStretch
lib = os.path.join(ROOT, "tests/cuda_emu/_build/libb200stretch_emu.so")
name = "config2_stereo_0p8x"
g = np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))
cfg, C, sr, ratio, _ = signals.CONFIGS[name]
r = lambda a: float(np.sqrt(np.mean(a.astype(np.float64) ** 2)))
res = {}
for exact in (1, 0):
    e = BatchStretch(1, lib_path=lib); cfg(e); e.set_tuning(3, exact)
    H = e.intervalSamples(); lat = e.outputLatency() + int(e.inputLatency() * ratio)
    n_out = g["hdr"].shape[-1]
    x = g["x"][None]
    y = signals.run_batch(e, x, float(g["ratio"]), int(g["chunk"]))
    d = y[0] - g["hdr"][:, :y.shape[-1]]
    res[exact] = [r(d[:, lat + k * H: lat + (k + 2) * H]) for k in range(0, 14, 2)]
    print("exact" if exact else "fast ", " ".join("%.2e" % v for v in res[exact]), " | first 8 blocks %.3e, all %.3e" % (r(d[:, :lat + 8 * H]), r(d)), flush=True)
