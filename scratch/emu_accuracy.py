"""Accuracy of the emulated kernels (same float arithmetic as the GPU build, FMA included) against the golden
vectors of the reference header: RMS over the first 8 blocks after the latency (the test's 1e-4 gate) and overall."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
from signalsmith_stretch_b200 import BatchStretch
lib = os.path.join(ROOT, "tests/cuda_emu/_build/libb200stretch_emu.so")
names = sys.argv[1:] or ["config4_formant", "config2_stereo_0p8x"]
for name in names:
    g = np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))
    cfg, C, sr, ratio, _ = signals.CONFIGS[name]
    e = BatchStretch(1, lib_path=lib); cfg(e)
    t0 = time.time()
    nblk = int(os.environ.get("NBLK", "14"))
    H = e.intervalSamples(); lat = e.outputLatency() + int(e.inputLatency() * ratio)
    n_out = min(g["hdr"].shape[-1], lat + nblk * H)
    n_in = int(round(n_out / float(g["ratio"])))
    x = g["x"][None, :, :n_in]
    y = signals.run_batch(e, x, float(g["ratio"]), int(g["chunk"]))
    d = y[0] - g["hdr"][:, :y.shape[-1]]
    r = lambda a: float(np.sqrt(np.mean(a.astype(np.float64) ** 2)))
    print(name, "rms first 8 blocks %.3e" % r(d[:, :lat + 8 * H]), "rms all(%d) %.3e" % (y.shape[-1], r(d)), "%.0fs" % (time.time() - t0), flush=True)
