"""ad-hoc: k_chain_direct6 (stereo generation 6, mono DUAL) under the emulator against the oracle, exact mode bit for bit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
from oracle.hdrref import CpuStretch
from signalsmith_stretch_b200 import BatchStretch
d = os.path.join(ROOT, "tests", "cuda_emu", "_build")
EX = os.path.join(d, "libb200stretch_emu_exactfft.so")

def orc(cfg, x, ratio, chunk):
    outs = []
    for s in range(x.shape[0]):
        o = CpuStretch("orc"); cfg(o)
        outs.append(signals.run_single(o, x[s], ratio, chunk))
    return np.stack(outs)

def run(name, cfg, S, C, ratio, n, chunk, gen=None, exact=1, same=False):
    x = signals.batch("harmonic", S, C, n, 48000)
    if same: x[1::2] = x[0::2][: len(x[1::2])] * 0.7
    e = BatchStretch(S, lib_path=EX)
    e.set_tuning(3, exact)
    if gen: e.set_tuning(0, gen)
    cfg(e)
    t = time.time()
    y = signals.run_batch(e, x, ratio, chunk)
    ref = orc(cfg, x, ratio, chunk)
    d = np.abs(y - ref).max()
    print("%-40s exact=%d gen=%s maxdiff %.3g  equal=%s  (%.1fs)" % (name, exact, gen, d, np.array_equal(y, ref), time.time() - t), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "stereo"):
    run("stereo small 0.8x", lambda o: o.configure(2, 512, 128), 2, 2, 0.8, 4000, 640, gen=6)
    run("stereo small 0.8x 2 rounds", lambda o: o.configure(2, 512, 128), 2, 2, 0.8, 14000, 9000, gen=6)
    run("stereo small 0.55x (far)", lambda o: o.configure(2, 512, 128), 2, 2, 2.4, 3000, 900, gen=6)
    run("stereo L3 cheaper 1.5x", lambda o: o.presetCheaper(2, 48000.0), 2, 2, 1.5, 16000, 11520, gen=6)
    run("stereo L8", lambda o: o.configure(2, 512, 64), 2, 2, 1.25, 4000, 640, gen=6)
    run("stereo L1", lambda o: o.configure(2, 256, 256), 2, 2, 0.8, 6000, 1024, gen=6)
    run("stereo default 0.8x 32 blocks", lambda o: o.presetDefault(2, 48000.0), 2, 2, 0.8, 2 * 57600, 46080, gen=6)
if which in ("all", "mono"):
    run("mono small 1.25x (pairs)", lambda o: o.configure(1, 384, 96), 4, 1, 1.25, 9000, 8000)
    run("mono small 0.8x 3 streams", lambda o: o.configure(1, 512, 128), 3, 1, 0.8, 4000, 640)
    run("mono default 1.25x odd chunks", lambda o: o.presetDefault(1, 48000.0), 2, 1, 1.25, 12000, 4999)
    run("mono cheaper 0.5x", lambda o: o.presetCheaper(1, 48000.0), 2, 1, 0.5, 30000, 7000)
if which in ("all", "fast"):
    run("stereo default 0.8x fast", lambda o: o.presetDefault(2, 48000.0), 2, 2, 0.8, 57600, 46080, gen=6, exact=0)
    run("stereo default 0.8x fast gen4", lambda o: o.presetDefault(2, 48000.0), 2, 2, 0.8, 57600, 46080, gen=4, exact=0)
    run("mono default 0.8x fast", lambda o: o.presetDefault(1, 48000.0), 2, 1, 0.8, 57600, 46080, exact=0)
if which in ("all", "monoL"):
    for (B, H) in ((256, 256), (512, 256), (384, 128), (500, 100), (512, 64)):
        run("mono %d/%d L=%d 0.8x" % (B, H, round(max(B, 1) * 1.0 / H)), (lambda B, H: (lambda o: o.configure(1, B, H)))(B, H), 4, 1, 0.8, 5000, 1500)
        run("mono %d/%d 1.6x fast" % (B, H), (lambda B, H: (lambda o: o.configure(1, B, H)))(B, H), 2, 1, 1.6, 3000, 1500, exact=0)
