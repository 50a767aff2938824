"""Small runs of k_chain_direct6 for compute-sanitizer: stereo (generation 6) and mono stream pairs (tuning key 5), checked against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
from oracle.hdrref import CpuStretch
from signalsmith_stretch_b200 import BatchStretch

for C, S, ratio in ((2, 3, 0.8), (1, 5, 0.8)):
    e = BatchStretch(S)
    e.presetDefault(C, 48000.0)
    if C == 2:
        e.set_tuning(0, 6)
    else:
        e.set_tuning(5, 1)
    n_out = 40 * e.intervalSamples()  # more than 32 blocks: a second group of lanes
    x = signals.batch("harmonic", S, C, int(round(n_out / ratio)), 48000)
    if C == 1:
        x[3, :, 20000:40000] = 0.0  # a pair that stops sharing its schedule
    y = signals.run_batch(e, x, ratio, n_out)
    ref = []
    for s in range(S):
        o = CpuStretch("orc"); o.presetDefault(C, 48000.0)
        ref.append(signals.run_single(o, x[s], ratio, n_out))
    err = float(np.sqrt(np.mean((y - np.stack(ref)) ** 2)))
    print("C=%d S=%d: %d launches, rms vs oracle %.3e" % (C, S, e.kernel_launches(), err))
    assert err <= 1e-3
