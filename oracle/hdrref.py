"""ctypes driver shared by the two CPU checkers that expose the same C surface:

  * oracle/_ref/libhdr_stretch.so   (prefix hdr_)  -- the UNMODIFIED reference header compiled
    against the oracle's stand-in STFT (oracle/ref_header_shim.cpp), and
  * oracle/_build/liboracle_stretch.so (prefix orc_) -- the oracle restatement (stretch_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

STATE = {"input": 0, "prevInput": 1, "output": 2, "inputEnergy": 3, "predEnergy": 4, "outputMap": 5,
         "energy": 6, "smoothedEnergy": 7, "window": 8, "windowProducts": 9, "predInput": 10,
         "formantMetric": 11}


def lib_path(kind):
    if kind == "hdr":
        return os.path.join(_HERE, "_ref", "libhdr_stretch.so")
    return os.path.join(_HERE, "_build", "liboracle_stretch.so")


def available(kind):
    return os.path.exists(lib_path(kind))


def _lib(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    L = ctypes.CDLL(lib_path(kind))
    p = kind + "_" if kind == "hdr" else "orc_"
    vp, ci, cf, cd, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_long
    fp = ctypes.POINTER(ctypes.c_float)
    sigs = {
        "new": (vp, [cl]), "free": (None, [vp]),
        "preset_default": (None, [vp, ci, cf, ci]), "preset_cheaper": (None, [vp, ci, cf, ci]),
        "configure": (None, [vp, ci, ci, ci, ci]), "reset": (None, [vp]),
        "block_samples": (ci, [vp]), "interval_samples": (ci, [vp]), "input_latency": (ci, [vp]),
        "output_latency": (ci, [vp]), "split_computation": (ci, [vp]), "seek_length": (ci, [vp]),
        "output_seek_length": (ci, [vp, cf]), "bands": (ci, [vp]), "fft_samples": (ci, [vp]),
        "set_transpose_factor": (None, [vp, cf, cf]), "set_transpose_semitones": (None, [vp, cf, cf]),
        "set_formant_factor": (None, [vp, cf, ci]), "set_formant_semitones": (None, [vp, cf, ci]),
        "set_formant_base": (None, [vp, cf]), "set_freq_map_quadratic": (None, [vp, cf, cf]), "set_freq_map_table": (None, [vp, fp, fp, ci]),
        "seek": (None, [vp, fp, ci, cd]), "output_seek": (None, [vp, fp, ci]),
        "process": (None, [vp, fp, ci, fp, ci]), "flush": (None, [vp, fp, ci, cf]),
        "exact": (ci, [vp, fp, ci, fp, ci]),
        "get_state": (ci, [vp, ci, fp]), "num_peaks": (ci, [vp]), "get_peaks": (ci, [vp, fp]),
    }
    fns = {}
    for k, (res, args) in sigs.items():
        fn = getattr(L, p + k)
        fn.restype = res
        fn.argtypes = args
        fns[k] = fn
    _LIBS[kind] = fns
    return fns


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class CpuStretch:
    """Same method names as signalsmith::stretch::SignalsmithStretch (signalsmith-stretch.h:34-491)."""

    def __init__(self, kind="hdr", seed=1):
        self.f = _lib(kind)
        self.h = self.f["new"](seed)
        self.channels = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.f["free"](self.h)
            self.h = None

    def presetDefault(self, ch, sr, split=False):
        self.channels = ch
        self.f["preset_default"](self.h, ch, sr, int(split))

    def presetCheaper(self, ch, sr, split=True):
        self.channels = ch
        self.f["preset_cheaper"](self.h, ch, sr, int(split))

    def configure(self, ch, block, interval, split=False):
        self.channels = ch
        self.f["configure"](self.h, ch, block, interval, int(split))

    def reset(self):
        self.f["reset"](self.h)

    def blockSamples(self):
        return self.f["block_samples"](self.h)

    def intervalSamples(self):
        return self.f["interval_samples"](self.h)

    def inputLatency(self):
        return self.f["input_latency"](self.h)

    def outputLatency(self):
        return self.f["output_latency"](self.h)

    def splitComputation(self):
        return bool(self.f["split_computation"](self.h))

    def seekLength(self):
        return self.f["seek_length"](self.h)

    def outputSeekLength(self, rate):
        return self.f["output_seek_length"](self.h, rate)

    def bands(self):
        return self.f["bands"](self.h)

    def fftSamples(self):
        return self.f["fft_samples"](self.h)

    def setTransposeFactor(self, m, tonality=0.0):
        self.f["set_transpose_factor"](self.h, m, tonality)

    def setTransposeSemitones(self, s, tonality=0.0):
        self.f["set_transpose_semitones"](self.h, s, tonality)

    def setFormantFactor(self, m, comp=False):
        self.f["set_formant_factor"](self.h, m, int(comp))

    def setFormantSemitones(self, s, comp=False):
        self.f["set_formant_semitones"](self.h, s, int(comp))

    def setFormantBase(self, f):
        self.f["set_formant_base"](self.h, f)

    def setFreqMapQuadratic(self, a, b):
        self.f["set_freq_map_quadratic"](self.h, a, b)

    def setFreqMapTable(self, fin, fout):
        """setFreqMap with a piecewise-linear function given by its break points (freq as a multiple of the sample rate)."""
        fin = np.ascontiguousarray(fin, np.float32)
        fout = np.ascontiguousarray(fout, np.float32)
        self.f["set_freq_map_table"](self.h, _fp(fin), _fp(fout), len(fin))

    def _in(self, x):
        x = np.ascontiguousarray(np.asarray(x, np.float32).reshape(self.channels, -1))
        return x, x.shape[1]

    def seek(self, x, rate):
        x, n = self._in(x)
        self.f["seek"](self.h, _fp(x), n, float(rate))

    def outputSeek(self, x):
        x, n = self._in(x)
        self.f["output_seek"](self.h, _fp(x), n)

    def process(self, x, n_out):
        x, n = self._in(x)
        out = np.zeros((self.channels, max(n_out, 1)), np.float32)
        self.f["process"](self.h, _fp(x), n, _fp(out), n_out)
        return out[:, :n_out]

    def flush(self, n_out, rate=0.0):
        out = np.zeros((self.channels, max(n_out, 1)), np.float32)
        self.f["flush"](self.h, _fp(out), n_out, rate)
        return out[:, :n_out]

    def exact(self, x, n_out):
        x, n = self._in(x)
        out = np.zeros((self.channels, max(n_out, 1)), np.float32)
        ok = self.f["exact"](self.h, _fp(x), n, _fp(out), n_out)
        return bool(ok), out[:, :n_out]

    def state(self, name):
        K, C, B = self.bands(), self.channels, self.blockSamples()
        buf = np.zeros(2 * K * C + 2 * B + 16, np.float32)
        n = self.f["get_state"](self.h, STATE[name], _fp(buf))
        assert n >= 0
        v = buf[:n].copy()
        if name in ("input", "prevInput", "output", "predInput"):
            return v.view(np.complex64).reshape(C, K)
        if name in ("inputEnergy", "predEnergy"):
            return v.reshape(C, K)
        if name == "outputMap":
            return v.reshape(K, 2)
        return v

    def signal_state(self):
        """(oracle restatement only) the buffers the CUDA engine also keeps, for teacher-forced tests."""
        K, C = self.bands(), self.channels
        out = {}
        for name, what in (("history", 20), ("pending", 21), ("pendingWp", 22)):
            buf = np.zeros(C * (self.blockSamples() + 2 * self.intervalSamples()) + 16, np.float32)
            n = self.f["get_state"](self.h, what, _fp(buf))
            out[name] = buf[:n].reshape(C, -1).copy()
        for name in ("input", "prevInput", "output", "predEnergy"):
            out[name] = self.state(name)
        return out

    def peaks(self):
        n = self.f["num_peaks"](self.h)
        buf = np.zeros(2 * n + 2, np.float32)
        self.f["get_peaks"](self.h, _fp(buf))
        return buf[:2 * n].reshape(n, 2)
