"""ctypes driver for the natively-compiled reference WASM build (oracle/_ref/libwasm_stretch.so).

TEST INFRASTRUCTURE ONLY.  The library is the reference's own shipped binary
(/root/reference/web/emscripten/main.js:9, wrapper web/emscripten/main.cpp:15-77) translated
to C by oracle/wasm2c.py.  Export letters are emscripten's minified names.

Known limits of that binary: built -O3 -ffast-math; presetCheaper hard-wires split=true,
presetDefault split=false (main.cpp:43-48); older formant + flush(>interval) code than the
header in the tree.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "_ref", "libwasm_stretch.so")


def available():
    return os.path.exists(lib_path())


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(lib_path())
        vp, i32, f32, f64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float, ctypes.c_double
        L.wasm_new.restype = vp
        L.wasm_free.argtypes = [vp]
        L.wasm_mem.restype = vp
        L.wasm_mem.argtypes = [vp]
        sigs = {
            "f": (None, []), "y": (i32, [i32, i32]), "h": (i32, [i32, i32]),
            "i": (i32, []), "j": (i32, []), "k": (i32, []), "l": (i32, []), "m": (None, []),
            "n": (None, [i32, f32]), "o": (None, [i32, f32]), "p": (None, [i32, i32, i32, i32]),
            "q": (None, [f32, f32]), "r": (None, [f32, f32]), "s": (None, [f32, i32]),
            "t": (None, [f32, i32]), "u": (None, [f32]), "v": (None, [i32, f64]),
            "w": (None, [i32, i32]), "x": (None, [i32]),
        }
        for k, (res, args) in sigs.items():
            fn = getattr(L, "wasm_export_" + k)
            fn.restype = res
            fn.argtypes = [vp] + args
        _LIB = L
    return _LIB


class WasmStretch:
    """One instance of the reference module (one `Stretch stretch` singleton, main.cpp:9)."""

    def __init__(self):
        L = _lib()
        self.L = L
        self.w = L.wasm_new()
        assert self.w
        L.wasm_export_f(self.w)
        L.wasm_export_y(self.w, 0, 0)
        self.channels = 0
        self.buflen = 0
        self.ptr = 0

    def __del__(self):
        if getattr(self, "w", None):
            self.L.wasm_free(self.w)
            self.w = None

    def _call(self, k, *a):
        return getattr(self.L, "wasm_export_" + k)(self.w, *a)

    def presetDefault(self, ch, sr):
        self.channels = ch
        self._call("n", ch, sr)

    def presetCheaper(self, ch, sr):
        self.channels = ch
        self._call("o", ch, sr)

    def configure(self, ch, block, interval, split=False):
        self.channels = ch
        self._call("p", ch, block, interval, int(split))

    def blockSamples(self):
        return self._call("i")

    def intervalSamples(self):
        return self._call("j")

    def inputLatency(self):
        return self._call("k")

    def outputLatency(self):
        return self._call("l")

    def reset(self):
        self._call("m")

    def setTransposeFactor(self, mult, tonality=0.0):
        self._call("q", mult, tonality)

    def setTransposeSemitones(self, semis, tonality=0.0):
        self._call("r", semis, tonality)

    def setFormantFactor(self, mult, comp=False):
        self._call("s", mult, int(comp))

    def setFormantSemitones(self, semis, comp=False):
        self._call("t", semis, int(comp))

    def setFormantBase(self, f):
        self._call("u", f)

    def _buffers(self, n):
        if n > self.buflen or self.ptr == 0:
            self.buflen = max(n, 1)
            self.ptr = self._call("h", self.channels, self.buflen)
        return self.ptr

    def _view(self):
        base = self.L.wasm_mem(self.w)
        n = self.buflen * self.channels * 2
        arr = (ctypes.c_float * n).from_address(base + self.ptr)
        return np.frombuffer(arr, dtype=np.float32).reshape(2, self.channels, self.buflen)

    def mem_f32(self, addr, n):
        base = self.L.wasm_mem(self.w)
        arr = (ctypes.c_float * n).from_address(int(base) + int(addr))
        return np.frombuffer(arr, dtype=np.float32).copy()

    def mem_u32(self, addr, n=1):
        base = self.L.wasm_mem(self.w)
        arr = (ctypes.c_uint32 * n).from_address(int(base) + int(addr))
        return np.frombuffer(arr, dtype=np.uint32).copy()

    def seek(self, x, rate):
        x = np.asarray(x, np.float32).reshape(self.channels, -1)
        n = x.shape[1]
        self._buffers(n)
        self._view()[0, :, :n] = x
        self._call("v", n, float(rate))

    def process(self, x, n_out):
        """x: [channels][n_in] float32 -> [channels][n_out]."""
        x = np.asarray(x, np.float32).reshape(self.channels, -1)
        n_in = x.shape[1]
        self._buffers(max(n_in, n_out))
        v = self._view()
        v[0, :, :n_in] = x
        self._call("w", n_in, n_out)
        return self._view()[1, :, :n_out].copy()

    def flush(self, n_out):
        self._buffers(n_out)
        self._call("x", n_out)
        return self._view()[1, :, :n_out].copy()
