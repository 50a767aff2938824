"""signalsmith_stretch_b200 -- host-side mirror of the reference API over the B200 C-ABI library.

The product is `libb200stretch.so` (hand-written sm_100a CUDA behind the `extern "C"` ABI declared
in include/b200_stretch.h).  This module is a thin ctypes binding whose method names, argument
meaning and defaults follow `signalsmith::stretch::SignalsmithStretch<float>`
(/root/reference/signalsmith-stretch.h:34-491), batched: buffers are `[batch][channels][samples]`.

There is NO CPU fallback: importing is cheap, but constructing a `BatchStretch` raises
`StretchError` if the CUDA library is missing or no GPU is usable.
"""
import ctypes
import os

import numpy as np

__all__ = ["BatchStretch", "StretchError", "library_path", "build_library", "ABI_SYMBOLS"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB_NAME = "libb200stretch.so"

# every symbol include/b200_stretch.h declares
ABI_SYMBOLS = [
    "b200s_create", "b200s_destroy", "b200s_last_error", "b200s_version", "b200s_set_stream", "b200s_synchronize", "b200s_set_sub_batches", "b200s_set_tuning",
    "b200s_preset_default", "b200s_preset_cheaper", "b200s_configure", "b200s_reset", "b200s_reserve",
    "b200s_batch", "b200s_channels", "b200s_block_samples", "b200s_interval_samples", "b200s_input_latency",
    "b200s_output_latency", "b200s_split_computation", "b200s_seek_length", "b200s_output_seek_length",
    "b200s_fft_samples", "b200s_bands",
    "b200s_set_transpose_factor", "b200s_set_transpose_semitones", "b200s_set_formant_factor",
    "b200s_set_formant_semitones", "b200s_set_formant_base", "b200s_set_freq_map_table",
    "b200s_seek", "b200s_seek_rates", "b200s_live_seek", "b200s_output_seek", "b200s_process", "b200s_process_async", "b200s_process_pcm16", "b200s_flush", "b200s_exact",
    "b200s_seek_device", "b200s_process_device", "b200s_flush_device",
    "b200s_timer_start", "b200s_timer_stop", "b200s_kernel_launches", "b200s_device_allocations", "b200s_unserved_random_blocks", "b200s_profile_begin", "b200s_profile_end",
    "b200s_selftest_divsqrt",
    "b200s_state_size", "b200s_get_state", "b200s_set_state",
]

STATE = {"input": 0, "prevInput": 1, "output": 2, "predEnergy": 4, "history": 20, "pending": 21, "pendingWp": 22}


class StretchError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, _LIB_NAME)


def build_library(verbose=False):
    """Compile csrc/engine.cu for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    import subprocess

    src = os.path.join(_HERE, "csrc", "engine.cu")
    out = library_path()
    deps = [src] + [os.path.join(_HERE, "csrc", f) for f in ("kernels.cuh", "fft.cuh", "common.cuh", "chain_direct.cuh", "chain_direct2.cuh", "fft2.cuh", "stft2.cuh", "chain_direct3.cuh", "chain_direct4.cuh", "chain_ws.cuh", "chain_t.cuh", "chain_direct6.cuh")]
    deps.append(os.path.join(_ROOT, "include", "b200_stretch.h"))
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", out, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("B200S_EXTRA_NVCC"):  # profiling builds (e.g. -DB200S_CHAIN_PROBES); never set for the shipped library
        cmd[1:1] = os.environ["B200S_EXTRA_NVCC"].split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise StretchError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


def _bind(lib):
    vp, ci, cf, cd, cl, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_long, ctypes.c_longlong
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    sig = {
        "b200s_create": (ci, [ci, cl, ci, ctypes.POINTER(vp)]), "b200s_destroy": (None, [vp]),
        "b200s_last_error": (ctypes.c_char_p, [vp]), "b200s_version": (ci, [ip, ip, ip]),
        "b200s_set_stream": (ci, [vp, vp]), "b200s_synchronize": (ci, [vp]), "b200s_set_sub_batches": (ci, [vp, ci]), "b200s_set_tuning": (ci, [vp, ci, ci]),
        "b200s_preset_default": (ci, [vp, ci, cf, ci]), "b200s_preset_cheaper": (ci, [vp, ci, cf, ci]),
        "b200s_configure": (ci, [vp, ci, ci, ci, ci]), "b200s_reset": (ci, [vp]), "b200s_reserve": (ci, [vp, ci, ci]),
        "b200s_output_seek_length": (ci, [vp, cf]),
        "b200s_set_transpose_factor": (ci, [vp, cf, cf]), "b200s_set_transpose_semitones": (ci, [vp, cf, cf]),
        "b200s_set_formant_factor": (ci, [vp, cf, ci]), "b200s_set_formant_semitones": (ci, [vp, cf, ci]),
        "b200s_set_formant_base": (ci, [vp, cf]), "b200s_set_freq_map_table": (ci, [vp, fp, fp, ci]),
        "b200s_seek": (ci, [vp, vp, ci, cd]), "b200s_seek_rates": (ci, [vp, vp, ci, ctypes.POINTER(cd)]), "b200s_live_seek": (ci, [vp, vp, cll, ctypes.POINTER(cll), ci, ctypes.POINTER(cd)]), "b200s_output_seek": (ci, [vp, vp, ci]),
        "b200s_process": (ci, [vp, vp, ci, vp, ci]), "b200s_process_async": (ci, [vp, vp, ci, vp, ci]), "b200s_process_pcm16": (ci, [vp, vp, ci, vp, ci, ci]), "b200s_flush": (ci, [vp, vp, ci, cf]),
        "b200s_exact": (ci, [vp, vp, ci, vp, ci, ip]),
        "b200s_seek_device": (ci, [vp, vp, ci, cd]), "b200s_process_device": (ci, [vp, vp, ci, vp, ci]),
        "b200s_flush_device": (ci, [vp, vp, ci, cf]),
        "b200s_timer_start": (ci, [vp]), "b200s_timer_stop": (ci, [vp, fp]), "b200s_kernel_launches": (cll, [vp]), "b200s_device_allocations": (cll, [vp]), "b200s_unserved_random_blocks": (cll, [vp]),
        "b200s_profile_begin": (ci, [vp]), "b200s_profile_end": (ci, [vp, fp, ip, ci]),
        "b200s_selftest_divsqrt": (ci, [vp, cll, cll, ctypes.POINTER(cll), ctypes.POINTER(cll)]),
        "b200s_state_size": (ci, [vp, ci]), "b200s_get_state": (ci, [vp, ci, vp]), "b200s_set_state": (ci, [vp, ci, vp]),
    }
    for name in ("batch", "channels", "block_samples", "interval_samples", "input_latency", "output_latency",
                 "split_computation", "seek_length", "fft_samples", "bands"):
        sig["b200s_" + name] = (ci, [vp])
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIBS = {}


def _load(path):
    if path not in _LIBS:
        if not os.path.exists(path):
            raise StretchError(
                "CUDA library %s not found: build it with signalsmith_stretch_b200.build_library() "
                "(there is no CPU fallback)" % path)
        _LIBS[path] = _bind(ctypes.CDLL(path))
    return _LIBS[path]


def _host(a, shape):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if a.size != int(np.prod(shape)):
        raise ValueError("expected %s floats, got %s" % (shape, a.shape))
    return a.reshape(shape)


def _is_torch_cuda(t):
    return hasattr(t, "data_ptr") and hasattr(t, "is_cuda") and t.is_cuda


class BatchStretch:
    """A batch of independent streams; one `SignalsmithStretch<float>` per stream (signalsmith-stretch.h:34)."""

    version = (1, 3, 2)  # :36

    def __init__(self, batch=1, seed=0, device=0, lib_path=None):
        self._lib = _load(lib_path or library_path())
        h = ctypes.c_void_p()
        rc = self._lib.b200s_create(int(batch), int(seed), int(device), ctypes.byref(h))
        if rc != 0:
            raise StretchError("b200s_create failed (%d): %s" % (rc, self._lib.b200s_last_error(None).decode()))
        self._h = h
        self.batch = int(batch)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200s_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise StretchError("b200_stretch error %d: %s" % (rc, self._lib.b200s_last_error(self._h).decode()))

    # ---- configuration (:49-104) ----
    def presetDefault(self, nChannels, sampleRate, splitComputation=False):
        self._ck(self._lib.b200s_preset_default(self._h, nChannels, sampleRate, int(splitComputation)))

    def presetCheaper(self, nChannels, sampleRate, splitComputation=True):
        self._ck(self._lib.b200s_preset_cheaper(self._h, nChannels, sampleRate, int(splitComputation)))

    def configure(self, nChannels, blockSamples, intervalSamples, splitComputation=False):
        self._ck(self._lib.b200s_configure(self._h, nChannels, blockSamples, intervalSamples, int(splitComputation)))

    def reset(self):
        self._ck(self._lib.b200s_reset(self._h))

    def reserve(self, maxInputSamples, maxOutputSamples):
        self._ck(self._lib.b200s_reserve(self._h, maxInputSamples, maxOutputSamples))

    def channels(self):
        return self._lib.b200s_channels(self._h)

    def blockSamples(self):
        return self._lib.b200s_block_samples(self._h)

    def intervalSamples(self):
        return self._lib.b200s_interval_samples(self._h)

    def inputLatency(self):
        return self._lib.b200s_input_latency(self._h)

    def outputLatency(self):
        return self._lib.b200s_output_latency(self._h)

    def splitComputation(self):
        return bool(self._lib.b200s_split_computation(self._h))

    def seekLength(self):
        return self._lib.b200s_seek_length(self._h)

    def outputSeekLength(self, playbackRate):
        return self._lib.b200s_output_seek_length(self._h, playbackRate)

    def fftSamples(self):
        return self._lib.b200s_fft_samples(self._h)

    def bands(self):
        return self._lib.b200s_bands(self._h)

    # ---- parameters (:107-135) ----
    def setTransposeFactor(self, multiplier, tonalityLimit=0.0):
        self._ck(self._lib.b200s_set_transpose_factor(self._h, multiplier, tonalityLimit))

    def setTransposeSemitones(self, semitones, tonalityLimit=0.0):
        self._ck(self._lib.b200s_set_transpose_semitones(self._h, semitones, tonalityLimit))

    def setFormantFactor(self, multiplier, compensatePitch=False):
        self._ck(self._lib.b200s_set_formant_factor(self._h, multiplier, int(compensatePitch)))

    def setFormantSemitones(self, semitones, compensatePitch=False):
        self._ck(self._lib.b200s_set_formant_semitones(self._h, semitones, int(compensatePitch)))

    def setFormantBase(self, baseFreq=0.0):
        self._ck(self._lib.b200s_set_formant_base(self._h, baseFreq))

    def setFreqMap(self, inputToOutput, points=2049):
        """setFreqMap(std::function) (:120): the callable is tabulated on [0, 0.5] (piecewise linear)."""
        if inputToOutput is None:
            self._ck(self._lib.b200s_set_freq_map_table(self._h, None, None, 0))
            return
        fin = np.linspace(0.0, 0.5, points).astype(np.float32)
        fout = np.array([inputToOutput(float(f)) for f in fin], np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        self._ck(self._lib.b200s_set_freq_map_table(self._h, fin.ctypes.data_as(fp), fout.ctypes.data_as(fp), points))

    def setFreqMapTable(self, fin, fout):
        """The piecewise-linear map itself (b200s_set_freq_map_table): break points fin (ascending) -> fout."""
        fin = np.ascontiguousarray(fin, np.float32)
        fout = np.ascontiguousarray(fout, np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        self._ck(self._lib.b200s_set_freq_map_table(self._h, fin.ctypes.data_as(fp), fout.ctypes.data_as(fp), len(fin)))

    # ---- the hot path ----
    def _shape(self, n):
        c = self.channels()
        if c <= 0:
            raise StretchError("engine not configured: call presetDefault / presetCheaper / configure first")
        return (self.batch, c, n)

    def _adopt_torch_stream(self):
        """Device-pointer calls are enqueued on torch's CURRENT stream (b200s_set_stream), so they are ordered after the
        producers of `inputs` and before any torch op on the returned tensor, like a torch op.  (The engine's own side
        streams fork from and join that stream.)"""
        import torch

        st = torch.cuda.current_stream().cuda_stream
        if getattr(self, "_torch_stream", None) != st:
            self._ck(self._lib.b200s_set_stream(self._h, ctypes.c_void_p(st)))
            self._torch_stream = st

    def seek(self, inputs, playbackRate):
        if _is_torch_cuda(inputs):
            self._adopt_torch_stream()
            n = inputs.shape[-1]
            self._ck(self._lib.b200s_seek_device(self._h, inputs.data_ptr(), n, float(playbackRate)))
            return
        x = np.asarray(inputs, np.float32)
        n = x.shape[-1]
        x = _host(x, self._shape(n))
        if np.ndim(playbackRate) > 0:  # one rate per stream (b200s_seek_rates)
            r = np.ascontiguousarray(playbackRate, np.float64)
            if r.shape != (self.batch,):
                raise ValueError("expected %d playback rates" % self.batch)
            self._ck(self._lib.b200s_seek_rates(self._h, x.ctypes.data, n, r.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
            return
        self._ck(self._lib.b200s_seek(self._h, x.ctypes.data, n, float(playbackRate)))

    def live_seek(self, bank_ptr, bank_len, window_end, window, rates):
        """b200s_live_seek: seek windows cut from a device-resident audio bank (pointer), one end index and rate per stream."""
        we = np.ascontiguousarray(window_end, np.int64)
        r = np.ascontiguousarray(rates, np.float64)
        if we.shape != (self.batch,) or r.shape != (self.batch,):
            raise ValueError("expected %d window ends and rates" % self.batch)
        self._ck(self._lib.b200s_live_seek(self._h, ctypes.c_void_p(int(bank_ptr)), int(bank_len), we.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                           int(window), r.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))

    def outputSeek(self, inputs):
        x = np.asarray(inputs, np.float32)
        n = x.shape[-1]
        x = _host(x, self._shape(n))
        self._ck(self._lib.b200s_output_seek(self._h, x.ctypes.data, n))

    def process(self, inputs, outputSamples, out=None):
        """inputs: [batch][channels][inputSamples] (numpy -> host API; torch CUDA tensor -> device API)."""
        if _is_torch_cuda(inputs):
            import torch

            self._adopt_torch_stream()
            n_in = inputs.shape[-1]
            assert inputs.is_contiguous() and inputs.dtype == torch.float32
            if out is None:
                out = torch.empty(self._shape(outputSamples), dtype=torch.float32, device=inputs.device)
            self._ck(self._lib.b200s_process_device(self._h, inputs.data_ptr(), n_in, out.data_ptr(), outputSamples))
            return out
        x = np.asarray(inputs, np.float32)
        n_in = x.shape[-1]
        x = _host(x, self._shape(n_in))
        if out is None:
            out = np.empty(self._shape(max(outputSamples, 0)), np.float32)
        self._ck(self._lib.b200s_process(self._h, x.ctypes.data, n_in, out.ctypes.data, outputSamples))
        return out

    def process_host_ptr(self, in_ptr, n_in, out_ptr, n_out):
        """Raw host pointers (e.g. pinned torch tensors): the end-to-end call bench.py times."""
        self._ck(self._lib.b200s_process(self._h, in_ptr, n_in, out_ptr, n_out))

    def process_host_ptr_async(self, in_ptr, n_in, out_ptr, n_out):
        """b200s_process_async: enqueue and return; buffers (pinned) belong to the engine until synchronize()."""
        self._ck(self._lib.b200s_process_async(self._h, in_ptr, n_in, out_ptr, n_out))

    def process_pcm16(self, inputs, outputSamples):
        """16-bit PCM in / out (int16 arrays [batch][channels][n]); conversion on the device."""
        x = np.ascontiguousarray(inputs, dtype=np.int16)
        n_in = x.shape[-1]
        out = np.empty(self._shape(max(outputSamples, 0)), np.int16)
        self._ck(self._lib.b200s_process_pcm16(self._h, x.ctypes.data, n_in, out.ctypes.data, outputSamples, 1))
        return out

    def process_pcm16_ptr(self, in_ptr, n_in, out_ptr, n_out, wait=False):
        self._ck(self._lib.b200s_process_pcm16(self._h, in_ptr, n_in, out_ptr, n_out, 1 if wait else 0))

    def flush(self, outputSamples, playbackRate=0.0):
        out = np.empty(self._shape(max(outputSamples, 0)), np.float32)
        self._ck(self._lib.b200s_flush(self._h, out.ctypes.data, outputSamples, playbackRate))
        return out

    def exact(self, inputs, outputSamples):
        x = np.asarray(inputs, np.float32)
        n_in = x.shape[-1]
        x = _host(x, self._shape(n_in))
        out = np.empty(self._shape(outputSamples), np.float32)
        ok = ctypes.c_int(0)
        self._ck(self._lib.b200s_exact(self._h, x.ctypes.data, n_in, out.ctypes.data, outputSamples, ctypes.byref(ok)))
        return bool(ok.value), out

    # ---- plumbing ----
    def set_stream(self, cuda_stream):
        self._ck(self._lib.b200s_set_stream(self._h, ctypes.c_void_p(int(cuda_stream))))

    def set_tuning(self, key, value):
        """Implementation selectors (include/b200_stretch.h): 0 chain kernel generation, 1 scalar FFT kernels, 2 host pipeline groups."""
        self._ck(self._lib.b200s_set_tuning(self._h, int(key), int(value)))

    def set_sub_batches(self, n):
        self._ck(self._lib.b200s_set_sub_batches(self._h, n))

    def synchronize(self):
        self._ck(self._lib.b200s_synchronize(self._h))

    def timer_start(self):
        self._ck(self._lib.b200s_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        self._ck(self._lib.b200s_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def kernel_launches(self):
        return int(self._lib.b200s_kernel_launches(self._h))

    def unserved_random_blocks(self):
        return int(self._lib.b200s_unserved_random_blocks(self._h))

    def device_allocations(self):
        return int(self._lib.b200s_device_allocations(self._h))

    KERNELS = ("plan", "analyse", "prep", "chain", "synth", "commit")

    def profile_begin(self):
        self._ck(self._lib.b200s_profile_begin(self._h))

    def profile_end(self):
        """{kernel: (total ms, launches)} for the process() calls since profile_begin()."""
        ms = (ctypes.c_float * 6)()
        cnt = (ctypes.c_int * 6)()
        self._ck(self._lib.b200s_profile_end(self._h, ms, cnt, 6))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    def selftest_divsqrt(self, n=1 << 27, seed=1):
        a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
        self._ck(self._lib.b200s_selftest_divsqrt(self._h, n, seed, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def get_state(self, name):
        what = STATE[name]
        n = self._lib.b200s_state_size(self._h, what)
        buf = np.empty((self.batch, n), np.float32)
        self._ck(self._lib.b200s_get_state(self._h, what, buf.ctypes.data))
        C, K = self.channels(), self.bands()
        if name in ("input", "prevInput", "output"):
            return buf.view(np.complex64).reshape(self.batch, C, K)
        return buf.reshape(self.batch, C, -1)

    def set_state(self, name, value):
        what = STATE[name]
        n = self._lib.b200s_state_size(self._h, what)
        v = np.ascontiguousarray(value)
        if v.dtype == np.complex64:
            v = v.view(np.float32)
        v = np.ascontiguousarray(v, np.float32).reshape(self.batch, n)
        self._ck(self._lib.b200s_set_state(self._h, what, v.ctypes.data))
