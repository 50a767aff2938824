"""Pins the oracle (oracle/stretch_oracle.cpp) to the reference: golden vectors generated from the
reference's own binary / header (tests/golden/make_golden.py) and, when oracle/_ref is present,
live comparisons with both reference builds.  CPU only."""
import os

import numpy as np
import pytest

import signals

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


@pytest.mark.parametrize("name", list(signals.CONFIGS))
def test_oracle_matches_reference_header_golden(oracle_port, name):
    """Reference header (unmodified) on the stand-in STFT: the oracle is bit-exact against it."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = oracle_port()
    signals.CONFIGS[name][0](o)
    y = signals.run_single(o, g["x"], float(g["ratio"]), int(g["chunk"]))
    assert y.shape == g["hdr"].shape
    # same libm / compiler on the GPU box image; allow 1e-6 in case sinf/cosf differ by an ulp
    assert np.abs(y - g["hdr"]).max() <= 1e-6


@pytest.mark.parametrize("name", [n for n in signals.CONFIGS if "formant" not in n])
def test_oracle_matches_reference_binary_golden(oracle_port, name):
    """The reference's shipped WASM binary (-O3 -ffast-math, float FFT): agreement to float
    precision over the short horizon; the identity config at any length."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = oracle_port()
    cfg, C, sr, ratio, _ = signals.CONFIGS[name]
    cfg(o)
    y = signals.run_single(o, g["x"], float(g["ratio"]), int(g["chunk"]))
    H = o.intervalSamples()
    lat = o.outputLatency() + int(o.inputLatency() * ratio)
    d = y - g["wasm"]
    if name == "identity":
        assert rms(d) <= 1e-6
    else:
        first = d[:, : lat + 8 * H]
        assert rms(first) <= 1e-4, "short-horizon disagreement with the reference binary"
        assert rms(d) <= 1e-3  # the reference's own regression criterion (-60 dB, cmd/main-dev.cpp:215-232)


def test_kats(oracle_port):
    k = np.load(os.path.join(GOLD, "kats.npz"))
    o = oracle_port()
    o.presetDefault(1, 48000.0)
    win = o.state("window")
    assert np.abs(win[[0, 1440, 2879, 2880, 5759]] - k["window_48k_default_wasm"]).max() < 5e-6
    assert int(win.argmax()) == int(k["window_argmax"][0]) == 2879
    assert abs(win[0] - 0.015485264) < 5e-6 and abs(win[2880] - 0.8161286) < 5e-6
    wp = o.state("windowProducts")  # pending order: index i <-> ring position i + interval
    ref = k["wp_after_configure_wasm"]  # ring positions [0, 1439, 1440, 2880, 5759]
    assert abs(wp[0] - ref[2]) / ref[2] < 1e-5 and abs(wp[1440] - ref[3]) / ref[3] < 1e-5
    assert abs(wp[5759 - 1440] - ref[4]) / ref[4] < 1e-5
    assert ref[0] == np.float32(1e-30) and wp[5759] == np.float32(1e-30)
    presets = [("presetDefault", 44100.0), ("presetDefault", 48000.0), ("presetCheaper", 44100.0), ("presetCheaper", 48000.0),
               ("presetDefault", 96000.0), ("presetDefault", 16000.0)]
    for (preset, sr), row in zip(presets, k["latency_table_wasm"]):
        o = oracle_port()
        getattr(o, preset)(1, sr)
        assert [o.blockSamples(), o.intervalSamples(), o.inputLatency(), o.outputLatency()] == list(row)


def test_seek_process_flush_sequence(oracle_port):
    k = np.load(os.path.join(GOLD, "kats.npz"))
    x = k["seek_process_flush_x"]
    o = oracle_port()
    o.presetDefault(1, 48000.0)
    o.setTransposeSemitones(3, 0)
    o.seek(x[:, :2880], 1.0)
    y = np.concatenate([o.process(x[:, 2880:2880 + 7200], 7200), o.flush(1440, 1.0)], axis=1)
    assert np.abs(y - k["seek_process_flush_hdr"]).max() <= 1e-6
    assert rms(y - k["seek_process_flush_wasm"]) <= 1e-4


def test_identity_is_a_pure_delay(oracle_port):
    o = oracle_port()
    o.presetDefault(1, 48000.0)
    x = signals.harmonic(3 * 5760, 48000)[None]
    y = signals.run_single(o, x, 1.0, 1000)
    lat = o.inputLatency() + o.outputLatency()
    assert rms(y[:, lat:] - x[:, :-lat]) <= 1e-6
    assert np.abs(y[:, :lat]).max() <= 1e-6


def test_chunk_size_invariance(oracle_port):
    x = signals.harmonic(20000, 48000)[None]
    ys = []
    for chunk in (64, 480, 5000):
        o = oracle_port()
        signals.cfg_config3(o)
        ys.append(signals.run_single(o, x, 1.0, chunk))
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])


# ---- live checks against the real reference builds (only where oracle/_ref exists) ----
def _have_ref():
    from oracle import hdrref, wasmref

    return hdrref.available("hdr") and wasmref.available()


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_oracle_bit_exact_vs_reference_header(oracle_port):
    from oracle.hdrref import CpuStretch

    x = np.stack([signals.harmonic(30000, 48000, 1, 0), signals.harmonic(30000, 48000, 1, 1)])
    cases = [
        (lambda o: (o.presetDefault(2, 48000.0), o.setTransposeSemitones(-5, 0)), 1.5, 4800),
        (lambda o: (o.presetCheaper(2, 48000.0, False), o.setTransposeSemitones(4, 0.2), o.setFormantSemitones(3, False), o.setFormantBase(0)), 1.0, 480),
        (lambda o: (o.configure(2, 1000, 250, True), o.setFreqMapQuadratic(1.2, 0.5)), 0.9, 333),
        (lambda o: o.presetDefault(2, 44100.0), 2.5, 441),  # > 2x: exercises the RNG path of the header
        # setFreqMap with a piecewise-linear function: monotone, and one that folds back (non-monotone output map)
        (lambda o: (o.configure(2, 1000, 250, False), o.setFreqMapTable(*signals.PWL_MONOTONE)), 1.0, 500),
        (lambda o: (o.presetDefault(2, 48000.0), o.setFreqMapTable(*signals.PWL_FOLDING)), 1.25, 2880),
    ]
    for cfg, ratio, chunk in cases:
        h, o = CpuStretch("hdr"), oracle_port()
        cfg(h)
        cfg(o)
        assert np.array_equal(signals.run_single(h, x, ratio, chunk), signals.run_single(o, x, ratio, chunk))


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_api_sequence_vs_reference_header(oracle_port):
    from oracle.hdrref import CpuStretch

    x = signals.harmonic(60000, 48000)[None]

    def seq(o):
        o.presetDefault(1, 48000.0)
        o.setTransposeSemitones(3, 0)
        outs = []
        o.seek(x[:, :3000], 1.0)
        outs.append(o.process(x[:, 3000:7800], 4800))
        z = np.zeros((1, 30000), np.float32)
        outs += [o.process(z[:, :12000], 12000), o.process(z[:, :4000], 4000), o.process(z[:, :4000], 5000)]
        outs.append(o.process(x[:, 8000:17600], 9000))
        outs.append(o.flush(1000, 1.0))
        outs.append(o.process(x[:, 20000:24800], 4800))
        outs.append(o.flush(5000, 1.1))
        o.reset()
        outs.append(o.process(x[:, 20000:24800], 2400))
        o.outputSeek(x[:, : o.outputSeekLength(1.3)])
        outs.append(o.process(x[:, 5000:11240], 4800))
        outs.append(o.exact(x[:, :40000], 50000)[1])
        return np.concatenate(outs, axis=1)

    assert np.array_equal(seq(CpuStretch("hdr")), seq(oracle_port()))
