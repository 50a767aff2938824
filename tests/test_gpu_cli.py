"""The file-mode caller (cmd/stretch_cli.cpp, the role of the reference's cmd/main.cpp) on a GPU: WAV in, the
reference tool's stages (outputSeek / process / flush) through the drop-in C++ facade, WAV out -- against the oracle
driven through the same stages.  `pytest -m gpu`."""
import os
import struct
import subprocess

import numpy as np
import pytest

import signals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


def write_wav_f32(path, x, sr):  # x: [C][n] float32
    C, n = x.shape
    data = np.ascontiguousarray(x.T.astype("<f4")).tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, C, sr, sr * C * 4, C * 4, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def read_wav_i16(path):
    d = open(path, "rb").read()
    assert d[:4] == b"RIFF" and d[8:12] == b"WAVE" and d[12:16] == b"fmt " and d[36:40] == b"data"
    C = struct.unpack("<H", d[22:24])[0]
    n = struct.unpack("<I", d[40:44])[0]
    return (np.frombuffer(d[44:44 + n], dtype="<i2").reshape(-1, C).T / 32768.0).astype(np.float32)


@pytest.fixture(scope="module")
def cli(cuda_lib):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "cmd"), "stretch_cli"], check=True)
    return os.path.join(ROOT, "cmd", "stretch_cli")


@pytest.mark.parametrize("C,semitones,time", [(1, 4.0, 1.25), (2, 0.0, 0.8)])
def test_cli_matches_the_oracle_through_the_reference_tools_stages(cli, oracle_port, tmp_path, C, semitones, time):
    sr, n = 48000, 30000
    x = signals.batch("harmonic", 1, C, n, sr)[0]
    x = (np.round(x * 32768) / 32768).astype(np.float32)  # values a 16-bit file would hold as well
    src, dst = str(tmp_path / "in.wav"), str(tmp_path / "out.wav")
    write_wav_f32(src, x, sr)
    r = subprocess.run([cli, src, dst, "--semitones=%g" % semitones, "--time=%g" % time], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    y = read_wav_i16(dst)
    # the same stages on the oracle (cmd/main.cpp:58-82)
    o = oracle_port()
    o.presetDefault(C, float(sr))
    o.setTransposeSemitones(semitones, 8000.0 / sr)
    o.setFormantSemitones(0.0, False)
    o.setFormantBase(100.0 / sr)
    n_out = int(round(n * time))
    seek = o.outputSeekLength(np.float32(1 / time))
    out_index = n_out - o.intervalSamples()
    in_index = int(round((out_index + o.outputLatency()) / time)) + o.inputLatency()
    xp = np.zeros((C, max(in_index, seek)), np.float32)
    xp[:, :n] = x
    o.outputSeek(xp[:, :seek])
    ref = np.concatenate([o.process(xp[:, seek:in_index], out_index), o.flush(n_out - out_index)], axis=1)
    assert y.shape == ref.shape
    q = np.clip(np.round(ref * 32768), -32768, 32767) / 32768.0
    H = o.intervalSamples()
    assert rms(y[:, : 8 * H] - q[:, : 8 * H]) <= 1e-4  # first blocks: north-star tolerance (+ 16-bit rounding)
    assert rms(y - q) <= 1e-3                           # whole file: the reference's own -60 dB criterion
