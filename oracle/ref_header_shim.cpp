// oracle/ref_header_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI shim around the UNMODIFIED reference header /root/reference/signalsmith-stretch.h,
// compiled against the oracle's stand-in for its missing dependency
// (oracle/signalsmith-linear/stft.h).  Built by oracle/Makefile into oracle/_ref/ (git-ignored);
// reference sources are #included from where they lie, never copied.
// `private` is made public AFTER the standard headers are in, purely to export white-box state
// (Band / Prediction / outputMap) for teacher-forced parity tests.
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstring>
#include <functional>
#include <limits>
#include <random>
#include <type_traits>
#include <vector>

#include "signalsmith-linear/stft.h"
#define private public
#include "signalsmith-stretch.h"
#undef private

using Stretch = signalsmith::stretch::SignalsmithStretch<float>;

namespace {
struct Planar {
	float *p;
	int n;
	float *operator[](int c) { return p + (size_t)c * n; }
};
struct ConstPlanar {
	const float *p;
	int n;
	const float *operator[](int c) { return p + (size_t)c * n; }
};
} // namespace

extern "C" {
void *hdr_new(long seed) { return new Stretch(seed); }
void hdr_free(void *h) { delete (Stretch *)h; }
void hdr_preset_default(void *h, int ch, float sr, int split) { ((Stretch *)h)->presetDefault(ch, sr, split != 0); }
void hdr_preset_cheaper(void *h, int ch, float sr, int split) { ((Stretch *)h)->presetCheaper(ch, sr, split != 0); }
void hdr_configure(void *h, int ch, int block, int interval, int split) { ((Stretch *)h)->configure(ch, block, interval, split != 0); }
void hdr_reset(void *h) { ((Stretch *)h)->reset(); }
int hdr_block_samples(void *h) { return ((Stretch *)h)->blockSamples(); }
int hdr_interval_samples(void *h) { return ((Stretch *)h)->intervalSamples(); }
int hdr_input_latency(void *h) { return ((Stretch *)h)->inputLatency(); }
int hdr_output_latency(void *h) { return ((Stretch *)h)->outputLatency(); }
int hdr_split_computation(void *h) { return ((Stretch *)h)->splitComputation(); }
int hdr_seek_length(void *h) { return ((Stretch *)h)->seekLength(); }
int hdr_output_seek_length(void *h, float rate) { return ((Stretch *)h)->outputSeekLength(rate); }
int hdr_bands(void *h) { return ((Stretch *)h)->bands; }
int hdr_fft_samples(void *h) { return (int)((Stretch *)h)->stft.fftSamples(); }
void hdr_set_transpose_factor(void *h, float m, float t) { ((Stretch *)h)->setTransposeFactor(m, t); }
void hdr_set_transpose_semitones(void *h, float s, float t) { ((Stretch *)h)->setTransposeSemitones(s, t); }
void hdr_set_formant_factor(void *h, float m, int comp) { ((Stretch *)h)->setFormantFactor(m, comp != 0); }
void hdr_set_formant_semitones(void *h, float s, int comp) { ((Stretch *)h)->setFormantSemitones(s, comp != 0); }
void hdr_set_formant_base(void *h, float f) { ((Stretch *)h)->setFormantBase(f); }
// custom map kind 1: piecewise "octave fold" used by the tests: f -> a*f + b*f*f (monotone for small b)
void hdr_set_freq_map_quadratic(void *h, float a, float b) {
	((Stretch *)h)->setFreqMap([a, b](float f) { return a * f + b * f * f; });
}
// custom map kind 2: piecewise-linear table (same function as oracle/stretch_oracle.cpp pwl_map and the product's map_freq)
static inline float pwl_map(const float *in, const float *out, int n, float freq) {
	if (n == 1) return out[0] + (freq - in[0]);
	int lo = 0, hi = n - 1;
	while (hi - lo > 1) {
		int mid = (lo + hi) >> 1;
		if (in[mid] <= freq) lo = mid;
		else hi = mid;
	}
	float x0 = in[lo], x1 = in[lo + 1], y0 = out[lo], y1 = out[lo + 1];
	return y0 + (freq - x0) * ((y1 - y0) / (x1 - x0));
}
void hdr_set_freq_map_table(void *h, const float *fin, const float *fout, int n) {
	if (n <= 0) {
		((Stretch *)h)->setFreqMap(nullptr);
		return;
	}
	std::vector<float> a(fin, fin + n), b(fout, fout + n);
	((Stretch *)h)->setFreqMap([a, b](float f) { return pwl_map(a.data(), b.data(), (int)a.size(), f); });
}
void hdr_seek(void *h, const float *in, int n, double rate) {
	ConstPlanar p{in, n};
	((Stretch *)h)->seek(p, n, rate);
}
void hdr_output_seek(void *h, const float *in, int n) {
	ConstPlanar p{in, n};
	((Stretch *)h)->outputSeek(p, n);
}
void hdr_process(void *h, const float *in, int nIn, float *out, int nOut) {
	ConstPlanar pi{in, nIn};
	Planar po{out, nOut};
	((Stretch *)h)->process(pi, nIn, po, nOut);
}
void hdr_flush(void *h, float *out, int nOut, float rate) {
	Planar po{out, nOut};
	((Stretch *)h)->flush(po, nOut, rate);
}
int hdr_exact(void *h, const float *in, int nIn, float *out, int nOut) {
	ConstPlanar pi{in, nIn};
	Planar po{out, nOut};
	return ((Stretch *)h)->exact(pi, nIn, po, nOut) ? 1 : 0;
}

// ---- white-box state (teacher forcing) ----
// what: 0 input, 1 prevInput, 2 output (complex, 2*bands*channels floats)
//       3 inputEnergy, 4 prediction energy (bands*channels floats)
//       5 outputMap (2*bands: inputBin, freqGrad)  6 energy  7 smoothedEnergy (bands)
//       8 analysis window (block)  9 windowProducts (block)  10 prediction input (complex)
//       11 formantMetric (bands+2)
int hdr_get_state(void *h, int what, float *dst) {
	Stretch &s = *(Stretch *)h;
	int K = s.bands, C = s.channels, n = 0;
	switch (what) {
	case 0: case 1: case 2:
		for (int i = 0; i < K * C; ++i) {
			auto &b = s.channelBands[i];
			std::complex<float> v = what == 0 ? b.input : what == 1 ? b.prevInput : b.output;
			dst[n++] = v.real();
			dst[n++] = v.imag();
		}
		break;
	case 3: for (int i = 0; i < K * C; ++i) dst[n++] = s.channelBands[i].inputEnergy; break;
	case 4: for (int i = 0; i < K * C; ++i) dst[n++] = s.channelPredictions[i].energy; break;
	case 5: for (int i = 0; i < K; ++i) { dst[n++] = s.outputMap[i].inputBin; dst[n++] = s.outputMap[i].freqGrad; } break;
	case 6: for (int i = 0; i < K; ++i) dst[n++] = s.energy[i]; break;
	case 7: for (int i = 0; i < K; ++i) dst[n++] = s.smoothedEnergy[i]; break;
	case 8: for (float v : s.stft.analysisWindow()) dst[n++] = v; break;
	case 9: for (float v : s.stft.output.windowProducts) dst[n++] = v; break;
	case 10:
		for (int i = 0; i < K * C; ++i) { dst[n++] = s.channelPredictions[i].input.real(); dst[n++] = s.channelPredictions[i].input.imag(); }
		break;
	case 11: for (float v : s.formantMetric) dst[n++] = v; break;
	default: return -1;
	}
	return n;
}
int hdr_num_peaks(void *h) { return (int)((Stretch *)h)->peaks.size(); }
int hdr_get_peaks(void *h, float *dst) {
	Stretch &s = *(Stretch *)h;
	int n = 0;
	for (auto &p : s.peaks) { dst[n++] = p.input; dst[n++] = p.output; }
	return n;
}
}
