// oracle/stretch_oracle.cpp -- THE ORACLE (test infrastructure; never linked into the product).
//
// A from-scratch CPU restatement of the reference's STFT phase-vocoder hot path
//   /root/reference/signalsmith-stretch.h  (SignalsmithStretch<float>)  +
//   the used subset of its un-vendored dependency signalsmith-linear 0.2.6 (DynamicSTFT<float,false,true>)
// in the *frame-batched* decomposition the CUDA implementation uses:
//     plan (block scheduler)  ->  analyse  ->  spectral stage  ->  synthesise / overlap-add
// on linear history / pending buffers instead of rings, SoA spectra instead of AoS `Band`s.
// Every function cites the reference lines it follows.  All arithmetic is IEEE float in the
// reference's association order (build with -ffp-contract=off), the FFT is oracle/fft_ref.h
// (double, rounded once), so this file is BIT-EXACT against the unmodified reference header
// compiled on the oracle's stand-in STFT (oracle/_ref/libhdr_stretch.so) -- that equality, and
// the agreement of both with the reference's shipped WASM binary, is what pins it
// (tests/test_oracle_pinning.py, golden vectors under tests/golden/).
//
// Parity status: pinned (a) bit-exactly to reference header + stand-in STFT, (b) to float
// precision against the reference's own binary for STFT / non-formant process / seek / reset
// (the shipped binary predates the in-tree formant + flush code).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "fft_ref.h"

namespace oracle {

typedef std::complex<float> cfloat;

// signalsmith-stretch.h:16-32 (_impl::mul / norm)
static inline cfloat cmul(cfloat a, cfloat b) {
	return cfloat(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
static inline cfloat cmulConj(cfloat a, cfloat b) { // a * conj(b)
	return cfloat(b.real() * a.real() + b.imag() * a.imag(), b.real() * a.imag() - b.imag() * a.real());
}
static inline float cnorm(cfloat a) { return a.real() * a.real() + a.imag() * a.imag(); }

static const float noiseFloor = 1e-15f;    // :508
static const float maxCleanStretch = 2.0f; // :509
static const float almostZero = 1e-30f;    // dependency constant, measured
static const int64_t NEVER = std::numeric_limits<int64_t>::max() / 4; // samplesSinceLast initial (:496)

// One block of the schedule (signalsmith-stretch.h:281-319)
struct Frame {
	int t;           // output index (within the call) at which the block is triggered
	int inputOffset; // :288
	bool newSpectrum, reanalysePrev, mapped, formants;
	float timeFactor;
};

// custom map kind 2: a piecewise-linear table, i.e. what a caller passes to setFreqMap (:120) when the map is tabulated:
// `n` points (in[i] ascending), linear between neighbours and extrapolated beyond the ends.  The same float expression,
// segment search and operation order as the product's map_freq (csrc/kernels.cuh) -- it is the USER's function, so the
// reference, the oracle and the product must all be handed the identical one.
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

struct Stretch {
	// ---------------- configuration (signalsmith-stretch.h:63-94, dependency contract App. B) ----
	int C = 0, B = 0, H = 0, N = 0, K = 0;
	bool split = false;
	std::vector<float> window;   // analysis == synthesis window
	std::vector<float> winProd;  // window[i]*window[i]*N, the per-block windowProducts increment
	std::vector<cfloat> rot;     // per-bin rotation table, float recurrence of :647-655
	ModifiedRealFFT fft;

	// ---------------- parameters (:107-135) ----------------
	float freqMultiplier = 1, freqTonalityLimit = 0.5f;
	bool formantCompensation = false;
	float formantMultiplier = 1, invFormantMultiplier = 1;
	float formantBaseFreq = 0;
	int customMap = 0; // 0 none, 1 quadratic a*f+b*f*f (test stand-in for setFreqMap, :120), 2 piecewise-linear table
	std::vector<float> mapTabIn, mapTabOut;
	float mapA = 1, mapB = 0;

	// ---------------- scheduler state (:494-529) ----------------
	int64_t samplesSinceLast = NEVER;
	int prevInputOffset = -1;
	bool didSeek = false;
	float seekTimeFactor = 1;
	int64_t silenceCounter = 0;
	bool silenceFirst = true;
	float freqEstimateWeighted = 0, freqEstimateWeight = 0; // :927-928
	std::default_random_engine randomEngine;

	// ---------------- signal state ----------------
	int histLen = 0;                 // B + H samples of input history per channel (ring length B+H+1 in the reference)
	std::vector<float> hist;         // [C][histLen], most recent sample last
	int pendLen = 0, addOff = 0;     // pending output: B (+H when split); frames land at addOff
	std::vector<float> pend;         // [C][pendLen]  overlap-add accumulator for the NEXT output samples
	std::vector<float> pendWp;       // [pendLen]     matching windowProducts
	std::vector<cfloat> input, prevInput, output; // [C][K]   Band::input / prevInput / output (:538-542)
	std::vector<float> inputEnergy, predEnergy;   // [C][K]   Band::inputEnergy, Prediction::energy
	std::vector<cfloat> predInput;                // [C][K]   Prediction::input
	std::vector<float> energy, smoothed;          // [K]
	std::vector<float> mapBin, mapGrad;           // [K]      PitchMapPoint
	std::vector<float> peakIn, peakOut;
	std::vector<float> formantMetric;             // [K+2]
	float freqEstimate = 0;

	explicit Stretch(long seed) : randomEngine(seed) {}

	// ---- dependency contract ----
	float binToFreq(float b) const { return (b + 0.5f) / float(N); }
	float freqToBin(float f) const { return f * float(N) - 0.5f; }

	// signalsmith-stretch.h:71-94 + DynamicSTFT::configure/setInterval (App. B)
	void configure(int channels, int block, int interval, bool splitComputation) {
		split = splitComputation;
		C = channels;
		B = block;
		H = interval;
		N = 2 * fastSizeAbove((block + 1) / 2);
		K = N / 2;
		fft.resize(N);
		// Kaiser window with the dependency's bandwidth heuristic, then forced perfect reconstruction
		double bw = double(B) / double(H);
		bw += 8 / ((bw + 3) * (bw + 3)) + 0.25 * std::max(3 - bw, 0.0);
		bw = std::max(bw, 2.0);
		double beta = M_PI * std::sqrt(bw * bw * 0.25 - 1);
		std::vector<double> w(B);
		for (int i = 0; i < B; ++i) {
			double r = (2.0 * i + 1) / B - 1;
			w[i] = bessel0(beta * std::sqrt(std::max(0.0, 1 - r * r))) / bessel0(beta);
		}
		for (int i = 0; i < H && i < B; ++i) {
			double s2 = 0;
			for (int k = i; k < B; k += H) s2 += w[k] * w[k];
			double f = 1 / std::sqrt(s2);
			for (int k = i; k < B; k += H) w[k] *= f;
		}
		window.resize(B);
		winProd.resize(B);
		for (int i = 0; i < B; ++i) {
			window[i] = float(w[i]);
			winProd[i] = window[i] * window[i] * float(N);
		}
		// :647-655 -- rotation by one interval, accumulated with the reference's float recurrence
		rot.resize(K);
		{
			cfloat r = std::polar(1.0f, binToFreq(0) * float(H) * float(2 * M_PI));
			float freqStep = binToFreq(1) - binToFreq(0);
			cfloat step = std::polar(1.0f, freqStep * float(H) * float(2 * M_PI));
			for (int b = 0; b < K; ++b) {
				rot[b] = r;
				r = cmul(r, step);
			}
		}
		histLen = B + H;
		pendLen = B + (split ? H : 0);
		addOff = split ? H : 0;
		hist.assign(size_t(C) * histLen, 0);
		pend.assign(size_t(C) * pendLen, 0);
		pendWp.assign(pendLen, 0);
		input.assign(size_t(C) * K, 0);
		prevInput.assign(size_t(C) * K, 0);
		output.assign(size_t(C) * K, 0);
		inputEnergy.assign(size_t(C) * K, 0);
		predEnergy.assign(size_t(C) * K, 0);
		predInput.assign(size_t(C) * K, 0);
		energy.assign(K, 0);
		smoothed.assign(K, 0);
		mapBin.assign(K, 0);
		mapGrad.assign(K, 0);
		formantMetric.assign(K + 2, 0);
		resetStft();
		samplesSinceLast = NEVER; // blockProcess = {} (:89)
	}
	static double bessel0(double x) {
		double sum = 1, term = 1, q = x * x * 0.25;
		for (int k = 1; k < 200; ++k) {
			term *= q / (double(k) * k);
			sum += term;
			if (term < sum * 1e-17) break;
		}
		return sum;
	}
	// DynamicSTFT::reset(0.1) (called at :50,76,456), in linear "pending" form:
	// windowProducts[i] = 0.1*N*sum_{k>=0} w^2[i+kH] + 1e-30, then moveOutput(H).
	void resetStft() {
		std::fill(hist.begin(), hist.end(), 0.0f);
		std::fill(pend.begin(), pend.end(), 0.0f);
		for (int i = 0; i < pendLen; ++i) {
			int ring = i + H;
			if (ring < B) {
				float sum = 0;
				for (int k = ring; k < B; k += H) sum += window[k] * window[k];
				pendWp[i] = 0.1f * float(N) * sum + almostZero;
			} else {
				pendWp[i] = almostZero;
			}
		}
	}
	// signalsmith-stretch.h:49-60
	void reset() {
		resetStft();
		prevInputOffset = -1;
		std::fill(input.begin(), input.end(), cfloat(0));
		std::fill(prevInput.begin(), prevInput.end(), cfloat(0));
		std::fill(output.begin(), output.end(), cfloat(0));
		std::fill(inputEnergy.begin(), inputEnergy.end(), 0.0f); // Band() value-initialises
		silenceCounter = 0;
		didSeek = false;
		samplesSinceLast = NEVER;
		freqEstimateWeighted = freqEstimateWeight = 0;
	}
	int inputLatency() const { return B - B / 2; }                 // :42-44
	int outputLatency() const { return B / 2 + (split ? H : 0); }  // :45-47
	int seekLength() const { return B + H; }                       // :166-168
	int outputSeekLength(float rate) const { return int(inputLatency() + rate * outputLatency()); } // :205-207

	// :107-135
	void setTransposeFactor(float mult, float tonality) {
		freqMultiplier = mult;
		if (tonality > 0) freqTonalityLimit = tonality / std::sqrt(mult);
		else freqTonalityLimit = 1;
		customMap = 0;
	}
	void setTransposeSemitones(float semis, float tonality) { setTransposeFactor(std::pow(2, semis / 12), tonality); }
	void setFormantFactor(float mult, bool comp) {
		formantMultiplier = mult;
		invFormantMultiplier = 1 / mult;
		formantCompensation = comp;
	}
	void setFormantSemitones(float semis, bool comp) { setFormantFactor(std::pow(2, semis / 12), comp); }

	// ---------------- input history ----------------
	// The reference keeps a ring of B+H+1 samples that copyInput()/seek() feed (:215-229,:141-158);
	// only the last B+H samples are ever read, so a linear history is equivalent.
	void appendHistory(const float *in, int stride, int n) {
		for (int c = 0; c < C; ++c) {
			float *h = hist.data() + size_t(c) * histLen;
			const float *x = in + size_t(c) * stride;
			if (n >= histLen) {
				std::copy(x + n - histLen, x + n, h);
			} else {
				std::memmove(h, h + n, sizeof(float) * (histLen - n));
				std::copy(x, x + n, h + histLen - n);
			}
		}
	}
	// sample `i` of the stream "history ++ this call's input"; i<0 reaches into the history
	float streamSample(const float *in, int stride, int c, int i) const {
		if (i >= 0) return in[size_t(c) * stride + i];
		int h = histLen + i;
		return h >= 0 ? hist[size_t(c) * histLen + h] : 0.0f;
	}

	// :139-165
	void seek(const float *in, int n, double playbackRate) {
		int len = B + H;
		int startIndex = std::max(0, n - len);
		int padStart = len + startIndex - n;
		float totalEnergy = 0;
		std::vector<float> tmp(size_t(C) * len, 0.0f);
		for (int c = 0; c < C; ++c) {
			for (int i = startIndex; i < n; ++i) {
				float s = in[size_t(c) * n + i];
				totalEnergy += s * s;
				tmp[size_t(c) * len + (i - startIndex + padStart)] = s;
			}
		}
		appendHistory(tmp.data(), len, len);
		if (totalEnergy >= noiseFloor) {
			silenceCounter = 0;
			silenceFirst = true;
		}
		didSeek = true;
		seekTimeFactor = (playbackRate * H > 1) ? float(1 / playbackRate) : float(H);
	}

	// ---------------- plan: the block scheduler (:231-319, :406, :418-419) ----------------
	// Returns false when the call is a silence bypass (:240-271).
	bool plan(const float *in, int nIn, int nOut, std::vector<Frame> &frames) {
		frames.clear();
		float totalEnergy = 0;
		for (int c = 0; c < C; ++c)
			for (int i = 0; i < nIn; ++i) {
				float s = in[size_t(c) * nIn + i];
				totalEnergy += s * s;
			}
		if (totalEnergy < noiseFloor) {
			if (silenceCounter >= 2 * int64_t(B)) {
				if (silenceFirst) {
					silenceFirst = false;
					samplesSinceLast = NEVER; // blockProcess = {}
					std::fill(input.begin(), input.end(), cfloat(0));
					std::fill(prevInput.begin(), prevInput.end(), cfloat(0));
					std::fill(output.begin(), output.end(), cfloat(0));
					std::fill(inputEnergy.begin(), inputEnergy.end(), 0.0f);
				}
				return false;
			}
			silenceCounter += nIn;
		} else {
			silenceCounter = 0;
			silenceFirst = true;
		}
		for (int outputIndex = 0; outputIndex < nOut; ++outputIndex) {
			if (samplesSinceLast >= H) {
				samplesSinceLast = 0;
				Frame f;
				f.t = outputIndex;
				f.inputOffset = int(std::round(outputIndex * float(nIn) / nOut)); // :288 (float!)
				int inputInterval = f.inputOffset - prevInputOffset;
				prevInputOffset = f.inputOffset;
				f.newSpectrum = didSeek || inputInterval > 0;                             // :299
				f.mapped = customMap != 0 || freqMultiplier != 1;                         // :300
				f.reanalysePrev = f.newSpectrum && (didSeek || std::abs(inputInterval - H) > 1); // :303
				f.formants = formantMultiplier != 1 || (formantCompensation && f.mapped); // :310
				f.timeFactor = didSeek ? seekTimeFactor : float(H) / std::max<float>(1, float(inputInterval)); // :312
				didSeek = false;
				frames.push_back(f);
			}
			++samplesSinceLast; // :406
		}
		prevInputOffset -= nIn; // :419
		return true;
	}

	// ---------------- analysis (dependency analyseStep, :337,:359) ----------------
	void analyse(const float *in, int stride, int c, int endIndex, cfloat *spectrum) {
		std::vector<double> x(B);
		std::vector<cplx> X(K);
		for (int n = 0; n < B; ++n) x[n] = double(streamSample(in, stride, c, endIndex - B + n) * window[n]);
		fft.forward(x.data(), B, B / 2, X.data());
		for (int b = 0; b < K; ++b) spectrum[b] = cfloat(float(X[b].real()), float(X[b].imag()));
	}

	// fractional reads with zero outside [0,K) (:547-580)
	template <typename T>
	T getFractional(const T *arr, int lowIndex, float frac) const {
		T low = (lowIndex < 0 || lowIndex >= K) ? T(0) : arr[lowIndex];
		T high = (lowIndex + 1 < 0 || lowIndex + 1 >= K) ? T(0) : arr[lowIndex + 1];
		return low + (high - low) * frac;
	}
	template <typename T>
	T getFractional(const T *arr, float index) const {
		int low = int(std::floor(index));
		return getFractional(arr, low, index - low);
	}

	// :850-856
	float mapFreq(float freq) const {
		if (customMap == 1) return mapA * freq + mapB * freq * freq;
		if (customMap == 2) return pwl_map(mapTabIn.data(), mapTabOut.data(), (int)mapTabIn.size(), freq);
		if (freq > freqTonalityLimit) return freq + (freqMultiplier - 1) * freqTonalityLimit;
		return freq * freqMultiplier;
	}
	// :920-925
	float invMapFormant(float freq) const {
		if (freq * invFormantMultiplier > freqTonalityLimit) return freq + (1 - formantMultiplier) * freqTonalityLimit;
		return freq * invFormantMultiplier;
	}

	// :816-848 (all three steps)
	void smoothEnergy() {
		float smoothingBins = float(N) / H;
		float slew = 1 / (1 + smoothingBins * 0.5f);
		for (int b = 0; b < K; ++b) energy[b] = 0;
		for (int c = 0; c < C; ++c)
			for (int b = 0; b < K; ++b) {
				float e = cnorm(input[size_t(c) * K + b]);
				inputEnergy[size_t(c) * K + b] = e;
				energy[b] += e;
			}
		for (int b = 0; b < K; ++b) smoothed[b] = energy[b];
		float e = 0;
		for (int repeat = 0; repeat < 2; ++repeat) {
			for (int b = K - 1; b >= 0; --b) {
				e += (smoothed[b] - e) * slew;
				smoothed[b] = e;
			}
			for (int b = 0; b < K; ++b) {
				e += (smoothed[b] - e) * slew;
				smoothed[b] = e;
			}
		}
	}
	// :859-880
	void findPeaks() {
		peakIn.clear();
		peakOut.clear();
		int start = 0;
		while (start < K) {
			if (energy[start] > smoothed[start]) {
				int end = start;
				float bandSum = 0, energySum = 0;
				while (end < K && energy[end] > smoothed[end]) {
					bandSum += end * energy[end];
					energySum += energy[end];
					++end;
				}
				float avgBand = bandSum / energySum;
				float avgFreq = binToFreq(avgBand);
				peakIn.push_back(avgBand);
				peakOut.push_back(freqToBin(mapFreq(avgFreq)));
				start = end;
			}
			++start;
		}
	}
	// :882-917
	void updateOutputMap() {
		size_t P = peakIn.size();
		if (P == 0) {
			for (int b = 0; b < K; ++b) {
				mapBin[b] = float(b);
				mapGrad[b] = 1;
			}
			return;
		}
		float bottomOffset = peakIn[0] - peakOut[0];
		for (int b = 0; b < std::min<int>(K, int(std::ceil(peakOut[0]))); ++b) {
			mapBin[b] = b + bottomOffset;
			mapGrad[b] = 1;
		}
		for (size_t p = 1; p < P; ++p) {
			float prevIn = peakIn[p - 1], prevOut = peakOut[p - 1], nextIn = peakIn[p], nextOut = peakOut[p];
			float rangeScale = 1 / (nextOut - prevOut);
			float outOffset = prevIn - prevOut;
			float outScale = nextIn - nextOut - prevIn + prevOut;
			float gradScale = outScale * rangeScale;
			int startBin = std::max<int>(0, int(std::ceil(prevOut)));
			int endBin = std::min<int>(K, int(std::ceil(nextOut)));
			for (int b = startBin; b < endBin; ++b) {
				float r = (b - prevOut) * rangeScale;
				float h = r * r * (3 - 2 * r);
				float outB = b + outOffset + h * outScale;
				float gradH = 6 * r * (1 - r);
				float gradB = 1 + gradH * gradScale;
				mapBin[b] = outB;
				mapGrad[b] = gradB;
			}
		}
		float topOffset = peakIn[P - 1] - peakOut[P - 1];
		for (int b = std::max<int>(0, int(peakOut[P - 1])); b < K; ++b) {
			mapBin[b] = b + topOffset;
			mapGrad[b] = 1;
		}
	}
	// :929-966
	float estimateFrequency() {
		int p0 = 0, p1 = 0, p2 = 0;
		const std::vector<float> &m = formantMetric;
		for (int b = 1; b < K - 1; ++b) {
			float e = m[b];
			if (e < m[b - 1] || e <= m[b + 1]) continue;
			if (e > m[p0]) {
				if (e > m[p1]) {
					if (e > m[p2]) {
						p0 = p1;
						p1 = p2;
						p2 = b;
					} else {
						p0 = p1;
						p1 = b;
					}
				} else {
					p0 = b;
				}
			}
		}
		int peakEstimate = p2;
		if (m[p1] > m[p2] * 0.1) { // double constants, as in the reference
			int diff = std::abs(peakEstimate - p1);
			if (diff > peakEstimate / 8 && diff < peakEstimate * 7 / 8) peakEstimate = peakEstimate % diff;
			if (m[p0] > m[p2] * 0.01) {
				int diff2 = std::abs(peakEstimate - p0);
				if (diff2 > peakEstimate / 8 && diff2 < peakEstimate * 7 / 8) peakEstimate = peakEstimate % diff2;
			}
		}
		float weight = m[p2];
		freqEstimateWeighted += (peakEstimate * weight - freqEstimateWeighted) * 0.25;
		freqEstimateWeight += (weight - freqEstimateWeight) * 0.25;
		return freqEstimateWeighted / (freqEstimateWeight + 1e-30f);
	}
	// :972-1036 (all three steps)
	void updateFormants() {
		for (auto &e : formantMetric) e = 0;
		for (int c = 0; c < C; ++c)
			for (int b = 0; b < K; ++b) formantMetric[b] += inputEnergy[size_t(c) * K + b];
		freqEstimate = freqToBin(formantBaseFreq);
		if (formantBaseFreq <= 0) freqEstimate = estimateFrequency();

		float decay = float(1 - 1 / (freqEstimate * 0.5 + 1)); // :985 evaluates in double
		float e = 0;
		for (int repeat = 0; repeat < 2; ++repeat) {
			for (int b = K - 1; b >= 0; --b) {
				e = std::max(formantMetric[b], e * decay);
				formantMetric[b] = e;
			}
			for (int b = 0; b < K; ++b) {
				e = std::max(formantMetric[b], e * decay);
				formantMetric[b] = e;
			}
		}
		decay = 1 / decay;
		for (int repeat = 0; repeat < 2; ++repeat) {
			for (int b = K - 1; b >= 0; --b) {
				e = std::min(formantMetric[b], e * decay);
				formantMetric[b] = e;
			}
			for (int b = 0; b < K; ++b) {
				e = std::min(formantMetric[b], e * decay);
				formantMetric[b] = e;
			}
		}
		for (int b = 0; b < K; ++b) {
			float inputF = binToFreq(float(b));
			float outputF = formantCompensation ? mapFreq(inputF) : inputF;
			outputF = invMapFormant(outputF);
			float inputE = formantMetric[b];
			float band = freqToBin(outputF);
			float targetE;
			if (band < 0) {
				targetE = 0;
			} else {
				band = std::min<float>(band, float(K));
				int floorBand = int(std::floor(band));
				float fracBand = band - floorBand;
				float low = formantMetric[floorBand], high = formantMetric[floorBand + 1];
				targetE = low + (high - low) * fracBand;
			}
			float energyRatio = targetE / (inputE + 1e-30f);
			for (int c = 0; c < C; ++c) inputEnergy[size_t(c) * K + b] *= energyRatio;
		}
	}
	// Prediction::makeOutput, :596-603
	static cfloat makeOutput(cfloat phase, float energyValue, cfloat inputValue) {
		float phaseNorm = cnorm(phase);
		if (phaseNorm <= noiseFloor) {
			phase = inputValue;
			phaseNorm = cnorm(inputValue) + noiseFloor;
		}
		return phase * std::sqrt(energyValue / phaseNorm);
	}

	// ---------------- the spectral stage for one block: processSpectrum(), :633-813 ----------------
	void processSpectrum(const Frame &f) {
		float timeFactor = f.timeFactor;
		float smoothingBins = float(N) / H;
		int L = int(std::round(smoothingBins)); // longVerticalStep, :637
		timeFactor = std::max<float>(timeFactor, 1 / maxCleanStretch);
		bool randomTimeFactor = timeFactor > maxCleanStretch;
		std::uniform_real_distribution<float> timeFactorDist(maxCleanStretch * 2 * randomTimeFactor - timeFactor, timeFactor);

		if (f.newSpectrum) { // :642-660
			for (int c = 0; c < C; ++c)
				for (int b = 0; b < K; ++b) {
					output[size_t(c) * K + b] = cmul(output[size_t(c) * K + b], rot[b]);
					prevInput[size_t(c) * K + b] = cmul(prevInput[size_t(c) * K + b], rot[b]);
				}
		}
		if (f.mapped) { // :661-674
			smoothEnergy();
			findPeaks();
			updateOutputMap();
		} else { // :675-686
			for (size_t i = 0; i < size_t(C) * K; ++i) inputEnergy[i] = cnorm(input[i]);
			for (int b = 0; b < K; ++b) {
				mapBin[b] = float(b);
				mapGrad[b] = 1;
			}
		}
		if (f.formants) updateFormants(); // :689-695

		for (int c = 0; c < C; ++c) { // preliminary prediction, :696-719
			const cfloat *in = input.data() + size_t(c) * K, *prev = prevInput.data() + size_t(c) * K;
			const float *inE = inputEnergy.data() + size_t(c) * K;
			for (int b = 0; b < K; ++b) {
				int lowIndex = int(std::floor(mapBin[b]));
				float fracIndex = mapBin[b] - lowIndex;
				float prevEnergy = predEnergy[size_t(c) * K + b];
				float e = getFractional(inE, lowIndex, fracIndex);
				e *= std::max<float>(0, mapGrad[b]);
				predEnergy[size_t(c) * K + b] = e;
				cfloat pin = getFractional(in, lowIndex, fracIndex);
				predInput[size_t(c) * K + b] = pin;
				cfloat pprev = getFractional(prev, lowIndex, fracIndex);
				cfloat freqTwist = cmulConj(pin, pprev);
				cfloat phase = cmul(output[size_t(c) * K + b], freqTwist);
				output[size_t(c) * K + b] = phase / (std::max(prevEnergy, e) + noiseFloor);
			}
		}
		for (int b = 0; b < K; ++b) { // main prediction, :722-804 (serial over b)
			int maxChannel = 0;
			float maxEnergy = predEnergy[b];
			for (int c = 1; c < C; ++c) {
				float e = predEnergy[size_t(c) * K + b];
				if (e > maxEnergy) {
					maxChannel = c;
					maxEnergy = e;
				}
			}
			const cfloat *in = input.data() + size_t(maxChannel) * K;
			cfloat *out = output.data() + size_t(maxChannel) * K;
			const cfloat *pin = predInput.data() + size_t(maxChannel) * K;
			cfloat phase = 0;
			if (b > 0) {
				float binTimeFactor = randomTimeFactor ? timeFactorDist(randomEngine) : timeFactor;
				cfloat downInput = getFractional(in, mapBin[b] - binTimeFactor);
				cfloat shortTwist = cmulConj(pin[b], downInput);
				phase += cmul(out[b - 1], shortTwist);
				if (b >= L) {
					cfloat longDownInput = getFractional(in, mapBin[b] - L * binTimeFactor);
					cfloat longTwist = cmulConj(pin[b], longDownInput);
					phase += cmul(out[b - L], longTwist);
				}
			}
			if (b < K - 1) {
				float binTimeFactor = randomTimeFactor ? timeFactorDist(randomEngine) : timeFactor;
				cfloat downInput = getFractional(in, mapBin[b + 1] - binTimeFactor);
				cfloat shortTwist = cmulConj(pin[b + 1], downInput);
				phase += cmulConj(out[b + 1], shortTwist);
				if (b < K - L) {
					cfloat longDownInput = getFractional(in, mapBin[b + L] - L * binTimeFactor);
					cfloat longTwist = cmulConj(pin[b + L], longDownInput);
					phase += cmulConj(out[b + L], longTwist);
				}
			}
			out[b] = makeOutput(phase, predEnergy[size_t(maxChannel) * K + b], pin[b]);
			for (int c = 0; c < C; ++c) {
				if (c == maxChannel) continue;
				cfloat cin = predInput[size_t(c) * K + b];
				cfloat channelTwist = cmulConj(cin, pin[b]);
				cfloat channelPhase = cmul(out[b], channelTwist);
				output[size_t(c) * K + b] = makeOutput(channelPhase, predEnergy[size_t(c) * K + b], cin);
			}
		}
		if (f.newSpectrum) prevInput = input; // :806-812
	}

	// ---------------- synthesis + overlap-add (dependency synthesiseStep, :397-399) ----------------
	void synthesise() {
		for (int i = 0; i < B; ++i) pendWp[addOff + i] += winProd[i];
		std::vector<cplx> Y(K);
		std::vector<double> y(B);
		for (int c = 0; c < C; ++c) {
			for (int b = 0; b < K; ++b) Y[b] = cplx(output[size_t(c) * K + b].real(), output[size_t(c) * K + b].imag());
			fft.inverse(Y.data(), y.data(), B, B / 2);
			float *p = pend.data() + size_t(c) * pendLen + addOff;
			for (int i = 0; i < B; ++i) p[i] += float(y[i]) * window[i];
		}
	}
	// readOutput(c,1)+moveOutput(1) for n samples (:408-414): emit pending/windowProducts, shift in zeros
	void emit(float *out, int stride, int at, int n) {
		if (n <= 0) return;
		for (int c = 0; c < C; ++c) {
			float *p = pend.data() + size_t(c) * pendLen;
			for (int i = 0; i < n; ++i) out[size_t(c) * stride + at + i] = (i < pendLen) ? p[i] / pendWp[i] : 0.0f / almostZero;
			if (n >= pendLen) {
				std::fill(p, p + pendLen, 0.0f);
			} else {
				std::memmove(p, p + n, sizeof(float) * (pendLen - n));
				std::fill(p + pendLen - n, p + pendLen, 0.0f);
			}
		}
		if (n >= pendLen) {
			std::fill(pendWp.begin(), pendWp.end(), almostZero);
		} else {
			std::memmove(pendWp.data(), pendWp.data() + n, sizeof(float) * (pendLen - n));
			std::fill(pendWp.begin() + (pendLen - n), pendWp.end(), almostZero);
		}
	}

	// ---------------- process(), :209-423 ----------------
	void process(const float *in, int nIn, float *out, int nOut) {
		std::vector<Frame> frames;
		if (!plan(in, nIn, nOut, frames)) {
			// silence bypass :252-271
			for (int c = 0; c < C; ++c)
				for (int i = 0; i < nOut; ++i) out[size_t(c) * nOut + i] = nIn > 0 ? in[size_t(c) * nIn + (i % nIn)] : 0.0f;
			appendHistory(in, nIn, nIn);
			return;
		}
		int emitted = 0;
		for (const Frame &f : frames) {
			emit(out, nOut, emitted, f.t - emitted);
			emitted = f.t;
			if (f.newSpectrum) {
				if (f.reanalysePrev)
					for (int c = 0; c < C; ++c) analyse(in, nIn, c, f.inputOffset - H, prevInput.data() + size_t(c) * K);
				for (int c = 0; c < C; ++c) analyse(in, nIn, c, f.inputOffset, input.data() + size_t(c) * K);
			}
			processSpectrum(f);
			synthesise();
		}
		emit(out, nOut, emitted, nOut - emitted);
		appendHistory(in, nIn, nIn);
	}

	// ---------------- flush(), :426-464 ----------------
	void flush(float *out, int nOut, float playbackRate) {
		int outputBlock = std::max(0, nOut - H);
		if (outputBlock > 0) {
			int nIn = int(outputBlock * playbackRate);
			std::vector<float> zeros(size_t(C) * std::max(nIn, 1), 0.0f), tmp(size_t(C) * outputBlock);
			process(zeros.data(), nIn, tmp.data(), outputBlock);
			for (int c = 0; c < C; ++c) std::copy(tmp.begin() + size_t(c) * outputBlock, tmp.begin() + size_t(c + 1) * outputBlock, out + size_t(c) * nOut);
		}
		int tail = nOut - outputBlock;
		// in split mode stft.output sits one interval ahead of the samples still being read from
		// the stashed copy (:294-297); at a block boundary that is exactly our pending index 0.
		int base = split ? int(H - std::min<int64_t>(samplesSinceLast, H)) : 0;
		// finishOutput(1): running max over the ring from output.pos
		std::vector<float> wp(B);
		float mx = 0;
		for (int i = 0; i < B; ++i) {
			float v = (base + i < pendLen) ? pendWp[base + i] : almostZero;
			mx = std::max(mx, v);
			v += (mx - v) * 1.0f;
			wp[i] = v;
		}
		for (int c = 0; c < C; ++c) {
			const float *p = pend.data() + size_t(c) * pendLen;
			auto rd = [&](int i) { int k = i % B; return ((base + k < pendLen) ? p[base + k] : 0.0f) / wp[k]; };
			for (int i = 0; i < tail; ++i) out[size_t(c) * nOut + outputBlock + i] = rd(i);
			for (int i = 0; i < tail; ++i) out[size_t(c) * nOut + outputBlock + tail - 1 - i] -= rd(tail + i);
		}
		resetStft();
		std::fill(prevInput.begin(), prevInput.end(), cfloat(0));
		std::fill(output.begin(), output.end(), cfloat(0));
	}

	// ---------------- outputSeek(), :172-204 ----------------
	void outputSeek(const float *in, int stride, int inputLength) {
		reset();
		int surplus = std::max(inputLength - inputLatency(), 0);
		float playbackRate = surplus / float(outputLatency());
		int seekSamples = inputLength - surplus;
		std::vector<float> head(size_t(C) * std::max(seekSamples, 1)), rest(size_t(C) * std::max(surplus, 1));
		for (int c = 0; c < C; ++c) {
			std::copy(in + size_t(c) * stride, in + size_t(c) * stride + seekSamples, head.begin() + size_t(c) * seekSamples);
			std::copy(in + size_t(c) * stride + seekSamples, in + size_t(c) * stride + inputLength, rest.begin() + size_t(c) * surplus);
		}
		seek(head.data(), seekSamples, playbackRate);
		int len = outputLatency();
		std::vector<float> pre(size_t(C) * len);
		process(rest.data(), surplus, pre.data(), len);
		// negate, reverse, addOutput (dependency addOutput [recalled]: buffer += v*windowProducts)
		for (int c = 0; c < C; ++c) {
			float *p = pend.data() + size_t(c) * pendLen;
			int base = split ? int(H - std::min<int64_t>(samplesSinceLast, H)) : 0;
			for (int i = 0; i < len && i < B; ++i) {
				float v = -pre[size_t(c) * len + (len - 1 - i)];
				if (base + i < pendLen) p[base + i] += v * pendWp[base + i];
			}
		}
	}
	// ---------------- exact(), :467-491 ----------------
	bool exact(const float *in, int nIn, float *out, int nOut) {
		float playbackRate = nIn / float(nOut);
		int seekLen = outputSeekLength(playbackRate);
		if (nIn < seekLen) {
			std::fill(out, out + size_t(C) * nOut, 0.0f);
			return false;
		}
		outputSeek(in, nIn, seekLen);
		int outputIndex = int(nOut - seekLen / playbackRate);
		int restIn = nIn - seekLen;
		std::vector<float> rest(size_t(C) * std::max(restIn, 1)), body(size_t(C) * std::max(outputIndex, 1));
		for (int c = 0; c < C; ++c) std::copy(in + size_t(c) * nIn + seekLen, in + size_t(c + 1) * nIn, rest.begin() + size_t(c) * restIn);
		process(rest.data(), restIn, body.data(), outputIndex);
		int tailN = nOut - outputIndex;
		std::vector<float> tail(size_t(C) * std::max(tailN, 1));
		flush(tail.data(), tailN, playbackRate);
		for (int c = 0; c < C; ++c) {
			std::copy(body.begin() + size_t(c) * outputIndex, body.begin() + size_t(c + 1) * outputIndex, out + size_t(c) * nOut);
			std::copy(tail.begin() + size_t(c) * tailN, tail.begin() + size_t(c + 1) * tailN, out + size_t(c) * nOut + outputIndex);
		}
		return true;
	}
};

} // namespace oracle

using oracle::Stretch;

extern "C" {
void *orc_new(long seed) { return new Stretch(seed); }
void orc_free(void *h) { delete (Stretch *)h; }
void orc_configure(void *h, int ch, int block, int interval, int split) { ((Stretch *)h)->configure(ch, block, interval, split != 0); }
// presets: :63-68 (float products truncated by the int parameters of configure)
void orc_preset_default(void *h, int ch, float sr, int split) { ((Stretch *)h)->configure(ch, int(sr * 0.12), int(sr * 0.03), split != 0); }
void orc_preset_cheaper(void *h, int ch, float sr, int split) { ((Stretch *)h)->configure(ch, int(sr * 0.1), int(sr * 0.04), split != 0); }
void orc_reset(void *h) { ((Stretch *)h)->reset(); }
int orc_block_samples(void *h) { return ((Stretch *)h)->B; }
int orc_interval_samples(void *h) { return ((Stretch *)h)->H; }
int orc_input_latency(void *h) { return ((Stretch *)h)->inputLatency(); }
int orc_output_latency(void *h) { return ((Stretch *)h)->outputLatency(); }
int orc_split_computation(void *h) { return ((Stretch *)h)->split; }
int orc_seek_length(void *h) { return ((Stretch *)h)->seekLength(); }
int orc_output_seek_length(void *h, float rate) { return ((Stretch *)h)->outputSeekLength(rate); }
int orc_bands(void *h) { return ((Stretch *)h)->K; }
int orc_fft_samples(void *h) { return ((Stretch *)h)->N; }
void orc_set_transpose_factor(void *h, float m, float t) { ((Stretch *)h)->setTransposeFactor(m, t); }
void orc_set_transpose_semitones(void *h, float s, float t) { ((Stretch *)h)->setTransposeSemitones(s, t); }
void orc_set_formant_factor(void *h, float m, int comp) { ((Stretch *)h)->setFormantFactor(m, comp != 0); }
void orc_set_formant_semitones(void *h, float s, int comp) { ((Stretch *)h)->setFormantSemitones(s, comp != 0); }
void orc_set_formant_base(void *h, float f) { ((Stretch *)h)->formantBaseFreq = f; }
void orc_set_freq_map_quadratic(void *h, float a, float b) {
	Stretch &s = *(Stretch *)h;
	s.customMap = 1;
	s.mapA = a;
	s.mapB = b;
}
void orc_set_freq_map_table(void *h, const float *fin, const float *fout, int n) {
	Stretch &s = *(Stretch *)h;
	s.customMap = n > 0 ? 2 : 0;
	s.mapTabIn.assign(fin, fin + n);
	s.mapTabOut.assign(fout, fout + n);
}
void orc_seek(void *h, const float *in, int n, double rate) { ((Stretch *)h)->seek(in, n, rate); }
void orc_output_seek(void *h, const float *in, int n) { ((Stretch *)h)->outputSeek(in, n, n); }
void orc_process(void *h, const float *in, int nIn, float *out, int nOut) { ((Stretch *)h)->process(in, nIn, out, nOut); }
void orc_flush(void *h, float *out, int nOut, float rate) { ((Stretch *)h)->flush(out, nOut, rate); }
int orc_exact(void *h, const float *in, int nIn, float *out, int nOut) { return ((Stretch *)h)->exact(in, nIn, out, nOut) ? 1 : 0; }

// white-box state, same numbering as oracle/ref_header_shim.cpp
int orc_get_state(void *h, int what, float *dst) {
	Stretch &s = *(Stretch *)h;
	int K = s.K, C = s.C, n = 0;
	auto putc = [&](const std::vector<oracle::cfloat> &v) {
		for (int i = 0; i < K * C; ++i) { dst[n++] = v[i].real(); dst[n++] = v[i].imag(); }
	};
	switch (what) {
	case 0: putc(s.input); break;
	case 1: putc(s.prevInput); break;
	case 2: putc(s.output); break;
	case 3: for (int i = 0; i < K * C; ++i) dst[n++] = s.inputEnergy[i]; break;
	case 4: for (int i = 0; i < K * C; ++i) dst[n++] = s.predEnergy[i]; break;
	case 5: for (int i = 0; i < K; ++i) { dst[n++] = s.mapBin[i]; dst[n++] = s.mapGrad[i]; } break;
	case 6: for (int i = 0; i < K; ++i) dst[n++] = s.energy[i]; break;
	case 7: for (int i = 0; i < K; ++i) dst[n++] = s.smoothed[i]; break;
	case 8: for (float v : s.window) dst[n++] = v; break;
	case 9: { // windowProducts in ring order relative to output.pos is implementation detail; export pending order
		for (int i = 0; i < s.B; ++i) dst[n++] = s.pendWp[i];
		break;
	}
	case 10: putc(s.predInput); break;
	case 11: for (float v : s.formantMetric) dst[n++] = v; break;
	// 20..22: the linear signal buffers, same layout as the CUDA engine's state (teacher forcing)
	case 20: for (float v : s.hist) dst[n++] = v; break;
	case 21: for (float v : s.pend) dst[n++] = v; break;
	case 22: for (int c = 0; c < C; ++c) for (float v : s.pendWp) dst[n++] = v; break;
	default: return -1;
	}
	return n;
}
int orc_state_size(void *h, int what) {
	Stretch &s = *(Stretch *)h;
	switch (what) {
	case 20: return s.C * s.histLen;
	case 21: return s.C * s.pendLen;
	case 22: return s.C * s.pendLen;
	}
	return -1;
}
int orc_num_peaks(void *h) { return (int)((Stretch *)h)->peakIn.size(); }
int orc_get_peaks(void *h, float *dst) {
	Stretch &s = *(Stretch *)h;
	int n = 0;
	for (size_t i = 0; i < s.peakIn.size(); ++i) { dst[n++] = s.peakIn[i]; dst[n++] = s.peakOut[i]; }
	return n;
}
}
