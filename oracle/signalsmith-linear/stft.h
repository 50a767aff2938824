// oracle/signalsmith-linear/stft.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Stand-in for the reference's un-vendored dependency Signalsmith-Audio/linear @ 0.2.6
// (`#include "signalsmith-linear/stft.h"`, /root/reference/signalsmith-stretch.h:4, pinned by
// /root/reference/CMakeLists.txt:6-13; absent from the container, no network).  It restates ONLY
// the members of `signalsmith::linear::DynamicSTFT<Sample,false,true>` that the reference header
// calls (SURVEY.md Appendix B lists every call site), so that the UNMODIFIED reference header
// compiles with `-I oracle -I /root/reference`.  Semantics were inferred from those call sites and
// pinned numerically against the reference's shipped WASM binary (tests/test_oracle_pinning.py).
// It is written from the contract, not from the dependency's source (which is not available).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <vector>

#include "../fft_ref.h"

namespace signalsmith { namespace linear {

template <typename Sample, bool splitComputation = false, bool modified = false>
class DynamicSTFT {
	static_assert(!splitComputation && modified, "oracle stand-in only covers DynamicSTFT<Sample,false,true>");

public:
	typedef std::complex<Sample> Complex;
	enum WindowShape { ignore, acg, kaiser };

	struct Input {
		size_t pos = 0;
		std::vector<Sample> buffer; // [channel][ringLength]
		void swap(Input &o) {
			std::swap(pos, o.pos);
			buffer.swap(o.buffer);
		}
	};
	struct Output {
		size_t pos = 0;
		std::vector<Sample> buffer;         // [channel][block]
		std::vector<Sample> windowProducts; // [block]
		void swap(Output &o) {
			std::swap(pos, o.pos);
			buffer.swap(o.buffer);
			windowProducts.swap(o.windowProducts);
		}
	};
	Input input;
	Output output;

	// reference call site: signalsmith-stretch.h:74
	void configure(size_t inChannels, size_t outChannels, size_t blockSamples, size_t extraInputHistory = 0, size_t intervalSamples = 0) {
		_inCh = inChannels;
		_outCh = outChannels;
		_block = blockSamples;
		_fft = 2 * (size_t)oracle::fastSizeAbove(int((blockSamples + 1) / 2));
		_bands = _fft / 2;
		_inLen = blockSamples + extraInputHistory;
		real.resize(int(_fft));
		input.buffer.assign(_inLen * _inCh, 0);
		output.buffer.assign(_block * _outCh, 0);
		output.windowProducts.assign(_block, 0);
		spectrumBuffer.assign(_bands * std::max(_inCh, _outCh), 0);
		timeD.resize(_block);
		specD.resize(_bands);
		_analysisWindow.assign(_block, 0);
		_synthesisWindow.assign(_block, 0);
		_analysisOffset = _synthesisOffset = _block / 2;
		if (intervalSamples) setInterval(intervalSamples, kaiser);
	}
	// signalsmith-stretch.h:75 -- Kaiser window, heuristic bandwidth, forced perfect reconstruction
	void setInterval(size_t interval, WindowShape shape = ignore) {
		_interval = interval;
		if (shape == ignore) return;
		double bw = double(_block) / double(interval);
		bw += 8 / ((bw + 3) * (bw + 3)) + 0.25 * std::max(3 - bw, 0.0);
		bw = std::max(bw, 2.0);
		double beta = M_PI * std::sqrt(bw * bw * 0.25 - 1);
		std::vector<double> w(_block);
		double invI0b = 1 / bessel0(beta);
		for (size_t i = 0; i < _block; ++i) {
			double r = (2.0 * i + 1) / _block - 1;
			w[i] = bessel0(beta * std::sqrt(std::max(0.0, 1 - r * r))) * invI0b;
		}
		for (size_t i = 0; i < interval; ++i) { // force perfect reconstruction
			double sum2 = 0;
			for (size_t k = i; k < _block; k += interval) sum2 += w[k] * w[k];
			double f = 1 / std::sqrt(sum2);
			for (size_t k = i; k < _block; k += interval) w[k] *= f;
		}
		for (size_t i = 0; i < _block; ++i) _analysisWindow[i] = _synthesisWindow[i] = Sample(w[i]);
	}
	// signalsmith-stretch.h:50,76,456
	void reset(Sample productWeight = 1) {
		input.pos = _block;
		output.pos = 0;
		std::fill(input.buffer.begin(), input.buffer.end(), Sample(0));
		std::fill(output.buffer.begin(), output.buffer.end(), Sample(0));
		std::fill(spectrumBuffer.begin(), spectrumBuffer.end(), Complex(0));
		for (size_t i = 0; i < _block; ++i) {
			Sample sum = 0;
			for (size_t k = i; k < _block; k += _interval) sum += _analysisWindow[k] * _synthesisWindow[k];
			output.windowProducts[i] = productWeight * Sample(_fft) * sum + almostZero;
		}
		moveOutput(_interval);
	}

	size_t blockSamples() const { return _block; }
	size_t fftSamples() const { return _fft; }
	size_t defaultInterval() const { return _interval; }
	size_t bands() const { return _bands; }
	size_t analysisLatency() const { return _block - _analysisOffset; }
	size_t synthesisLatency() const { return _synthesisOffset; }
	Sample binToFreq(Sample b) const { return (b + Sample(0.5)) / Sample(_fft); }
	Sample freqToBin(Sample f) const { return f * Sample(_fft) - Sample(0.5); }

	// signalsmith-stretch.h:156,225 / :158,227
	void writeInput(size_t channel, size_t length, const Sample *data) {
		Sample *ring = input.buffer.data() + channel * _inLen;
		for (size_t i = 0; i < length; ++i) ring[(input.pos + i) % _inLen] = data[i];
	}
	void moveInput(size_t samples) { input.pos = (input.pos + samples) % _inLen; }

	size_t analyseSteps() const { return _inCh; }
	// signalsmith-stretch.h:337,359
	void analyseStep(size_t channel, size_t samplesInPast = 0) {
		const Sample *ring = input.buffer.data() + channel * _inLen;
		size_t start = (input.pos + 2 * _inLen - samplesInPast - _block) % _inLen;
		for (size_t n = 0; n < _block; ++n) timeD[n] = double(ring[(start + n) % _inLen] * _analysisWindow[n]);
		real.forward(timeD.data(), int(_block), int(_analysisOffset), specD.data());
		Complex *s = spectrum(channel);
		for (size_t b = 0; b < _bands; ++b) s[b] = Complex(Sample(specD[b].real()), Sample(specD[b].imag()));
	}
	Complex *spectrum(size_t channel) { return spectrumBuffer.data() + channel * _bands; }

	size_t synthesiseSteps() const { return _outCh + 1; }
	// signalsmith-stretch.h:397-398
	void synthesiseStep(size_t step) {
		if (step == 0) {
			for (size_t i = 0; i < _block; ++i) {
				output.windowProducts[(output.pos + i) % _block] += _analysisWindow[i] * _synthesisWindow[i] * Sample(_fft);
			}
			return;
		}
		size_t channel = step - 1;
		const Complex *s = spectrum(channel);
		for (size_t b = 0; b < _bands; ++b) specD[b] = oracle::cplx(s[b].real(), s[b].imag());
		real.inverse(specD.data(), timeD.data(), int(_block), int(_synthesisOffset));
		Sample *ring = output.buffer.data() + channel * _block;
		for (size_t i = 0; i < _block; ++i) ring[(output.pos + i) % _block] += Sample(timeD[i]) * _synthesisWindow[i];
	}

	// signalsmith-stretch.h:411,446 / :451
	void readOutput(size_t channel, size_t length, Sample *data) { readOutput(channel, 0, length, data); }
	void readOutput(size_t channel, size_t offset, size_t length, Sample *data) {
		const Sample *ring = output.buffer.data() + channel * _block;
		for (size_t i = 0; i < length; ++i) {
			size_t p = (output.pos + offset + i) % _block;
			data[i] = ring[p] / output.windowProducts[p];
		}
	}
	// signalsmith-stretch.h:296,414
	void moveOutput(size_t samples) {
		for (size_t i = 0; i < samples; ++i) {
			size_t p = (output.pos + i) % _block;
			for (size_t c = 0; c < _outCh; ++c) output.buffer[c * _block + p] = 0;
			output.windowProducts[p] = almostZero;
		}
		output.pos = (output.pos + samples) % _block;
	}
	// signalsmith-stretch.h:444
	void finishOutput(Sample strength = 1) {
		Sample maxProduct = 0;
		for (size_t i = 0; i < _block; ++i) {
			size_t p = (output.pos + i) % _block;
			Sample &wp = output.windowProducts[p];
			maxProduct = std::max(maxProduct, wp);
			wp += (maxProduct - wp) * strength;
		}
	}
	// signalsmith-stretch.h:202 -- [recalled, unverifiable: not exported by the WASM build]
	void addOutput(size_t channel, size_t length, const Sample *data) {
		Sample *ring = output.buffer.data() + channel * _block;
		for (size_t i = 0; i < length; ++i) {
			size_t p = (output.pos + i) % _block;
			ring[p] += data[i] * output.windowProducts[p];
		}
	}

	const std::vector<Sample> &analysisWindow() const { return _analysisWindow; }

private:
	static constexpr Sample almostZero = Sample(1e-30);
	static double bessel0(double x) {
		double sum = 1, term = 1, q = x * x * 0.25;
		for (int k = 1; k < 200; ++k) {
			term *= q / (double(k) * k);
			sum += term;
			if (term < sum * 1e-17) break;
		}
		return sum;
	}
	size_t _inCh = 0, _outCh = 0, _block = 0, _fft = 0, _bands = 0, _inLen = 0, _interval = 1;
	size_t _analysisOffset = 0, _synthesisOffset = 0;
	std::vector<Sample> _analysisWindow, _synthesisWindow;
	std::vector<Complex> spectrumBuffer;
	oracle::ModifiedRealFFT real;
	std::vector<double> timeD;
	std::vector<oracle::cplx> specD;
};

}} // namespace
