// oracle/fft_ref.h -- TEST INFRASTRUCTURE ONLY (CPU oracle), never linked into the product.
//
// Plain double-precision mixed-radix complex FFT (factors 2, 3, 5) plus the "modified"
// (half-bin shifted) real transform pair that the reference obtains from its un-vendored
// dependency signalsmith-linear 0.2.6 (`DynamicSTFT<Sample,false,true>`, used at
// /root/reference/signalsmith-stretch.h:337,359,398).  The transform convention was pinned
// against the reference's shipped WASM binary:
//
//   forward:  X[b] = sum_n x[n] * exp(-2*pi*i*(b+1/2)*(n-o)/N),  b in [0,N/2),  o = offset
//   inverse:  y[n] = sum_b 2*Re( Y[b] * exp(+2*pi*i*(b+1/2)*(n-o)/N) )   (unnormalised)
//
// Everything is computed in double and rounded once to float by the caller, so this is the
// "mathematically exact, rounded to float" answer any float FFT is compared against.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

namespace oracle {

typedef std::complex<double> cplx;

class FFT {
public:
	explicit FFT(int n = 0) { resize(n); }
	void resize(int n) {
		size = n;
		tw.resize(n > 0 ? n : 0);
		for (int k = 0; k < n; ++k) tw[k] = std::polar(1.0, -2.0 * M_PI * k / n);
		tmp.resize(n > 0 ? n : 0);
	}
	// out[k] = sum_n in[n] * exp(sign * 2*pi*i*n*k/size), sign=-1 forward, +1 inverse (unnormalised)
	void run(const cplx *in, cplx *out, bool inverse) {
		rec(size, in, 1, out, inverse);
	}
	int size = 0;

private:
	std::vector<cplx> tw, tmp;
	cplx twiddle(int n, long idx, bool inverse) const {
		cplx w = tw[(idx % n) * (size / n)];
		return inverse ? std::conj(w) : w;
	}
	void rec(int n, const cplx *in, int stride, cplx *out, bool inverse) {
		if (n == 1) {
			out[0] = in[0];
			return;
		}
		int r = (n % 2 == 0) ? 2 : (n % 3 == 0) ? 3 : (n % 5 == 0) ? 5 : n;
		int m = n / r;
		if (r == n && n > 5) { // generic O(n^2) fallback (never used for 2^k*{1,3,5})
			std::vector<cplx> t(n);
			for (int k = 0; k < n; ++k) {
				cplx s = 0;
				for (int j = 0; j < n; ++j) s += in[j * stride] * twiddle(n, (long)j * k, inverse);
				t[k] = s;
			}
			for (int k = 0; k < n; ++k) out[k] = t[k];
			return;
		}
		for (int q = 0; q < r; ++q) rec(m, in + q * stride, stride * r, out + q * m, inverse);
		cplx y[5], z[5];
		for (int k = 0; k < m; ++k) {
			for (int q = 0; q < r; ++q) y[q] = out[q * m + k] * twiddle(n, (long)q * k, inverse);
			for (int j = 0; j < r; ++j) {
				cplx s = y[0];
				for (int q = 1; q < r; ++q) s += y[q] * twiddle(r, (long)q * j, inverse);
				z[j] = s;
			}
			for (int j = 0; j < r; ++j) out[j * m + k] = z[j];
		}
	}
};

// The modified real FFT pair, via one complex FFT of size N/2.
class ModifiedRealFFT {
public:
	void resize(int fftSamples) {
		N = fftSamples;
		half = N / 2;
		fft.resize(half);
		z.resize(half);
		Z.resize(half);
		t.resize(N);
		pre.resize(half);
		for (int n = 0; n < half; ++n) pre[n] = std::polar(1.0, -M_PI * n / N);
	}
	// x: `len` (<= N) real samples (already windowed), offset o: the sample x[o] sits at time 0.
	// spectrum: N/2 complex bins.
	void forward(const double *x, int len, int o, cplx *spectrum) {
		for (int n = 0; n < N; ++n) t[n] = 0;
		for (int n = 0; n < len; ++n) {
			int m = n - o;
			if (m >= 0) t[m] = x[n];
			else t[m + N] = -x[n]; // a shift by N flips the sign of the half-bin kernel
		}
		for (int n = 0; n < half; ++n) z[n] = cplx(t[n], -t[n + half]) * pre[n];
		fft.run(z.data(), Z.data(), false);
		for (int k = 0; k < half / 2; ++k) {
			spectrum[2 * k] = Z[k];
			spectrum[2 * k + 1] = std::conj(Z[half - 1 - k]);
		}
	}
	void inverse(const cplx *spectrum, double *y, int len, int o) {
		for (int k = 0; k < half / 2; ++k) {
			Z[k] = spectrum[2 * k];
			Z[half - 1 - k] = std::conj(spectrum[2 * k + 1]);
		}
		fft.run(Z.data(), z.data(), true);
		for (int n = 0; n < half; ++n) {
			cplx v = z[n] * std::conj(pre[n]);
			t[n] = 2 * v.real();
			t[n + half] = -2 * v.imag();
		}
		for (int n = 0; n < len; ++n) {
			int m = n - o;
			y[n] = (m >= 0) ? t[m] : -t[m + N];
		}
	}
	int N = 0, half = 0;

private:
	FFT fft;
	std::vector<cplx> z, Z, pre;
	std::vector<double> t;
};

// complex FFT size chosen by the dependency for a block: 2^k * {1,3,5} (SURVEY.md App. B,
// measured on 46/46 block sizes of the reference binary).
inline int fastSizeAbove(int n) {
	int p = 1;
	while (p < 16 && p < n) p *= 2;
	while (8 * p < n) p *= 2;
	int m = (n + p - 1) / p;
	if (m == 7) m = 8;
	return m * p;
}

} // namespace oracle
