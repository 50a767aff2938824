// fft.cuh -- shared-memory Stockham FFT for the K = N/2 point complex transform behind the
// reference's "modified" (half-bin shifted) real FFT (dependency DynamicSTFT::analyseStep /
// synthesiseStep; reference call sites signalsmith-stretch.h:337,359,398; convention pinned in
// SURVEY.md section 8(a) row 6 / App. F).
//
// Design (B200): the whole working set lives in shared memory (two ping-pong buffers of K float2,
// padded by one element every 16 so that the strided stores of the early passes are bank-conflict
// free); every pass is a LARGE radix (16, 12, 10, 8, ...) done entirely in registers by one thread
// per butterfly, so K = 3072 = 16*16*12 needs only THREE shared-memory round trips and three
// barriers.  The in-register small DFTs are composed at compile time from radix-2/3/4/5 kernels
// with constant twiddles (trivial ones vanish); the inter-pass twiddles come from one exact table
// (double-precision values rounded once): w^1, w^2, w^4, w^8 are loaded, the other powers are
// products of at most three of them.  No cuFFT, no tensor cores.
// Sizes are always 2^a * {1,3,5} (fastSizeAbove); the radix carrying the odd
// factor is the LAST pass so every pass's sub-transform length Ns is a power of two.
#pragma once
#include <utility>

#include "common.cuh"

namespace b200s {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { // FMA allowed: FFT arithmetic only
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulcf(float2 a, float2 b) { // a * conj(b)
	return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
	return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// index into a padded shared-memory FFT buffer
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int fft_buf_len(int K) { return K + (K >> 4) + 1; }

// ---- compile-time twiddle constants: exp(-2*pi*i*M/R) ----
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;
__host__ __device__ constexpr double sin_small(double x) { // |x| <= pi/4
	double x2 = x * x, term = x, sum = x;
	for (int k = 1; k < 14; ++k) {
		term *= -x2 / double((2 * k) * (2 * k + 1));
		sum += term;
	}
	return sum;
}
__host__ __device__ constexpr double cos_small(double x) {
	double x2 = x * x, term = 1, sum = 1;
	for (int k = 1; k < 14; ++k) {
		term *= -x2 / double((2 * k - 1) * (2 * k));
		sum += term;
	}
	return sum;
}
// cos/sin of 2*pi*m/r with exact quadrant handling
__host__ __device__ constexpr double cosq(int m, int r) {
	m = ((m % r) + r) % r;
	int quad = (4 * m) / r;          // 0..3
	int rem4 = 4 * m - quad * r;     // angle within the quadrant = (pi/2) * rem4 / r
	double phi = (kPi / 2) * double(rem4) / double(r);
	double c = (2 * rem4 <= r) ? cos_small(phi) : sin_small(kPi / 2 - phi);
	double s = (2 * rem4 <= r) ? sin_small(phi) : cos_small(kPi / 2 - phi);
	return quad == 0 ? c : quad == 1 ? -s : quad == 2 ? -c : s;
}
__host__ __device__ constexpr double sinq(int m, int r) {
	m = ((m % r) + r) % r;
	int quad = (4 * m) / r;
	int rem4 = 4 * m - quad * r;
	double phi = (kPi / 2) * double(rem4) / double(r);
	double c = (2 * rem4 <= r) ? cos_small(phi) : sin_small(kPi / 2 - phi);
	double s = (2 * rem4 <= r) ? sin_small(phi) : cos_small(kPi / 2 - phi);
	return quad == 0 ? s : quad == 1 ? c : quad == 2 ? -s : -c;
}
} // namespace ct

// a * exp(-+2*pi*i*M/R) with M, R compile-time: trivial factors cost nothing
template <int R, int M, bool INV>
__device__ __forceinline__ float2 twmul(float2 a) {
	constexpr int m = ((M % R) + R) % R;
	if constexpr (m == 0) {
		return a;
	} else if constexpr (2 * m == R) {
		return make_float2(-a.x, -a.y);
	} else if constexpr (4 * m == R) { // exp(-i*pi/2) = -i  (forward)
		return rot90<INV>(a);
	} else if constexpr (4 * m == 3 * R) { // exp(-3i*pi/2) = +i (forward)
		return rot90<!INV>(a);
	} else {
		constexpr float c = float(ct::cosq(m, R));
		constexpr float s = INV ? float(ct::sinq(m, R)) : -float(ct::sinq(m, R));
		return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
	}
}

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
	(f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
	static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// ---- in-register small DFTs ----
template <int R, bool INV>
struct SmallDFT;

template <bool INV>
struct SmallDFT<1, INV> {
	static __device__ __forceinline__ void run(float2 (&)[1]) {}
};
template <bool INV>
struct SmallDFT<2, INV> {
	static __device__ __forceinline__ void run(float2 (&v)[2]) {
		float2 a = v[0], b = v[1];
		v[0] = cadd(a, b);
		v[1] = csub(a, b);
	}
};
template <bool INV>
struct SmallDFT<3, INV> {
	static __device__ __forceinline__ void run(float2 (&v)[3]) {
		const float s60 = 0.86602540378443864676f;
		float2 t1 = cadd(v[1], v[2]);
		float2 m1 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
		float2 d = csub(v[1], v[2]);
		float2 t2 = rot90<INV>(make_float2(d.x * s60, d.y * s60));
		v[0] = cadd(v[0], t1);
		v[1] = cadd(m1, t2);
		v[2] = csub(m1, t2);
	}
};
template <bool INV>
struct SmallDFT<4, INV> {
	static __device__ __forceinline__ void run(float2 (&v)[4]) {
		float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = rot90<INV>(csub(v[1], v[3]));
		v[0] = cadd(a, c);
		v[1] = cadd(b, d);
		v[2] = csub(a, c);
		v[3] = csub(b, d);
	}
};
template <bool INV>
struct SmallDFT<5, INV> {
	static __device__ __forceinline__ void run(float2 (&v)[5]) {
		const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
		const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
		float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
		float2 r1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
		float2 r2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
		float2 i1 = rot90<INV>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
		float2 i2 = rot90<INV>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
		v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
		v[1] = cadd(r1, i1);
		v[2] = cadd(r2, i2);
		v[3] = csub(r2, i2);
		v[4] = csub(r1, i1);
	}
};
// Cooley-Tukey R = R1*R2 entirely in registers: input index n = R2*n1 + n2, output k = k1 + R1*k2
template <int R1, int R2, bool INV>
struct CompositeDFT {
	static __device__ __forceinline__ void run(float2 (&v)[R1 * R2]) {
		constexpr int R = R1 * R2;
		float2 t[R2][R1];
		static_for<R2>([&](auto n2c) {
			constexpr int n2 = decltype(n2c)::value;
			float2 a[R1];
			static_for<R1>([&](auto n1c) { a[decltype(n1c)::value] = v[R2 * decltype(n1c)::value + n2]; });
			SmallDFT<R1, INV>::run(a);
			static_for<R1>([&](auto k1c) {
				constexpr int k1 = decltype(k1c)::value;
				t[n2][k1] = twmul<R, n2 * k1, INV>(a[k1]);
			});
		});
		static_for<R1>([&](auto k1c) {
			constexpr int k1 = decltype(k1c)::value;
			float2 b[R2];
			static_for<R2>([&](auto n2c) { b[decltype(n2c)::value] = t[decltype(n2c)::value][k1]; });
			SmallDFT<R2, INV>::run(b);
			static_for<R2>([&](auto k2c) { v[k1 + R1 * decltype(k2c)::value] = b[decltype(k2c)::value]; });
		});
	}
};
template <bool INV> struct SmallDFT<6, INV> : CompositeDFT<2, 3, INV> {};
template <bool INV> struct SmallDFT<8, INV> : CompositeDFT<4, 2, INV> {};
template <bool INV> struct SmallDFT<10, INV> : CompositeDFT<2, 5, INV> {};
template <bool INV> struct SmallDFT<12, INV> : CompositeDFT<4, 3, INV> {};
template <bool INV> struct SmallDFT<16, INV> : CompositeDFT<4, 4, INV> {};

template <bool INV>
__device__ __forceinline__ float2 twiddle_at(const float2 *__restrict__ tw, int idx) {
	float2 w = __ldg(tw + idx);
	if (INV) w.y = -w.y;
	return w;
}

// One Stockham pass of radix R over `src` -> `dst` (padded shared-memory buffers, K points).
// Ns = product of the radices of earlier passes (a power of two).  KC/NSC > 0: K and Ns are
// compile-time constants (the preset sizes), which turns every per-element index into
// base + immediate: fpad(j + q*nb) = fpad(j) + q*(nb + nb/16) when 16 | nb, and likewise for the
// scattered stores (Ns == 1: fpad(R*j + q) with R == 16 is 17*j + q).
template <int R, bool INV, int KC = 0, int NSC = 0, int NT = 0>
__device__ __forceinline__ void fft_pass(int Krt, int NsRt, const float2 *src, float2 *dst, const float2 *__restrict__ tw, int tid, int nthrRt) {
	const int K = KC ? KC : Krt, Ns = NSC ? NSC : NsRt, nthr = NT ? NT : nthrRt;
	const int nb = K / R;
	const int twStep = K / (Ns * R);
	const int mask = Ns - 1;
	constexpr bool kConstStride = KC > 0 && NSC > 0 && ((KC / R) % 16 == 0) && (NSC % 16 == 0 || (NSC == 1 && R == 16));
	for (int j = tid; j < nb; j += nthr) {
		const int k = j & mask;
		float2 v[R];
		if constexpr (kConstStride) {
			constexpr int S1 = (KC / R) + (KC / R) / 16;
			const float2 *sp = src + fpad(j);
			static_for<R>([&](auto qc) { v[decltype(qc)::value] = sp[decltype(qc)::value * S1]; });
		} else {
			static_for<R>([&](auto qc) { v[decltype(qc)::value] = src[fpad(j + decltype(qc)::value * nb)]; });
		}
		if (Ns > 1) {
			// w^q = exp(-+2*pi*i*q*k/(Ns*R)): exact table entries for q = 1,2,4,8, products for the rest
			const int t = k * twStep;
			float2 wp[4];
			wp[0] = twiddle_at<INV>(tw, t);
			if (R > 2) wp[1] = twiddle_at<INV>(tw, 2 * t);
			if (R > 4) wp[2] = twiddle_at<INV>(tw, 4 * t);
			if (R > 8) wp[3] = twiddle_at<INV>(tw, 8 * t);
			static_for<R>([&](auto qc) {
				constexpr int q = decltype(qc)::value;
				if constexpr (q > 0) {
					constexpr int lowBit = (q & 1) ? 0 : (q & 2) ? 1 : (q & 4) ? 2 : 3;
					float2 w = wp[lowBit];
					constexpr int rest = q & ~(1 << lowBit);
					if constexpr (rest & 2) w = cmulf(w, wp[1]);
					if constexpr (rest & 4) w = cmulf(w, wp[2]);
					if constexpr (rest & 8) w = cmulf(w, wp[3]);
					v[q] = cmulf(v[q], w);
				}
			});
		}
		SmallDFT<R, INV>::run(v);
		const int j0 = (j - k) * R + k;
		if constexpr (kConstStride) {
			if constexpr (NSC == 1) {
				float2 *dp = dst + 17 * j;
				static_for<R>([&](auto qc) { dp[decltype(qc)::value] = v[decltype(qc)::value]; });
			} else {
				constexpr int S2 = NSC + NSC / 16;
				float2 *dp = dst + fpad(j0);
				static_for<R>([&](auto qc) { dp[decltype(qc)::value * S2] = v[decltype(qc)::value]; });
			}
		} else {
			static_for<R>([&](auto qc) { dst[fpad(j0 + decltype(qc)::value * Ns)] = v[decltype(qc)::value]; });
		}
	}
}

// Full K-point FFT of bufA (ping-pong with bufB, both fft_buf_len(K) float2, indexed through
// fpad()).  Returns the buffer holding the result.  Every thread of the block must call this (it
// contains __syncthreads()); the input must be complete and visible (caller syncs before).
// KT = 3072 / 2560 (presetDefault / presetCheaper at 44.1-48 kHz): fully specialised 3-pass plans
// 16*16*12 / 16*16*10 with 256 threads; KT = 0: run-time plan from cfg.radix[].
template <bool INV, int KT = 0>
__device__ float2 *fft_run(const Cfg &cfg, float2 *bufA, float2 *bufB, const float2 *__restrict__ tw, int tid, int nthr) {
	if constexpr (KT == 3072 || KT == 2560) {
		constexpr int RL = KT == 3072 ? 12 : 10;
		fft_pass<16, INV, KT, 1, 256>(KT, 1, bufA, bufB, tw, tid, 256);
		__syncthreads();
		fft_pass<16, INV, KT, 16, 256>(KT, 16, bufB, bufA, tw, tid, 256);
		__syncthreads();
		fft_pass<RL, INV, KT, 256, 256>(KT, 256, bufA, bufB, tw, tid, 256);
		__syncthreads();
		return bufB;
	} else {
		float2 *src = bufA, *dst = bufB;
		int Ns = 1;
		for (int st = 0; st < cfg.nStages; ++st) {
			const int R = cfg.radix[st];
			switch (R) {
			case 16: fft_pass<16, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 12: fft_pass<12, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 10: fft_pass<10, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 8: fft_pass<8, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 6: fft_pass<6, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 5: fft_pass<5, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 4: fft_pass<4, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			case 3: fft_pass<3, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			default: fft_pass<2, INV>(cfg.K, Ns, src, dst, tw, tid, nthr); break;
			}
			__syncthreads();
			float2 *t = src;
			src = dst;
			dst = t;
			Ns *= R;
		}
		return src;
	}
}

} // namespace b200s
