// fft.cuh -- shared-memory Stockham FFT (radix 4/2 stages + one radix-3 or radix-5 stage) for the
// K = N/2 point complex transform behind the reference's "modified" (half-bin shifted) real FFT
// (dependency DynamicSTFT::analyseStep / synthesiseStep; reference call sites
// signalsmith-stretch.h:337,359,398; convention pinned in SURVEY.md section 8(a) row 6 / App. F).
//
// The whole working set (K float2, ping-pong) lives in shared memory; no cuFFT, no tensor cores.
// Sizes are always 2^a * {1,3,5} (fastSizeAbove); the odd radix is the LAST
// stage so that every stage's sub-transform length Ns is a power of two (index math by masks).
#pragma once
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

template <bool INV>
__device__ __forceinline__ float2 twiddle_at(const float2 *__restrict__ tw, int idx) {
	float2 w = __ldg(tw + idx);
	if (INV) w.y = -w.y;
	return w;
}

// One Stockham stage of radix R over `src` -> `dst` (both shared memory, M points).
// Ns = product of the radices of earlier stages (a power of two).
template <bool INV>
__device__ void fft_stage(int R, int M, int Ns, const float2 *src, float2 *dst, const float2 *__restrict__ tw, int tid, int nthr) {
	const int nb = M / R;
	const int twStep = M / (Ns * R);
	const int mask = Ns - 1;
	if (R == 4) {
		for (int j = tid; j < nb; j += nthr) {
			int k = j & mask;
			float2 v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb], v3 = src[j + 3 * nb];
			if (Ns > 1) {
				int t = k * twStep;
				v1 = cmulf(v1, twiddle_at<INV>(tw, t));
				v2 = cmulf(v2, twiddle_at<INV>(tw, 2 * t));
				v3 = cmulf(v3, twiddle_at<INV>(tw, 3 * t));
			}
			float2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = rot90<INV>(csub(v1, v3));
			int j0 = ((j - k) << 2) + k; // (j / Ns) * Ns * 4 + k
			dst[j0] = cadd(a, c);
			dst[j0 + Ns] = cadd(b, d);
			dst[j0 + 2 * Ns] = csub(a, c);
			dst[j0 + 3 * Ns] = csub(b, d);
		}
	} else if (R == 2) {
		for (int j = tid; j < nb; j += nthr) {
			int k = j & mask;
			float2 v0 = src[j], v1 = src[j + nb];
			if (Ns > 1) v1 = cmulf(v1, twiddle_at<INV>(tw, k * twStep));
			int j0 = ((j - k) << 1) + k;
			dst[j0] = cadd(v0, v1);
			dst[j0 + Ns] = csub(v0, v1);
		}
	} else if (R == 3) {
		const float s60 = 0.86602540378443864676f;
		for (int j = tid; j < nb; j += nthr) {
			int k = j & mask;
			float2 v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb];
			if (Ns > 1) {
				int t = k * twStep;
				v1 = cmulf(v1, twiddle_at<INV>(tw, t));
				v2 = cmulf(v2, twiddle_at<INV>(tw, 2 * t));
			}
			float2 t1 = cadd(v1, v2);
			float2 m1 = make_float2(v0.x - 0.5f * t1.x, v0.y - 0.5f * t1.y);
			float2 d = csub(v1, v2);
			float2 t2 = rot90<INV>(make_float2(d.x * s60, d.y * s60));
			int j0 = (j - k) * 3 + k;
			dst[j0] = cadd(v0, t1);
			dst[j0 + Ns] = cadd(m1, t2);
			dst[j0 + 2 * Ns] = csub(m1, t2);
		}
	} else { // R == 5
		const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
		const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
		for (int j = tid; j < nb; j += nthr) {
			int k = j & mask;
			float2 v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb], v3 = src[j + 3 * nb], v4 = src[j + 4 * nb];
			if (Ns > 1) {
				int t = k * twStep;
				v1 = cmulf(v1, twiddle_at<INV>(tw, t));
				v2 = cmulf(v2, twiddle_at<INV>(tw, 2 * t));
				v3 = cmulf(v3, twiddle_at<INV>(tw, 3 * t));
				v4 = cmulf(v4, twiddle_at<INV>(tw, 4 * t));
			}
			float2 a1 = cadd(v1, v4), a2 = cadd(v2, v3), b1 = csub(v1, v4), b2 = csub(v2, v3);
			float2 r1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
			float2 r2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
			float2 i1 = rot90<INV>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
			float2 i2 = rot90<INV>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
			int j0 = (j - k) * 5 + k;
			dst[j0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
			dst[j0 + Ns] = cadd(r1, i1);
			dst[j0 + 2 * Ns] = cadd(r2, i2);
			dst[j0 + 3 * Ns] = csub(r2, i2);
			dst[j0 + 4 * Ns] = csub(r1, i1);
		}
	}
}

// Full K-point FFT of bufA (ping-pong with bufB).  Returns the buffer holding the result.
// Every thread of the block must call this (it contains __syncthreads()); the input must be
// complete and visible (caller syncs before).
template <bool INV>
__device__ float2 *fft_run(const Cfg &cfg, float2 *bufA, float2 *bufB, const float2 *__restrict__ tw, int tid, int nthr) {
	float2 *src = bufA, *dst = bufB;
	int Ns = 1;
	for (int st = 0; st < cfg.nStages; ++st) {
		int R = cfg.radix[st];
		fft_stage<INV>(R, cfg.K, Ns, src, dst, tw, tid, nthr);
		__syncthreads();
		float2 *t = src;
		src = dst;
		dst = t;
		Ns *= R;
	}
	return src;
}

} // namespace b200s
