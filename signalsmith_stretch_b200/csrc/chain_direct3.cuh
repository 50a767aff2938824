// chain_direct3.cuh -- k_chain_direct3: the frame-wavefront phase prediction (reference :722-804, see
// the wavefront comment at k_chain in kernels.cuh) for STEREO calls without frequency map / formants,
// with the two channels of a block processed by PACKED sm_100 arithmetic.
//
//   * lane = block (32 blocks of the call per warp, one stream per warp), as in the first generation,
//     but every per-channel quantity is an f32x2 register pair {channel 0, channel 1}: the preliminary
//     prediction, the twists and the phase locking of the quieter channel run once per lane for both
//     channels (FMUL2 / FADD2 / FFMA2).  The packed instructions round each element exactly like their
//     scalar forms, and every product / sum below is issued in the reference's association order, so
//     the results are bit-identical to the scalar kernels (multiply-adds are NOT fused, except inside the
//     correctly-rounded division / square-root sequences, where the scalar code fuses them too).
//   * To make the pairs free, the analysis kernel writes the spectra of a stereo call channel-interleaved:
//     one float4 {re0, re1, im0, im1} per bin (k_analyse2 transforms the two channels as one pair anyway),
//     so the chunk fill is one 16-byte cp.async per (block, bin) and every tile read is one LDS.128 that
//     lands directly in register pairs.  Band::output rows (Y) stay planar for the synthesis kernel.
//   * Measured motivation (profiles/r01_v10_ncu_summary.md): the scalar chain kernels are bound by the
//     number of instructions on each warp's serial stream (311-455 per bin step at ~5-6 cycles per issue);
//     packing the channels roughly halves that stream.
#pragma once
#include <type_traits>

#include "chain_direct.cuh"
#include "fft2.cuh"
#include "kernels.cuh"

namespace b200s {

#define CH3_RING 32
#define CH3_RS 34 // float4 row stride: step reads [k - D*lane][lane] and the fill below are conflict-free (scratch/bank_check3.py)

struct Chain3Tiles {
	float4 in[CH3_RING][CH3_RS]; // rolling window of each block's interleaved input spectrum, [bin & 31][lane]
	// previous-input spectrum at the chunk's bins; overwritten by the finals of the same step.  Two buffers: the
	// fill of chunk c+1 is in flight (cp.async) while chunk c is computed
	float4 pvy[2][CHAIN_CH][CH3_RS];
	float4 p0Out[2][CHAIN_CH];    // lane 0's predecessor block: {c0.re, c0.im, c1.re, c1.im}
	float2 p0E[2][CHAIN_CH];      // its Prediction::energy {c0, c1}
	const float4 *rowIn[32], *rowPv[32];
	float2 *rowY[32]; // channel 0 row of Band::output; channel 1 row follows K bins later
};

// ---- packed counterparts of the exact helpers of kernels.cuh (same operations, same order, per element) ----
#ifdef B200S_EMU
__device__ __forceinline__ f2 fma2(f2 x, f2 y, f2 z) { return f2{std::fma(x.a, y.a, z.a), std::fma(x.b, y.b, z.b)}; }
__device__ __forceinline__ f2 rcp_approx2(f2 b) { return f2{1.0f / b.a, 1.0f / b.b}; }
__device__ __forceinline__ f2 rsqrt_approx2(f2 b) { return f2{1.0f / std::sqrt(b.a), 1.0f / std::sqrt(b.b)}; }
__device__ __forceinline__ f2 fdivq2(f2 a, f2 b) { return f2{a.a / b.a, a.b / b.b}; }
__device__ __forceinline__ f2 fsqrtq2(f2 a) { return f2{std::sqrt(a.a), std::sqrt(a.b)}; }
#else
__device__ __forceinline__ f2 fma2(f2 x, f2 y, f2 z) {
	f2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(x.v), "l"(y.v), "l"(z.v));
	return r;
}
__device__ __forceinline__ f2 neg2(f2 a) { return f2_make(-f2_lo(a), -f2_hi(a)); } // folded into operand modifiers by ptxas
__device__ __forceinline__ f2 fdivq2(f2 a, f2 b) { // fdivq (kernels.cuh) on both elements
	float r0, r1;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(f2_lo(b)));
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(f2_hi(b)));
	f2 r = f2_make(r0, r1);
	const f2 nb = neg2(b);
	r = fma2(r, fma2(nb, r, f2_make(1.0f, 1.0f)), r);
	const f2 q = mul2(a, r);
	return fma2(fma2(nb, q, a), r, q);
}
__device__ __forceinline__ f2 fsqrtq2(f2 a) { // fsqrtq (kernels.cuh) on both elements
	float r0, r1;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(f2_lo(a)));
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(f2_hi(a)));
	const f2 r = f2_make(r0, r1);
	const f2 g = mul2(a, r), h = muls(r, 0.5f);
	const f2 s = fma2(fma2(neg2(g), g, a), h, g);
	return f2_make(f2_lo(a) == 0.f ? 0.f : f2_lo(s), f2_hi(a) == 0.f ? 0.f : f2_hi(s));
}
#endif
__device__ __forceinline__ f2 sel_f2(bool p, f2 a) { return f2_make(p ? f2_lo(a) : 0.f, p ? f2_hi(a) : 0.f); }
__device__ __forceinline__ c2 sel_c2(bool p, c2 a) { return c2{sel_f2(p, a.re), sel_f2(p, a.im)}; }

// EXACT sum / difference of two packed values of which at least one is a product.  ptxas (12.9) contracts
// mul.rn.f32x2 followed by add/sub.rn.f32x2 into FMUL2 + FFMA2 even with --fmad=false (measured; the scalar
// .rn forms are never contracted), which would round differently from the reference's separate multiply and
// add.  p*one + q with `one` = 1.0f from a kernel parameter (opaque to the compiler) is the same correctly
// rounded p + q in one FFMA2 and cannot absorb the multiply that produced p.
#ifdef B200S_EMU
__device__ __forceinline__ f2 padd(f2 p, f2 q, float) { return p + q; }
__device__ __forceinline__ f2 psub(f2 p, f2 q, float) { return p - q; }
#else
__device__ __forceinline__ f2 padd(f2 p, f2 q, float one) { return fma2(p, f2_make(one, one), q); }
__device__ __forceinline__ f2 psub(f2 p, f2 q, float one) { return fma2(p, f2_make(one, one), neg2(q)); }
#endif
// xmul: a * b, both packed
__device__ __forceinline__ c2 xmul2(c2 a, c2 b, float one) {
	return c2{psub(mul2(a.re, b.re), mul2(a.im, b.im), one), padd(mul2(a.re, b.im), mul2(a.im, b.re), one)};
}
// xmul with a scalar complex factor on the right (rotation): a * r
__device__ __forceinline__ c2 xmul2s(c2 a, float2 r, float one) {
	return c2{psub(muls(a.re, r.x), muls(a.im, r.y), one), padd(muls(a.re, r.y), muls(a.im, r.x), one)};
}
// xmulc: a * conj(b) in the scalar helper's operand order (b.x*a.x + b.y*a.y, b.x*a.y - b.y*a.x)
__device__ __forceinline__ c2 xmulc2(c2 a, c2 b, float one) {
	return c2{padd(mul2(b.re, a.re), mul2(b.im, a.im), one), psub(mul2(b.re, a.im), mul2(b.im, a.re), one)};
}
__device__ __forceinline__ f2 xnorm2(c2 a, float one) { return padd(mul2(a.re, a.re), mul2(a.im, a.im), one); }
// low + (high - low)*frac, frac common to both channels
__device__ __forceinline__ c2 xlerp2p(c2 lo, c2 hi, float fr, float one) {
	return c2{padd(muls(hi.re - lo.re, fr), lo.re, one), padd(muls(hi.im - lo.im, fr), lo.im, one)};
}
// Prediction::makeOutput (:596-603) for both channels
__device__ __forceinline__ c2 make_output_q2(c2 phase, f2 energy, c2 input, float one) {
	const f2 pn = xnorm2(phase, one);
	const bool w0 = f2_lo(pn) <= B200S_NOISE_FLOOR, w1 = f2_hi(pn) <= B200S_NOISE_FLOOR;
	const f2 pni = xnorm2(input, one) + f2_make(B200S_NOISE_FLOOR, B200S_NOISE_FLOOR);
	const c2 ph = c2{f2_make(w0 ? f2_lo(input.re) : f2_lo(phase.re), w1 ? f2_hi(input.re) : f2_hi(phase.re)),
	                 f2_make(w0 ? f2_lo(input.im) : f2_lo(phase.im), w1 ? f2_hi(input.im) : f2_hi(phase.im))};
	const f2 pn2 = f2_make(w0 ? f2_lo(pni) : f2_lo(pn), w1 ? f2_hi(pni) : f2_hi(pn));
	const f2 g = fsqrtq2(fdivq2(energy, pn2));
	return c2{mul2(ph.re, g), mul2(ph.im, g)};
}
__device__ __forceinline__ c2 ld_c2s(const float4 *p) { // shared / global float4 {re0, re1, im0, im1}
	const float4 v = *p;
	return c2{f2_make(v.x, v.y), f2_make(v.z, v.w)};
}
__device__ __forceinline__ float2 pick(bool second, c2 a) { // one channel of a packed complex
	return make_float2(second ? f2_hi(a.re) : f2_lo(a.re), second ? f2_hi(a.im) : f2_lo(a.im));
}

template <int LT>
__global__ void __launch_bounds__(32) k_chain_direct3(Ctx x) {
	const Cfg &g = x.cfg;
	const int K = g.K;
	B200S_DYN_SHARED
	const int lane = threadIdx.x & 31;
	const int s = x.sBase + blockIdx.x;
	const Call cl = x.call[s];
	if (cl.nFrames == 0) return;
	if (cl.hasRandom && x.randomPathOn) return; // a block beyond 2x stretch draws random time factors: k_prep + k_chain take the stream
	constexpr int D = LT + 1;
	Chain3Tiles &U = *(Chain3Tiles *)dyn_smem;
	// chunk fill: lane -> (bin offset, row within a group of 4); a quarter-warp covers 4 bins (64 B) of 2 rows
	const int fillI = (lane & 3) | (((lane >> 3) & 1) << 2), fillF = ((lane >> 2) & 1) | (((lane >> 4) & 1) << 1);
	const float2 rot0 = x.rot0, rotStep = x.rotStep;
	const float one = x.one; // 1.0f, opaque to the compiler (see padd / psub)

	for (int base = 0; base < cl.nFrames; base += 32) {
		__syncwarp();
		const int f = base + lane;
		const bool active = f < cl.nFrames;
		const Frame fr = x.frames[(size_t)s * x.maxFrames + (active ? f : base)];
		const bool rotOn = fr.flags & FR_NEW_SPECTRUM;
		const int nAct = min(32, cl.nFrames - base);
		const float tf = fmaxf(fr.timeFactor, 1.0f / B200S_MAX_CLEAN_STRETCH); // :638
		const float longTf = fmul((float)LT, tf);
		const bool farAny = __any_sync(0xffffffffu, active && longTf > (float)(CH3_RING - 2 * CHAIN_CH - 3)); // 8 bins of the next chunk are in flight
		const float2 *prevOut[2];
		const float *prevE[2];
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			prevOut[c] = base == 0 ? x.stOut + ((size_t)s * 2 + c) * K : x.Y + coef_off(x, s, base - 1, c);
			prevE[c] = x.stPredE + ((size_t)s * 2 + c) * K; // base == 0 only; later groups recompute it, see the fill
		}
		// Prediction::energy of a block on this path is |input|^2 of its own spectrum (:679,:708): the chain never stores
		// it -- the next group recomputes it from the predecessor's input row, k_commit from the final input spectrum
		const float4 *prevIn = base == 0 ? nullptr : il_row(x, s, x.frames[(size_t)s * x.maxFrames + base - 1].inSlot);
		const float4 *myIn = il_row(x, s, fr.inSlot);
		U.rowIn[lane] = myIn;
		U.rowPv[lane] = il_row(x, s, fr.prevSlot);
		U.rowY[lane] = x.Y + coef_off(x, s, active ? f : base, 0);
		__syncwarp();
		// register FIFOs (channel pairs); at the start of a step (q = prelim bin, b = q - L = final bin):
		//   pre/eF/t2F/inF[i] <-> prelim output / energy / long twist / input at bin b+i
		//   oh[i] <-> final output at bin b-1-i;   t1P <-> short twist at bin b
		const f2 z2 = f2_make(0.f, 0.f);
		const c2 zc = c2{z2, z2};
		c2 oh[LT], pre[LT], t2F[LT], inF[LT], t1P = zc, lastFinal = zc;
		f2 eF[LT], lastE = z2;
#pragma unroll
		for (int i = 0; i < LT; ++i) {
			oh[i] = pre[i] = t2F[i] = inF[i] = zc;
			eF[i] = z2;
		}
		float2 rotq = rotOn ? rot0 : make_float2(1.f, 0.f); // rot[q] by the reference's float recurrence (:647-655)
		const float2 rotS = rotOn ? rotStep : make_float2(1.f, 0.f);
		const int steps = K + LT + D * (nAct - 1);
		// first chunk start from which every ACTIVE lane has q - L*tf - 1 >= 0 and b = q - L >= L (inactive lanes only
		// produce values nobody consumes): q >= D*(nAct-1) + 2L + ceil(max L*tf) + 1
		int interiorFrom;
		{
			float mx = active ? longTf : 0.f;
			for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
			interiorFrom = D * (nAct - 1) + 2 * LT + (int)ceilf(mx) + 2;
		}
		// asynchronous fill of the chunk starting at kf into buffer `buf`: 8 new bins per block, both channels per
		// 16-byte copy.  The ring slots it writes ((q .. q+7) & 31 per lane) are disjoint from what the chunk in
		// progress reads (at most L*tf + 1 <= 9 bins behind its own 8).
		auto fill = [&](int kf, int buf) {
#pragma unroll
			for (int it = 0; it < 8; ++it) {
				const int fl = fillF + 4 * it;
				const int q = kf + fillI - D * fl;
				if (base + fl < cl.nFrames && (unsigned)q < (unsigned)K) {
					cp_async16(&U.in[q & (CH3_RING - 1)][fl], U.rowIn[fl] + q);
					cp_async16(&U.pvy[buf][fillI][fl], U.rowPv[fl] + q);
				}
			}
			if (lane < 2 * CHAIN_CH) { // lane 0's predecessor: planar state / previous group rows -> {c0, c1} slots
				const int qq = kf + (lane >> 1), c = lane & 1;
				if (qq < K) {
					cp_async8((float2 *)&U.p0Out[buf][lane >> 1] + c, (c ? prevOut[1] : prevOut[0]) + qq);
					if (base == 0) {
						cp_async4((float *)&U.p0E[buf][lane >> 1] + c, (c ? prevE[1] : prevE[0]) + qq);
					} else {
						const float4 v = prevIn[qq];
						((float *)&U.p0E[buf][lane >> 1])[c] = c ? xnorm(make_float2(v.y, v.w)) : xnorm(make_float2(v.x, v.z));
					}
				}
			}
		};
		fill(0, 0);
		int cb = 0; // buffer of the chunk being computed
		for (int k0 = 0; k0 < steps; k0 += CHAIN_CH, cb ^= 1) {
			cp_async_wait_all(); // this chunk's tiles (issued one chunk ago)
			__syncwarp();
			if (k0 + CHAIN_CH < steps) fill(k0 + CHAIN_CH, cb ^ 1); // next chunk: in flight during the 8 steps below
			// ---------------- CHAIN_CH steps ----------------
			// INTERIOR: every lane's q, b and interpolation points are inside [0, K) for the whole chunk, so all the
			// edge masks below are identities and are compiled out (about nine chunks in ten)
			auto step = [&](const int i, auto farTag, auto intTag) {
				constexpr bool FAR = decltype(farTag)::value, INTERIOR = decltype(intTag)::value;
				const int q = k0 + i - D * lane;
				const int b = q - LT;
				const bool qIn = INTERIOR || (active && (unsigned)q < (unsigned)K);
				// the twists need input interpolated at q - L*tf and (b+1) - tf  (:750,:757)
				const float i2 = fsub((float)q, longTf);
				const int l2 = (int)floorf(i2);
				const float f2s = fsub(i2, (float)l2);
				const float i1 = fsub((float)(b + 1), tf);
				const int l1 = (int)floorf(i1);
				const float f1s = fsub(i1, (float)l1);
				// previous block's final output / energy at bin q: finalised by lane-1 in the last step
				c2 ro;
				f2 re;
				{
					const float a0 = __shfl_up_sync(0xffffffffu, f2_lo(lastFinal.re), 1), a1 = __shfl_up_sync(0xffffffffu, f2_hi(lastFinal.re), 1);
					const float b0 = __shfl_up_sync(0xffffffffu, f2_lo(lastFinal.im), 1), b1 = __shfl_up_sync(0xffffffffu, f2_hi(lastFinal.im), 1);
					const float e0 = __shfl_up_sync(0xffffffffu, f2_lo(lastE), 1), e1 = __shfl_up_sync(0xffffffffu, f2_hi(lastE), 1);
					const float4 p0 = U.p0Out[cb][i];
					const float2 p0e = U.p0E[cb][i];
					const bool first = lane == 0;
					ro = c2{f2_make(first ? p0.x : a0, first ? p0.z : a1), f2_make(first ? p0.y : b0, first ? p0.w : b1)};
					re = f2_make(first ? p0e.x : e0, first ? p0e.y : e1);
				}
				const c2 inq = ld_c2s(&U.in[q & (CH3_RING - 1)][lane]);
				c2 pv = ld_c2s(&U.pvy[cb][i][lane]);
				c2 lo2, hi2, lo1, hi1;
				if constexpr (!FAR) {
					lo2 = sel_c2(INTERIOR || l2 >= 0, ld_c2s(&U.in[l2 & (CH3_RING - 1)][lane]));
					hi2 = sel_c2(INTERIOR || l2 >= -1, ld_c2s(&U.in[(l2 + 1) & (CH3_RING - 1)][lane]));
					lo1 = sel_c2(INTERIOR || l1 >= 0, ld_c2s(&U.in[l1 & (CH3_RING - 1)][lane]));
					hi1 = sel_c2(INTERIOR || l1 >= -1, ld_c2s(&U.in[(l1 + 1) & (CH3_RING - 1)][lane]));
				} else { // extreme stretch (> 2x): gather straight from the spectrum row
					lo2 = (l2 < 0 || l2 >= K) ? zc : ld_c2s(myIn + l2);
					hi2 = (l2 + 1 < 0 || l2 + 1 >= K) ? zc : ld_c2s(myIn + l2 + 1);
					lo1 = (l1 < 0 || l1 >= K) ? zc : ld_c2s(myIn + l1);
					hi1 = (l1 + 1 < 0 || l1 + 1 >= K) ? zc : ld_c2s(myIn + l1 + 1);
				}
				pv = xmul2s(pv, rotq, one); // :653-654 rotate Band::output and Band::prevInput by one interval
				ro = xmul2s(ro, rotq, one);
				const f2 e = xnorm2(inq, one);                       // :679 (identity map: energy = |input|^2)
				const c2 ph0 = xmul2(ro, xmulc2(inq, pv, one), one);      // :714-715
				const f2 den = f2_make(fmaxf(f2_lo(re), f2_lo(e)), fmaxf(f2_hi(re), f2_hi(e))) + f2_make(B200S_NOISE_FLOOR, B200S_NOISE_FLOOR);
				const c2 newPre = sel_c2(qIn, c2{fdivq2(ph0.re, den), fdivq2(ph0.im, den)}); // :716
				const f2 newE = sel_f2(qIn, e);
				const c2 newIn = sel_c2(qIn, inq);
				const c2 newT2 = sel_c2(qIn, xmulc2(inq, xlerp2p(lo2, hi2, f2s, one), one)); // long twist at q (:758)
				// short twist at b+1 (:751,:771): Prediction::input[b+1] is inF[1] before the shift
				const c2 t1N = xmulc2(LT > 1 ? inF[LT > 1 ? 1 : 0] : newIn, xlerp2p(lo1, hi1, f1s, one), one);
				{
					const float2 rn = xmul(rotq, rotS);
					rotq = make_float2((INTERIOR || q >= 0) ? rn.x : rotq.x, (INTERIOR || q >= 0) ? rn.y : rotq.y);
				}
				// ---- FIFO rotation (pure renaming after unrolling): what falls out belongs to bin b
				const f2 eB = eF[0];
				const c2 t2B = t2F[0], inB = inF[0];
#pragma unroll
				for (int u = 0; u + 1 < LT; ++u) {
					pre[u] = pre[u + 1];
					eF[u] = eF[u + 1];
					t2F[u] = t2F[u + 1];
					inF[u] = inF[u + 1];
				}
				pre[LT - 1] = newPre;
				eF[LT - 1] = newE;
				t2F[LT - 1] = newT2;
				inF[LT - 1] = newIn;
				// ---- main prediction at bin b (:727-800): the louder channel (first on ties, :733) leads
				const bool m = f2_hi(eB) > f2_lo(eB);
				const float maxE = m ? f2_hi(eB) : f2_lo(eB);
				// the phase sum of :754-784 is formed for both channels at once (packed, each channel from its own
				// registers, exactly as if it were the leader) and the leader's is picked afterwards
				c2 ph2 = zc;
				ph2 = ph2 + sel_c2(INTERIOR || b > 0, xmul2(oh[0], t1P, one));                      // :754
				ph2 = ph2 + sel_c2(INTERIOR || b >= LT, xmul2(oh[LT - 1], t2B, one));               // :761
				ph2 = ph2 + sel_c2(INTERIOR || b < K - 1, xmulc2(pre[0], t1N, one));                // :774
				ph2 = ph2 + sel_c2(INTERIOR || b < K - LT, xmulc2(pre[LT - 1], t2F[LT - 1], one)); // :784
				const float2 phase = pick(m, ph2), pinM = pick(m, inB);
				const float2 outM = make_output_q(phase, maxE, pinM);     // :788
				// the other channel is locked in phase (:791-799); computed for both, the leader keeps outM
				//   cph = xmul(outM, xmulc(inB_c, pinM)), operand order of the scalar helpers
				const c2 tw = c2{padd(muls(inB.re, pinM.x), muls(inB.im, pinM.y), one), psub(muls(inB.im, pinM.x), muls(inB.re, pinM.y), one)};
				const c2 cph = c2{psub(muls(tw.re, outM.x), muls(tw.im, outM.y), one), padd(muls(tw.im, outM.x), muls(tw.re, outM.y), one)};
				const c2 other = make_output_q2(cph, eB, inB, one);
				const c2 oc = c2{f2_make(m ? f2_lo(other.re) : outM.x, m ? outM.x : f2_hi(other.re)),
				                 f2_make(m ? f2_lo(other.im) : outM.y, m ? outM.y : f2_hi(other.im))};
				// unconditional: out-of-range steps only produce values that every consumer masks
#pragma unroll
				for (int u = LT - 1; u > 0; --u) oh[u] = oh[u - 1];
				oh[0] = oc;
				lastFinal = oc;
				lastE = eB;
				t1P = t1N;
				U.pvy[cb][i][lane] = make_float4(f2_lo(oc.re), f2_hi(oc.re), f2_lo(oc.im), f2_hi(oc.im));
			};
			// unrolled by 4 (for L = 4 the register FIFOs rotate by pure renaming) so that the hot loop stays in the
			// instruction cache; branch-free inside
			auto run_chunk = [&](auto farTag, auto intTag) {
#pragma unroll 1
				for (int h = 0; h < CHAIN_CH; h += 4) {
#pragma unroll
					for (int u = 0; u < 4; ++u) step(h + u, farTag, intTag);
				}
			};
			if (farAny) run_chunk(std::true_type{}, std::false_type{});
			else if (k0 >= interiorFrom && k0 + CHAIN_CH <= K) run_chunk(std::false_type{}, std::true_type{});
			else run_chunk(std::false_type{}, std::false_type{});
			__syncwarp();
			// ---------------- write the chunk's finals back: planar Band::output rows, 32 B per row and quarter-warp ----------------
#pragma unroll
			for (int it = 0; it < 8; ++it) {
				const int fl = fillF + 4 * it;
				const int b = k0 + fillI - D * fl - LT;
				if (base + fl < cl.nFrames && (unsigned)b < (unsigned)K) {
					const float4 v = U.pvy[cb][fillI][fl];
					U.rowY[fl][b] = make_float2(v.x, v.z);
					U.rowY[fl][K + b] = make_float2(v.y, v.w);
				}
			}
			__syncwarp();
		}
	}
}

} // namespace b200s
