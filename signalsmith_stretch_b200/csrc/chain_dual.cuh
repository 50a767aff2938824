// chain_dual.cuh -- k_chain_dual: TWO MONO STREAMS per warp on the packed frame wavefront of k_chain_direct4 (round 2).
//
// The mono plain path ran k_chain_direct2 (scalar arithmetic, one stream per warp): per channel 1.6x the cost of the
// packed stereo kernel, half of every point of the ratio x preset sweep (profiles/r02_sweep.jsonl).  A mono batch has no
// second channel to pack -- but it has a second STREAM: here lane j runs block j of stream 2p in the low halves of its
// f32x2 registers and block j of stream 2p + 1 in the high halves.  Everything of k_chain_direct4 carries over (decoupled
// lane skew, 8-step chunks, fast / exact arithmetic) except that the halves never meet: no loudest-channel choice, no
// phase lock, each half its own makeOutput (:788).  The spectra stay planar (mono analysis is unchanged): the chunk fill
// copies 8 bytes per stream into the two halves of a 16-byte tile slot {A.re, A.im, B.re, B.im}.
// Packing needs the two streams to walk the same schedule in this call -- same blocks, flags, time factors and spectrum
// slots (streams of a batch do, unless one of them was silent or sought differently): dual_pair_ok checks exactly that;
// pairs that fail it, and the odd stream of an odd batch, are left to k_chain_direct2, which skips the pairs taken here.
#pragma once
#include "chain_direct4.cuh"

namespace b200s {

// both streams of the pair starting at stream s (s - sBase even) are there, free of random blocks, and share the schedule
__device__ __forceinline__ bool dual_pair_ok(const Ctx &x, int s, int lane) {
	if (s + 1 >= x.sBase + x.sCount) return false;
	const Call a = x.call[s], b = x.call[s + 1];
	if (a.nFrames == 0 || a.nFrames != b.nFrames || a.bypass || b.bypass) return false;
	if ((a.hasRandom || b.hasRandom) && x.randomPathOn) return false;
	bool bad = false;
	for (int f = lane; f < a.nFrames; f += 32) {
		const Frame fa = x.frames[(size_t)s * x.maxFrames + f], fb = x.frames[(size_t)(s + 1) * x.maxFrames + f];
		bad = bad || fa.flags != fb.flags || fa.timeFactor != fb.timeFactor || fa.inSlot != fb.inSlot || fa.prevSlot != fb.prevSlot;
	}
	return !__any_sync(0xffffffffu, bad);
}

struct ChainDualTiles { // as Chain4Tiles; every float4 slot holds {A.re, A.im, B.re, B.im} of the pair's two streams
	float4 in[CH3_RING][CH3_RS];
	float4 pvy[2][CHAIN_CH][CH3_RS];
	float4 p0Out[2][CHAIN_CH]; // lane 0's predecessor blocks: {A.re, A.im, B.re, B.im}
	float2 p0E[2][CHAIN_CH];   // their Prediction::energy {A, B}
	const float2 *rowIn[2][32], *rowPv[2][32];
};
__device__ __forceinline__ c2 ld_c2d(const float4 *p) { // {A.re, A.im, B.re, B.im} -> packed pairs {A, B}
	const float4 v = *p;
	return c2{f2_make(v.x, v.z), f2_make(v.y, v.w)};
}
__device__ __forceinline__ c2 gather2(const float2 *const rows[2], int b, int K) { // bin b of both streams' rows, zero outside
	if (b < 0 || b >= K) return c2{f2_make(0.f, 0.f), f2_make(0.f, 0.f)};
	const float2 a = rows[0][b], bb = rows[1][b];
	return c2{f2_make(a.x, bb.x), f2_make(a.y, bb.y)};
}

template <int LT, bool FAST>
__global__ void __launch_bounds__(32) k_chain_dual(Ctx x) {
	const Cfg &g = x.cfg;
	const int K = g.K;
	B200S_DYN_SHARED
	const int lane = threadIdx.x & 31;
	const int s = x.sBase + 2 * blockIdx.x, sB = s + 1; // the pair: stream s in the low halves, stream s + 1 in the high halves
	if (!dual_pair_ok(x, s, lane)) return;             // k_chain_direct2 takes these streams one by one
	const Call cl = x.call[s];
	const int sH[2] = {s, sB};
	constexpr int G = LT + 2; // lane skew in bins
	constexpr int NF = LT + 1; // FIFO entries: bins b .. b+L of the lane's block
	ChainDualTiles &U = *(ChainDualTiles *)dyn_smem;
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
		const bool farAny = __any_sync(0xffffffffu, active && longTf > (float)CH4_FAR);
		const float2 *prevOut[2];
		const float *prevE[2];
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			prevOut[c] = base == 0 ? x.stOut + (size_t)sH[c] * K : x.Y + coef_off(x, sH[c], base - 1, 0);
			prevE[c] = x.stPredE + (size_t)sH[c] * K; // base == 0 only; later groups recompute it, see the fill
		}
		// Prediction::energy of a block on this path is |input|^2 of its own spectrum (:679,:708): the chain never stores
		// it -- the next group recomputes it from the predecessor's input row, k_commit from the final input spectrum
		const float2 *prevIn[2], *myIn[2];
		float2 *yBase[2]; // Band::output row of block base + r of stream h: yBase[h] + r * K (mono: one row per block)
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			prevIn[h] = base == 0 ? nullptr : spec_slot(x, sH[h], x.frames[(size_t)s * x.maxFrames + base - 1].inSlot, 0);
			myIn[h] = spec_slot(x, sH[h], fr.inSlot, 0);
			U.rowIn[h][lane] = myIn[h];
			U.rowPv[h][lane] = spec_slot(x, sH[h], fr.prevSlot, 0);
			yBase[h] = x.Y + coef_off(x, sH[h], base, 0);
		}
		const bool lastFrame = active && f == cl.nFrames - 1; // its Prediction::energy is the state the next call continues from (k_commit)
		float *const eRow[2] = {x.cE + coef_off(x, s, active ? f : base, 0), x.cE + coef_off(x, sB, active ? f : base, 0)};
		__syncwarp();
		// register FIFOs (channel pairs); at the start of a step (q = prelim bin, b = q - L - 1 = final bin):
		//   pre/eF/t2F/inF[i] <-> prelim output / energy / long twist / input at bin b+i, i = 0..L (b+L = q-1)
		//   oh[i] <-> final output at bin b-1-i;   t1P <-> short twist at bin b
		const f2 z2 = f2_make(0.f, 0.f);
		const c2 zc = c2{z2, z2};
		c2 oh[LT], pre[NF], t2F[NF], inF[NF], t1P = zc, lastFinal = zc;
		f2 eF[NF], lastE = z2;
#pragma unroll
		for (int i = 0; i < LT; ++i) oh[i] = zc;
#pragma unroll
		for (int i = 0; i < NF; ++i) {
			pre[i] = t2F[i] = inF[i] = zc;
			eF[i] = z2;
		}
		float2 rotq = rotOn ? rot0 : make_float2(1.f, 0.f); // rot[q] by the reference's float recurrence (:647-655)
		const float2 rotS = rotOn ? rotStep : make_float2(1.f, 0.f);
		const int steps = K + LT + 1 + G * (nAct - 1);
		// first chunk start from which every ACTIVE lane has q - L*tf - 1 >= 0, (b+1) - tf - 1 >= 0 and b = q - L - 1 >= L
		// (inactive lanes only produce values nobody consumes): q >= G*(nAct-1) + 2L + ceil(max L*tf) + 2
		int interiorFrom;
		{
			float mx = active ? longTf : 0.f;
			for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
			interiorFrom = G * (nAct - 1) + 2 * LT + (int)ceilf(mx) + 3;
		}
		// asynchronous fill of the chunk starting at kf into buffer `buf`: 8 new bins per block, both channels per
		// 16-byte copy.  The ring slots it writes (the 8 bins after the chunk's own 8, per lane) are disjoint from what
		// the chunk in progress reads (at most CH4_FAR + 2 bins behind its own 8: 8 + 8 + CH4_FAR + 2 <= CH3_RING).
		// The row pointers are read first, all at once: the copies are asm volatile and would otherwise serialise the
		// shared-memory latency of every pointer load (measured: the fill was 15 % of the kernel's stall samples).
		auto fill = [&](int kf, int buf) {
			const float2 *rIn[2][8], *rPv[2][8];
#pragma unroll
			for (int it = 0; it < 8; ++it) {
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					rIn[h][it] = U.rowIn[h][fillF + 4 * it];
					rPv[h][it] = U.rowPv[h][fillF + 4 * it];
				}
			}
#pragma unroll
			for (int it = 0; it < 8; ++it) {
				const int fl = fillF + 4 * it;
				const int q = kf + fillI - G * fl;
				if (fl < nAct && (unsigned)q < (unsigned)K) {
#pragma unroll
					for (int h = 0; h < 2; ++h) { // slot = {A.re, A.im, B.re, B.im}
						cp_async8((float2 *)&U.in[q & (CH3_RING - 1)][fl] + h, rIn[h][it] + q);
						cp_async8((float2 *)&U.pvy[buf][fillI][fl] + h, rPv[h][it] + q);
					}
				}
			}
			if (lane < 2 * CHAIN_CH) { // lane 0's predecessor: planar state / previous group rows -> {c0, c1} slots
				const int qq = kf + (lane >> 1), c = lane & 1;
				if (qq < K) {
					cp_async8((float2 *)&U.p0Out[buf][lane >> 1] + c, (c ? prevOut[1] : prevOut[0]) + qq);
					if (base == 0) {
						cp_async4((float *)&U.p0E[buf][lane >> 1] + c, (c ? prevE[1] : prevE[0]) + qq);
					} else {
						((float *)&U.p0E[buf][lane >> 1])[c] = xnorm((c ? prevIn[1] : prevIn[0])[qq]);
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
				const int q = k0 + i - G * lane;
				const int b = q - LT - 1;
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
				const c2 inq = ld_c2d(&U.in[q & (CH3_RING - 1)][lane]);
				c2 pv = ld_c2d(&U.pvy[cb][i][lane]);
				c2 lo2, hi2, lo1, hi1;
				if constexpr (!FAR) {
					lo2 = sel_c2(INTERIOR || l2 >= 0, ld_c2d(&U.in[l2 & (CH3_RING - 1)][lane]));
					hi2 = sel_c2(INTERIOR || l2 >= -1, ld_c2d(&U.in[(l2 + 1) & (CH3_RING - 1)][lane]));
					lo1 = sel_c2(INTERIOR || l1 >= 0, ld_c2d(&U.in[l1 & (CH3_RING - 1)][lane]));
					hi1 = sel_c2(INTERIOR || l1 >= -1, ld_c2d(&U.in[(l1 + 1) & (CH3_RING - 1)][lane]));
				} else { // extreme stretch (> 2x): gather straight from the spectrum row
					lo2 = gather2(myIn, l2, K);
					hi2 = gather2(myIn, l2 + 1, K);
					lo1 = gather2(myIn, l1, K);
					hi1 = gather2(myIn, l1 + 1, K);
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
				// short twist at b+1 (:751,:771): Prediction::input[b+1] is inF[1]
				const c2 t1N = xmulc2(inF[1], xlerp2p(lo1, hi1, f1s, one), one);
				{
					const float2 rn = xmul(rotq, rotS);
					rotq = make_float2((INTERIOR || q >= 0) ? rn.x : rotq.x, (INTERIOR || q >= 0) ? rn.y : rotq.y);
				}
				// ---- the FIFO heads belong to bin b; the new preliminary values (bin q = b+L+1) enter after the main
				//      prediction below, which only reads entries computed in EARLIER steps
				const f2 eB = eF[0];
				const c2 t2B = t2F[0], inB = inF[0];
				const c2 preN = pre[1], preL = pre[LT], t2L = t2F[LT]; // prelim output at b+1, b+L; long twist at b+L
#pragma unroll
				for (int u = 0; u + 1 < NF; ++u) {
					pre[u] = pre[u + 1];
					eF[u] = eF[u + 1];
					t2F[u] = t2F[u + 1];
					inF[u] = inF[u + 1];
				}
				pre[NF - 1] = newPre;
				eF[NF - 1] = newE;
				t2F[NF - 1] = newT2;
				inF[NF - 1] = newIn;
				// ---- main prediction at bin b (:727-800): the louder channel (first on ties, :733) leads
				// the phase sum of :754-784 is formed for both channels at once (packed, each channel from its own
				// registers, exactly as if it were the leader) and the leader's is picked afterwards
				c2 ph2 = zc;
				ph2 = ph2 + sel_c2(INTERIOR || b > 0, xmul2(oh[0], t1P, one));                      // :754
				ph2 = ph2 + sel_c2(INTERIOR || b >= LT, xmul2(oh[LT - 1], t2B, one));               // :761
				ph2 = ph2 + sel_c2(INTERIOR || b < K - 1, xmulc2(preN, t1N, one));                  // :774
				ph2 = ph2 + sel_c2(INTERIOR || b < K - LT, xmulc2(preL, t2L, one));                // :784
				const c2 oc = make_output_q2(ph2, eB, inB, one); // :788, each half its own (mono) stream
				// unconditional: out-of-range steps only produce values that every consumer masks
#pragma unroll
				for (int u = LT - 1; u > 0; --u) oh[u] = oh[u - 1];
				oh[0] = oc;
				lastFinal = oc;
				lastE = eB;
				t1P = t1N;
				U.pvy[cb][i][lane] = make_float4(f2_lo(oc.re), f2_lo(oc.im), f2_hi(oc.re), f2_hi(oc.im));
				if (lastFrame && (unsigned)b < (unsigned)K) {
					eRow[0][b] = f2_lo(eB);
					eRow[1][b] = f2_hi(eB);
				}
			};
			// FAST arithmetic (see the top of the file): same data flow, fused operations, re-associated phase sum
			auto step_fast = [&](const int i, auto farTag, auto intTag) {
				constexpr bool FAR = decltype(farTag)::value, INTERIOR = decltype(intTag)::value;
				const int q = k0 + i - G * lane;
				const int b = q - LT - 1;
				const bool qIn = INTERIOR || (active && (unsigned)q < (unsigned)K);
				const float i2 = fsub((float)q, longTf);
				const int l2 = (int)floorf(i2);
				const float f2s = fsub(i2, (float)l2);
				const float i1 = fsub((float)(b + 1), tf);
				const int l1 = (int)floorf(i1);
				const float f1s = fsub(i1, (float)l1);
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
				const c2 inq = ld_c2d(&U.in[q & (CH3_RING - 1)][lane]);
				c2 pv = ld_c2d(&U.pvy[cb][i][lane]);
				c2 lo2, hi2, lo1, hi1;
				if constexpr (!FAR) {
					lo2 = sel_c2(INTERIOR || l2 >= 0, ld_c2d(&U.in[l2 & (CH3_RING - 1)][lane]));
					hi2 = sel_c2(INTERIOR || l2 >= -1, ld_c2d(&U.in[(l2 + 1) & (CH3_RING - 1)][lane]));
					lo1 = sel_c2(INTERIOR || l1 >= 0, ld_c2d(&U.in[l1 & (CH3_RING - 1)][lane]));
					hi1 = sel_c2(INTERIOR || l1 >= -1, ld_c2d(&U.in[(l1 + 1) & (CH3_RING - 1)][lane]));
				} else {
					lo2 = gather2(myIn, l2, K);
					hi2 = gather2(myIn, l2 + 1, K);
					lo1 = gather2(myIn, l1, K);
					hi1 = gather2(myIn, l1 + 1, K);
				}
				pv = fmul_s(pv, rotq); // :653-654
				ro = fmul_s(ro, rotq);
				const f2 e = fnorm2(inq);                      // :679
				const c2 ph0 = fmul_c(ro, fmulc_c(inq, pv));     // :714-715
				const f2 den = f2_make(fmaxf(f2_lo(re), f2_lo(e)), fmaxf(f2_hi(re), f2_hi(e))) + f2_make(B200S_NOISE_FLOOR, B200S_NOISE_FLOOR);
				const f2 rden = f2_make(rcp_fast(f2_lo(den)), rcp_fast(f2_hi(den)));
				const c2 newPre = sel_c2(qIn, c2{mul2(ph0.re, rden), mul2(ph0.im, rden)}); // :716
				const f2 newE = sel_f2(qIn, e);
				const c2 newIn = sel_c2(qIn, inq);
				const c2 newT2 = sel_c2(qIn, fmulc_c(inq, flerp2(lo2, hi2, f2s))); // long twist at q (:758)
				const c2 t1N = fmulc_c(inF[1], flerp2(lo1, hi1, f1s));             // short twist at b+1 (:751,:771)
				{
					const float2 rn = xmul(rotq, rotS); // the table recurrence stays in the reference's own arithmetic
					rotq = make_float2((INTERIOR || q >= 0) ? rn.x : rotq.x, (INTERIOR || q >= 0) ? rn.y : rotq.y);
				}
				const f2 eB = eF[0];
				const c2 t2B = t2F[0], inB = inF[0];
				const c2 preN = pre[1], preL = pre[LT], t2L = t2F[LT];
#pragma unroll
				for (int u = 0; u + 1 < NF; ++u) {
					pre[u] = pre[u + 1];
					eF[u] = eF[u + 1];
					t2F[u] = t2F[u + 1];
					inF[u] = inF[u + 1];
				}
				pre[NF - 1] = newPre;
				eF[NF - 1] = newE;
				t2F[NF - 1] = newT2;
				inF[NF - 1] = newIn;
				// ---- main prediction at bin b (:727-800); the terms that do not depend on the previous bin first
				c2 ph2;
				if constexpr (INTERIOR) {
					ph2 = fmul_c(oh[LT - 1], t2B);        // :761
					ph2 = fmulc_acc(ph2, preN, t1N);      // :774
					ph2 = fmulc_acc(ph2, preL, t2L);      // :784
					ph2 = fmul_acc(ph2, oh[0], t1P);      // :754
				} else {
					ph2 = sel_c2(b >= LT, fmul_c(oh[LT - 1], t2B));
					ph2 = ph2 + sel_c2(b < K - 1, fmulc_c(preN, t1N));
					ph2 = ph2 + sel_c2(b < K - LT, fmulc_c(preL, t2L));
					ph2 = ph2 + sel_c2(b > 0, fmul_c(oh[0], t1P));
				}
				const float2 oA = make_output_fast(pick(false, ph2), f2_lo(eB), pick(false, inB)); // :788, each half its own (mono) stream
				const float2 oB = make_output_fast(pick(true, ph2), f2_hi(eB), pick(true, inB));
				const c2 oc = c2{f2_make(oA.x, oB.x), f2_make(oA.y, oB.y)};
#pragma unroll
				for (int u = LT - 1; u > 0; --u) oh[u] = oh[u - 1];
				oh[0] = oc;
				lastFinal = oc;
				lastE = eB;
				t1P = t1N;
				U.pvy[cb][i][lane] = make_float4(f2_lo(oc.re), f2_lo(oc.im), f2_hi(oc.re), f2_hi(oc.im));
				if (lastFrame && (unsigned)b < (unsigned)K) {
					eRow[0][b] = f2_lo(eB);
					eRow[1][b] = f2_hi(eB);
				}
			};
			// unrolled by 4 (for L = 4 the register FIFOs rotate by pure renaming) so that the hot loop stays in the
			// instruction cache; branch-free inside
			auto run_chunk = [&](auto farTag, auto intTag) {
#pragma unroll 1
				for (int h = 0; h < CHAIN_CH; h += 4) { // (measured: unrolling the whole chunk gains nothing, 1.69 vs 1.67 ms)
#pragma unroll
					for (int u = 0; u < 4; ++u) {
						if constexpr (FAST) step_fast(h + u, farTag, intTag);
						else step(h + u, farTag, intTag);
					}
				}
			};
			if (farAny) run_chunk(std::true_type{}, std::false_type{});
			else if (k0 >= interiorFrom && k0 + CHAIN_CH <= K) run_chunk(std::false_type{}, std::true_type{});
			else run_chunk(std::false_type{}, std::false_type{});
			__syncwarp();
			// ---------------- write the chunk's finals back: planar Band::output rows, 32 B per row and quarter-warp;
			//                  all tile reads first, then the stores (row addresses are arithmetic) ----------------
			{
				float4 v[8];
#pragma unroll
				for (int it = 0; it < 8; ++it) v[it] = U.pvy[cb][fillI][fillF + 4 * it];
#pragma unroll
				for (int it = 0; it < 8; ++it) {
					const int fl = fillF + 4 * it;
					const int b = k0 + fillI - G * fl - LT - 1;
					if (fl < nAct && (unsigned)b < (unsigned)K) {
						yBase[0][(size_t)fl * K + b] = make_float2(v[it].x, v[it].y);
						yBase[1][(size_t)fl * K + b] = make_float2(v[it].z, v[it].w);
					}
				}
			}
			__syncwarp();
		}
	}
}

static inline size_t smem_chain_dual() { return sizeof(ChainDualTiles); }

} // namespace b200s
