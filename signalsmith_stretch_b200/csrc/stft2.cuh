// stft2.cuh -- analysis and synthesis kernels for the preset FFT sizes on the PAIRED in-place FFT
// (fft2.cuh).  Same results as k_analyse / k_synth (kernels.cuh), which stay for generic sizes.
//
//   k_analyse2  windowed modified real FFT of TWO analyses at a time (dependency analyseStep,
//               reference call sites signalsmith-stretch.h:337,359).  Persistent CTAs walk the
//               (stream, job pair) items of the call; the raw samples of the NEXT pair are copied to
//               shared memory with 16-byte cp.async while the current FFT runs.
//   k_synth2    inverse FFT of TWO consecutive blocks of one stream-channel at a time, synthesis
//               window and overlap-add (synthesiseStep / readOutput / moveOutput, :397-414).  One
//               sweep over the pending ring applies, per ring slot and in the reference's order,
//               block A's contribution, the emission of the samples between the two blocks, and
//               block B's contribution -- one read-modify-write per slot instead of three.
#pragma once
#include "fft2.cuh"
#include "kernels.cuh"

namespace b200s {


__host__ __device__ __forceinline__ int stage_len(int B) { return (B + 8 + 3) & ~3; }

// streaming accesses (read once / written once): keep them out of L1, which -- next to two 95 KB CTAs -- is only
// large enough for the window / twiddle tables that every item re-reads
#ifdef B200S_EMU
__device__ __forceinline__ float2 ld_stream(const float2 *p) { return *p; }
__device__ __forceinline__ void st_stream(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_stream4(float *p, float4 v) { *(float4 *)p = v; }
#else
__device__ __forceinline__ void st_stream4(float *p, float4 v) { __stcs((float4 *)p, v); }
__device__ __forceinline__ float2 ld_stream(const float2 *p) {
	float2 v;
	asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
#endif

// ---------------------------------------------------------------------------------------------
// k_analyse2: grid = persistent CTAs (a multiple of the SM count), 256 threads.
// dyn smem: PairGeo::LEN float4 (FFT pair) + 2 * stage_len(B) floats (raw samples of the next pair).
// ---------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256, 2) k_analyse2(Ctx x) {
	using G = PairGeo<KT>;
	const Cfg &g = x.cfg;
	B200S_DYN_SHARED
	float4 *buf = (float4 *)dyn_smem;
	const int B = g.B, o = g.o, tid = threadIdx.x;
	float *stA = (float *)(buf + G::LEN), *stB = stA + stage_len(B);
	const PairTw tw = pair_tw_dit<KT>(x.twiddle, tid);
	const int PJ = g.C * x.maxFrames;         // pair slots per stream (2*C*maxFrames jobs)
	const long long total = (long long)x.sCount * PJ;

	struct Item {
		int s, hasB;
		Job a, b;
	};
	// item -> jobs; false when the stream has fewer pairs than slots
	// (the three loads are independent of each other: the job slots exist even when the stream has fewer jobs)
	// fetch_raw only issues the loads (m.hasB temporarily holds the stream's job count); fetch_done turns them into
	// an item.  Keeping the two apart lets the loads of the item after next fly for a whole iteration.
	auto fetch_raw = [&](long long it, Item &m) {
		m.s = x.sBase + (int)(it / PJ);
		const Job *jb = x.jobs + (size_t)m.s * 2 * g.C * x.maxFrames + 2 * (int)(it % PJ);
		m.hasB = x.call[m.s].nJobs;
		m.a = jb[0];
		m.b = jb[1];
	};
	auto fetch_done = [&](long long it, Item &m) -> bool {
		const int p = (int)(it % PJ), nJ = m.hasB;
		m.hasB = 2 * p + 1 < nJ;
		if (!m.hasB) m.b = m.a;
		return 2 * p < nJ;
	};
	auto fetch = [&](long long it, Item &m) -> bool {
		fetch_raw(it, m);
		return fetch_done(it, m);
	};
	auto next_valid = [&](long long it, Item &m) -> long long {
		while (it < total && !fetch(it, m)) it += gridDim.x;
		return it;
	};
	// asynchronously copy the B samples (history ++ input) of one job into a staging buffer.
	// Aligned mode copies the enclosing 4-aligned range (sample i lands at st[i + (start & 3)]) with at most two BULK
	// copies issued by one thread -- the part that lies in the history and the part that lies in this call's input --
	// which complete on the CTA's mbarrier (two arrivals per item: one per job); whatever lies outside both (before the
	// history, after the input) is zero-filled by all threads.  (The per-16-byte cp.async loop this replaces was 17 % of
	// the kernel's instructions, profiles/r01_v12.)
	B200S_SHARED unsigned long long stageBar;
	if (tid == 0) mbar_init(&stageBar, 2);
	fence_async_proxy();
	__syncthreads();
	unsigned stagePar = 0u;
	auto stage_job = [&](float *st, int s, const Job &j, int t0, int nT) {
		const float *ib = x.in + (size_t)s * x.inStreamStride + (size_t)j.c * x.inChanStride;
		const float *he = x.histCur + ((size_t)s * g.C + j.c) * g.histLen + g.histLen;
		if (x.inAligned) {
			const int sh = j.start & 3, a00 = j.start - sh, nCh = (B + sh + 3) >> 2;
			// chunk q holds samples a00 + 4q ..: [qh0, z) comes from the history, [z, qi1) from the input
			const int z = min(max((-a00) >> 2, 0), nCh);
			const int qh0 = min(max((-g.histLen - a00) >> 2, 0), z), qi1 = min(max((x.nIn - a00) >> 2, z), nCh);
			if (tid == t0) {
				fence_async_proxy(); // the staging buffer was last read through the generic proxy
				mbar_expect(&stageBar, 16u * (unsigned)(qi1 - qh0));
				if (z > qh0) bulk_g2s(st + 4 * qh0, he + a00 + 4 * qh0, 16u * (unsigned)(z - qh0), &stageBar);
				if (qi1 > z) bulk_g2s(st + 4 * z, ib + a00 + 4 * z, 16u * (unsigned)(qi1 - z), &stageBar);
			}
			for (int q = tid; q < qh0; q += 256) *(float4 *)(st + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
			for (int q = qi1 + tid; q < nCh; q += 256) *(float4 *)(st + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
		} else {
			for (int i = tid; i < B; i += nT) {
				const int a = j.start + i;
				if (a < x.nIn && a >= -g.histLen) cp_async4(st + i, a >= 0 ? ib + a : he + a);
				else st[i] = 0.f;
			}
		}
	};
	// Items are looked up TWO iterations ahead: the descriptor loads of the item after next are issued while the
	// current pair is transformed and are only consumed at the end of the iteration, so that the copies of the next
	// item's input can start the moment the staging buffers are free.  (Measured alternative, profiles/r01_v12: letting
	// the two warps that idle during the radix-16 passes do the look-up and the copies was slower -- the other six
	// waited for them at the next barrier.)
	Item cur, nxt;
	long long item = next_valid(blockIdx.x, cur);
	if (item < total) {
		stage_job(stA, cur.s, cur.a, 0, 256);
		stage_job(stB, cur.s, cur.b, 32, 256);
	}
	long long nitem = item < total ? next_valid(item + gridDim.x, nxt) : total;
	while (item < total) {
		float2 *dstA = x.spec + ((size_t)cur.s * 2 * x.maxFrames * g.C + cur.a.row) * g.K;
		float2 *dstB = x.spec + ((size_t)cur.s * 2 * x.maxFrames * g.C + cur.b.row) * g.K;
		// window samples and pre-twiddles of this thread's R3 elements: issued before the wait so that their latency overlaps it.
		// (The values are the same for every pair; `wofs` is opaque to the compiler so that they are re-loaded,
		// L1/L2 hits, instead of being hoisted into 2*R3 permanently live registers.)
		float4 tabr[G::R3]; // {window[n+o], window[n+o-K], pre-twiddle exp(-i*pi*n/N)} of element n = tid + 256*it
		{
			int wofs = 0;
#ifndef B200S_EMU
			asm volatile("" : "+r"(wofs));
#endif
			static_for<G::R3>([&](auto itc) {
				constexpr int it = decltype(itc)::value;
				tabr[it] = __ldg(x.anaTab + tid + 256 * it + wofs);
			});
		}
		cp_async_wait_all();
		if (x.inAligned) {
			mbar_wait(&stageBar, stagePar);
			stagePar ^= 1u;
		}
		__syncthreads(); // staging complete and visible; previous unpack finished with buf
#ifdef B200S_EMU_EXACT_FFT // test builds only: swap in the oracle's double FFT to isolate the non-FFT logic
		{
			const int shA = x.inAligned ? (cur.a.start & 3) : 0, shB = x.inAligned ? (cur.b.start & 3) : 0;
			if (tid == 0) {
				float *xw = (float *)buf;
				// interleaved mode: transform into scratch rows (the Y rows of block 0: the chain writes them only later), then interleave
				float2 *tA = x.specIl ? x.Y + coef_off(x, cur.s, 0, 0) : dstA, *tB = x.specIl ? tA + KT : dstB;
				for (int i = 0; i < B; ++i) xw[i] = fmul(stA[i + shA], x.window[i]);
				emu_exact_forward(xw, B, o, 2 * KT, tA);
				if (cur.hasB) {
					for (int i = 0; i < B; ++i) xw[i] = fmul(stB[i + shB], x.window[i]);
					emu_exact_forward(xw, B, o, 2 * KT, tB);
				}
				if (x.specIl) {
					float4 *d4 = (float4 *)x.spec + ((size_t)cur.s * 2 * x.maxFrames + (cur.a.row >> 1)) * g.K;
					for (int b = 0; b < KT; ++b) d4[b] = make_float4(tA[b].x, tB[b].x, tA[b].y, tB[b].y);
				}
			}
			__syncthreads();
			item = nitem;
			cur = nxt;
			if (item < total) {
				stage_job(stA, cur.s, cur.a, 0, 256);
				stage_job(stB, cur.s, cur.b, 32, 256);
				nitem = next_valid(item + gridDim.x, nxt);
			}
			continue;
		}
#endif
		{ // ---- load stage fused with the first FFT pass: window, wrap-sign fold, half-bin
		  //      pre-twiddle of the thread's R3 samples n = tid + 256*n3, radix-R3 butterfly over n3 in registers
			const int shA = x.inAligned ? (cur.a.start & 3) : 0, shB = x.inAligned ? (cur.b.start & 3) : 0;
			c2 v[G::R3];
			static_for<G::R3>([&](auto itc) {
				constexpr int it = decltype(itc)::value;
				const int n = tid + 256 * it, i0 = n + o, i1 = n + o - KT;
				const bool in0 = i0 < B, in1 = i1 >= 0;
				const float w0 = tabr[it].x, w1 = tabr[it].y;
				const float a0 = in0 ? stA[i0 + shA] : 0.f, a1 = in1 ? stA[i1 + shA] : 0.f;
				const float b0 = in0 ? stB[i0 + shB] : 0.f, b1 = in1 ? stB[i1 + shB] : 0.f;
				const float2 pw = make_float2(tabr[it].z, tabr[it].w); // exp(-i*pi*n/N), exactly rounded table entry
				const c2 t = c2{muls(f2_make(a0, b0), w0), muls(f2_make(a1, b1), w1)}; // (x0*w0) + i*(x1*w1)
				v[it] = cmulw(t, pw.x, pw.y);
			});
			PairDFT<G::R3, false>::run(v);
			float4 *p = buf + (tid & 15) * G::P1 + (tid >> 4) * G::P2;
			static_for<G::R3>([&](auto qc) { st_c2(p + decltype(qc)::value, v[decltype(qc)::value]); });
		}
		__syncthreads(); // staging consumed, buf complete
		Item n2;
		long long cand = nitem + gridDim.x;
		bool ok2 = false;
		if (nitem < total) { // start the next pair's input on its way while this FFT runs
			stage_job(stA, nxt.s, nxt.a, 0, 256);
			stage_job(stB, nxt.s, nxt.b, 32, 256);
			if (cand < total) fetch_raw(cand, n2); // descriptor of the item after next: loads in flight until the end of the iteration
		}
		pair_dit_stage_b<KT>(buf, tw, tid);
		__syncthreads();
		if (tid < G::M1) { // ---- last pass fused with the store: thread holds X[tid + M1*q], q = 0..15.
			// bin b = Z[b/2] (b even) or conj(Z[K-1-b/2]) (b odd): Z[k] with k < K/2 (q < 8) is bin 2k, the rest bin 2(K-1-k)+1
			c2 v[16];
			pair_dit_stage_c<KT>(buf, tw, tid, v);
			const bool hasB = cur.hasB, il = x.specIl;
			// stereo direct path: the pair is (channel 0, channel 1) of one analysis and is stored interleaved
			float4 *dst4 = (float4 *)x.spec + ((size_t)cur.s * 2 * x.maxFrames + (cur.a.row >> 1)) * g.K;
			static_for<16>([&](auto qc) {
				constexpr int q = decltype(qc)::value;
				const int k = tid + G::M1 * q;
				const int b = q < 8 ? 2 * k : 2 * (KT - 1 - k) + 1;
				const float sg = q < 8 ? 1.f : -1.f;
				const float ra = f2_lo(v[q].re), rb = f2_hi(v[q].re), ia = sg * f2_lo(v[q].im), ib = sg * f2_hi(v[q].im);
				if (il) {
					dst4[b] = make_float4(ra, rb, ia, ib);
				} else {
					dstA[b] = make_float2(ra, ia);
					if (hasB) dstB[b] = make_float2(rb, ib);
				}
			});
		}
		item = nitem;
		cur = nxt;
		if (nitem < total) {
			if (cand < total) ok2 = fetch_done(cand, n2);
			if (cand < total && !ok2) cand = next_valid(cand + gridDim.x, n2); // rare: a stream with fewer pairs than slots
			nitem = cand < total ? cand : total;
			nxt = n2;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// k_synth2: grid (C, S), 256 threads, one CTA per stream-channel; blocks of the call two at a time.
// dyn smem: PairGeo::LEN float4 (FFT pair) + pendLen floats (pending ring) + pendLen floats
// (windowProducts ring; kept per channel so that no two CTAs share writable state).
// ---------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256, 2) k_synth2(Ctx x) {
	using G = PairGeo<KT>;
	const Cfg &g = x.cfg;
	B200S_DYN_SHARED
	float4 *buf = (float4 *)dyn_smem;
	float *pend = (float *)(buf + G::LEN), *wp = pend + g.pendLen;
	const int c = blockIdx.x, s = x.sBase + blockIdx.y, tid = threadIdx.x;
	const Call cl = x.call[s];
	float *out = x.out + (size_t)s * x.outStreamStride + (size_t)c * x.outChanStride;
	if (cl.bypass) { // :252-267
		const float *in = x.in + (size_t)s * x.inStreamStride + (size_t)c * x.inChanStride;
		for (int i = tid; i < x.nOut; i += 256) out[i] = x.nIn > 0 ? in[i % x.nIn] : 0.f;
		return;
	}
	const int P = g.pendLen, B = g.B, o = g.o, addOff = g.addOff;
	const float fN = (float)g.N;
	float *gp = x.pend + ((size_t)s * g.C + c) * P, *gw = x.pendWp + ((size_t)s * g.C + c) * P;
	for (int i = tid; i < P; i += 256) {
		pend[i] = gp[i];
		wp[i] = gw[i];
	}
	const PairTw tw = pair_tw_dif<KT>(x.twiddle, tid);
	const Frame *frames = x.frames + (size_t)s * x.maxFrames;
	__syncthreads();
	int head = 0, emitted = 0;
	// emit n samples from the ring head, zero them behind (:408-414)
	// VECTOR form of the ring loops: when the ring head, the counts and the output row are multiples of four floats
	// (the presets at 48 kHz with chunks that are multiples of four), four consecutive ring slots are one 16-byte
	// shared-memory access each and never straddle the wrap; element for element the same operations in the same order.
	const bool ioVec = ((uintptr_t)out & 15) == 0 && ((P | B | addOff) & 3) == 0;
	auto emit = [&](int n) {
		if (ioVec && ((head | n | emitted) & 3) == 0 && n <= P) {
			for (int i = 4 * tid; i < n; i += 1024) {
				int p = head + i;
				if (p >= P) p -= P;
				const float4 a = *(float4 *)(pend + p), w = *(float4 *)(wp + p);
				st_stream4(out + emitted + i, make_float4(fdiv(a.x, w.x), fdiv(a.y, w.y), fdiv(a.z, w.z), fdiv(a.w, w.w)));
				*(float4 *)(pend + p) = make_float4(0.f, 0.f, 0.f, 0.f);
				*(float4 *)(wp + p) = make_float4(B200S_ALMOST_ZERO, B200S_ALMOST_ZERO, B200S_ALMOST_ZERO, B200S_ALMOST_ZERO);
			}
			head = (head + n) % P;
			emitted += n;
			return;
		}
		for (int i = tid; i < n; i += 256) {
			float v = 0.f;
			if (i < P) {
				int p = head + i;
				if (p >= P) p -= P;
				v = fdiv(pend[p], wp[p]);
				pend[p] = 0.f;
				wp[p] = B200S_ALMOST_ZERO;
			}
			st_stream(out + emitted + i, v);
		}
		head = (head + (n < P ? n : P)) % P; // n >= P leaves an all-clear ring; any head is fine
		emitted += n;
	};
	for (int f = 0; f < cl.nFrames; f += 2) {
		const bool hasB = f + 1 < cl.nFrames;
		const int tA = frames[f].t;
		// samples emitted between the two blocks; without a second block nothing is emitted inside the sweep
		const int gap = hasB ? frames[f + 1].t - tA : 0;
		const float2 *YA = x.Y + coef_off(x, s, f, c), *YB = x.Y + coef_off(x, s, hasB ? f + 1 : f, c);
#ifndef B200S_EMU
		// the spectra of the NEXT pair of blocks start their way from HBM to L2 now (two rows of 8*K bytes, one 128-byte
		// line per thread and step), so that the loads at the top of the next iteration find them there
		if (f + 2 < cl.nFrames) {
			const char *nA = (const char *)(x.Y + coef_off(x, s, f + 2, c)), *nB = (const char *)(x.Y + coef_off(x, s, min(f + 3, cl.nFrames - 1), c));
			for (int o2 = tid * 128; o2 < KT * 8; o2 += 256 * 128) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(nA + o2));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(nB + o2));
			}
		}
#endif
#ifndef B200S_EMU_EXACT_FFT
		// ---- first inverse-FFT pass takes its inputs straight from HBM (in flight while the ring is being emitted):
		//      Z'[k] = Y[2k] for k < K/2, conj(Y[2(K-1-k)+1]) otherwise; thread tid < M1 needs k = q*M1 + tid, q = 0..15
		c2 v[16];
		if (tid < G::M1) {
			static_for<16>([&](auto qc) {
				constexpr int q = decltype(qc)::value;
				const int k = q * G::M1 + tid;
				const int b = q < 8 ? 2 * k : 2 * (KT - 1 - k) + 1;
				const float sg = q < 8 ? 1.f : -1.f;
				const float2 a = ld_stream(YA + b), bb = ld_stream(YB + b);
				v[q] = c2{f2_make(a.x, bb.x), f2_make(sg * a.y, sg * bb.y)};
			});
		}
#endif
		emit(tA - emitted);
#ifdef B200S_EMU_EXACT_FFT // test builds only
		float *yTimeA = (float *)buf, *yTimeB = yTimeA + B;
		__syncthreads();
		if (tid == 0) {
			emu_exact_inverse(YA, B, o, g.N, yTimeA);
			if (hasB) emu_exact_inverse(YB, B, o, g.N, yTimeB);
			for (int i = 0; i < B; ++i) { // the synthesis window is applied where the block is written (see below)
				yTimeA[i] = fmul(yTimeA[i], x.window[i]);
				if (hasB) yTimeB[i] = fmul(yTimeB[i], x.window[i]);
			}
		}
		__syncthreads();
#else
		float *yTimeA = (float *)buf, *yTimeB = yTimeA + B; // the FFT buffer is reused for the real time-domain blocks
		if (tid < G::M1) pair_dif_pass1<KT>(buf, tw, tid, v);
		__syncthreads(); // buf complete; the emission above is complete too
		{
			// last pass in registers, then the half-bin post-twiddle: time sample i of the block is
			// 2*Re(z[i-o] * conj(pre[i-o])) for i >= o and 2*Im(z[i-o+K] * conj(pre[i-o+K])) for i < o,
			// so z[n2] yields sample n2+o (if < B) from its real part and sample n2+o-K (if >= 0) from its imaginary part
			c2 z[G::R3];
			pair_dif_pass23<KT>(buf, tw, tid, z);
			__syncthreads(); // every thread has its outputs in registers: the buffer can take the real samples
			static_for<G::R3>([&](auto qc) {
				constexpr int q3 = decltype(qc)::value;
				const int n2 = (tid & 15) + 16 * (tid >> 4) + 256 * q3;
				// one 16-byte table entry per element: {window[n2+o], window[n2+o-K], pre-twiddle}; the synthesis window
				// (:397-399, y * window) is applied here, where the index is fixed per thread, not in the ring sweep
				const float4 tb = __ldg(x.anaTab + n2);
				const int iRe = n2 + o, iIm = n2 + o - KT;
				if (iRe < B) {
					const f2 y = muls(muls(fmas(z[q3].im, tb.w, muls(z[q3].re, tb.z)), 2.f), tb.x);
					yTimeA[iRe] = f2_lo(y);
					yTimeB[iRe] = f2_hi(y);
				}
				if (iIm >= 0) {
					const f2 y = muls(muls(fmas(z[q3].re, -tb.w, muls(z[q3].im, tb.z)), 2.f), tb.y);
					yTimeA[iIm] = f2_lo(y);
					yTimeB[iIm] = f2_hi(y);
				}
			});
			__syncthreads();
		}
#endif
		// ---- one sweep over the ring: slot j ahead of the head gets, in the reference's order,
		//      block A's sample j-addOff, the emission if j < gap, then block B's sample (relative to the new head)
		//      (yTimeA / yTimeB already carry the synthesis window).  Two slots per iteration, loads first, branch-free.
		if (ioVec && ((head | gap | emitted) & 3) == 0) {
			const float *winProd = x.winProd;
			auto add4 = [](float4 a, float4 b) { return make_float4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w)); };
#pragma unroll 2
			for (int j = 4 * tid; j < P; j += 1024) {
				int p = head + j;
				if (p >= P) p -= P;
				float4 pv = *(float4 *)(pend + p), wv = *(float4 *)(wp + p);
				if (j >= addOff) { // block A (:397-399): i = j - addOff in [0, B)
					const int iA = j - addOff;
					pv = add4(pv, *(const float4 *)(yTimeA + iA));
					wv = add4(wv, __ldg((const float4 *)(winProd + iA)));
				}
				if (hasB) {
					int iB = j - gap - addOff; // block B's samples for these slots, relative to the head after the emission
					if (j < gap) { // emitted between the blocks (:408-414)
						st_stream4(out + emitted + j, make_float4(fdiv(pv.x, wv.x), fdiv(pv.y, wv.y), fdiv(pv.z, wv.z), fdiv(pv.w, wv.w)));
						pv = make_float4(0.f, 0.f, 0.f, 0.f);
						wv = make_float4(B200S_ALMOST_ZERO, B200S_ALMOST_ZERO, B200S_ALMOST_ZERO, B200S_ALMOST_ZERO);
						iB += P;
					}
					if (iB >= 0 && iB < B) {
						pv = add4(pv, *(const float4 *)(yTimeB + iB));
						wv = add4(wv, __ldg((const float4 *)(winProd + iB)));
					}
				}
				*(float4 *)(pend + p) = pv;
				*(float4 *)(wp + p) = wv;
			}
		} else {
			const float *winProd = x.winProd;
			auto slot = [&](int j, float &pv, float &wv, float &ya, float &wa, float &yb, float &wb, int &p, bool &okA, bool &okB) {
				p = head + j;
				if (p >= P) p -= P;
				pv = pend[p];
				wv = wp[p];
				const int iA = j - addOff;
				okA = j >= addOff; // block A (:397-399): i = j - addOff in [0, B)
				ya = okA ? yTimeA[iA] : 0.f;
				wa = okA ? __ldg(winProd + iA) : 0.f;
				int iB = j - gap - addOff; // block B's sample for this slot, relative to the head after the emission
				if (j < gap) iB += P;
				okB = hasB && iB >= 0 && iB < B;
				yb = okB ? yTimeB[iB] : 0.f;
				wb = okB ? __ldg(winProd + iB) : 0.f;
			};
			auto finish = [&](int j, float pv, float wv, float ya, float wa, float yb, float wb, int p, bool okA, bool okB) {
				if (okA) {
					pv = fadd(pv, ya);
					wv = fadd(wv, wa);
				}
				if (hasB && j < gap) { // emitted between the blocks (:408-414)
					st_stream(out + emitted + j, fdiv(pv, wv));
					pv = 0.f;
					wv = B200S_ALMOST_ZERO;
				}
				if (okB) {
					pv = fadd(pv, yb);
					wv = fadd(wv, wb);
				}
				pend[p] = pv;
				wp[p] = wv;
			};
			for (int j = tid; j < P; j += 512) {
				float pv0, wv0, ya0, wa0, yb0, wb0, pv1 = 0.f, wv1 = 0.f, ya1 = 0.f, wa1 = 0.f, yb1 = 0.f, wb1 = 0.f;
				int p0, p1 = 0;
				bool a0, b0, a1 = false, b1 = false;
				const bool two = j + 256 < P;
				slot(j, pv0, wv0, ya0, wa0, yb0, wb0, p0, a0, b0);
				if (two) slot(j + 256, pv1, wv1, ya1, wa1, yb1, wb1, p1, a1, b1);
				finish(j, pv0, wv0, ya0, wa0, yb0, wb0, p0, a0, b0);
				if (two) finish(j + 256, pv1, wv1, ya1, wa1, yb1, wb1, p1, a1, b1);
			}
		}
		if (hasB) {
			// gap <= P: inside a call k_plan triggers block f+1 exactly H output samples after block f (samplesSinceLast
			// restarts at 0 on a trigger and a block fires the moment it reaches H, :281-286), and H <= B <= P (configure)
			head = (head + gap) % P;
			emitted += gap;
		}
		__syncthreads();
	}
	emit(x.nOut - emitted);
	__syncthreads();
	for (int i = tid; i < P; i += 256) {
		int p = head + i;
		if (p >= P) p -= P;
		gp[i] = pend[p];
		gw[i] = wp[p];
	}
}

static inline size_t smem_analyse2(const Cfg &g) {
	const int len = g.K == 3072 ? PairGeo<3072>::LEN : PairGeo<2560>::LEN;
	return sizeof(float4) * len + sizeof(float) * 2 * stage_len(g.B);
}
static inline size_t smem_synth2(const Cfg &g) {
	const int len = g.K == 3072 ? PairGeo<3072>::LEN : PairGeo<2560>::LEN;
	return sizeof(float4) * len + sizeof(float) * 2 * g.pendLen;
}

} // namespace b200s
