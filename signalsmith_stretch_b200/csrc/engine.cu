// engine.cu -- host side of the B200 stretch engine + the extern "C" ABI of include/b200_stretch.h.
//
// The host does no DSP: it sizes buffers, builds the constant tables once per configure() and
// enqueues kernels on one CUDA stream.  Even the block scheduler runs on the device (k_plan), so
// a process() call is a fixed, sync-free launch sequence:
//   plain (no frequency map / formants), presets:  k_plan -> k_analyse2 -> k_chain_direct4 (stereo) | k_chain_direct2 (mono)
//                                                  -> k_synth2 || k_commit (side stream)
//   mapped / formants:  k_plan -> k_analyse2 -> [k_pitch] -> k_energy -> k_passes -> k_prep (map-only) -> k_products
//                       -> k_chain_t -> k_synth2 || k_commit                                        (chain_t.cuh)
//   calls that may stretch beyond 2x additionally launch k_prep + k_chain in random-only mode: they take the streams
//   whose blocks draw random time factors (Call::hasRandom, decided per stream by k_plan); every other CTA exits at once
//   generic sizes (not 3072 / 2560 bands): k_analyse / k_synth (fft.cuh), k_chain_direct2 (plain) / k_chain (mapped)
//   selectable for A/B: k_chain_direct6 (stereo presets, tuning key 0 = 6; mono stream pairs, key 5), chain_direct6.cuh
// Reference for every step: /root/reference/signalsmith-stretch.h (cited per kernel in kernels.cuh).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_stretch.h"
#include "kernels.cuh"
#include "chain_direct.cuh"
#include "chain_direct2.cuh"
#include "stft2.cuh"
#include "chain_direct3.cuh"
#include "chain_direct4.cuh"
#include "chain_ws.cuh"
#include "chain_direct6.cuh"
#include "chain_t.cuh"

using namespace b200s;

static std::string g_createError;

struct b200s_engine {
	int S = 0, device = 0;
	long seed = 0;
	bool configured = false;
	Cfg cfg;
	Params prm;
	cudaStream_t stream = 0;
	bool ownStream = false;
	// sub-batch pipeline: process() splits the batch over these prioritised streams (see process_impl)
	static const int kMaxSub = 16;
	int nSub = 1, maxSub = 1;
	cudaStream_t subStream[kMaxSub] = {};
	cudaEvent_t evBegin = 0, evSubDone[kMaxSub] = {};
	int chainedGroups = 0; // > 0: the last operation was a host-buffer process() over this many stream groups (see process_impl)
	int chainedIn = -1, chainedOut = -1; // its per-channel sample counts: the groups' slabs of the staging buffers sit at sBase*C*n
	int chainV = 0, fftV1 = 0; // b200s_set_tuning overrides (0 = default)
	int exactMath = 0;         // b200s_set_tuning key 3: 1 = the phase chain in the reference's unfused IEEE arithmetic
	int nHostParts = 12; // (measured, batch 1024 stereo: 2 -> 15.1, 4 -> 13.0, 8 -> 12.2, 12 -> 11.2, 16 -> 11.8 ms per step) host-buffer API: stream groups whose H2D copy / kernels / D2H copy are pipelined
	cudaEvent_t evStart = 0, evStop = 0;
	// k_commit (state carry + history append) only needs the chain's output, not the synthesis: it runs on a side stream
	// beside k_synth2 (fork after the chain, join after both)
	cudaStream_t commitStream = 0;
	cudaEvent_t evFork = 0, evJoin = 0;
	long long launches = 0, allocs = 0;
	std::string err;
	// optional per-kernel CUDA-event timing (b200s_profile_begin/end)
	bool profiling = false;
	std::vector<cudaEvent_t> profEv; // pairs
	std::vector<int> profKind;

	// tables
	float *dWindow = 0, *dWinProd = 0, *dWpReset = 0;
	float2 *dRot = 0, *dTwiddle = 0, *dPretw = 0;
	float4 *dAnaTab = 0;
	float2 rot0, rotStep;
	float *dMapIn = 0, *dMapOut = 0;
	// state
	Sched *dSched = 0;
	float *dHist[2] = {0, 0};
	int histCur = 0;
	float *dPend = 0, *dPendWp = 0;
	float2 *dStIn = 0, *dStPrev = 0, *dStOut = 0;
	float *dStPredE = 0;
	float4 *dStIl = 0;
	// call scratch
	int maxFrames = 0, coefFrames = 0; // coefFrames: block capacity of the coefficient rows (mapped / formant calls only)
	Frame *dFrames = 0;
	Call *dCall = 0;
	Job *dJobs = 0;
	int numSMs = 148;
	float2 *dSpec = 0, *dY = 0, *dPI = 0, *dFT = 0, *dT1 = 0, *dT2 = 0;
	float *dE = 0, *dS = 0, *dM = 0;
	// step-major path of the mapped / formant configurations (chain_t.cuh)
	float *dMapB = 0, *dMapG = 0, *dRatio = 0, *dTE = 0;
	float2 *dTPI = 0, *dTFT = 0, *dTT1 = 0, *dTT2 = 0;
	int tFrames = 0;
	int dual = -1;     // b200s_set_tuning key 5: mono plain path, pairs of streams on the packed wavefront (-1: B200S_DUAL or on)
	int stepMajor = 1; // b200s_set_tuning key 4: 0 = the round-1 kernels (k_prep + k_chain) for every stream
	// random time factors beyond 2x stretch (:639-640): engine state per stream (lives as long as the handle, like the
	// reference's randomEngine member), powers of the multiplier, the upwards twists of random blocks
	unsigned *dRng = 0, *dRngPow = 0;
	unsigned long long *dDiag = 0;
	float *dSeekStf = 0; // per-stream seek time factors (b200s_seek_rates)
	std::vector<float> hSeekStf;
	long long *dSeekEnd = 0; // per-stream window ends in the audio bank (b200s_live_seek)
	unsigned rngJump = 1;
	float2 *dT1u = 0, *dT2u = 0;
	int randFrames = 0;
	bool prevCallMayRandom = false, seekMayRandom = false;
	float *dStPitch = 0, *dPitch = 0;
	// staging for the host-buffer API and for flush/outputSeek
	float *dIn = 0, *dOut = 0, *dZero = 0, *dTmp = 0;
	size_t inCap = 0, outCap = 0, zeroCap = 0, tmpCap = 0;
	short *dIn16 = 0, *dOut16 = 0; // 16-bit PCM staging of b200s_process_pcm16
	size_t in16Cap = 0, out16Cap = 0;
};

#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t _e = (call);                                                                   \
		if (_e != cudaSuccess) {                                                                   \
			e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
			return B200S_ECUDA;                                                                    \
		}                                                                                          \
	} while (0)
#define CKL()                                                                                      \
	do {                                                                                           \
		++e->launches;                                                                             \
		CK(cudaGetLastError());                                                                    \
	} while (0)
#define NEED_CFG()                                                                                 \
	do {                                                                                           \
		if (!e) return B200S_EINVAL;                                                               \
		if (!e->configured) {                                                                      \
			e->err = "engine not configured";                                                      \
			return B200S_EINVAL;                                                                   \
		}                                                                                          \
	} while (0)

template <typename T>
static int dalloc(b200s_engine *e, T **p, size_t n) {
	if (*p) cudaFree(*p);
	*p = 0;
	++e->allocs;
	CK(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
	return 0;
}
template <typename T>
static void dfree(T *&p) {
	if (p) cudaFree(p);
	p = 0;
}

static void free_all(b200s_engine *e) {
	dfree(e->dWindow); dfree(e->dWinProd); dfree(e->dWpReset); dfree(e->dRot); dfree(e->dTwiddle); dfree(e->dPretw); dfree(e->dAnaTab);
	dfree(e->dSched); dfree(e->dHist[0]); dfree(e->dHist[1]); dfree(e->dPend); dfree(e->dPendWp);
	dfree(e->dStIn); dfree(e->dStPrev); dfree(e->dStOut); dfree(e->dStPredE); dfree(e->dStIl);
	dfree(e->dFrames); dfree(e->dCall); dfree(e->dJobs); dfree(e->dSpec); dfree(e->dY); dfree(e->dPI); dfree(e->dFT); dfree(e->dT1); dfree(e->dT2); dfree(e->dE); dfree(e->dS); dfree(e->dM); dfree(e->dT1u); dfree(e->dT2u); dfree(e->dRngPow); dfree(e->dStPitch);
	dfree(e->dMapB); dfree(e->dMapG); dfree(e->dRatio); dfree(e->dTE); dfree(e->dTPI); dfree(e->dTFT); dfree(e->dTT1); dfree(e->dTT2);
	e->randFrames = e->tFrames = 0; dfree(e->dPitch);
	dfree(e->dIn); dfree(e->dOut); dfree(e->dZero); dfree(e->dTmp); dfree(e->dIn16); dfree(e->dOut16);
	e->in16Cap = e->out16Cap = 0;
	e->maxFrames = e->coefFrames = 0;
	e->inCap = e->outCap = e->zeroCap = e->tmpCap = 0;
}

static Ctx make_ctx(b200s_engine *e) {
	Ctx x;
	memset(&x, 0, sizeof(x));
	x.cfg = e->cfg;
	x.prm = e->prm;
	x.prm.mapIn = e->dMapIn;
	x.prm.mapOut = e->dMapOut;
	x.window = e->dWindow; x.winProd = e->dWinProd; x.wpReset = e->dWpReset;
	x.rot = e->dRot; x.twiddle = e->dTwiddle; x.pretw = e->dPretw; x.anaTab = e->dAnaTab;
	x.rot0 = e->rot0; x.rotStep = e->rotStep;
	x.one = 1.0f;
	x.sched = e->dSched;
	x.histCur = e->dHist[e->histCur]; x.histNext = e->dHist[e->histCur ^ 1];
	x.pend = e->dPend; x.pendWp = e->dPendWp;
	x.stIn = e->dStIn; x.stPrev = e->dStPrev; x.stOut = e->dStOut; x.stPredE = e->dStPredE; x.stIl = e->dStIl;
	x.maxFrames = e->maxFrames;
	x.sBase = 0; x.sCount = e->S;
	x.frames = e->dFrames; x.call = e->dCall; x.jobs = e->dJobs;
	x.spec = e->dSpec; x.Y = e->dY; x.cPI = e->dPI; x.cFT = e->dFT; x.cT1 = e->dT1; x.cT2 = e->dT2; x.cE = e->dE; x.cS = e->dS; x.cM = e->dM; x.cT1u = e->dT1u; x.cT2u = e->dT2u;
	x.cMapB = e->dMapB; x.cMapG = e->dMapG; x.cRatio = e->dRatio; x.tE = e->dTE; x.tPI = e->dTPI; x.tFT = e->dTFT; x.tT1 = e->dTT1; x.tT2 = e->dTT2;
	x.tRows = t_rows(e->cfg); x.tGroups = (e->maxFrames + 31) / 32;
	x.rngState = e->dRng; x.rngPow = e->dRngPow; x.rngJump = e->rngJump; x.diag = e->dDiag; x.stPitch = e->dStPitch; x.cPitch = e->dPitch;
	return x;
}

// complex FFT size the dependency picks for a block: 2^k * {1,3,5} (measured)
static int fast_size_above(int n) {
	int p = 1;
	while (p < 16 && p < n) p *= 2;
	while (8 * p < n) p *= 2;
	int m = (n + p - 1) / p;
	if (m == 7) m = 8;
	return m * p;
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

static const int kThreads = 256;
static size_t smem_analyse(const Cfg &g) { return sizeof(float2) * 2 * fft_buf_len(g.K) + sizeof(float) * (g.B + 4); }
static size_t smem_prep(const Cfg &g, bool formants = true) { return sizeof(float) * ((formants ? 7 : 6) * (size_t)g.K + 8); }
static size_t smem_synth(const Cfg &g) { return sizeof(float2) * 2 * fft_buf_len(g.K) + sizeof(float) * 2 * g.pendLen; }

enum { PK_PLAN = 0, PK_ANALYSE, PK_PREP, PK_CHAIN, PK_SYNTH, PK_COMMIT, PK_COUNT };
static int prof_mark(b200s_engine *e, int kind, bool begin) {
	if (!e->profiling) return 0;
	cudaEvent_t ev;
	CK(cudaEventCreate(&ev));
	CK(cudaEventRecord(ev, e->stream));
	e->profEv.push_back(ev);
	if (begin) e->profKind.push_back(kind);
	return 0;
}
#define PROF(kind, stmt)                          \
	do {                                          \
		int _rc;                                  \
		if ((_rc = prof_mark(e, kind, true))) return _rc;  \
		stmt;                                     \
		CKL();                                    \
		if ((_rc = prof_mark(e, kind, false))) return _rc; \
	} while (0)

typedef void (*ChainKernel)(Ctx);
// Superseded kernel generations (first-generation FFT kernels specialised for the preset sizes, chain generations 1, 3 and 5)
// are only instantiated with -DB200S_KEEP_OLD_KERNELS (the emulator test builds: cross-checks between generations); the
// shipped library carries the generic FFT kernels (any size), k_chain (any configuration), k_chain_direct2 / 4 / 6, k_chain_t.
#ifdef B200S_KEEP_OLD_KERNELS
static ChainKernel analyse_kernel(const Cfg &g) { return g.K == 3072 ? k_analyse<3072> : g.K == 2560 ? k_analyse<2560> : k_analyse<0>; }
#else
static ChainKernel analyse_kernel(const Cfg &) { return k_analyse<0>; }
#endif
// paired in-place FFT kernels (stft2.cuh) for the preset sizes
static bool use_pair_fft(const Cfg &g, int forceV1 = 0) {
	static int v1 = -1;
	if (v1 < 0) v1 = getenv("B200S_FFT_V1") ? 1 : 0; // A/B switch for profiling the first-generation kernels
	return !v1 && !forceV1 && (g.K == 3072 || g.K == 2560);
}
static ChainKernel analyse2_kernel(const Cfg &g) { return g.K == 3072 ? k_analyse2<3072> : k_analyse2<2560>; }
static ChainKernel synth2_kernel(const Cfg &g) { return g.K == 3072 ? k_synth2<3072> : k_synth2<2560>; }
#ifdef B200S_KEEP_OLD_KERNELS
static ChainKernel synth_kernel(const Cfg &g) { return g.K == 3072 ? k_synth<3072> : g.K == 2560 ? k_synth<2560> : k_synth<0>; }
#else
static ChainKernel synth_kernel(const Cfg &) { return k_synth<0>; }
#endif
template <int CT>
static ChainKernel chain_kernel_for(int L, bool direct) {
#ifdef B200S_KEEP_OLD_KERNELS
	if (direct) switch (L) {
		case 1: return k_chain_direct<CT, 1>;
		case 2: return k_chain_direct<CT, 2>;
		case 3: return k_chain_direct<CT, 3>;
		case 4: return k_chain_direct<CT, 4>;
		case 5: return k_chain_direct<CT, 5>;
		case 6: return k_chain_direct<CT, 6>;
		case 7: return k_chain_direct<CT, 7>;
		default: return k_chain_direct<CT, 8>;
		}
#endif
	(void)direct; // (without the old generations chain_version() never asks for the first-generation direct kernel)
	switch (L) {
	case 1: return k_chain<CT, 1, false>;
	case 2: return k_chain<CT, 2, false>;
	case 3: return k_chain<CT, 3, false>;
	case 4: return k_chain<CT, 4, false>;
	case 5: return k_chain<CT, 5, false>;
	case 6: return k_chain<CT, 6, false>;
	case 7: return k_chain<CT, 7, false>;
	default: return k_chain<CT, 8, false>;
	}
}
// second-generation direct chain: lane = (block, channel), several warps per stream (chain_direct2.cuh)
template <int CT, bool FAST>
static ChainKernel chain2_kernel_for(int L) {
	switch (L) {
	case 1: return k_chain_direct2<CT, 1, FAST>;
	case 2: return k_chain_direct2<CT, 2, FAST>;
	case 3: return k_chain_direct2<CT, 3, FAST>;
	case 4: return k_chain_direct2<CT, 4, FAST>;
	case 5: return k_chain_direct2<CT, 5, FAST>;
	case 6: return k_chain_direct2<CT, 6, FAST>;
	case 7: return k_chain_direct2<CT, 7, FAST>;
	default: return k_chain_direct2<CT, 8, FAST>;
	}
}
// fast = the fused arithmetic (mono only: the stereo default is the packed kernel, whose exact cross-check this one is)
static ChainKernel chain2_kernel(const Cfg &g, bool fast = false) {
	return g.C == 1 ? (fast ? chain2_kernel_for<1, true>(g.L) : chain2_kernel_for<1, false>(g.L)) : chain2_kernel_for<2, false>(g.L);
}
// warps per stream for a call of nOut output samples: one per 32/C blocks (blocks trigger every H samples)
static int chain2_warps(const Cfg &g, int nOut) {
	const int blocks = (nOut + g.H - 1) / g.H, bpw = 32 / g.C;
	return std::max(1, std::min(CH2_MAXW, (blocks + bpw - 1) / bpw));
}
// which direct chain kernel: 3 = packed stereo (chain_direct3.cuh, needs the paired analysis kernel),
// 2 = lane-per-(block,channel) multi-warp (chain_direct2.cuh), 1 = first generation; B200S_CHAIN_V overrides (A/B profiling)
static int chain_version(const Cfg &g, int override, int forceFftV1) {
	static int env = -1;
	if (env < 0) {
		const char *v = getenv("B200S_CHAIN_V");
		env = v ? atoi(v) : 0;
	}
	// default 4: measured on B200 (profiles/r02_chain_ws_ncu_summary.md) the warp-specialised kernel (5) is slower,
	// 1.95 vs 1.68 ms -- its tile hand-off doubles the shared-memory traffic per step and the consumer waits for the producer
	int want = override ? override : env ? env : 4;
	if (want >= 3 && !(g.C == 2 && use_pair_fft(g, forceFftV1))) want = 2;
	if (want == 5 && g.L > 4) want = 4; // k_chain_ws is laid out for L <= 4 (both presets)
#ifndef B200S_KEEP_OLD_KERNELS
	if (want == 1) want = 2;
	if (want == 3 || want == 5) want = 4;
#endif
	return want;
}
// 5 = warp-specialised producer / consumer chain (chain_ws.cuh), followed by k_chain_direct4 for the streams it leaves
#ifdef B200S_KEEP_OLD_KERNELS
template <bool FAST>
static ChainKernel chain_ws_kernel(int L) {
	switch (L) {
	case 1: return k_chain_ws<1, FAST>;
	case 2: return k_chain_ws<2, FAST>;
	case 3: return k_chain_ws<3, FAST>;
	default: return k_chain_ws<4, FAST>;
	}
}
#endif
template <bool FAST>
static ChainKernel chain4_kernel(int L) {
	switch (L) {
	case 1: return k_chain_direct4<1, FAST>;
	case 2: return k_chain_direct4<2, FAST>;
	case 3: return k_chain_direct4<3, FAST>;
	case 4: return k_chain_direct4<4, FAST>;
	case 5: return k_chain_direct4<5, FAST>;
	case 6: return k_chain_direct4<6, FAST>;
	case 7: return k_chain_direct4<7, FAST>;
	default: return k_chain_direct4<8, FAST>;
	}
}
template <bool FAST, bool DUAL>
static ChainKernel chain6_kernel(int L) {
	switch (L) {
	case 1: return k_chain_direct6<1, FAST, DUAL>;
	case 2: return k_chain_direct6<2, FAST, DUAL>;
	case 3: return k_chain_direct6<3, FAST, DUAL>;
	case 4: return k_chain_direct6<4, FAST, DUAL>;
	case 5: return k_chain_direct6<5, FAST, DUAL>;
	case 6: return k_chain_direct6<6, FAST, DUAL>;
	case 7: return k_chain_direct6<7, FAST, DUAL>;
	default: return k_chain_direct6<8, FAST, DUAL>;
	}
}
// mono plain path: pairs of streams on the packed wavefront (k_chain_direct6<.., DUAL>) instead of k_chain_direct2.  Off by
// default: measured on B200 (profiles/r02_chain6_ab.md) it is no faster (4.99 vs 4.85 ms for 4096 streams) although it issues
// 40 % fewer instructions per stream -- the wavefront kernels are bound by dependent latency at 7 warps per SM, not by issue.
// B200S_DUAL=1 / b200s_set_tuning(e, 5, 1) switch it on (cross-check in the tests).
static bool dual_enabled() {
	static int env = -1;
	if (env < 0) {
		const char *v = getenv("B200S_DUAL");
		env = v ? (atoi(v) != 0) : 0;
	}
	return env != 0;
}
static ChainKernel chain3_kernel(const Cfg &g, int v, bool fast) {
#ifdef B200S_CHAIN_PROBES // profiling builds only: ablations of k_chain_direct6 (wrong results), B200S_CHAIN_PROBE=1..7
	if (v >= 6 && fast && g.L == 4) {
		static int probe = -1;
		if (probe < 0) {
			const char *pv = getenv("B200S_CHAIN_PROBE");
			probe = pv ? atoi(pv) : 0;
		}
		switch (probe) {
		case 1: return k_chain_direct6<4, true, false, 1>;
		case 2: return k_chain_direct6<4, true, false, 2>;
		case 3: return k_chain_direct6<4, true, false, 3>;
		case 4: return k_chain_direct6<4, true, false, 4>;
		case 5: return k_chain_direct6<4, true, false, 5>;
		case 6: return k_chain_direct6<4, true, false, 6>;
		case 7: return k_chain_direct6<4, true, false, 7>;
		default: break;
		}
	}
#endif
	if (v >= 6) return fast ? chain6_kernel<true, false>(g.L) : chain6_kernel<false, false>(g.L);
	if (v >= 4) return fast ? chain4_kernel<true>(g.L) : chain4_kernel<false>(g.L);
#ifdef B200S_KEEP_OLD_KERNELS
	switch (g.L) {
	case 1: return k_chain_direct3<1>;
	case 2: return k_chain_direct3<2>;
	case 3: return k_chain_direct3<3>;
	case 4: return k_chain_direct3<4>;
	case 5: return k_chain_direct3<5>;
	case 6: return k_chain_direct3<6>;
	case 7: return k_chain_direct3<7>;
	default: return k_chain_direct3<8>;
	}
#else
	return fast ? chain4_kernel<true>(g.L) : chain4_kernel<false>(g.L);
#endif
}
template <int CT>
static ChainKernel chain_t_kernel_for(int L) {
	switch (L) {
	case 1: return k_chain_t<CT, 1>;
	case 2: return k_chain_t<CT, 2>;
	case 3: return k_chain_t<CT, 3>;
	case 4: return k_chain_t<CT, 4>;
	case 5: return k_chain_t<CT, 5>;
	case 6: return k_chain_t<CT, 6>;
	case 7: return k_chain_t<CT, 7>;
	default: return k_chain_t<CT, 8>;
	}
}
static ChainKernel chain_t_kernel(const Cfg &g) { return g.C == 1 ? chain_t_kernel_for<1>(g.L) : chain_t_kernel_for<2>(g.L); }
static ChainKernel chain_kernel(const Cfg &g, bool direct) {
	return g.C == 1 ? chain_kernel_for<1>(g.L, direct) : chain_kernel_for<2>(g.L, direct);
}
static const int kChainWarps = 1; // one stream per CTA: 1024 streams spread evenly over the 148 SMs
static size_t smem_chain(const Cfg &g, bool direct, bool randTiles = false) {
	size_t per = direct ? (g.C == 1 ? sizeof(DirectTiles2<1>) : sizeof(DirectTiles2<2>)) : (g.C == 1 ? sizeof(ChainTiles<1>) : sizeof(ChainTiles<2>));
	if (!direct && randTiles) per += g.C == 1 ? sizeof(ChainRandTiles<1>) : sizeof(ChainRandTiles<2>);
	return per * kChainWarps;
}

static int reset_impl(b200s_engine *e, bool full) {
	e->chainedGroups = 0;
	Ctx x = make_ctx(e);
	B200S_LAUNCH(k_reset_stft, dim3(e->S), dim3(kThreads), 0, e->stream, x);
	CKL();
	B200S_LAUNCH(k_reset_bands, dim3(e->S), dim3(kThreads), 0, e->stream, x, full ? 15 : 6);
	CKL();
	return 0;
}

static int configure_impl(b200s_engine *e, int channels, int block, int interval, int split) {
	if (!e) return B200S_EINVAL;
	if (channels < 1 || block < 4 || interval < 1 || interval > block) {
		e->err = "configure: need channels >= 1, block >= 4, 1 <= interval <= block";
		return B200S_EINVAL;
	}
	Cfg g;
	memset(&g, 0, sizeof(g));
	g.S = e->S;
	g.C = channels;
	g.B = block;
	g.H = interval;
	g.N = 2 * fast_size_above((block + 1) / 2); // DynamicSTFT::configure (App. B)
	g.K = g.N / 2;
	g.L = (int)std::round(float(g.N) / float(g.H)); // :636-637
	g.split = split ? 1 : 0;
	g.histLen = g.B + g.H;
	g.pendLen = g.B + (g.split ? g.H : 0);
	g.addOff = g.split ? g.H : 0;
	g.o = g.B / 2;
	{ // Stockham plan: power-of-two radices first (16s, then 8/4/2), the radix carrying the odd factor last
		int k = g.K, odd = 1, n = 0;
		if (k % 3 == 0) { odd = 3; k /= 3; }
		else if (k % 5 == 0) { odd = 5; k /= 5; }
		int last = odd;
		if (odd == 3) { int c = 0; while (c < 2 && k % 2 == 0 && k > 1) { last *= 2; k /= 2; ++c; } } // 12, 6 or 3
		if (odd == 5) { if (k % 2 == 0 && k > 1) { last *= 2; k /= 2; } }                            // 10 or 5
		if (k & (k - 1)) { e->err = "unsupported FFT size"; return B200S_EUNSUPPORTED; }
		while (k % 16 == 0 && k > 16) { g.radix[n++] = 16; k /= 16; }
		if (k == 16 && last > 1) { g.radix[n++] = 16; k = 1; }
		else if (k == 16) { g.radix[n++] = 4; last = 4; k = 1; }
		if (k == 8) { g.radix[n++] = 8; k = 1; }
		if (k == 4) { g.radix[n++] = 4; k = 1; }
		if (k == 2) { g.radix[n++] = 2; k = 1; }
		if (last > 1) g.radix[n++] = last;
		g.nStages = n;
		int prod = 1;
		for (int i = 0; i < n; ++i) prod *= g.radix[i];
		if (prod != g.K) { e->err = "internal: FFT plan does not factor the size"; return B200S_EUNSUPPORTED; }
	}
	if (g.L < 1 || g.L > 8 || (channels > 2)) {
		e->err = "GPU path supports 1-2 channels and block/interval ratios with round(fftSamples/interval) in 1..8";
		return B200S_EUNSUPPORTED;
	}
	if (smem_prep(g) > 220 * 1024 || smem_synth(g) > 220 * 1024) {
		e->err = "block too large for the shared-memory FFT (bands must be <= ~6144)";
		return B200S_EUNSUPPORTED;
	}
	CK(cudaSetDevice(e->device));
	CK(cudaStreamSynchronize(e->stream));
	free_all(e);
	e->cfg = g;

	// ---- tables.  Window: Kaiser, the dependency's bandwidth heuristic, forced perfect reconstruction (App. B)
	std::vector<float> window(g.B), winProd(g.B), wpReset(g.pendLen);
	{
		double bw = double(g.B) / double(g.H);
		bw += 8 / ((bw + 3) * (bw + 3)) + 0.25 * std::max(3 - bw, 0.0);
		bw = std::max(bw, 2.0);
		double beta = M_PI * std::sqrt(bw * bw * 0.25 - 1);
		std::vector<double> w(g.B);
		for (int i = 0; i < g.B; ++i) {
			double r = (2.0 * i + 1) / g.B - 1;
			w[i] = bessel0(beta * std::sqrt(std::max(0.0, 1 - r * r))) / bessel0(beta);
		}
		for (int i = 0; i < g.H && i < g.B; ++i) {
			double s2 = 0;
			for (int k = i; k < g.B; k += g.H) s2 += w[k] * w[k];
			double f = 1 / std::sqrt(s2);
			for (int k = i; k < g.B; k += g.H) w[k] *= f;
		}
		for (int i = 0; i < g.B; ++i) {
			window[i] = float(w[i]);
			winProd[i] = window[i] * window[i] * float(g.N);
		}
		for (int i = 0; i < g.pendLen; ++i) { // DynamicSTFT::reset(0.1) then moveOutput(H)
			int ring = i + g.H;
			if (ring < g.B) {
				float sum = 0;
				for (int k = ring; k < g.B; k += g.H) sum += window[k] * window[k];
				wpReset[i] = 0.1f * float(g.N) * sum + B200S_ALMOST_ZERO;
			} else {
				wpReset[i] = B200S_ALMOST_ZERO;
			}
		}
	}
	std::vector<float2> rot(g.K), tw(g.K), pre(g.K);
	{ // :647-655, the reference's own float recurrence (its rounding is part of the result)
		auto b2f = [&](float b) { return (b + 0.5f) / float(g.N); };
		std::complex<float> r = std::polar(1.0f, b2f(0) * float(g.H) * float(2 * M_PI));
		float freqStep = b2f(1) - b2f(0);
		std::complex<float> step = std::polar(1.0f, freqStep * float(g.H) * float(2 * M_PI));
		e->rot0 = make_float2(r.real(), r.imag());
		e->rotStep = make_float2(step.real(), step.imag());
		for (int b = 0; b < g.K; ++b) {
			rot[b] = make_float2(r.real(), r.imag());
			r = std::complex<float>(r.real() * step.real() - r.imag() * step.imag(), r.real() * step.imag() + r.imag() * step.real());
		}
		for (int t = 0; t < g.K; ++t) {
			tw[t] = make_float2(float(std::cos(-2.0 * M_PI * t / g.K)), float(std::sin(-2.0 * M_PI * t / g.K)));
			pre[t] = make_float2(float(std::cos(-M_PI * t / g.N)), float(std::sin(-M_PI * t / g.N)));
		}
	}
	int rc;
	if ((rc = dalloc(e, &e->dWindow, g.B))) return rc;
	if ((rc = dalloc(e, &e->dWinProd, g.B))) return rc;
	if ((rc = dalloc(e, &e->dWpReset, g.pendLen))) return rc;
	if ((rc = dalloc(e, &e->dRot, g.K))) return rc;
	if ((rc = dalloc(e, &e->dTwiddle, g.K))) return rc;
	if ((rc = dalloc(e, &e->dPretw, g.K))) return rc;
	CK(cudaMemcpy(e->dWindow, window.data(), sizeof(float) * g.B, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dWinProd, winProd.data(), sizeof(float) * g.B, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dWpReset, wpReset.data(), sizeof(float) * g.pendLen, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dRot, rot.data(), sizeof(float2) * g.K, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dTwiddle, tw.data(), sizeof(float2) * g.K, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dPretw, pre.data(), sizeof(float2) * g.K, cudaMemcpyHostToDevice));
	{ // per-element constants of the analysis load stage, one float4 each
		std::vector<float4> tab(g.K);
		for (int n = 0; n < g.K; ++n) {
			const int i0 = n + g.o, i1 = n + g.o - g.K;
			tab[n] = make_float4(i0 < g.B ? window[i0] : 0.f, i1 >= 0 ? window[i1] : 0.f, pre[n].x, pre[n].y);
		}
		if ((rc = dalloc(e, &e->dAnaTab, g.K))) return rc;
		CK(cudaMemcpy(e->dAnaTab, tab.data(), sizeof(float4) * g.K, cudaMemcpyHostToDevice));
	}

	{ // powers of the random engine's multiplier: draw i of a block is state * 16807^i mod 2^31-1 (kernels.cuh)
		std::vector<unsigned> pw(2 * (size_t)g.K);
		unsigned v = 1;
		for (size_t i = 0; i < pw.size(); ++i) {
			pw[i] = v;
			v = rng_mulmod(v, B200S_RNG_A);
		}
		e->rngJump = pw[2 * (size_t)g.K - 2];
		if ((rc = dalloc(e, &e->dRngPow, pw.size()))) return rc;
		CK(cudaMemcpy(e->dRngPow, pw.data(), sizeof(unsigned) * pw.size(), cudaMemcpyHostToDevice));
		if (!e->dRng) { // std::default_random_engine(seed): state = seed mod (2^31-1), 1 if that is 0; NOT touched by configure / reset
			if ((rc = dalloc(e, &e->dRng, (size_t)g.S))) return rc;
			unsigned x0 = (unsigned)((unsigned long long)e->seed % (unsigned long long)B200S_RNG_M);
			if (x0 == 0) x0 = 1;
			std::vector<unsigned> st((size_t)g.S, x0);
			CK(cudaMemcpy(e->dRng, st.data(), sizeof(unsigned) * st.size(), cudaMemcpyHostToDevice));
			if ((rc = dalloc(e, &e->dDiag, (size_t)1))) return rc;
			CK(cudaMemset(e->dDiag, 0, sizeof(unsigned long long)));
		}
	}

	// ---- state
	const size_t SC = (size_t)g.S * g.C;
	if ((rc = dalloc(e, &e->dSched, g.S))) return rc;
	if ((rc = dalloc(e, &e->dHist[0], SC * g.histLen))) return rc;
	if ((rc = dalloc(e, &e->dHist[1], SC * g.histLen))) return rc;
	if ((rc = dalloc(e, &e->dPend, SC * g.pendLen))) return rc;
	if ((rc = dalloc(e, &e->dPendWp, SC * g.pendLen))) return rc;
	if ((rc = dalloc(e, &e->dStIn, SC * g.K))) return rc;
	if ((rc = dalloc(e, &e->dStPrev, SC * g.K))) return rc;
	if ((rc = dalloc(e, &e->dStOut, SC * g.K))) return rc;
	if ((rc = dalloc(e, &e->dStPredE, SC * g.K))) return rc;
	if ((rc = dalloc(e, &e->dStIl, g.C == 2 ? (size_t)g.S * 2 * g.K : 1))) return rc;
	if ((rc = dalloc(e, &e->dCall, g.S))) return rc;
	CK(cudaMemset(e->dStPredE, 0, sizeof(float) * SC * g.K));
	if ((rc = dalloc(e, &e->dStPitch, (size_t)g.S * 2))) return rc;
	CK(cudaMemset(e->dStPitch, 0, sizeof(float) * (size_t)g.S * 2));
	CK(cudaMemset(e->dHist[1], 0, sizeof(float) * SC * g.histLen));
	{
		std::vector<Sched> sc(g.S);
		for (auto &s : sc) {
			s.samplesSinceLast = B200S_NEVER;
			s.silenceCounter = 0;
			s.prevInputOffset = -1;
			s.didSeek = 0;
			s.silenceFirst = 1;
			s.seekTimeFactor = 1;
			s.zeroRun = B200S_NEVER;
		}
		CK(cudaMemcpy(e->dSched, sc.data(), sizeof(Sched) * g.S, cudaMemcpyHostToDevice));
	}
	e->histCur = 0;
#ifndef B200S_EMU
	CK(cudaFuncSetAttribute(analyse_kernel(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_analyse(g)));
	CK(cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep(g)));
	CK(cudaFuncSetAttribute(synth_kernel(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_synth(g)));
	CK(cudaFuncSetAttribute(k_flush_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * g.B)));
	CK(cudaFuncSetAttribute(chain_kernel(g, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain(g, false, true)));
#ifdef B200S_KEEP_OLD_KERNELS // (without the old generations chain_kernel(g, true) is the same function: do not shrink its limit again)
	CK(cudaFuncSetAttribute(chain_kernel(g, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain(g, true)));
#endif
	CK(cudaFuncSetAttribute(chain2_kernel(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain2(g.C, CH2_MAXW)));
	CK(cudaFuncSetAttribute(chain2_kernel(g, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain2(g.C, CH2_MAXW)));
#ifdef B200S_KEEP_OLD_KERNELS
	if (g.L <= 4) { // seven two-warp CTAs of 30.6 KB per SM: ask for the full shared-memory carve-out
		CK(cudaFuncSetAttribute(chain_ws_kernel<true>(g.L), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
		CK(cudaFuncSetAttribute(chain_ws_kernel<false>(g.L), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
	}
#endif
	// k_chain_direct6: seven one-warp CTAs of 31.5 KB per SM need the full shared-memory carve-out
	CK(cudaFuncSetAttribute(g.C == 1 ? chain6_kernel<true, true>(g.L) : chain6_kernel<true, false>(g.L), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
	CK(cudaFuncSetAttribute(g.C == 1 ? chain6_kernel<false, true>(g.L) : chain6_kernel<false, false>(g.L), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
	if (use_pair_fft(g)) {
		CK(cudaFuncSetAttribute(analyse2_kernel(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_analyse2(g)));
		CK(cudaFuncSetAttribute(synth2_kernel(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_synth2(g)));
	}
	{
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && n > 0) e->numSMs = n;
	}
#endif
	e->configured = true;
	return reset_impl(e, true);
}

static int frames_bound(const Cfg &g, int nOut) { return nOut <= 0 ? 0 : (nOut + g.H - 1) / g.H + 1; }

static int ensure_scratch(b200s_engine *e, int nOut) {
	const Cfg &g = e->cfg;
	int need = std::max(frames_bound(g, nOut), 1);
	if (need <= e->maxFrames) return 0;
	CK(cudaStreamSynchronize(e->stream));
	const size_t n = (size_t)g.S * need * g.C * g.K;
	int rc;
	if ((rc = dalloc(e, &e->dFrames, (size_t)g.S * need))) return rc;
	if ((rc = dalloc(e, &e->dJobs, (size_t)g.S * need * 2 * g.C))) return rc;
	if ((rc = dalloc(e, &e->dSpec, 2 * n))) return rc;
	if ((rc = dalloc(e, &e->dY, n))) return rc;
	if ((rc = dalloc(e, &e->dE, n))) return rc; // Prediction::energy rows: also the state carry of the generic direct chains
	if ((rc = dalloc(e, &e->dPitch, (size_t)g.S * need))) return rc;
	e->maxFrames = need;
	// the complex coefficient rows follow maxFrames (their row offsets use it) but only exist once a call needed them
	dfree(e->dPI); dfree(e->dFT); dfree(e->dT1); dfree(e->dT2); dfree(e->dS); dfree(e->dM); dfree(e->dT1u); dfree(e->dT2u);
	dfree(e->dMapB); dfree(e->dMapG); dfree(e->dRatio); dfree(e->dTE); dfree(e->dTPI); dfree(e->dTFT); dfree(e->dTT1); dfree(e->dTT2);
	e->coefFrames = e->randFrames = e->tFrames = 0;
	return 0;
}
// Complex coefficient rows of k_prep (32 B per bin-channel and block: 6.6 GB at batch 1024 x 33 blocks stereo): only calls
// with a frequency map or formant processing launch k_prep, so only they pay for them (b200s_reserve() allocates them
// when the parameters set at that time need them).
static bool params_plain(const b200s_engine *e) {
	return !(e->prm.mapN > 0 || e->prm.freqMultiplier != 1.0f) && e->prm.formantMultiplier == 1.0f;
}
static int ensure_coef(b200s_engine *e) {
	if (e->coefFrames == e->maxFrames && e->dPI) return 0;
	const Cfg &g = e->cfg;
	CK(cudaStreamSynchronize(e->stream));
	const size_t n = (size_t)g.S * e->maxFrames * g.C * g.K;
	int rc;
	if ((rc = dalloc(e, &e->dPI, n))) return rc;
	if ((rc = dalloc(e, &e->dFT, n))) return rc;
	if ((rc = dalloc(e, &e->dT1, n))) return rc;
	if ((rc = dalloc(e, &e->dT2, n))) return rc;
	if ((rc = dalloc(e, &e->dS, n / g.C))) return rc; // smoothed energy / formant envelope rows: one per block (k_passes)
	if ((rc = dalloc(e, &e->dM, n / g.C))) return rc;
	e->coefFrames = e->maxFrames;
	return 0;
}
// rows of the step-major path (mapped / formant calls): map, ratio, smoothed energy, envelope, and the transposed terms
static int ensure_stepmajor(b200s_engine *e) {
	if (e->tFrames == e->maxFrames && e->dTFT) return 0;
	const Cfg &g = e->cfg;
	CK(cudaStreamSynchronize(e->stream));
	const size_t nRow = (size_t)g.S * e->maxFrames * g.K;
	const size_t nT = (size_t)g.S * ((e->maxFrames + 31) / 32) * t_rows(g) * g.C * 32;
	int rc;
	if ((rc = dalloc(e, &e->dMapB, nRow))) return rc;
	if ((rc = dalloc(e, &e->dMapG, nRow))) return rc;
	if ((rc = dalloc(e, &e->dRatio, nRow))) return rc;
	if (!e->dS && (rc = dalloc(e, &e->dS, nRow))) return rc;
	if (!e->dM && (rc = dalloc(e, &e->dM, nRow))) return rc;
	if ((rc = dalloc(e, &e->dTE, nT))) return rc;
	if ((rc = dalloc(e, &e->dTPI, nT))) return rc;
	if ((rc = dalloc(e, &e->dTFT, nT))) return rc;
	if ((rc = dalloc(e, &e->dTT1, nT))) return rc;
	if ((rc = dalloc(e, &e->dTT2, nT))) return rc;
	e->tFrames = e->maxFrames;
	return 0;
}
// the two extra twist rows of blocks with random time factors (calls that may stretch beyond 2x only)
static int ensure_rand(b200s_engine *e) {
	if (e->randFrames == e->maxFrames && e->dT1u) return 0;
	const Cfg &g = e->cfg;
	CK(cudaStreamSynchronize(e->stream));
	const size_t n = (size_t)g.S * e->maxFrames * g.C * g.K;
	int rc;
	if ((rc = dalloc(e, &e->dT1u, n))) return rc;
	if ((rc = dalloc(e, &e->dT2u, n))) return rc;
	e->randFrames = e->maxFrames;
	return 0;
}
static int ensure_buf(b200s_engine *e, float **p, size_t *cap, size_t n, bool zero) {
	if (n <= *cap && *p) return 0;
	CK(cudaStreamSynchronize(e->stream));
	int rc;
	if ((rc = dalloc(e, p, n))) return rc;
	*cap = n;
	if (zero) CK(cudaMemset(*p, 0, sizeof(float) * std::max<size_t>(n, 1)));
	return 0;
}

// process() on device buffers with explicit strides (floats between channels / streams)
// hIn / hOut (host-buffer API only): packed [S][C][n] host arrays mirrored by dIn / dOut; each stream group
// then copies its own slab in, runs its kernels and copies its slab out on its own CUDA stream, so that the
// PCIe transfers of one group overlap the kernels of the others.
static int process_impl(b200s_engine *e, const float *dIn, int inChanStride, long long inStreamStride, int nIn,
                        float *dOut, int outChanStride, long long outStreamStride, int nOut,
                        const float *hIn = nullptr, float *hOut = nullptr, const short *hIn16 = nullptr, short *hOut16 = nullptr) {
	const Cfg &g = e->cfg;
	if (nIn < 0 || nOut < 0) {
		e->err = "process: negative sample count";
		return B200S_EINVAL;
	}
	int rc;
	if ((rc = ensure_scratch(e, nOut))) return rc;
	// Can a block of this call stretch beyond 2x (timeFactor = interval / inputInterval > 2, :312,:639)?  Inside a call the
	// input interval is interval * nIn / nOut give or take a sample; the first block's mixes in the previous call's ratio;
	// after a seek it is 1 / playbackRate.  Conservative on purpose: the decision itself is taken per stream on the device
	// (k_plan sets Call::hasRandom); this only says whether the kernels of the random path have to be launched at all.
	const bool lowRatio = nOut > 0 && (long long)nIn * 2 * g.H < (long long)nOut * (g.H + 4);
	const bool mayRandom = nOut > 0 && (lowRatio || e->prevCallMayRandom || e->seekMayRandom);
	if (nOut > 0) {
		e->prevCallMayRandom = lowRatio;
		e->seekMayRandom = false;
	}
	const bool useStepMajor = !params_plain(e) && e->stepMajor; // mapped / formant call on k_products + k_chain_t
	if (((!params_plain(e) && !useStepMajor) || mayRandom) && (rc = ensure_coef(e))) return rc;
	if (useStepMajor && (rc = ensure_stepmajor(e))) return rc;
	if (mayRandom && (rc = ensure_rand(e))) return rc;
	Ctx x = make_ctx(e);
	if (!mayRandom) x.cT1u = x.cT2u = nullptr; // (stale rows of an earlier call are never read)
	x.randomPathOn = (mayRandom || (!params_plain(e) && !useStepMajor)) ? 1 : 0; // the round-1 mapped path runs k_prep + k_chain for every stream anyway
	x.in = dIn; x.out = dOut; x.nIn = nIn; x.nOut = nOut;
	x.inChanStride = inChanStride; x.inStreamStride = inStreamStride;
	x.outChanStride = outChanStride; x.outStreamStride = outStreamStride;
	x.inAligned = ((uintptr_t)dIn % 16 == 0) && inChanStride % 4 == 0 && inStreamStride % 4 == 0 && nIn % 4 == 0 && g.histLen % 4 == 0;
	const int F = frames_bound(g, nOut);
	// Sub-batch pipeline: the streams of the batch are independent, so the launch sequence is issued
	// per sub-batch on its own CUDA stream (descending priority).  Sub-batch 0's chain then runs while
	// sub-batch 1 is still in its FFT kernel, and so on: the latency-bound wavefront kernel and the
	// throughput-bound FFT kernels share the SMs instead of taking turns.  Per-kernel profiling
	// (b200s_profile_begin) runs unsplit on the main stream so that the event pairs time one kernel each.
	const bool hostIO = hIn || hOut || hIn16 || hOut16;
	const int wantSub = hostIO ? e->nHostParts : e->nSub;
	// (host-buffer pipeline: at least two streams per group; device-resident sub-batches: only for large batches)
	const int nSub = (e->profiling || wantSub <= 1 || g.S < (hostIO ? 2 * wantSub : 64)) ? 1 : std::min(wantSub, e->maxSub);
	// Consecutive host-buffer calls over the SAME stream groups need no join between them: every group's copies and
	// kernels are ordered on the group's own stream and touch only that group's streams of the state, so group g of
	// call n+1 may start its H2D copy while other groups still finish call n (b200s_process_async pipelines calls this
	// way).  Anything else first joins: the main stream has waited for every group at the end of the previous call.
	// The groups' slabs of the shared staging buffers sit at sBase*C*nIn / sBase*C*nOut, so they only coincide from call to
	// call when the sample counts do: a call with different counts must first join (otherwise group g's copy of call n+1
	// could overwrite the slab a neighbouring group still uses for call n); evBegin below is that join, since the main
	// stream has waited for every group's evSubDone at the end of the previous call.
	const bool chained = hostIO && nSub > 1 && e->chainedGroups == nSub && e->chainedIn == nIn && e->chainedOut == nOut;
	e->chainedGroups = (hostIO && nSub > 1) ? nSub : 0;
	e->chainedIn = nIn;
	e->chainedOut = nOut;
	if (nSub > 1 && !chained) CK(cudaEventRecord(e->evBegin, e->stream));
	const bool plain = params_plain(e);
	const bool formantsOn = e->prm.formantMultiplier != 1.0f || (e->prm.formantCompensation && (e->prm.mapN > 0 || e->prm.freqMultiplier != 1.0f)); // :310
	const int chainV = chain_version(g, e->chainV, e->fftV1);
	const bool pairFft = use_pair_fft(g, e->fftV1);
	x.specIl = (plain && chainV >= 3) ? 1 : 0;
	for (int sub = 0; sub < nSub; ++sub) {
		cudaStream_t st = nSub > 1 ? e->subStream[sub] : e->stream;
		if (nSub > 1 && !chained) CK(cudaStreamWaitEvent(st, e->evBegin, 0));
		x.sBase = (int)((long long)g.S * sub / nSub);
		x.sCount = (int)((long long)g.S * (sub + 1) / nSub) - x.sBase;
		if (hIn && nIn > 0)
			CK(cudaMemcpyAsync((float *)dIn + (size_t)x.sBase * g.C * nIn, hIn + (size_t)x.sBase * g.C * nIn, sizeof(float) * (size_t)x.sCount * g.C * nIn, cudaMemcpyHostToDevice, st));
		if (hIn16 && nIn > 0) { // 16-bit PCM: 2 bytes per sample over PCIe, converted on the device
			const size_t off = (size_t)x.sBase * g.C * nIn, cnt = (size_t)x.sCount * g.C * nIn;
			CK(cudaMemcpyAsync(e->dIn16 + off, hIn16 + off, sizeof(short) * cnt, cudaMemcpyHostToDevice, st));
			B200S_LAUNCH(k_pcm16_in, dim3((unsigned)std::min<size_t>((cnt + 1023) / 1024, 4096)), dim3(256), 0, st, e->dIn16 + off, (float *)dIn + off, cnt);
			CKL();
		}
		{ // (profiling implies nSub == 1, i.e. st == e->stream, which is where PROF() records its events)
			PROF(PK_PLAN, B200S_LAUNCH(k_plan, dim3(x.sCount), dim3(kThreads), 0, st, x));
			if (F > 0) {
				if (pairFft) { // persistent CTAs, two per SM, over the (stream, job pair) items
					const long long items = (long long)x.sCount * g.C * x.maxFrames;
					const int grid = (int)std::min<long long>(items, 2LL * e->numSMs);
					PROF(PK_ANALYSE, B200S_LAUNCH(analyse2_kernel(g), dim3(grid), dim3(256), smem_analyse2(g), st, x));
				} else {
					ChainKernel ka = analyse_kernel(g);
					PROF(PK_ANALYSE, B200S_LAUNCH(ka, dim3(std::min(2 * F * g.C, ANALYSE_CTAS_PER_STREAM), 1, x.sCount), dim3(kThreads), smem_analyse(g), st, x));
				}
				// Pure time-stretch (no frequency map, no formants): the chain forms its terms directly from
				// the spectra and k_prep is skipped; otherwise k_prep produces the coefficient arrays.
				// formants with setFormantBase(0): the automatic pitch estimate of every block (serial over blocks) first
				if (!plain && formantsOn && !(e->prm.formantBaseFreq > 0)) {
					B200S_LAUNCH(k_pitch, dim3(x.sCount), dim3(kThreads), sizeof(float) * g.K, st, x);
					CKL();
				}
				if (!plain) {
					int _rc;
					if ((_rc = prof_mark(e, PK_PREP, true))) return _rc;
					// the serial one-pole passes over the bins, one lane per block (kernels.cuh), then the per-bin products
					B200S_LAUNCH(k_energy, dim3(F, x.sCount), dim3(kThreads), 0, st, x);
					CKL();
					B200S_LAUNCH(k_passes, dim3(x.sCount), dim3(32), 0, st, x);
					CKL();
					if (useStepMajor) { // peaks / map / ratio rows, then the per-bin terms in step-major order (chain_t.cuh)
						x.mapOnly = 1;
						B200S_LAUNCH(k_prep, dim3(F, x.sCount), dim3(kThreads), smem_prep(g, formantsOn), st, x);
						CKL();
						x.mapOnly = 0;
						B200S_LAUNCH(k_products, dim3(x.tRows / 32, x.tGroups, x.sCount), dim3(256), 0, st, x);
						CKL();
						if (mayRandom) { // streams with a random block: the full k_prep (random-only mode), consumed by k_chain below
							x.randomOnly = 1;
							B200S_LAUNCH(k_prep, dim3(F, x.sCount), dim3(kThreads), smem_prep(g, formantsOn), st, x);
							CKL();
							x.randomOnly = 0;
						}
					} else {
						B200S_LAUNCH(k_prep, dim3(F, x.sCount), dim3(kThreads), smem_prep(g, formantsOn), st, x);
						CKL();
					}
					if ((_rc = prof_mark(e, PK_PREP, false))) return _rc;
				}
				if (useStepMajor) {
					int _rc;
					if ((_rc = prof_mark(e, PK_CHAIN, true))) return _rc;
					ChainKernel kt = chain_t_kernel(g);
					B200S_LAUNCH(kt, dim3(x.sCount), dim3(32), 0, st, x);
					CKL();
					if (mayRandom) { // the streams k_chain_t left (Call::hasRandom)
						x.randomOnly = 1;
						dim3 grid((x.sCount + kChainWarps - 1) / kChainWarps), block(32 * kChainWarps);
						ChainKernel kc = chain_kernel(g, false);
						B200S_LAUNCH(kc, grid, block, smem_chain(g, false, true), st, x);
						CKL();
						x.randomOnly = 0;
					}
					if ((_rc = prof_mark(e, PK_CHAIN, false))) return _rc;
#ifdef B200S_KEEP_OLD_KERNELS
				} else if (plain && chainV == 5) {
					int _rc;
					if ((_rc = prof_mark(e, PK_CHAIN, true))) return _rc;
					const bool fast = !e->exactMath;
					ChainKernel kw = fast ? chain_ws_kernel<true>(g.L) : chain_ws_kernel<false>(g.L), k4 = chain3_kernel(g, 4, fast);
					B200S_LAUNCH(kw, dim3(x.sCount), dim3(64), fast ? smem_chain_ws<true>() : smem_chain_ws<false>(), st, x);
					CKL();
					x.wsRan = 1; // streams with a block beyond the 2x stretch limit: k_chain_direct4 (every other CTA exits at once)
					B200S_LAUNCH(k4, dim3(x.sCount), dim3(32), smem_chain4(g.L), st, x);
					CKL();
					x.wsRan = 0;
					if ((_rc = prof_mark(e, PK_CHAIN, false))) return _rc;
#endif
				} else if (plain && chainV >= 3) {
					PROF(PK_CHAIN, B200S_LAUNCH(chain3_kernel(g, chainV, !e->exactMath), dim3(x.sCount), dim3(32), chainV >= 6 ? smem_chain6() : chainV == 4 ? smem_chain4(g.L) : sizeof(Chain3Tiles), st, x));
				} else if (plain && chainV == 2) {
					const int W = chain2_warps(g, nOut);
					ChainKernel kc = chain2_kernel(g, g.C == 1 && !e->exactMath);
					int _rc;
					if ((_rc = prof_mark(e, PK_CHAIN, true))) return _rc;
					if (g.C == 1 && (e->dual < 0 ? dual_enabled() : e->dual != 0)) { // mono: two streams per warp on the packed wavefront
						ChainKernel kd = e->exactMath ? chain6_kernel<false, true>(g.L) : chain6_kernel<true, true>(g.L);
						B200S_LAUNCH(kd, dim3((x.sCount + 1) / 2), dim3(32), smem_chain6(), st, x);
					} else {
						B200S_LAUNCH(kc, dim3(x.sCount), dim3(32 * W), smem_chain2(g.C, W), st, x);
					}
					CKL();
					if ((_rc = prof_mark(e, PK_CHAIN, false))) return _rc;
				} else {
					dim3 grid((x.sCount + kChainWarps - 1) / kChainWarps), block(32 * kChainWarps);
					ChainKernel kc = chain_kernel(g, plain);
					PROF(PK_CHAIN, B200S_LAUNCH(kc, grid, block, smem_chain(g, plain, mayRandom), st, x));
				}
			}
			if (F > 0 && plain && mayRandom) {
				// streams with a block beyond 2x stretch (Call::hasRandom): the direct chain kernels above returned at once for
				// them; k_prep forms their coefficient rows with the per-bin random time factors and the generic chain runs
				// them.  Every other CTA of these two launches exits immediately.
				x.randomOnly = 1;
				B200S_LAUNCH(k_prep, dim3(F, x.sCount), dim3(kThreads), smem_prep(g, false), st, x);
				CKL();
				dim3 grid((x.sCount + kChainWarps - 1) / kChainWarps), block(32 * kChainWarps);
				ChainKernel kc = chain_kernel(g, false);
				B200S_LAUNCH(kc, grid, block, smem_chain(g, false, true), st, x);
				CKL();
				x.randomOnly = 0;
			}
			// fork: the commit of this call runs beside the synthesis (both only read what the chain wrote)
			const bool forkCommit = nSub == 1 && !e->profiling && e->commitStream;
			cudaStream_t cst = forkCommit ? e->commitStream : st;
			if (forkCommit) {
				CK(cudaEventRecord(e->evFork, st));
				CK(cudaStreamWaitEvent(cst, e->evFork, 0));
			}
			if (pairFft) {
				PROF(PK_SYNTH, B200S_LAUNCH(synth2_kernel(g), dim3(g.C, x.sCount), dim3(256), smem_synth2(g), st, x));
			} else {
				ChainKernel ks = synth_kernel(g);
				PROF(PK_SYNTH, B200S_LAUNCH(ks, dim3(g.C, x.sCount), dim3(kThreads), smem_synth(g), st, x));
			}
			// (measured: folding the commit into k_synth2's CTAs costs more there than the launch saves: 4.76 vs 4.68 ms)
			PROF(PK_COMMIT, B200S_LAUNCH(k_commit, dim3(g.C, x.sCount), dim3(kThreads), 0, cst, x));
			if (forkCommit) {
				CK(cudaEventRecord(e->evJoin, cst));
				CK(cudaStreamWaitEvent(st, e->evJoin, 0));
			}
		}
		if (hOut && nOut > 0)
			CK(cudaMemcpyAsync(hOut + (size_t)x.sBase * g.C * nOut, dOut + (size_t)x.sBase * g.C * nOut, sizeof(float) * (size_t)x.sCount * g.C * nOut, cudaMemcpyDeviceToHost, st));
		if (hOut16 && nOut > 0) {
			const size_t off = (size_t)x.sBase * g.C * nOut, cnt = (size_t)x.sCount * g.C * nOut;
			B200S_LAUNCH(k_pcm16_out, dim3((unsigned)std::min<size_t>((cnt + 1023) / 1024, 4096)), dim3(256), 0, st, dOut + off, e->dOut16 + off, cnt);
			CKL();
			CK(cudaMemcpyAsync(hOut16 + off, e->dOut16 + off, sizeof(short) * cnt, cudaMemcpyDeviceToHost, st));
		}
		if (nSub > 1) {
			CK(cudaEventRecord(e->evSubDone[sub], st));
			CK(cudaStreamWaitEvent(e->stream, e->evSubDone[sub], 0));
		}
	}
	e->histCur ^= 1;
	return 0;
}

static int seek_impl(b200s_engine *e, const float *dIn, int chanStride, long long streamStride, int n, double playbackRate, const double *rates = nullptr,
                     const long long *bankEnds = nullptr, long long bankLen = 0) {
	const Cfg &g = e->cfg;
	e->chainedGroups = 0;
	Ctx x = make_ctx(e);
	x.in = dIn; x.nIn = n; x.inChanStride = chanStride; x.inStreamStride = streamStride;
	if (bankEnds) { // the windows are cut out of a device-resident audio bank by k_seek itself
		int rc;
		if (!e->dSeekEnd && (rc = dalloc(e, &e->dSeekEnd, (size_t)g.S))) return rc;
		CK(cudaMemcpyAsync(e->dSeekEnd, bankEnds, sizeof(long long) * g.S, cudaMemcpyHostToDevice, e->stream));
		x.seekEnd = e->dSeekEnd;
		x.bankLen = bankLen;
	}
	float stf = (playbackRate * g.H > 1) ? float(1 / playbackRate) : float(g.H); // :164
	e->seekMayRandom = stf > 1.9f; // the next block takes this as its time factor (:312)
	if (rates) { // one playback rate per stream (a server whose streams follow their own time maps)
		int rc;
		if (!e->dSeekStf && (rc = dalloc(e, &e->dSeekStf, (size_t)g.S))) return rc;
		e->hSeekStf.resize(g.S);
		e->seekMayRandom = false;
		for (int s = 0; s < g.S; ++s) {
			e->hSeekStf[s] = (rates[s] * g.H > 1) ? float(1 / rates[s]) : float(g.H);
			if (e->hSeekStf[s] > 1.9f) e->seekMayRandom = true;
		}
		CK(cudaMemcpyAsync(e->dSeekStf, e->hSeekStf.data(), sizeof(float) * g.S, cudaMemcpyHostToDevice, e->stream));
		x.seekStf = e->dSeekStf;
	}
	B200S_LAUNCH(k_seek, dim3(g.S), dim3(kThreads), 0, e->stream, x, stf);
	CKL();
	return 0;
}

static int flush_impl(b200s_engine *e, float *dOut, int outChanStride, long long outStreamStride, int nOut, float playbackRate) {
	const Cfg &g = e->cfg;
	e->chainedGroups = 0;
	int rc;
	int outputBlock = std::max(0, nOut - g.H); // :439
	if (outputBlock > 0) {
		int nIn = int(outputBlock * playbackRate); // :440
		if (nIn < 0) nIn = 0;
		size_t zn = (size_t)g.S * g.C * std::max(nIn, 1);
		if ((rc = ensure_buf(e, &e->dZero, &e->zeroCap, zn, true))) return rc;
		if ((rc = process_impl(e, e->dZero, nIn, (long long)g.C * nIn, nIn, dOut, outChanStride, outStreamStride, outputBlock))) return rc;
	}
	int tail = nOut - outputBlock;
	Ctx x = make_ctx(e);
	x.out = dOut; x.outChanStride = outChanStride; x.outStreamStride = outStreamStride;
	if (tail > 0) {
		B200S_LAUNCH(k_flush_tail, dim3(g.C, g.S), dim3(kThreads), sizeof(float) * g.B, e->stream, x, outputBlock, tail);
		CKL();
	}
	return reset_impl(e, false); // stft.reset(0.1); prevInput = output = 0 (:456-463)
}

static int output_seek_impl(b200s_engine *e, const float *dIn, int chanStride, long long streamStride, int inputLength) {
	const Cfg &g = e->cfg;
	int rc;
	if ((rc = reset_impl(e, true))) return rc; // :175
	const int inLat = g.B - g.B / 2, outLat = g.B / 2 + (g.split ? g.H : 0);
	int surplus = std::max(inputLength - inLat, 0);            // :177
	float playbackRate = surplus / float(outLat);              // :178
	int seekSamples = inputLength - surplus;                   // :181
	if ((rc = seek_impl(e, dIn, chanStride, streamStride, seekSamples, playbackRate))) return rc;
	size_t tn = (size_t)g.S * g.C * outLat;
	if ((rc = ensure_buf(e, &e->dTmp, &e->tmpCap, tn, false))) return rc;
	if ((rc = process_impl(e, dIn + seekSamples, chanStride, streamStride, surplus, e->dTmp, outLat, (long long)g.C * outLat, outLat))) return rc; // :196
	Ctx x = make_ctx(e);
	B200S_LAUNCH(k_add_output, dim3(g.C, g.S), dim3(kThreads), 0, e->stream, x, e->dTmp, outLat); // :199-203
	CKL();
	return 0;
}

static int stage_in(b200s_engine *e, const float *in, int n) {
	size_t cnt = (size_t)e->cfg.S * e->cfg.C * std::max(n, 1);
	int rc;
	if ((rc = ensure_buf(e, &e->dIn, &e->inCap, cnt, false))) return rc;
	if (n > 0 && in) CK(cudaMemcpyAsync(e->dIn, in, sizeof(float) * (size_t)e->cfg.S * e->cfg.C * n, cudaMemcpyHostToDevice, e->stream));
	return 0;
}
static int stage_out(b200s_engine *e, int n) {
	size_t cnt = (size_t)e->cfg.S * e->cfg.C * std::max(n, 1);
	return ensure_buf(e, &e->dOut, &e->outCap, cnt, false);
}
static int fetch_out(b200s_engine *e, float *out, int n) {
	if (n > 0) CK(cudaMemcpyAsync(out, e->dOut, sizeof(float) * (size_t)e->cfg.S * e->cfg.C * n, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}

// =================================================================================================
// extern "C" ABI
// =================================================================================================
extern "C" {

int b200s_create(int batch, long seed, int device, b200s_engine **out) {
	if (!out || batch < 1) {
		g_createError = "b200s_create: bad arguments";
		return B200S_EINVAL;
	}
	*out = 0;
	int n = 0;
	cudaError_t ce = cudaGetDeviceCount(&n);
	if (ce != cudaSuccess || n <= 0 || device < 0 || device >= n) {
		g_createError = std::string("no usable CUDA device (this library has no CPU path): ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device index out of range");
		return B200S_ENODEVICE;
	}
	b200s_engine *e = new b200s_engine();
	e->S = batch;
	e->seed = seed;
	e->device = device;
	memset(&e->cfg, 0, sizeof(e->cfg));
	memset(&e->prm, 0, sizeof(e->prm));
	e->prm.freqMultiplier = 1;
	e->prm.freqTonalityLimit = 0.5f; // :513
	e->prm.formantMultiplier = e->prm.invFormantMultiplier = 1;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreate(&e->evStart) != cudaSuccess || cudaEventCreate(&e->evStop) != cudaSuccess) {
		g_createError = "CUDA stream/event creation failed";
		delete e;
		return B200S_ECUDA;
	}
	e->ownStream = true;
	{
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		if (cudaStreamCreateWithPriority(&e->commitStream, cudaStreamNonBlocking, hi) != cudaSuccess ||
		    cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&e->evJoin, cudaEventDisableTiming) != cudaSuccess)
			e->commitStream = 0; // (no fork then: the commit follows the synthesis on the main stream)
	}
	{ // sub-batch streams: highest priority first
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		bool ok = cudaEventCreateWithFlags(&e->evBegin, cudaEventDisableTiming) == cudaSuccess;
		for (int i = 0; i < b200s_engine::kMaxSub && ok; ++i) {
			int pr = hi + i;
			if (pr > lo) pr = lo;
			ok = cudaStreamCreateWithPriority(&e->subStream[i], cudaStreamNonBlocking, pr) == cudaSuccess &&
			     cudaEventCreateWithFlags(&e->evSubDone[i], cudaEventDisableTiming) == cudaSuccess;
		}
		// Measured on B200 (profiles/r01_bench_v7_subbatch.json): splitting the 1024-stream batch over 4
		// prioritised streams is SLOWER (11.9 vs 9.7 ms/step) -- the wavefront kernel is bound by per-warp
		// latency, not by occupancy, so co-running FFT CTAs only steal its issue slots.  Kept (off) for
		// batches large enough to oversubscribe the SMs; b200s_set_sub_batches() turns it on.
		e->nSub = 1;
		e->maxSub = ok ? b200s_engine::kMaxSub : 1;
		if (const char *hp = getenv("B200S_HOST_PARTS")) e->nHostParts = std::max(1, atoi(hp)); // tuning knob of the host-buffer pipeline
	}
	*out = e;
	return 0;
}
void b200s_destroy(b200s_engine *e) {
	if (!e) return;
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->stream);
	free_all(e);
	dfree(e->dMapIn);
	dfree(e->dMapOut);
	dfree(e->dRng);
	dfree(e->dDiag);
	dfree(e->dSeekStf);
	dfree(e->dSeekEnd);
	for (int i = 0; i < b200s_engine::kMaxSub; ++i) {
		if (e->subStream[i]) cudaStreamDestroy(e->subStream[i]);
		if (e->evSubDone[i]) cudaEventDestroy(e->evSubDone[i]);
	}
	if (e->evBegin) cudaEventDestroy(e->evBegin);
	if (e->commitStream) cudaStreamDestroy(e->commitStream);
	if (e->evFork) cudaEventDestroy(e->evFork);
	if (e->evJoin) cudaEventDestroy(e->evJoin);
	if (e->evStart) cudaEventDestroy(e->evStart);
	if (e->evStop) cudaEventDestroy(e->evStop);
	if (e->ownStream && e->stream) cudaStreamDestroy(e->stream);
	delete e;
}
const char *b200s_last_error(const b200s_engine *e) { return e ? e->err.c_str() : g_createError.c_str(); }
int b200s_version(int *major, int *minor, int *patch) {
	if (major) *major = 1;
	if (minor) *minor = 3;
	if (patch) *patch = 2;
	return 0;
}
int b200s_set_stream(b200s_engine *e, void *cuda_stream) {
	if (!e) return B200S_EINVAL;
	CK(cudaStreamSynchronize(e->stream));
	if (e->ownStream && e->stream) cudaStreamDestroy(e->stream);
	e->stream = (cudaStream_t)cuda_stream;
	e->ownStream = false;
	return 0;
}
int b200s_set_sub_batches(b200s_engine *e, int n) {
	if (!e || n < 1) return B200S_EINVAL;
	e->nSub = n > e->maxSub ? e->maxSub : n;
	return 0;
}
int b200s_set_tuning(b200s_engine *e, int key, int value) {
	if (!e) return B200S_EINVAL;
	if (key == 0 && value >= 0 && value <= 6) {
#ifndef B200S_KEEP_OLD_KERNELS
		if (value == 1 || value == 3 || value == 5) {
			e->err = "b200s_set_tuning: chain generations 1, 3 and 5 are superseded and not part of this build (-DB200S_KEEP_OLD_KERNELS)";
			return B200S_EUNSUPPORTED;
		}
#endif
		e->chainV = value;
	}
	else if (key == 1 && (value == 0 || value == 1)) e->fftV1 = value;
	else if (key == 2 && value >= 1) e->nHostParts = std::min(value, (int)b200s_engine::kMaxSub);
	else if (key == 3 && (value == 0 || value == 1)) e->exactMath = value;
	else if (key == 4 && (value == 0 || value == 1)) e->stepMajor = value;
	else if (key == 5 && (value == 0 || value == 1)) e->dual = value;
	else {
		e->err = "b200s_set_tuning: unknown key or value";
		return B200S_EINVAL;
	}
	return 0;
}
int b200s_synchronize(b200s_engine *e) {
	if (!e) return B200S_EINVAL;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}

int b200s_preset_default(b200s_engine *e, int channels, float sr, int split) {
	return configure_impl(e, channels, int(sr * 0.12), int(sr * 0.03), split); // :63-65
}
int b200s_preset_cheaper(b200s_engine *e, int channels, float sr, int split) {
	return configure_impl(e, channels, int(sr * 0.1), int(sr * 0.04), split); // :66-68
}
int b200s_configure(b200s_engine *e, int channels, int block, int interval, int split) { return configure_impl(e, channels, block, interval, split); }
int b200s_reset(b200s_engine *e) {
	NEED_CFG();
	return reset_impl(e, true);
}
int b200s_reserve(b200s_engine *e, int maxIn, int maxOut) {
	NEED_CFG();
	int rc;
	if ((rc = ensure_scratch(e, maxOut))) return rc;
	if (!params_plain(e) && (rc = e->stepMajor ? ensure_stepmajor(e) : ensure_coef(e))) return rc;
	if ((rc = stage_in(e, 0, 0))) return rc;
	size_t ci = (size_t)e->cfg.S * e->cfg.C * std::max(maxIn, 1), co = (size_t)e->cfg.S * e->cfg.C * std::max(maxOut, 1);
	if ((rc = ensure_buf(e, &e->dIn, &e->inCap, ci, false))) return rc;
	return ensure_buf(e, &e->dOut, &e->outCap, co, false);
}

int b200s_batch(const b200s_engine *e) { return e ? e->S : 0; }
int b200s_channels(const b200s_engine *e) { return e && e->configured ? e->cfg.C : 0; }
int b200s_block_samples(const b200s_engine *e) { return e && e->configured ? e->cfg.B : 0; }
int b200s_interval_samples(const b200s_engine *e) { return e && e->configured ? e->cfg.H : 0; }
int b200s_input_latency(const b200s_engine *e) { return e && e->configured ? e->cfg.B - e->cfg.B / 2 : 0; }                       // :42-44
int b200s_output_latency(const b200s_engine *e) { return e && e->configured ? e->cfg.B / 2 + (e->cfg.split ? e->cfg.H : 0) : 0; } // :45-47
int b200s_split_computation(const b200s_engine *e) { return e && e->configured ? e->cfg.split : 0; }
int b200s_seek_length(const b200s_engine *e) { return e && e->configured ? e->cfg.B + e->cfg.H : 0; } // :166-168
int b200s_output_seek_length(const b200s_engine *e, float rate) {                                       // :205-207
	return e && e->configured ? int(b200s_input_latency(e) + rate * b200s_output_latency(e)) : 0;
}
int b200s_fft_samples(const b200s_engine *e) { return e && e->configured ? e->cfg.N : 0; }
int b200s_bands(const b200s_engine *e) { return e && e->configured ? e->cfg.K : 0; }

int b200s_set_transpose_factor(b200s_engine *e, float multiplier, float tonality) { // :107-115
	if (!e) return B200S_EINVAL;
	e->prm.freqMultiplier = multiplier;
	if (tonality > 0) e->prm.freqTonalityLimit = tonality / std::sqrt(multiplier);
	else e->prm.freqTonalityLimit = 1;
	e->prm.mapN = 0;
	return 0;
}
int b200s_set_transpose_semitones(b200s_engine *e, float semitones, float tonality) { // :116-118
	return b200s_set_transpose_factor(e, std::pow(2, semitones / 12), tonality);
}
int b200s_set_formant_factor(b200s_engine *e, float multiplier, int comp) { // :124-128
	if (!e) return B200S_EINVAL;
	e->prm.formantMultiplier = multiplier;
	e->prm.invFormantMultiplier = 1 / multiplier;
	e->prm.formantCompensation = comp ? 1 : 0;
	return 0;
}
int b200s_set_formant_semitones(b200s_engine *e, float semitones, int comp) { // :129-131
	return b200s_set_formant_factor(e, std::pow(2, semitones / 12), comp);
}
int b200s_set_formant_base(b200s_engine *e, float f) { // :133-135
	if (!e) return B200S_EINVAL;
	e->prm.formantBaseFreq = f;
	return 0;
}
int b200s_set_freq_map_table(b200s_engine *e, const float *fin, const float *fout, int n) {
	if (!e || n < 0 || (n > 0 && (!fin || !fout))) return B200S_EINVAL;
	CK(cudaStreamSynchronize(e->stream));
	dfree(e->dMapIn);
	dfree(e->dMapOut);
	e->prm.mapN = 0;
	if (n == 0) return 0;
	int rc;
	if ((rc = dalloc(e, &e->dMapIn, n))) return rc;
	if ((rc = dalloc(e, &e->dMapOut, n))) return rc;
	CK(cudaMemcpy(e->dMapIn, fin, sizeof(float) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->dMapOut, fout, sizeof(float) * n, cudaMemcpyHostToDevice));
	e->prm.mapN = n;
	return 0;
}

int b200s_seek_device(b200s_engine *e, const float *dIn, int n, double rate) {
	NEED_CFG();
	if (n < 0) return B200S_EINVAL;
	return seek_impl(e, dIn, n, (long long)e->cfg.C * n, n, rate);
}
int b200s_process_device(b200s_engine *e, const float *dIn, int nIn, float *dOut, int nOut) {
	NEED_CFG();
	int rc;
	return process_impl(e, dIn, nIn, (long long)e->cfg.C * nIn, nIn, dOut, nOut, (long long)e->cfg.C * nOut, nOut);
}
int b200s_flush_device(b200s_engine *e, float *dOut, int nOut, float rate) {
	NEED_CFG();
	if (nOut < 0) return B200S_EINVAL;
	return flush_impl(e, dOut, nOut, (long long)e->cfg.C * nOut, nOut, rate);
}
int b200s_seek(b200s_engine *e, const float *in, int n, double rate) {
	NEED_CFG();
	if (n < 0) return B200S_EINVAL;
	int rc;
	if ((rc = stage_in(e, in, n))) return rc;
	if ((rc = seek_impl(e, e->dIn, n, (long long)e->cfg.C * n, n, rate))) return rc;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
int b200s_seek_rates(b200s_engine *e, const float *in, int n, const double *rates) {
	NEED_CFG();
	if (n < 0 || !rates) return B200S_EINVAL;
	int rc;
	if ((rc = stage_in(e, in, n))) return rc;
	if ((rc = seek_impl(e, e->dIn, n, (long long)e->cfg.C * n, n, 1.0, rates))) return rc;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
int b200s_live_seek(b200s_engine *e, const float *d_bank, long long bank_len, const long long *window_end, int window, const double *rates) {
	NEED_CFG();
	if (window < 0 || bank_len < 0 || !d_bank || !window_end || !rates) return B200S_EINVAL;
	if (bank_len > 0x7fffffffLL) {
		e->err = "b200s_live_seek: bank rows longer than 2^31-1 samples are not supported";
		return B200S_EUNSUPPORTED;
	}
	return seek_impl(e, d_bank, (int)bank_len, (long long)e->cfg.C * bank_len, window, 1.0, rates, window_end, bank_len);
}
int b200s_process(b200s_engine *e, const float *in, int nIn, float *out, int nOut) {
	NEED_CFG();
	int rc;
	if ((rc = stage_in(e, nullptr, nIn))) return rc; // capacity only: the copies are issued per stream group
	if ((rc = stage_out(e, nOut))) return rc;
	if ((rc = process_impl(e, e->dIn, nIn, (long long)e->cfg.C * nIn, nIn, e->dOut, nOut, (long long)e->cfg.C * nOut, nOut, in, out))) return rc;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
// 16-bit PCM host buffers (asynchronous like b200s_process_async when `wait` is 0)
int b200s_process_pcm16(b200s_engine *e, const short *in, int nIn, short *out, int nOut, int wait) {
	NEED_CFG();
	int rc;
	const size_t ci = (size_t)e->cfg.S * e->cfg.C * std::max(nIn, 1), co = (size_t)e->cfg.S * e->cfg.C * std::max(nOut, 1);
	if (ci > e->inCap || co > e->outCap || ci > e->in16Cap || co > e->out16Cap || !e->dIn || !e->dOut) CK(cudaStreamSynchronize(e->stream)); // staging grows: drain first
	if ((rc = stage_in(e, nullptr, nIn))) return rc;
	if ((rc = stage_out(e, nOut))) return rc;
	if (ci > e->in16Cap) {
		if ((rc = dalloc(e, &e->dIn16, ci))) return rc;
		e->in16Cap = ci;
	}
	if (co > e->out16Cap) {
		if ((rc = dalloc(e, &e->dOut16, co))) return rc;
		e->out16Cap = co;
	}
	if ((rc = process_impl(e, e->dIn, nIn, (long long)e->cfg.C * nIn, nIn, e->dOut, nOut, (long long)e->cfg.C * nOut, nOut, nullptr, nullptr, in, out))) return rc;
	if (wait) CK(cudaStreamSynchronize(e->stream));
	return 0;
}
int b200s_process_async(b200s_engine *e, const float *in, int nIn, float *out, int nOut) {
	NEED_CFG();
	int rc;
	if ((size_t)e->cfg.S * e->cfg.C * std::max(nIn, 1) > e->inCap || (size_t)e->cfg.S * e->cfg.C * std::max(nOut, 1) > e->outCap || !e->dIn || !e->dOut) {
		// growing the staging buffers would invalidate copies still in flight: drain first
		CK(cudaStreamSynchronize(e->stream));
	}
	if ((rc = stage_in(e, nullptr, nIn))) return rc;
	if ((rc = stage_out(e, nOut))) return rc;
	return process_impl(e, e->dIn, nIn, (long long)e->cfg.C * nIn, nIn, e->dOut, nOut, (long long)e->cfg.C * nOut, nOut, in, out);
}
int b200s_flush(b200s_engine *e, float *out, int nOut, float rate) {
	NEED_CFG();
	if (nOut < 0) return B200S_EINVAL;
	int rc;
	if ((rc = stage_out(e, nOut))) return rc;
	if ((rc = flush_impl(e, e->dOut, nOut, (long long)e->cfg.C * nOut, nOut, rate))) return rc;
	return fetch_out(e, out, nOut);
}
int b200s_output_seek(b200s_engine *e, const float *in, int inputLength) {
	NEED_CFG();
	if (inputLength < 0) return B200S_EINVAL;
	int rc;
	if ((rc = stage_in(e, in, inputLength))) return rc;
	if ((rc = output_seek_impl(e, e->dIn, inputLength, (long long)e->cfg.C * inputLength, inputLength))) return rc;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
int b200s_exact(b200s_engine *e, const float *in, int nIn, float *out, int nOut, int *ok) { // :467-491
	NEED_CFG();
	if (nIn < 0 || nOut <= 0) return B200S_EINVAL;
	int rc;
	const Cfg &g = e->cfg;
	float playbackRate = nIn / float(nOut);
	int seekLen = b200s_output_seek_length(e, playbackRate);
	if (ok) *ok = 0;
	if (nIn < seekLen) {
		memset(out, 0, sizeof(float) * (size_t)g.S * g.C * nOut);
		return 0;
	}
	if ((rc = stage_in(e, in, nIn))) return rc;
	if ((rc = stage_out(e, nOut))) return rc;
	const long long inSS = (long long)g.C * nIn, outSS = (long long)g.C * nOut;
	if ((rc = output_seek_impl(e, e->dIn, nIn, inSS, seekLen))) return rc;
	int outputIndex = int(nOut - seekLen / playbackRate); // :484
	if ((rc = process_impl(e, e->dIn + seekLen, nIn, inSS, nIn - seekLen, e->dOut, nOut, outSS, outputIndex))) return rc;
	if ((rc = flush_impl(e, e->dOut + outputIndex, nOut, outSS, nOut - outputIndex, playbackRate))) return rc;
	if (ok) *ok = 1;
	return fetch_out(e, out, nOut);
}

int b200s_timer_start(b200s_engine *e) {
	if (!e) return B200S_EINVAL;
	CK(cudaEventRecord(e->evStart, e->stream));
	return 0;
}
int b200s_timer_stop(b200s_engine *e, float *ms) {
	if (!e || !ms) return B200S_EINVAL;
	CK(cudaEventRecord(e->evStop, e->stream));
	CK(cudaEventSynchronize(e->evStop));
	CK(cudaEventElapsedTime(ms, e->evStart, e->evStop));
	return 0;
}
long long b200s_kernel_launches(const b200s_engine *e) { return e ? e->launches : 0; }
long long b200s_device_allocations(const b200s_engine *e) { return e ? e->allocs : 0; }
long long b200s_unserved_random_blocks(b200s_engine *e) {
	if (!e || !e->dDiag) return 0;
	unsigned long long v = 0;
	if (cudaStreamSynchronize(e->stream) != cudaSuccess || cudaMemcpy(&v, e->dDiag, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
	return (long long)v;
}
int b200s_profile_begin(b200s_engine *e) {
	if (!e) return B200S_EINVAL;
	e->profiling = true;
	return 0;
}
int b200s_profile_end(b200s_engine *e, float *ms, int *counts, int n) {
	if (!e || !ms || !counts || n < PK_COUNT) return B200S_EINVAL;
	e->profiling = false;
	CK(cudaStreamSynchronize(e->stream));
	for (int i = 0; i < n; ++i) {
		ms[i] = 0;
		counts[i] = 0;
	}
	for (size_t i = 0; i < e->profKind.size(); ++i) {
		float t = 0;
		CK(cudaEventElapsedTime(&t, e->profEv[2 * i], e->profEv[2 * i + 1]));
		ms[e->profKind[i]] += t;
		counts[e->profKind[i]] += 1;
	}
	for (cudaEvent_t ev : e->profEv) cudaEventDestroy(ev);
	e->profEv.clear();
	e->profKind.clear();
	return 0;
}

// Device self-test of the branch-free division / square root used by the phase chain.
int b200s_selftest_divsqrt(b200s_engine *e, long long n, long long seed, long long *divMismatch, long long *sqrtMismatch) {
	if (!e || n <= 0) return B200S_EINVAL;
	unsigned long long *d = 0;
	int rc;
	if ((rc = dalloc(e, &d, 2))) return rc;
	CK(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
	const int threads = 256, blocks = 1184, per = (int)((n + (long long)threads * blocks - 1) / ((long long)threads * blocks));
	B200S_LAUNCH(k_selftest_divsqrt, dim3(blocks), dim3(threads), 0, e->stream, (unsigned long long)seed, per, d);
	CKL();
	unsigned long long h[2] = {0, 0};
	CK(cudaStreamSynchronize(e->stream));
	CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
	cudaFree(d);
	if (divMismatch) *divMismatch = (long long)h[0];
	if (sqrtMismatch) *sqrtMismatch = (long long)h[1];
	return 0;
}

// ---- white-box state ----
static int state_desc(const b200s_engine *e, int what, float **ptr, size_t *perStream) {
	const Cfg &g = e->cfg;
	switch (what) {
	case 0: *ptr = (float *)e->dStIn; *perStream = (size_t)2 * g.C * g.K; return 0;
	case 1: *ptr = (float *)e->dStPrev; *perStream = (size_t)2 * g.C * g.K; return 0;
	case 2: *ptr = (float *)e->dStOut; *perStream = (size_t)2 * g.C * g.K; return 0;
	case 4: *ptr = e->dStPredE; *perStream = (size_t)g.C * g.K; return 0;
	case 20: *ptr = e->dHist[e->histCur]; *perStream = (size_t)g.C * g.histLen; return 0;
	case 21: *ptr = e->dPend; *perStream = (size_t)g.C * g.pendLen; return 0;
	case 22: *ptr = e->dPendWp; *perStream = (size_t)g.C * g.pendLen; return 0;
	}
	return B200S_EINVAL;
}
int b200s_state_size(const b200s_engine *e, int what) {
	if (!e || !e->configured) return B200S_EINVAL;
	float *p;
	size_t n;
	if (state_desc(e, what, &p, &n)) return B200S_EINVAL;
	return (int)n;
}
int b200s_get_state(b200s_engine *e, int what, float *dst) {
	NEED_CFG();
	float *p;
	size_t n;
	if (state_desc(e, what, &p, &n)) return B200S_EINVAL;
	CK(cudaStreamSynchronize(e->stream));
	CK(cudaMemcpy(dst, p, sizeof(float) * n * e->S, cudaMemcpyDeviceToHost));
	return 0;
}
int b200s_set_state(b200s_engine *e, int what, const float *src) {
	NEED_CFG();
	float *p;
	size_t n;
	if (state_desc(e, what, &p, &n)) return B200S_EINVAL;
	CK(cudaStreamSynchronize(e->stream));
	CK(cudaMemcpy(p, src, sizeof(float) * n * e->S, cudaMemcpyHostToDevice));
	return 0;
}

} // extern "C"
