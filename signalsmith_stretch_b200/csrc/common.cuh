// common.cuh -- shared types of the B200 stretch engine (device + host).
//
// Data layout in HBM (all float32 / float2, per engine = per batch of S streams on one GPU):
//   hist     [S][C][B+H]        input history, most recent sample last        (reference: STFT input ring, B+H+1)
//   pend     [S][C][B(+H)]      pending overlap-add accumulator, linearised   (reference: STFT output ring)
//   pendWp   [S][C][B(+H)]      matching windowProducts (one identical copy per channel: single-writer CTAs)
//   stIn/stPrev/stOut [S][C][K] float2   Band::input / prevInput / output of the last block (:538-542)
//   stPredE  [S][C][K]          Prediction::energy of the last block (:593)
// and per process() call scratch, F = blocks in the call:
//   spec     [S][2F][C][K] float2   analysis spectra (slot 2f: block f, slot 2f+1: its re-analysed predecessor)
//   cS, cM   [S][F][K]              smoothed energy / formant envelope of the block (mapped / formant calls only)
//   cE       [S][F][C][K]           Prediction::energy
//   cPI,cFT,cT1,cT2 [S][F][C][K] float2   Prediction::input, freqTwist, short/long vertical twists
//   Y        [S][F][C][K] float2    final Band::output of every block
#pragma once

#ifdef B200S_EMU
#include "cuda_emu.h" // tests/cuda_emu: thread-per-CUDA-thread CPU emulator, TEST BUILDS ONLY
#else
#include <cuda_runtime.h>
#define B200S_SHARED __shared__
#define B200S_DYN_SHARED extern __shared__ float4 dyn_smem[];
#define B200S_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <stdint.h>

namespace b200s {

// ---- constants of the reference ----
#define B200S_NOISE_FLOOR 1e-15f      // signalsmith-stretch.h:508
#define B200S_MAX_CLEAN_STRETCH 2.0f  // :509
#define B200S_ALMOST_ZERO 1e-30f      // dependency's windowProducts floor (measured)
#define B200S_NEVER (1ll << 60)       // blockProcess.samplesSinceLast initial value (:496, SIZE_MAX there)

enum FrameFlags {
	FR_NEW_SPECTRUM = 1, // :299
	FR_REANALYSE = 2,    // :303
	FR_MAPPED = 4,       // :300
	FR_FORMANTS = 8,     // :310
	FR_RANDOM = 16       // :639
};

// static configuration of one engine
struct Cfg {
	int S, C;          // streams, channels
	int B, H, N, K;    // block, interval, fftSamples, bands (= N/2 = complex FFT size)
	int L;             // longVerticalStep = round(N/H) (:637)
	int split;         // splitComputation
	int histLen;       // B + H
	int pendLen;       // B (+ H when split)
	int addOff;        // where a block lands in `pend` (H when split, :294-297)
	int o;             // analysis/synthesis offset = B/2
	int nStages;       // Stockham stages of the K-point complex FFT
	int radix[20];
};

// user parameters (:107-135), broadcast to every stream
struct Params {
	float freqMultiplier, freqTonalityLimit;
	float formantMultiplier, invFormantMultiplier, formantBaseFreq;
	int formantCompensation;
	int mapN;                 // >0: custom piecewise-linear frequency map (setFreqMap, :120)
	const float *mapIn, *mapOut;
};

// per-stream block scheduler state (:494-529)
struct Sched {
	long long samplesSinceLast;
	long long silenceCounter;
	int prevInputOffset;
	int didSeek;
	int silenceFirst;
	float seekTimeFactor;
	long long zeroRun; // input samples, counted back from the newest, known to be exactly 0.0f (>= histLen: the history is all zeros)
};

// one block of the schedule (:281-319)
struct Frame {
	int t;           // output index within the call at which the block triggers
	int inputOffset; // :288
	int flags;
	float timeFactor; // :312, before the clamp of :638
	int inSlot, prevSlot; // spectrum slots: 0 = stIn, 1 = stPrev, 2+2f / 3+2f = this call's analyses
	unsigned rng;         // FR_RANDOM: state of the stream's random engine before this block's 2K-2 draws (:749,:769)
};

// per-stream result of the planner for one process() call
struct Call {
	int bypass;   // silence bypass (:240-271)
	int nFrames;
	int finalIn, finalPrev; // slots that become stIn / stPrev after the call
	int nJobs;              // analyses of this call (entries of the stream's Job list)
	int hasRandom;          // some block of the call stretches beyond 2x (:639): the stream takes the k_prep + k_chain path
};

// one windowed analysis FFT of the call (:333-376): block samples start at stream index `start`
// (history ++ input), the spectrum goes to row `row` = (2*f + w)*C + c of the stream's `spec` scratch
struct Job {
	int start, row, c;
};

// everything a kernel needs, passed by value
struct Ctx {
	Cfg cfg;
	Params prm;
	// tables
	const float *window, *winProd, *wpReset;
	const float2 *rot, *twiddle, *pretw;
	const float4 *anaTab; // [K] {window[n+o], window[n+o-K], pretw[n].x, pretw[n].y}: one load per element in k_analyse2
	float2 rot0, rotStep; // rot[b+1] = rot[b] * rotStep in float, the reference's own recurrence (:647-655)
	// state
	Sched *sched;
	float *histCur, *histNext;
	float *pend, *pendWp;
	float2 *stIn, *stPrev, *stOut;
	float *stPredE;
	float *stPitch; // [S][2] freqEstimateWeighted, freqEstimateWeight (:927-928): automatic pitch estimate of the formant envelope
	// stereo direct path: spectra are channel-interleaved float4 {re0, re1, im0, im1} per bin (chain_direct3.cuh);
	// stIl [S][2][K] holds such copies of stIn / stPrev, made by k_plan at the start of the call
	float4 *stIl;
	int specIl;
	int wsRan; // k_chain_ws ran before k_chain_direct4 in this call: the latter only takes the streams the former left (chain_ws.cuh)
	float one; // 1.0f; passed as data so that ptxas cannot fold the exact packed add p*one + q (chain_direct3.cuh)
	// call scratch
	int maxFrames;
	Frame *frames; // [S][maxFrames]
	Call *call;    // [S]
	Job *jobs;     // [S][2*C*maxFrames], consecutive entries are transformed as one PAIR (fft2.cuh)
	int inAligned; // input / history rows allow 16-byte cp.async (pointer, strides and lengths multiples of 4 floats)
	float2 *spec, *Y, *cPI, *cFT, *cT1, *cT2;
	float *cE;
	// random time factors beyond 2x stretch (:639-640,:749,:769).  The reference draws from std::default_random_engine
	// (libstdc++: minstd_rand0, x <- 16807 x mod 2^31-1) through uniform_real_distribution<float>, 2K-2 draws per block in
	// bin order; draw i of a block is state * 16807^i, so every bin finds its draws by one modular multiplication.
	unsigned *rngState;     // [S] engine state per stream (persists over reset / configure, like the reference's member)
	const unsigned *rngPow; // [2K] 16807^i mod 2^31-1
	unsigned rngJump;       // 16807^(2K-2): one block's worth of draws
	const long long *seekEnd; // k_seek from a device-resident audio bank (b200s_live_seek): per stream, the bank index at which the
	long long bankLen;        // seek window ends; samples outside [0, bankLen) read as zero.  Null: x.in holds the window itself
	const float *seekStf;   // k_seek: per-stream seekTimeFactor (b200s_seek_rates), or null: the launch argument for every stream
	int randomPathOn;       // the kernels of the random path run in this call: the direct chain kernels leave hasRandom streams to them
	unsigned long long *diag; // [1] blocks whose random time factors could not be honoured (random path not launched; see engine.cu)
	int randomOnly;         // k_prep / k_chain launched beside the direct chain kernels: only streams with Call::hasRandom
	float2 *cT1u, *cT2u;    // [S][F][C][K] the "upwards" twists of a random block (:769-781; the downwards ones are cT1 / cT2)
	// step-major path of the mapped / formant configurations (chain_t.cuh): map + formant-ratio rows of k_prep's map-only mode,
	// and the chain's terms transposed by k_products: [S][tGroups][tRows][C][32 lanes]
	int mapOnly, tRows, tGroups;
	float *cMapB, *cMapG, *cRatio; // [S][maxFrames][K]
	float2 *tPI, *tFT, *tT1, *tT2;
	float *tE;
	float *cS, *cM; // [S][maxFrames][K]: smoothed energy (:816-848) and formant envelope (:986-1007) of every block: k_energy / k_passes -> k_prep
	float *cPitch; // [S][maxFrames] freqEstimate of every block when formantBaseFreq <= 0 (k_pitch)
	// sub-batch of streams this launch covers (the batch is processed as a few sub-batches on
	// prioritised CUDA streams so that different kernels of the sequence overlap on the SMs)
	int sBase, sCount;
	// I/O of the current call
	const float *in;
	float *out;
	int nIn, nOut;
	long long inStreamStride, outStreamStride; // floats between streams
	int inChanStride, outChanStride;           // floats between channels
};

} // namespace b200s
