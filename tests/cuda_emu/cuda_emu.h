// tests/cuda_emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal "CUDA on CPU" shim: one cooperative fiber per CUDA thread, one block at a time, yielding barriers
// for __syncthreads / warp shuffles.  It exists because the build container has no GPU: it lets the
// `-m "not gpu"` tests execute the SAME kernel sources (signalsmith_stretch_b200/csrc/*.cu*)
// and check their logic against the oracle before they ever reach a B200.  It is compiled only by
// tests/cuda_emu/build.sh into tests/cuda_emu/_build/, is never shipped, and the product package
// cannot load it (signalsmith_stretch_b200 only loads the nvcc-built library).  Not a fallback.
#pragma once
#include <sched.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define B200S_SHARED static
#define B200S_DYN_SHARED

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
alignas(16) static float4 dyn_smem[(232 * 1024) / 16];

using std::max;
using std::min;

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }

namespace emu {
// ---- cooperative fibers: every CUDA thread of the block is a ucontext fiber on ONE OS thread, scheduled round
// robin; barriers and shuffles yield until everybody has arrived.  (The first version used one OS thread per CUDA
// thread and pthread barriers: 256 threads on a handful of cores spent almost all their time in futex calls.)
struct Fiber {
	ucontext_t uc;
	uint3 tIdx, bIdx;
	bool done;
};
struct Bar {
	int count = 0, gen = 0, n = 0;
};
static std::vector<Fiber> g_f;
static std::vector<char *> g_stacks;
static const size_t kStack = 512 * 1024;
static ucontext_t g_main;
static int g_cur = 0, g_T = 0;
static unsigned long long g_seed = getenv("B200S_EMU_SEED") ? strtoull(getenv("B200S_EMU_SEED"), nullptr, 10) : 0ull;
static Bar g_blockBar;
static std::vector<Bar> g_warpBar;
static std::vector<uint32_t> g_shfl;
static std::function<void()> g_fn;
static dim3 g_grid;
static thread_local int t_tid = 0;

static inline void load(int i) {
	g_cur = i;
	t_tid = i;
	threadIdx = g_f[i].tIdx;
	blockIdx = g_f[i].bIdx;
}
// give the processor to the next fiber that has not finished
static inline void yield() {
	const int from = g_cur;
	g_f[from].tIdx = threadIdx;
	g_f[from].bIdx = blockIdx;
	int nx = from;
	if (g_seed) { // B200S_EMU_SEED: pseudo-random choice of the next fiber, to shake out order-dependent code
		g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull;
		nx = (int)((g_seed >> 33) % (unsigned)g_T);
		if (nx == from) nx = nx + 1 == g_T ? 0 : nx + 1;
		nx = nx == 0 ? g_T - 1 : nx - 1; // the loop below starts one after nx
	}
	do {
		nx = nx + 1 == g_T ? 0 : nx + 1;
	} while (g_f[nx].done && nx != from);
	if (nx == from) return;
	load(nx);
	swapcontext(&g_f[from].uc, &g_f[nx].uc);
}
static inline void bar_wait(Bar &b) {
	const int gen = b.gen;
	if (++b.count == b.n) {
		b.count = 0;
		++b.gen;
	} else {
		while (b.gen == gen) yield();
	}
}
static inline void warp_barrier() { bar_wait(g_warpBar[t_tid >> 5]); }
template <typename T>
static inline T shfl_from(T v, int srcLane) {
	static_assert(sizeof(T) == 4, "32-bit shuffles only");
	uint32_t bits;
	memcpy(&bits, &v, 4);
	g_shfl[t_tid] = bits;
	warp_barrier();
	int base = t_tid & ~31, n = std::min<int>(32, (int)blockDim.x - base);
	uint32_t r = (srcLane >= 0 && srcLane < n) ? g_shfl[base + srcLane] : bits;
	warp_barrier();
	T out;
	memcpy(&out, &r, 4);
	return out;
}
static void trampoline() {
	for (unsigned bz = 0; bz < g_grid.z; ++bz)
		for (unsigned by = 0; by < g_grid.y; ++by)
			for (unsigned bx = 0; bx < g_grid.x; ++bx) {
				blockIdx = uint3{bx, by, bz};
				g_fn();
				bar_wait(g_blockBar); // one block at a time: the blocks share the emulated shared memory
			}
	const int me = g_cur;
	g_f[me].done = true;
	int nx = me;
	do {
		nx = nx + 1 == g_T ? 0 : nx + 1;
	} while (g_f[nx].done && nx != me);
	if (nx == me) {
		setcontext(&g_main); // the last fiber returns to the launcher
	} else {
		load(nx);
		setcontext(&g_f[nx].uc);
	}
}

template <typename F>
static void launch(dim3 grid, dim3 block, size_t smemBytes, F fn) {
	// canary behind the dynamic shared memory the launch asked for: a kernel that writes past its allocation (an
	// "illegal memory access" on the GPU) is caught here
	const size_t canaryBytes = std::min<size_t>(16384, sizeof(dyn_smem) - std::min(smemBytes, sizeof(dyn_smem)));
	unsigned char *canary = (unsigned char *)dyn_smem + std::min(smemBytes, sizeof(dyn_smem));
	memset(canary, 0xA5, canaryBytes);
	const int T = (int)(block.x * block.y * block.z);
	g_T = T;
	g_grid = grid;
	blockDim = block;
	gridDim = grid;
	g_fn = fn;
	g_blockBar = Bar();
	g_blockBar.n = T;
	const int nw = (T + 31) / 32;
	g_warpBar.assign(nw, Bar());
	for (int w = 0; w < nw; ++w) g_warpBar[w].n = std::min(32, T - 32 * w);
	g_shfl.assign(T, 0);
	g_f.resize(T);
	while ((int)g_stacks.size() < T) g_stacks.push_back((char *)malloc(kStack));
	for (int t = 0; t < T; ++t) {
		Fiber &f = g_f[t];
		f.done = false;
		f.tIdx = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
		f.bIdx = uint3{0, 0, 0};
		getcontext(&f.uc);
		f.uc.uc_stack.ss_sp = g_stacks[t];
		f.uc.uc_stack.ss_size = kStack;
		f.uc.uc_link = nullptr;
		makecontext(&f.uc, trampoline, 0);
	}
	if (T > 0 && grid.x * grid.y * grid.z > 0) {
		load(0);
		swapcontext(&g_main, &g_f[0].uc);
	}
	for (size_t i = 0; i < canaryBytes; ++i)
		if (canary[i] != 0xA5) {
			fprintf(stderr, "cuda_emu: a kernel wrote %zu bytes past its %zu bytes of dynamic shared memory\n", i + 1, smemBytes);
			abort();
		}
}
} // namespace emu
static inline void emu_yield() { emu::yield(); }

static inline void __syncthreads() { emu::bar_wait(emu::g_blockBar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) { return emu::shfl_from(v, (emu::t_tid & 31) - d); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d) { return emu::shfl_from(v, (emu::t_tid & 31) + d); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_from(v, src); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_from(v, (emu::t_tid & 31) ^ m); }
static inline int __any_sync(unsigned, int pred) {
	int r = 0;
	for (int l = 0; l < 32; ++l) r |= emu::shfl_from(pred ? 1 : 0, l);
	return r;
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline void atomicAdd(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

#define B200S_LAUNCH(kernel, grid, block, smem, stream, ...) \
	emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })

// ---- the handful of runtime calls engine.cu makes ----
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulator error"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 2; return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n + 64); return *p ? 0 : 1; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, int) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, int, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }

#ifdef B200S_EMU_EXACT_FFT
// Swap the oracle's double-precision modified real FFT into the analysis / synthesis kernels, so
// that everything EXCEPT the FFT can be compared bit-exactly with the oracle.
#include "../../oracle/fft_ref.h"
static inline void emu_exact_forward(const float *xw, int B, int o, int N, float2 *spec) {
	oracle::ModifiedRealFFT fft;
	fft.resize(N);
	std::vector<double> x(B);
	std::vector<oracle::cplx> X(N / 2);
	for (int i = 0; i < B; ++i) x[i] = xw[i];
	fft.forward(x.data(), B, o, X.data());
	for (int b = 0; b < N / 2; ++b) spec[b] = float2{float(X[b].real()), float(X[b].imag())};
}
static inline void emu_exact_inverse(const float2 *Y, int B, int o, int N, float *y) {
	oracle::ModifiedRealFFT fft;
	fft.resize(N);
	std::vector<oracle::cplx> X(N / 2);
	std::vector<double> t(B);
	for (int b = 0; b < N / 2; ++b) X[b] = oracle::cplx(Y[b].x, Y[b].y);
	fft.inverse(X.data(), t.data(), B, o);
	for (int i = 0; i < B; ++i) y[i] = float(t[i]);
}
#endif
