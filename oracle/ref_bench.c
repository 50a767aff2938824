/* oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY (CPU baseline driver).
 *
 * Native multi-threaded driver for the reference's own shipped binary (the WASM blob of
 * /root/reference/web/emscripten/main.js:9 translated to C by oracle/wasm2c.py): one module
 * instance per audio stream (the reference is one object per stream, single-threaded:
 * README.md:93-97), a pthread pool over the host cores, CLOCK_MONOTONIC wall time
 * (BASELINE.md section 3).  Linked into oracle/_ref/libwasm_stretch.so next to the translated code.
 * Export letters: SURVEY.md Appendix D (f ctors, y main, h setBuffers, n presetDefault,
 * o presetCheaper, r setTransposeSemitones, w process). */
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

typedef struct W W;
W *wasm_new(void);
void wasm_free(W *);
uint8_t *wasm_mem(W *);
void wasm_export_f(W *);
uint32_t wasm_export_y(W *, uint32_t, uint32_t);
uint32_t wasm_export_h(W *, uint32_t, uint32_t);
void wasm_export_n(W *, uint32_t, float);
void wasm_export_o(W *, uint32_t, float);
void wasm_export_r(W *, float, float);
void wasm_export_w(W *, uint32_t, uint32_t);

typedef struct {
	int streams, channels, preset, nIn, nOut, chunkOut;
	float sr, semitones, tonality;
	const float *x;
	int next;
	pthread_mutex_t mu;
	double checksum;
	double procMax; /* largest per-thread sum of the time spent inside the process() loops (instance set-up excluded) */
} Job;

static double now_s(void) {
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *worker(void *arg) {
	Job *j = (Job *)arg;
	double proc = 0;
	for (;;) {
		pthread_mutex_lock(&j->mu);
		int s = j->next++;
		pthread_mutex_unlock(&j->mu);
		if (s >= j->streams) {
			pthread_mutex_lock(&j->mu);
			if (proc > j->procMax) j->procMax = proc;
			pthread_mutex_unlock(&j->mu);
			break;
		}
		W *w = wasm_new();
		wasm_export_f(w);
		wasm_export_y(w, 0, 0);
		if (j->preset == 0) wasm_export_n(w, j->channels, j->sr);
		else wasm_export_o(w, j->channels, j->sr);
		if (j->semitones != 0) wasm_export_r(w, j->semitones, j->tonality);
		double ratio = (double)j->nIn / (double)j->nOut;
		int maxIn = (int)(j->chunkOut * ratio) + 2;
		int len = maxIn > j->chunkOut ? maxIn : j->chunkOut;
		uint32_t ptr = wasm_export_h(w, j->channels, len);
		int i = 0, done = 0;
		double acc = 0;
		const double tp0 = now_s(); /* timed: the process() calls only, not the instance creation / preset above */
		while (done < j->nOut) {
			int co = j->chunkOut < j->nOut - done ? j->chunkOut : j->nOut - done;
			int ci = (int)((double)(done + co) * ratio + 0.5) - i;
			if (i + ci > j->nIn) ci = j->nIn - i;
			float *mem = (float *)(wasm_mem(w) + ptr);
			for (int c = 0; c < j->channels; ++c)
				memcpy(mem + (size_t)c * len, j->x + ((size_t)s * j->channels + c) * j->nIn + i, sizeof(float) * ci);
			wasm_export_w(w, ci, co);
			mem = (float *)(wasm_mem(w) + ptr);
			acc += mem[(size_t)j->channels * len + co - 1];
			i += ci;
			done += co;
		}
		proc += now_s() - tp0;
		pthread_mutex_lock(&j->mu);
		j->checksum += acc;
		pthread_mutex_unlock(&j->mu);
		wasm_free(w);
	}
	return 0;
}

/* returns wall seconds (instance creation included); *checksum receives a data-dependent value so the work cannot be
 * elided; *procSeconds (may be null) receives the wall time of the process() calls alone: the largest per-thread sum of
 * the intervals spent inside the process() loops, i.e. what the pool would take if instances were created beforehand */
double refbench_run2(int threads, int streams, int channels, float sr, int preset, float semitones, float tonality,
                     int nIn, int nOut, int chunkOut, const float *x, double *checksum, double *procSeconds) {
	Job j;
	memset(&j, 0, sizeof j);
	j.streams = streams; j.channels = channels; j.preset = preset; j.nIn = nIn; j.nOut = nOut; j.chunkOut = chunkOut;
	j.sr = sr; j.semitones = semitones; j.tonality = tonality; j.x = x;
	pthread_mutex_init(&j.mu, 0);
	pthread_t th[256];
	if (threads > 256) threads = 256;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < threads; ++t) pthread_create(&th[t], 0, worker, &j);
	for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (checksum) *checksum = j.checksum;
	if (procSeconds) *procSeconds = j.procMax;
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

double refbench_run(int threads, int streams, int channels, float sr, int preset, float semitones, float tonality,
                    int nIn, int nOut, int chunkOut, const float *x, double *checksum) {
	return refbench_run2(threads, streams, channels, sr, preset, semitones, tonality, nIn, nOut, chunkOut, x, checksum, 0);
}
