// cmd/stretch_cli.cpp -- file-mode caller of the B200 path: the role of the reference's cmd/main.cpp
// (flags and processing stages of cmd/main.cpp:11-86; its WAV helper cmd/util/wav.h is replaced by the
// minimal reader / writer below).  Written against the drop-in facade, i.e. the SAME calls the reference's
// command-line tool makes: presetDefault, setTransposeSemitones, setFormantSemitones, setFormantBase,
// outputSeek, process, flush.
//
//   stretch_cli in.wav out.wav [--semitones=S] [--tonality=Hz] [--time=T] [--formant=S] [--formant-comp]
//                              [--formant-base=Hz] [--split-computation]
//   stretch_cli -v
//
// Reads 16/24/32-bit PCM or 32-bit float WAV, writes 16-bit PCM (what the reference's tool writes).
// Every file is one stream (batch of 1): the batched C ABI (include/b200_stretch.h) is the interface
// for throughput; this tool is the plumbing either side of the path.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "signalsmith-stretch/signalsmith-stretch.h"

namespace {

struct Audio {
	int channels = 0;
	double sampleRate = 0;
	std::vector<std::vector<float>> ch; // planar
	size_t offset = 0;                  // view start, so that `audio[c][i]` reads sample offset + i
	struct View {
		float *p;
		float &operator[](size_t i) const { return p[i]; }
	};
	View operator[](int c) { return View{ch[c].data() + offset}; }
	size_t length() const { return ch.empty() ? 0 : ch[0].size(); }
	void resize(size_t n) {
		for (auto &v : ch) v.resize(n, 0.f);
	}
};

uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

bool read_wav(const std::string &path, Audio &a, std::string &err) {
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) {
		err = "cannot open " + path;
		return false;
	}
	std::vector<unsigned char> d;
	unsigned char buf[65536];
	size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
	fclose(f);
	if (d.size() < 12 || memcmp(&d[0], "RIFF", 4) || memcmp(&d[8], "WAVE", 4)) {
		err = "not a RIFF/WAVE file";
		return false;
	}
	int fmt = 0, bits = 0, frameBytes = 0;
	size_t pos = 12;
	const unsigned char *data = nullptr;
	size_t dataLen = 0;
	while (pos + 8 <= d.size()) {
		const uint32_t len = rd32(&d[pos + 4]);
		const unsigned char *body = &d[pos + 8];
		if (pos + 8 + len > d.size() && memcmp(&d[pos], "data", 4)) break;
		if (!memcmp(&d[pos], "fmt ", 4) && len >= 16) {
			fmt = rd16(body);
			a.channels = rd16(body + 2);
			a.sampleRate = rd32(body + 4);
			frameBytes = rd16(body + 12);
			bits = rd16(body + 14);
			if (fmt == 0xFFFE && len >= 26) fmt = rd16(body + 24); // WAVE_FORMAT_EXTENSIBLE: sub-format
		} else if (!memcmp(&d[pos], "data", 4)) {
			data = body;
			dataLen = std::min<size_t>(len, d.size() - (pos + 8));
			break;
		}
		pos += 8 + len + (len & 1);
	}
	if (!data || a.channels < 1 || frameBytes < 1 || !((fmt == 1 && (bits == 16 || bits == 24 || bits == 32)) || (fmt == 3 && bits == 32))) {
		err = "unsupported WAV (need PCM 16/24/32 or float 32)";
		return false;
	}
	const size_t frames = dataLen / frameBytes, bps = bits / 8;
	a.ch.assign(a.channels, std::vector<float>(frames));
	for (size_t i = 0; i < frames; ++i)
		for (int c = 0; c < a.channels; ++c) {
			const unsigned char *p = data + i * frameBytes + c * bps;
			float v;
			if (fmt == 3) {
				uint32_t u = rd32(p);
				memcpy(&v, &u, 4);
			} else if (bits == 16) {
				v = (int16_t)rd16(p) / 32768.0f;
			} else if (bits == 24) {
				v = (float)((int32_t)((p[0] << 8) | (p[1] << 16) | ((uint32_t)p[2] << 24)) >> 8) / 8388608.0f;
			} else {
				v = (float)((int32_t)rd32(p) / 2147483648.0);
			}
			a.ch[c][i] = v;
		}
	return true;
}

bool write_wav16(const std::string &path, const Audio &a, std::string &err) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) {
		err = "cannot create " + path;
		return false;
	}
	const uint32_t frames = (uint32_t)a.length(), dataLen = frames * a.channels * 2, rate = (uint32_t)std::lround(a.sampleRate);
	auto w32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
	auto w16 = [&](uint16_t v) { fwrite(&v, 2, 1, f); };
	fwrite("RIFF", 1, 4, f);
	w32(36 + dataLen);
	fwrite("WAVEfmt ", 1, 8, f);
	w32(16);
	w16(1);
	w16((uint16_t)a.channels);
	w32(rate);
	w32(rate * a.channels * 2);
	w16((uint16_t)(a.channels * 2));
	w16(16);
	fwrite("data", 1, 4, f);
	w32(dataLen);
	std::vector<int16_t> row((size_t)frames * a.channels);
	for (uint32_t i = 0; i < frames; ++i)
		for (int c = 0; c < a.channels; ++c) {
			float v = a.ch[c][i] * 32768.0f;
			v = std::fmax(-32768.0f, std::fmin(32767.0f, std::round(v)));
			row[(size_t)i * a.channels + c] = (int16_t)v;
		}
	fwrite(row.data(), 2, row.size(), f);
	fclose(f);
	return true;
}

bool flag(int argc, char **argv, const char *name, double *val) { // --name=value or --name (val untouched)
	const std::string key = std::string("--") + name;
	for (int i = 1; i < argc; ++i) {
		const std::string s = argv[i];
		if (s == key) return true;
		if (s.compare(0, key.size() + 1, key + "=") == 0) {
			if (val) *val = atof(s.c_str() + key.size() + 1);
			return true;
		}
	}
	return false;
}

} // namespace

int main(int argc, char **argv) {
	using Stretch = signalsmith::stretch::SignalsmithStretch<float>;
	if (argc > 1 && std::string(argv[1]) == "-v") {
		printf("%zu.%zu.%zu\n", Stretch::version[0], Stretch::version[1], Stretch::version[2]);
		return 0;
	}
	std::vector<std::string> files;
	for (int i = 1; i < argc; ++i)
		if (argv[i][0] != '-') files.push_back(argv[i]);
	if (files.size() != 2) {
		fprintf(stderr, "usage: %s input.wav output.wav [--semitones=S] [--tonality=Hz] [--time=T] [--formant=S] [--formant-comp] "
		                "[--formant-base=Hz] [--split-computation]\n", argv[0]);
		return 2;
	}
	double semitones = 0, formants = 0, formantBase = 100, tonality = 8000, time = 1; // defaults of cmd/main.cpp:21-27
	flag(argc, argv, "semitones", &semitones);
	flag(argc, argv, "formant", &formants);
	flag(argc, argv, "formant-base", &formantBase);
	flag(argc, argv, "tonality", &tonality);
	flag(argc, argv, "time", &time);
	const bool formantComp = flag(argc, argv, "formant-comp", nullptr), split = flag(argc, argv, "split-computation", nullptr);

	Audio in, out;
	std::string err;
	if (!read_wav(files[0], in, err)) {
		fprintf(stderr, "failed to read WAV: %s\n", err.c_str());
		return 1;
	}
	printf("%s -> %s\n", files[0].c_str(), files[1].c_str());
	const size_t inputLength = in.length(), outputLength = (size_t)std::round(inputLength * time);
	out.channels = in.channels;
	out.sampleRate = in.sampleRate;
	out.ch.assign(in.channels, std::vector<float>(outputLength, 0.f));

	Stretch stretch;
	stretch.presetDefault(in.channels, (float)in.sampleRate, split);
	stretch.setTransposeSemitones((float)semitones, (float)(tonality / in.sampleRate));
	stretch.setFormantSemitones((float)formants, formantComp);
	stretch.setFormantBase((float)(formantBase / in.sampleRate));

	// the stages of cmd/main.cpp:58-82: output seek, one process() call to just before the end, flush
	const int seekLength = stretch.outputSeekLength((float)(1 / time));
	if ((size_t)seekLength > in.length()) in.resize(seekLength);
	stretch.outputSeek(in, seekLength);
	int outputIndex = (int)outputLength - stretch.intervalSamples();
	if (outputIndex < 0) outputIndex = 0;
	const int outputPos = outputIndex + stretch.outputLatency();
	const int inputPos = (int)std::round(outputPos / time);
	const int inputIndex = std::max(inputPos + stretch.inputLatency(), seekLength);
	in.resize(inputIndex); // zero padding past the end of the file
	in.offset = seekLength;
	stretch.process(in, inputIndex - seekLength, out, outputIndex);
	out.offset = outputIndex;
	stretch.flush(out, (int)outputLength - outputIndex);
	out.offset = 0;
	if (!stretch.lastError().empty()) {
		fprintf(stderr, "B200 stretch failed: %s\n", stretch.lastError().c_str());
		return 1;
	}
	if (!write_wav16(files[1], out, err)) {
		fprintf(stderr, "failed to write WAV: %s\n", err.c_str());
		return 1;
	}
	return 0;
}
