"""LiveBatch -- the reference's live / streaming caller (web/web-wrapper.js:215-332, its AudioWorklet processor in
buffer-playback mode) as a batched server loop over the B200 C ABI.

One `LiveBatch` drives S independent streams of one `BatchStretch` engine.  Every stream has what one worklet has:
  * audio buffers appended with `add_buffers` (web-wrapper.js:163-169) and dropped with `drop_buffers` (:137-162);
  * a TIME MAP: a list of segments {output, input, rate, active, loopStart, loopEnd} edited with `schedule` / `start` /
    `stop` (:45-108, including `adjustPrevious`), looked up per audio quantum (:227-231);
  * per quantum, the segment's input position at the quantum's output time (+ latencies, loop wrap :271-279), the
    `bufferLength = inputLatency + outputLatency` samples that END there gathered from the audio buffers with zero padding
    on both sides (:282-311), and then -- exactly like the worklet, which "constantly seeks" (:313) -- `seek(window, rate)`
    followed by `process(0, quantum)` (:314-315).
The batch makes ONE `b200s_seek_rates` + ONE `b200s_process(0, n)` call per quantum for all streams (each stream its own
window and its own rate).  When no stream is active the quantum is `process(zeros, n)` for the batch, as :247-253.
Difference from S separate worklets, by construction of a batch call: in a quantum where SOME streams are active, a stopped
stream is fed a silent seek window instead of `n` zero samples appended to its history (its tail is then the synthesis
ring's, up to one block shorter).  Parameters (transpose / formants) are batch-wide, as in the C ABI.
"""
import numpy as np

__all__ = ["LiveBatch"]


class _Stream:
    def __init__(self, channels):
        self.time_map = [dict(active=False, input=0.0, output=0.0, rate=1.0, loopStart=0.0, loopEnd=0.0)]  # :17-29
        self.audio = np.zeros((channels, 0), np.float32)
        self.audio_start = 0  # sample index of audio[:, 0] (audioBuffersStart)


class LiveBatch:
    def __init__(self, engine, sample_rate):
        self.eng, self.sr = engine, float(sample_rate)
        self.S, self.C = engine.batch, engine.channels()
        self.buffer_length = engine.inputLatency() + engine.outputLatency()  # :205-207
        self.in_lat_s = engine.inputLatency() / self.sr
        self.out_lat_s = engine.outputLatency() / self.sr
        self.streams = [_Stream(self.C) for _ in range(self.S)]
        self.current_sample = 0  # currentTime * sampleRate of the audio context
        self._win = np.zeros((self.S, self.C, self.buffer_length), np.float32)

    @property
    def current_time(self):
        return self.current_sample / self.sr

    # ---- remote methods of the worklet, per stream ----
    def add_buffers(self, s, samples):  # :163-169
        st = self.streams[s]
        x = np.asarray(samples, np.float32).reshape(self.C, -1)
        st.audio = np.concatenate([st.audio, x], axis=1)
        return (st.audio_start + st.audio.shape[1]) / self.sr

    def drop_buffers(self, s, to_seconds=None):  # :137-162 (sample-accurate instead of whole buffers)
        st = self.streams[s]
        if to_seconds is None:
            st.audio, st.audio_start = np.zeros((self.C, 0), np.float32), 0
            return 0.0, 0.0
        n = max(0, min(st.audio.shape[1], int(to_seconds * self.sr) - st.audio_start))
        st.audio, st.audio_start = st.audio[:, n:], st.audio_start + n
        return st.audio_start / self.sr, (st.audio_start + st.audio.shape[1]) / self.sr

    def schedule(self, s, obj_in, adjust_previous=False):  # :67-108
        tm = self.streams[s].time_map
        output_time = obj_in.get("outputTime", self.current_time)
        latest = tm[-1]
        while tm and tm[-1]["output"] >= output_time:
            latest = tm.pop()
        obj = dict(latest)
        obj.update(input=None, output=output_time)
        obj.update({k: v for k, v in obj_in.items() if k != "outputTime"})
        if obj["input"] is None:
            rate = latest["rate"] if latest["active"] else 0.0
            obj["input"] = latest["input"] + (obj["output"] - latest["output"]) * rate
        tm.append(obj)
        if adjust_previous and len(tm) > 1:
            prev = tm[-2]
            if prev["output"] < self.current_time:
                rate = prev["rate"] if prev["active"] else 0.0
                prev["input"] += (self.current_time - prev["output"]) * rate
                prev["output"] = self.current_time
            if obj["output"] != prev["output"]:  # (JavaScript would store Infinity / NaN here; a zero-length segment is never looked up)
                prev["rate"] = (obj["input"] - prev["input"]) / (obj["output"] - prev["output"])
        while len(tm) > 1 and tm[1]["output"] <= output_time:
            tm.pop(0)
        return obj

    def start(self, s, when=None, offset=0.0, duration=None, rate=1.0):  # :49-66
        obj = dict(active=True, input=offset, output=self.current_time + self.out_lat_s if when is None else when, rate=rate)
        res = self.schedule(s, obj)
        if duration is not None:
            self.stop(s, obj["output"] + duration)
        return res

    def stop(self, s, when=None):  # :45-48
        return self.schedule(s, dict(active=False, output=self.current_time if when is None else when))

    # ---- one audio quantum for every stream (:215-332) ----
    def _window(self, st, seg, output_time):
        """The seek window of one stream (:268-311): bufferLength samples ending at the segment's input position."""
        input_time = seg["input"] + (output_time - seg["output"]) * seg["rate"]
        loop = seg["loopEnd"] - seg["loopStart"]
        if loop > 0 and input_time >= seg["loopEnd"]:
            seg["input"] -= loop
            input_time -= loop
        input_time += self.in_lat_s
        end = int(np.floor(input_time * self.sr + 0.5))  # Math.round
        lo, hi = end - self.buffer_length, end
        a0, a1 = st.audio_start, st.audio_start + st.audio.shape[1]
        w = np.zeros((self.C, self.buffer_length), np.float32)
        c0, c1 = max(lo, a0), min(hi, a1)
        if c1 > c0:
            w[:, c0 - lo:c1 - lo] = st.audio[:, c0 - a0:c1 - a0]
        return w

    def process(self, n_out=128):
        """Render the next quantum: returns [S][C][n_out]."""
        output_time = self.current_time + self.out_lat_s  # :226
        segs = []
        for st in self.streams:
            tm = st.time_map
            while len(tm) > 1 and tm[1]["output"] <= output_time:
                tm.pop(0)
            segs.append(tm[0])
        if not any(seg["active"] for seg in segs):  # :247-253 for the whole batch
            y = self.eng.process(np.zeros((self.S, self.C, n_out), np.float32), n_out)
        else:
            rates = np.ones(self.S)
            for s, (st, seg) in enumerate(zip(self.streams, segs)):
                if seg["active"]:
                    self._win[s] = self._window(st, seg, output_time)
                    rates[s] = seg["rate"]
                else:
                    self._win[s] = 0
            self.eng.seek(self._win, rates)                                        # :314
            y = self.eng.process(np.zeros((self.S, self.C, 0), np.float32), n_out)  # :315
        self.current_sample += n_out
        return np.array(y)
