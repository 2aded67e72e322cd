/* TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
 *
 * C-callable harness around the UNMODIFIED reference class `vfo`
 * (/root/reference/publish/vfo.h:11-116, vfo.cpp:57-313). It is compiled together with the
 * reference's own publish/{vfo,oscillator,dsp,halfbanddecimator,firfilter,zmqpublisher}.cpp
 * (from where they lie; see oracle/Makefile) into oracle/_ref/libref_vfo.so.
 *
 * What it adds: (1) definitions of the six libzmq calls zmqpublisher.cpp makes, turning the PUB
 * socket into an in-memory sink that keeps frame 2 (rate) and frame 3 (payload) per topic;
 * (2) a flat C API so Python tests / bench.py can drive `vfo` exactly like
 * Publisher::loadSettings + demodData do (publisher.cpp:118-148,159-219,301-305).
 *
 * The reference leaves Fs/decimateCount/mixer_freq uninitialised in the constructor
 * (vfo.cpp:5-29), so every setter is called before init().
 */
#include "vfo.h"
#include "firfilter.h"

#include <map>
#include <string>
#include <vector>

namespace {
struct Msg {
  std::string topic; /* frame 1: first 5 bytes of the topic (zmqpublisher.cpp:69) */
  uint32_t rate;     /* frame 2 */
  std::vector<unsigned char> payload; /* frame 3 */
};
/* One sink per thread: the reference calls publish() on the thread that calls process(). */
thread_local std::vector<Msg> t_msgs;
thread_local int t_frame = 0;
thread_local Msg t_cur;
int g_sock_token;
} // namespace

extern "C" {
void *zmq_ctx_new(void) { return &g_sock_token; }
void *zmq_socket(void *, int) { return &g_sock_token; }
int zmq_setsockopt(void *, int, const void *, size_t) { return 0; }
int zmq_bind(void *, const char *) { return 0; }
int zmq_connect(void *, const char *) { return 0; }
int zmq_send(void *, const void *buf, size_t len, int flags) {
  const unsigned char *b = (const unsigned char *)buf;
  if (t_frame == 0) {
    t_cur = Msg();
    t_cur.topic.assign((const char *)b, len);
  } else if (t_frame == 1) {
    uint32_t r = 0;
    memcpy(&r, b, len < 4 ? len : 4);
    t_cur.rate = r;
  } else {
    t_cur.payload.assign(b, b + len);
  }
  if (flags & ZMQ_SNDMORE) {
    t_frame++;
  } else {
    t_msgs.push_back(t_cur);
    t_frame = 0;
  }
  return (int)len;
}
}

struct RefVfo {
  vfo *v;
  std::string topic;
  QVector<vfo *> subs; /* storage handed to setVFOs (publisher.cpp:147) */
  std::vector<RefVfo *> sub_handles;
  bool owned_by_parent;
  std::vector<cpx_typef> in;
  int D;
};

extern "C" {

/* Mirrors the setter sequence of publisher.cpp:121-147 (main VFO: demod_usb=0) and :159-217
 * (leaf VFO: demod_usb=1). topic must be exactly 5 characters (the wire keeps only 5). */
RefVfo *refvfo_create(int Fs, int decim_count, double mixer_freq, float gain, double filter_bw,
                      int demod_usb, int compress_style, int scale_comp, const char *topic5,
                      int samples_per_buffer, int late_decimate) {
  RefVfo *h = new RefVfo();
  h->v = new vfo();
  h->topic = topic5 ? topic5 : "";
  h->owned_by_parent = false;
  h->D = decim_count;
  h->v->setFs(Fs);
  h->v->setDecimationCount(decim_count);
  h->v->setMixerFreq(mixer_freq);
  h->v->setGain(gain);
  h->v->setFilterBandwidth(filter_bw);
  h->v->setDemodUSB(demod_usb != 0);
  h->v->setCompressonStyle(compress_style);
  if (scale_comp > 0)
    h->v->setScaleComp(scale_comp);
  h->v->setZmqAddress(QString("inproc://oracle"));
  h->v->setZmqTopic(QString(h->topic.c_str()));
  h->v->init(samples_per_buffer, false, late_decimate);
  return h;
}

/* Attach `sub` under `main` the way Publisher does (VFOsub[i] + setVFOs, publisher.cpp:147,219).
 * The reference's vfo destructor deletes its sub-VFOs (vfo.cpp:50-55). */
void refvfo_add_sub(RefVfo *main_h, RefVfo *sub) {
  main_h->subs.push_back(sub->v);
  main_h->sub_handles.push_back(sub);
  sub->owned_by_parent = true;
  main_h->v->setVFOs(&main_h->subs);
}

/* One vfo::process call (vfo.cpp:154-186) on n complex samples given as interleaved float I,Q,
 * converted exactly as Publisher::demodData does without DC correction (publisher.cpp:288-299).
 * Messages emitted by this VFO and its sub-VFOs land in the calling thread's sink. */
void refvfo_process(RefVfo *h, const float *iq, int n_complex) {
  h->in.resize(n_complex);
  for (int i = 0; i < n_complex; ++i)
    h->in[i] = cpx_typef(iq[2 * i], iq[2 * i + 1]);
  h->v->process(h->in);
}

/* Number of messages waiting in this thread's sink. */
int refvfo_pending(void) { return (int)t_msgs.size(); }

/* Pop the oldest message: returns payload length (or -1 if none / -2 if cap too small). */
int refvfo_pop(char *topic5_out, uint32_t *rate_out, unsigned char *payload_out, int cap) {
  if (t_msgs.empty())
    return -1;
  const Msg &m = t_msgs.front();
  if ((int)m.payload.size() > cap)
    return -2;
  if (topic5_out) {
    memset(topic5_out, 0, 6);
    memcpy(topic5_out, m.topic.data(), m.topic.size() < 5 ? m.topic.size() : 5);
  }
  if (rate_out)
    *rate_out = m.rate;
  memcpy(payload_out, m.payload.data(), m.payload.size());
  int n = (int)m.payload.size();
  t_msgs.erase(t_msgs.begin());
  return n;
}

void refvfo_clear(void) {
  t_msgs.clear();
  t_frame = 0;
}

/* Copy the public stage buffer decimate[stage] (vfo.h:39) as interleaved floats. */
int refvfo_stage(RefVfo *h, int stage, float *out, int cap_complex) {
  if (stage < 0 || stage > 8)
    return -1;
  const std::vector<cpx_typef> &d = h->v->decimate[stage];
  int n = (int)d.size() < cap_complex ? (int)d.size() : cap_complex;
  for (int i = 0; i < n; ++i) {
    out[2 * i] = d[i].real();
    out[2 * i + 1] = d[i].imag();
  }
  return (int)d.size();
}

int refvfo_out_rate(RefVfo *h) { return h->v->getOutRate(); }

void refvfo_destroy(RefVfo *h) {
  if (!h)
    return;
  if (!h->owned_by_parent) {
    delete h->v; /* deletes sub vfo objects too (vfo.cpp:50-55) */
    for (size_t i = 0; i < h->sub_handles.size(); ++i) {
      h->sub_handles[i]->v = nullptr;
      delete h->sub_handles[i];
    }
    delete h;
  }
}

/* Convenience for the CPU baseline timer: run `n_blocks` process() calls of the same block,
 * discarding messages; returns nothing. Keeps the timed loop free of Python overhead. */
void refvfo_process_repeat(RefVfo *h, const float *iq, int n_complex, int n_blocks) {
  h->in.resize(n_complex);
  for (int i = 0; i < n_complex; ++i)
    h->in[i] = cpx_typef(iq[2 * i], iq[2 * i + 1]);
  for (int b = 0; b < n_blocks; ++b) {
    h->v->process(h->in);
    t_msgs.clear();
  }
}

/* Component probes used to pin each piece of the restated oracle separately. */

/* firfilter::low_pass with the Hamming window, as vfo::init calls it (vfo.cpp:71-79,92-102). */
int ref_lowpass(double gain, double fs, double fc, double tw, float *out, int cap) {
  firfilter f;
  QVector<float> t = f.low_pass(gain, fs, fc, tw, firfilter::WIN_HAMMING, 0);
  int n = t.length() < cap ? t.length() : cap;
  for (int i = 0; i < n; ++i)
    out[i] = t[i];
  return t.length();
}

/* FIRHilbert coefficient table (dsp.cpp:181-215). */
void ref_hilbert(int len, int fs, float *out) {
  FIRHilbert h(len, fs);
  for (int i = 0; i < len; ++i)
    out[i] = h.points[i];
}

/* Oscillator values as vfo::process consumes them (vfo.cpp:155-161, oscillator.cpp:4-39):
 * out[k] = value of _vector used for sample first+k. */
void ref_nco(double fs, double f, long long first, int count, float *out) {
  Oscillator o(fs, f);
  for (long long i = 0; i < first; ++i)
    o.tick();
  for (int k = 0; k < count; ++k) {
    out[2 * k] = o._vector.real();
    out[2 * k + 1] = o._vector.imag();
    o.tick();
  }
}
}
