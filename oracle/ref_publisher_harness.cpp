// TEST INFRASTRUCTURE (oracle/) - never linked into the product.
//
// oracle/_ref/ref_publish: the reference's UNMODIFIED aero-publish core - publish/publisher.cpp (settings-file
// semantics of Publisher::loadSettings, the reader loop, demodData with its optional DC correction) on top of
// publish/{vfo,oscillator,dsp,halfbanddecimator,firfilter,zmqpublisher}.cpp - compiled from /root/reference with
// stand-ins for what it links against:
//   Qt        -> oracle/shim_publisher/qt_publisher_shim.h (QSettings IniFormat reader, QFileInfo, QtConcurrent::run
//                executing inline, ...) on top of oracle/shim_decode/qt_decode_shim.h
//   SoapySDR  -> an IQ-file device (below): CF32 out, exactly bufflen/2 samples per readStream
//   libzmq    -> an in-memory sink (below) that keeps frame 3 of every message per topic
//   moc       -> the body of the one signal, Publisher::completed
// It pins what the vfo-level oracle cannot: Publisher::loadSettings' arithmetic and matching rules on the real
// code, the main -> sub VFO tree exactly as the reference builds it, and the DC-correction recurrence.
//
// A second build of this file, -DAERODDC_GPU_VFO (oracle/_ref/ref_publish_gpuvfo), puts the PRODUCT's `vfo` class
// (aero-cli_b200/host/vfo.{h,cpp}, the CUDA bank behind it) under the same unmodified publisher.cpp instead of the
// reference's vfo.cpp: the reference's Publisher then drives the GPU through nothing but the class surface SURVEY.md
// section 8b lists - the drop-in claim at link level. Payloads are captured through ZmqPublisher::setSink.
//
//   ref_publish <settings.ini> <iq file> <cu8|cs16|cf32> <dcc 0|1> <out dir>
// writes <out dir>/<TOPIC>.i16 + <TOPIC>.meta ("rate bytes_per_message"), the layout of `aero-publish-b200 --dump`,
// and prints one JSON line with the VFO tree the reference built.
#include "qt_publisher_shim.h"

#include <unistd.h>

#include <SoapySDR/Device.hpp>

#ifndef AERODDC_GPU_VFO
#include "zmq.h"
#endif

#define private public   // the harness reads the VFO tree the constructor built and calls the reader loop's owner
#include "publisher.h"
#undef private

bool gMaxLogVerbosity = false;   // common/logger.h

// ---- moc's part ----------------------------------------------------------------------------------------------------
void Publisher::completed() {}

// ---- libzmq: in-memory sink ------------------------------------------------------------------------------------------
namespace {
struct Topic {
  std::string payload;
  uint32_t rate = 0, msg_bytes = 0;
  uint64_t msgs = 0;
};
std::map<std::string, Topic> g_topics;
std::vector<std::string> g_order;
int g_frame = 0;
std::string g_cur_topic;
uint32_t g_cur_rate = 0;
int g_token;
}  // namespace

static void keep_message(const std::string& topic, uint32_t rate, const void* buf, size_t len) {
  std::string key(topic.c_str());   // a short topic is padded with NULs on the wire (zmqpublisher.cpp:69)
  if (!g_topics.count(key)) g_order.push_back(key);
  Topic& t = g_topics[key];
  t.payload.append((const char*)buf, len);
  t.rate = rate;
  t.msg_bytes = (uint32_t)len;
  t.msgs++;
}

#ifndef AERODDC_GPU_VFO
extern "C" {
void* zmq_ctx_new(void) { return &g_token; }
void* zmq_socket(void*, int) { return &g_token; }
int zmq_setsockopt(void*, int, const void*, size_t) { return 0; }
int zmq_bind(void*, const char*) { return 0; }
int zmq_connect(void*, const char*) { return 0; }
int zmq_send(void*, const void* buf, size_t len, int flags) {
  if (g_frame == 0) g_cur_topic.assign((const char*)buf, len);
  else if (g_frame == 1) { g_cur_rate = 0; std::memcpy(&g_cur_rate, buf, len < 4 ? len : 4); }
  else keep_message(g_cur_topic, g_cur_rate, buf, len);
  g_frame = (flags & ZMQ_SNDMORE) ? g_frame + 1 : 0;
  return (int)len;
}
}
#endif

// ---- SoapySDR: IQ-file device ----------------------------------------------------------------------------------------
namespace {
struct FileDev {
  FILE* f = nullptr;
  int fmt = 2;           // 0 cu8, 1 cs16, 2 cf32
  size_t per_read = 0;   // complex samples per readStream
  std::vector<unsigned char> raw;
};
}  // namespace

namespace SoapySDR {
Device* Device::make(const std::string& args) {
  std::map<std::string, std::string> kv;
  std::istringstream in(args);
  std::string item;
  while (std::getline(in, item, ',')) {
    const size_t eq = item.find('=');
    if (eq != std::string::npos) kv[item.substr(0, eq)] = item.substr(eq + 1);
  }
  if (!kv.count("file")) return nullptr;
  FileDev* d = new FileDev;
  d->f = fopen(kv["file"].c_str(), "rb");
  if (!d->f) { delete d; return nullptr; }
  const std::string fm = kv.count("format") ? kv["format"] : "cf32";
  d->fmt = fm == "cu8" ? 0 : (fm == "cs16" ? 1 : 2);
  Device* dev = new Device;
  dev->impl = d;
  return dev;
}
void Device::unmake(Device* dev) {
  if (!dev) return;
  FileDev* d = (FileDev*)dev->impl;
  if (d->f) fclose(d->f);
  delete d;
  delete dev;
}
Stream* Device::setupStream(int, const std::string& format, const std::vector<size_t>&, const Kwargs& args) {
  FileDev* d = (FileDev*)impl;
  if (format != "CF32") return nullptr;
  auto it = args.find("bufflen");          // bytes of cu8 IQ per driver buffer (publisher.cpp:241-242) = 2 bytes per sample
  d->per_read = it == args.end() ? 0 : (size_t)atol(it->second.c_str()) / 2;
  return d->per_read ? (Stream*)d : nullptr;
}
int Device::readStream(Stream*, void* const* buffs, size_t numElems, int&, long long&, long) {
  FileDev* d = (FileDev*)impl;
  const size_t n = d->per_read < numElems ? d->per_read : numElems;
  const size_t bps = d->fmt == 0 ? 2 : (d->fmt == 1 ? 4 : 8);
  d->raw.resize(n * bps);
  if (fread(d->raw.data(), 1, n * bps, d->f) != n * bps) return 0;   // end of file: a partial block is dropped
  float* out = (float*)buffs[0];
  // the product's conversions (aero-cli_b200/csrc/ddc_kernels.cuh load_raw, tests/oracle_bind.unpack)
  if (d->fmt == 0) for (size_t i = 0; i < 2 * n; ++i) out[i] = ((float)d->raw[i] - 127.4f) / 128.0f;
  else if (d->fmt == 1) for (size_t i = 0; i < 2 * n; ++i) out[i] = (float)((const int16_t*)d->raw.data())[i] / 32768.0f;
  else std::memcpy(out, d->raw.data(), n * 8);
  return (int)n;
}
}  // namespace SoapySDR

// ---- driver ----------------------------------------------------------------------------------------------------------
static void print_vfo(vfo* v, const char* kind, int parent, bool& first) {
  printf("%s{\"kind\": \"%s\", \"parent\": %d, \"mixer\": %.3f, \"out_rate\": %d, \"usb\": %d}", first ? "" : ", ", kind, parent, v->getMixerFreq(),
         v->getOutRate(), v->getDemodUSB() ? 1 : 0);
  first = false;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: ref_publish <settings.ini> <iq file> <cu8|cs16|cf32> <dcc 0|1> <out dir>\n");
    return 2;
  }
  const std::string ini = argv[1], iq = argv[2], fmt = argv[3], out = argv[5];
  const bool dcc = atoi(argv[4]) != 0;
#ifdef AERODDC_GPU_VFO
  ZmqPublisher::setSink([](const std::string& topic5, uint32_t rate, const unsigned char* p, uint32_t n) { keep_message(topic5, rate, p, n); });
#endif
  Publisher* pub = new Publisher(QString(("file=" + iq + ",format=" + fmt).c_str()), false, dcc, QString(ini.c_str()));
  if (!pub->isRunning()) {
    printf("{\"error\": \"reference Publisher did not start (settings or source rejected)\"}\n");
    fflush(stdout);
    _exit(1);   // the reference's destructor reads members its constructor never set on this path
  }
  pub->run();   // QtConcurrent::run executes the reader loop inline until the file ends
  printf("{\"sample_rate\": %d, \"buflen\": %d, \"dcc\": %d, \"vfos\": [", pub->Fs, pub->buflen, pub->enableDcc ? 1 : 0);
  bool first = true;
  for (int a = 0; a < pub->VFOmain.length(); a++) {
    print_vfo(pub->VFOmain.at(a), "main", -1, first);
    for (int k = 0; k < pub->VFOsub[a].length(); k++) print_vfo(pub->VFOsub[a].at(k), "sub", a, first);
  }
  printf("], \"topics\": {");
  first = true;
  for (const std::string& name : g_order) {
    const Topic& t = g_topics[name];
    FILE* f = fopen((out + "/" + name + ".i16").c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write to %s\n", out.c_str()); return 2; }
    fwrite(t.payload.data(), 1, t.payload.size(), f);
    fclose(f);
    if (FILE* m = fopen((out + "/" + name + ".meta").c_str(), "w")) { fprintf(m, "%u %u\n", t.rate, t.msg_bytes); fclose(m); }
    printf("%s\"%s\": {\"rate\": %u, \"messages\": %llu, \"bytes\": %zu}", first ? "" : ", ", name.c_str(), t.rate, (unsigned long long)t.msgs, t.payload.size());
    first = false;
  }
  printf("}}\n");
  fflush(stdout);
  _exit(0);   // skip static destructors: the reference's static PUB socket object outlives main
}
