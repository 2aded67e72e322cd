// aero-publish-b200: command-line shell around Publisher, the counterpart of
// /root/reference/publish/main.cpp:11-64 (same -d / --enable-biast / --enable-dcc / <settings> arguments).
// Extras for replay and testing:
//   --plan            parse the settings only (no device, no GPU) and print the VFO tree as JSON
//   --hash            capture every ZeroMQ message in-process and print per-topic FNV-1a-64 / bytes as JSON
//   --dump DIR        additionally write every topic's payloads to DIR/<topic>.i16 (+ .meta) for tools/replay_payloads.py
//   --anchors         drive stand-alone `vfo` objects over the SURVEY.md section-8c anchor inputs and print the hashes
#include <cinttypes>
#include <csignal>
#include <ctime>
#include <vector>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "publisher.h"

namespace {
Publisher* g_pub = nullptr;
void on_signal(int) { if (g_pub) g_pub->handleInterrupt(); }

struct TopicAcc { uint64_t fnv = 1469598103934665603ull; uint64_t bytes = 0; uint32_t rate = 0; uint64_t msgs = 0; };
std::map<std::string, TopicAcc> g_acc;
std::string g_dump_dir;
std::map<std::string, FILE*> g_dump;
void sink(const std::string& topic5, uint32_t rate, const unsigned char* p, uint32_t n) {
  TopicAcc& a = g_acc[std::string(topic5.c_str())];
  if (!g_dump_dir.empty()) {   // DIR/<topic>.i16 + DIR/<topic>.meta ("rate bytes_per_message"), read by tools/replay_payloads.py
    const std::string t(topic5.c_str());
    FILE*& f = g_dump[t];
    if (!f) {
      f = fopen((g_dump_dir + "/" + t + ".i16").c_str(), "wb");
      if (FILE* m = fopen((g_dump_dir + "/" + t + ".meta").c_str(), "w")) { fprintf(m, "%u %u\n", rate, n); fclose(m); }
    }
    if (f) fwrite(p, 1, n, f);
  }
  for (uint32_t i = 0; i < n; ++i) a.fnv = (a.fnv ^ p[i]) * 1099511628211ull;
  a.bytes += n; a.rate = rate; a.msgs++;
}
void print_acc() {
  printf("{");
  bool first = true;
  for (auto& kv : g_acc) {
    printf("%s\"%s\": {\"fnv1a64\": \"%016" PRIx64 "\", \"bytes\": %" PRIu64 ", \"rate\": %u, \"messages\": %" PRIu64 "}", first ? "" : ", ",
           kv.first.c_str(), kv.second.fnv, kv.second.bytes, kv.second.rate, kv.second.msgs);
    first = false;
  }
  printf("}\n");
}

void print_vfo(vfo* v, const char* kind, int parent, bool& first) {
  printf("%s{\"kind\": \"%s\", \"parent\": %d, \"topic\": \"%s\", \"fs\": %d, \"decim\": %d, \"late\": %d, \"mixer\": %.3f, \"gain\": %.9g, \"filter_bw\": %d, \"block\": %d, \"out_rate\": %d, \"usb\": %d}",
         first ? "" : ", ", kind, parent, v->topic().c_str(), v->fs(), v->decimationCount(), v->lateDecimate(), v->getMixerFreq(), (double)v->gainValue(),
         v->filterBandwidth(), v->samplesPerBuffer(), v->getOutRate(), v->getDemodUSB() ? 1 : 0);
  first = false;
}

int run_anchors() {
  struct A { int Fs, B, D, L; double f; float g; int bw, blocks; };
  const A cases[] = {{1536000, 384000, 5, 0, 123456.0, 0.5f, 0, 6}, {288000, 57600, 1, 6, -34567.0, 0.5f, 0, 8},
                     {288000, 57600, 0, 6, 20000.0, 0.25f, 3000, 8}, {1920000, 480000, 3, 5, -250000.0, 0.5f, 0, 6}};
  ZmqPublisher::setSink(sink);
  int k = 0;
  for (const A& c : cases) {
    vfo v;   // exactly how the oracle harness drives the reference's vfo
    char topic[8];
    snprintf(topic, sizeof topic, "ANC%02d", k++);
    v.setFs(c.Fs); v.setDecimationCount(c.D); v.setMixerFreq(c.f); v.setGain(c.g); v.setFilterBandwidth(c.bw);
    v.setDemodUSB(true); v.setCompressonStyle(1); v.setZmqAddress("inproc://selftest"); v.setZmqTopic(topic);
    v.init(c.B, false, c.L);
    std::vector<cpx_typef> x(c.B);
    for (int b = 0; b < c.blocks; ++b) {
      for (int i = 0; i < c.B; ++i) {
        const uint64_t n = (uint64_t)b * c.B + i;
        x[i] = cpx_typef((float)((int)((n * 7919ull + 13ull) % 2001ull) - 1000) / 1000.0f, (float)((int)((n * 104729ull + 7ull) % 2001ull) - 1000) / 1000.0f);
      }
      v.process(x);
    }
  }
  print_acc();
  return 0;
}
}  // namespace

// publish a few known messages through the real socket path (libzmq via dlopen): wire-format self test, no GPU
int run_zmq_selftest(const std::string& addr) {
  ZmqPublisher pub;
  pub.setAddress(addr);
  pub.setBind(true);
  pub.connect();
  if (!pub.connected) { fprintf(stderr, "bind failed or libzmq not found\n"); return 1; }
  std::vector<int16_t> payload(600);
  for (int round = 0; round < 200; ++round) {   // PUB/SUB drops messages until the subscriber has joined: keep sending
    for (int i = 0; i < 600; ++i) payload[i] = (int16_t)(i * 7 - 1000 + round);
    pub.publish((unsigned char*)payload.data(), (uint32_t)(payload.size() * 2), "VFO42-long-topic", 48000);
    pub.publish((unsigned char*)payload.data(), 0, "EMPTY", 12000);   // len 0: nothing is sent (zmqpublisher.cpp:67)
    struct timespec ts = {0, 10 * 1000 * 1000};
    nanosleep(&ts, nullptr);
  }
  return 0;
}

int main(int argc, char** argv) {
  std::string device, ini, zmq_addr;
  bool biast = false, dcc = false, plan = false, hash = false, anchors = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "-d" || a == "--device") && i + 1 < argc) device = argv[++i];
    else if (a == "--enable-biast") biast = true;
    else if (a == "--enable-dcc") dcc = true;
    else if (a == "--plan") plan = true;
    else if (a == "--hash") hash = true;
    else if (a == "--dump" && i + 1 < argc) { g_dump_dir = argv[++i]; hash = true; }
    else if (a == "--anchors") anchors = true;
    else if (a == "--zmq-selftest" && i + 1 < argc) zmq_addr = argv[++i];
    else if (a == "-v" || a == "--verbose") {}
    else if (a == "-h" || a == "--help") {
      printf("usage: aero-publish-b200 -d <file=path,format=cu8|cs16|cf32[,repeat=N] | synthetic=seed[,format=..][,blocks=N]>[,throttle=X][,delay=S][,gpus=N][,mode=fast|tensor] [--enable-dcc] [--hash|--dump DIR] <settings.ini>\n"
             "       aero-publish-b200 --plan <settings.ini>\n");
      return 0;
    } else ini = a;
  }
  if (!zmq_addr.empty()) return run_zmq_selftest(zmq_addr);
  if (anchors) return run_anchors();
  if (ini.empty()) { fprintf(stderr, "settings file required\n"); return 2; }
  if (plan) {
    Publisher* p = nullptr;
    std::string err;
    if (!Publisher::parseOnly(ini, &p, &err)) { printf("{\"error\": \"%s\"}\n", err.c_str()); return 1; }
    printf("{\"sample_rate\": %d, \"block\": %d, \"dcc\": %d, \"vfos\": [", p->sampleRate(), p->blockLen(), p->dcc() ? 1 : 0);
    bool first = true;
    int mi = 0;
    for (vfo* m : p->mainVfos()) {
      print_vfo(m, "main", -1, first);
      for (vfo* s : p->subVfos(mi)) print_vfo(s, "sub", mi, first);
      ++mi;
    }
    for (vfo* f : p->flatVfos()) print_vfo(f, "flat", -1, first);
    printf("]}\n");
    delete p;
    return 0;
  }
  if (device.empty()) { fprintf(stderr, "-d <source> required\n"); return 2; }
  if (hash) ZmqPublisher::setSink(sink);
  Publisher pub(device, biast, dcc, ini);
  if (!pub.isRunning()) return 1;   // main.cpp:52-55
  g_pub = &pub;
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  pub.run();
  pub.wait();
  fprintf(stderr, "processed %lld blocks of %d samples\n", pub.blocksProcessed(), pub.blockLen());
  for (auto& kv : g_dump) if (kv.second) fclose(kv.second);
  if (hash) print_acc();
  return pub.lastError().empty() ? 0 : 1;
}
