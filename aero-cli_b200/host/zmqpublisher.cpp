#include "zmqpublisher.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {
// the libzmq entry points and option numbers zmqpublisher.cpp:14-51,69-71 uses (values from zmq.h 4.x)
struct ZmqApi {
  void* (*ctx_new)();
  void* (*socket)(void*, int);
  int (*setsockopt)(void*, int, const void*, size_t);
  int (*bind)(void*, const char*);
  int (*connect)(void*, const char*);
  int (*send)(void*, const void*, size_t, int);
  bool ok = false;
};
constexpr int kZMQ_PUB = 1, kZMQ_SNDMORE = 2;
constexpr int kZMQ_RECONNECT_IVL = 18, kZMQ_RECONNECT_IVL_MAX = 21;
constexpr int kZMQ_TCP_KEEPALIVE = 34, kZMQ_TCP_KEEPALIVE_CNT = 35, kZMQ_TCP_KEEPALIVE_IDLE = 36, kZMQ_TCP_KEEPALIVE_INTVL = 37;

ZmqApi& api() {
  static ZmqApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* env = getenv("AERODDC_LIBZMQ")) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (const char* name : {"libzmq.so.5", "libzmq.so"})
      if (!h) h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;   // no search beyond the loader path: AERODDC_LIBZMQ names a library living elsewhere (e.g. pyzmq's bundled copy)
    a.ctx_new = (void* (*)())dlsym(h, "zmq_ctx_new");
    a.socket = (void* (*)(void*, int))dlsym(h, "zmq_socket");
    a.setsockopt = (int (*)(void*, int, const void*, size_t))dlsym(h, "zmq_setsockopt");
    a.bind = (int (*)(void*, const char*))dlsym(h, "zmq_bind");
    a.connect = (int (*)(void*, const char*))dlsym(h, "zmq_connect");
    a.send = (int (*)(void*, const void*, size_t, int))dlsym(h, "zmq_send");
    a.ok = a.ctx_new && a.socket && a.setsockopt && a.bind && a.connect && a.send;
  });
  return a;
}
ZmqPublisher::Sink g_sink;
}  // namespace

void ZmqPublisher::setSink(Sink sink) { g_sink = std::move(sink); }

ZmqPublisher::ZmqPublisher() : connected(false), context(nullptr), publisher(nullptr), bindAddress("tcp://*:6002"), zmqStatus(0), bind(false) {}

bool ZmqPublisher::available() { return g_sink || api().ok; }

void ZmqPublisher::connect() {
  if (connected) return;
  if (g_sink) {   // test / replay hook installed: no socket, every message goes to the sink
    connected = true;
    return;
  }
  if (!api().ok) {
    // The reference links libzmq at build time, so this cannot happen there. Here it must not pass silently: the publisher
    // stays unconnected, the caller (vfo / Publisher) turns that into an error, and nothing is ever "published" into the void.
    static bool said = false;
    if (!said) fprintf(stderr, "[CRIT] libzmq could not be loaded (tried $AERODDC_LIBZMQ, libzmq.so.5, libzmq.so): no ZeroMQ output possible\n");
    said = true;
    return;
  }
  ZmqApi& z = api();
  context = z.ctx_new();
  publisher = z.socket(context, kZMQ_PUB);
  const int keepalive = 1, keepalivecnt = 10, keepaliveidle = 1, keepaliveintrv = 1, reconnectInterval = 1000, maxReconnectInterval = 0;
  z.setsockopt(publisher, kZMQ_TCP_KEEPALIVE, &keepalive, sizeof(int));
  z.setsockopt(publisher, kZMQ_TCP_KEEPALIVE_CNT, &keepalivecnt, sizeof(int));
  z.setsockopt(publisher, kZMQ_TCP_KEEPALIVE_IDLE, &keepaliveidle, sizeof(int));
  z.setsockopt(publisher, kZMQ_TCP_KEEPALIVE_INTVL, &keepaliveintrv, sizeof(int));
  z.setsockopt(publisher, kZMQ_RECONNECT_IVL, &reconnectInterval, sizeof(int));
  z.setsockopt(publisher, kZMQ_RECONNECT_IVL_MAX, &maxReconnectInterval, sizeof(int));
  if (bind) {
    zmqStatus = z.bind(publisher, bindAddress.c_str());
    if (zmqStatus < 0) return;   // the reference returns here too, leaving connected == false (zmqpublisher.cpp:44-48)
  } else {
    zmqStatus = z.connect(publisher, bindAddress.c_str());
  }
  connected = true;
}

void ZmqPublisher::setAddress(const std::string& address) { bindAddress = address; }
void ZmqPublisher::setBind(bool b) { bind = b; }

void ZmqPublisher::publish(unsigned char* buf, uint32_t len, const std::string& topic, uint32_t sampleRate) {
  if (len == 0) return;
  // frame 1 is exactly 5 bytes of the topic's C string (zmqpublisher.cpp:69); shorter topics are padded with NULs here
  char t5[5] = {0, 0, 0, 0, 0};
  memcpy(t5, topic.data(), topic.size() < 5 ? topic.size() : 5);
  if (g_sink) {
    g_sink(std::string(t5, 5), sampleRate, buf, len);
    return;
  }
  if (!publisher) return;
  unsigned char rate[4];
  memcpy(rate, &sampleRate, 4);
  ZmqApi& z = api();
  z.send(publisher, t5, 5, kZMQ_SNDMORE);
  z.send(publisher, rate, 4, kZMQ_SNDMORE);
  z.send(publisher, buf, len, 0);
}
