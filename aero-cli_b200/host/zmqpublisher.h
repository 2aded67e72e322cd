// ZmqPublisher: same public interface as /root/reference/publish/zmqpublisher.h:7-27. libzmq is
// loaded at run time with dlopen (no libzmq headers are needed to build): $AERODDC_LIBZMQ, else libzmq.so.5 /
// libzmq.so on the loader path. When a test sink is installed messages go to the sink instead of a socket. When
// neither a sink nor libzmq is there, connect() leaves `connected` false and available() is false: callers must
// treat that as an error (vfo::connectSockets throws) - messages are never dropped silently.
// Wire format (zmqpublisher.cpp:61-73): frame 1 = the first 5 bytes of the topic, frame 2 = uint32
// sample rate, frame 3 = payload; nothing is sent for an empty payload.
#pragma once
#include <cstdint>
#include <functional>
#include <string>

class ZmqPublisher {
 public:
  ZmqPublisher();
  void connect();
  void setAddress(const std::string& address);
  void setBind(bool b = false);
  void publish(unsigned char* buf, uint32_t len, const std::string& topic, uint32_t sampleRate);
  bool connected;

  // test / replay hook: when set, every message is handed to the sink (frame 1 already cut to 5 bytes)
  using Sink = std::function<void(const std::string& topic5, uint32_t rate, const unsigned char* payload, uint32_t len)>;
  static void setSink(Sink sink);
  static bool available();   // a sink is installed or libzmq could be loaded

 private:
  void* context;
  void* publisher;
  std::string bindAddress;
  int zmqStatus;
  bool bind;
};
