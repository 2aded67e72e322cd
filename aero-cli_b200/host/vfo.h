// Host-side mirror of the reference class `vfo` (/root/reference/publish/vfo.h:11-116) on top of the
// C ABI of include/aeroddc.h. Same method names and argument meaning, so code (and tests) written
// against the reference's class read the same here; the arithmetic runs in libaeroddc.so on the GPU.
//
// Differences that follow from batching every VFO into one GPU bank:
//  * init() only records the configuration (and opens no socket: connectSockets() does, when the VFO is about to run,
//    so that parsing a settings file with --plan binds nothing). The bank is built lazily: by Publisher for all VFOs of a
//    settings file (Publisher::start), or - for a vfo driven on its own like the reference's object -
//    at its first process() call, together with the sub-VFOs attached by setVFOs().
//  * process() on a main VFO processes its sub-VFOs in the same GPU pass (the reference recurses,
//    vfo.cpp:167-172) and then publishes every leaf's payload through ZmqPublisher like
//    vfo::transmitData (vfo.cpp:289-313).
//  * Qt types are replaced by the standard ones (QString -> std::string, QVector -> std::vector).
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/aeroddc.h"
#include "zmqpublisher.h"

typedef std::complex<float> cpx_typef;

namespace aero {
// RAII holder of one aeroddc_fleet (one GPU bank per device; a single device is the common case); shared by
// the VFOs that were batched into it.
class DdcBank {
 public:
  DdcBank(int sample_rate, int block_len, int in_format, const std::vector<int>& devices = {0});
  ~DdcBank();
  DdcBank(const DdcBank&) = delete;
  DdcBank& operator=(const DdcBank&) = delete;
  aeroddc_fleet* handle() const { return bank_; }
  int blockLen() const { return block_len_; }
  int format() const { return format_; }
  bool finalized() const { return finalized_; }
  void finalize();                                  // throws std::runtime_error with aeroddc_last_error()
  void process(const void* host_iq, size_t n_complex);   // submit + wait
  void submit(const void* host_iq, size_t n_complex);    // async H2D + kernels + payload D2H; at most two blocks in flight
  void wait();                                           // oldest block in flight; its payloads stay valid until the next wait()
 private:
  aeroddc_fleet* bank_ = nullptr;
  int block_len_, format_;
  bool finalized_ = false;
};
}  // namespace aero

class vfo {
 public:
  ~vfo();
  vfo();

  void init(int samplesPerBuffer, bool bind, int lateDecimate = 0);
  void process(const std::vector<cpx_typef>& samples);
  void setZmqAddress(const std::string& bind);
  void setZmqTopic(const std::string& topic);
  void setScaleComp(int scale);
  void setFs(int samplerate);
  void setDecimationCount(int count);
  void setMixerFreq(double freq);
  double getMixerFreq();
  int getOutRate();
  void setOffsetBandwidth(double bw);
  void setFilterBandwidth(double bw);
  void setGain(float g);
  void setDemodUSB(bool usb);
  bool getDemodUSB();
  void setCompressonStyle(int st);
  void setVFOs(std::vector<vfo*>* pVFOs);
  std::vector<vfo*>* mpVFOs;

  // --- batching interface used by Publisher (no counterpart in the reference) ---
  // register this VFO (and, recursively, its sub-VFOs) in `bank`; parent = bank index of the main VFO or -1
  void addToBank(const std::shared_ptr<aero::DdcBank>& bank, int parent);
  // hand the last completed block's payload to ZmqPublisher (vfo::transmitData); recurses into sub-VFOs
  void transmitData();
  // open the ZeroMQ socket(s) this VFO and its sub-VFOs publish on (the socket half of vfo::init, vfo.cpp:128-136);
  // throws std::runtime_error when neither libzmq nor a sink is available - messages are never dropped silently
  void connectSockets();
  int bankIndex() const { return index_; }
  int samplesPerBuffer() const { return samplesPerBuffer_; }
  int lateDecimate() const { return lateDecimate_; }
  int decimationCount() const { return decimateCount; }
  int fs() const { return Fs; }
  int filterBandwidth() const { return filterbw; }
  float gainValue() const { return gain; }
  const std::string& topic() const { return zmqTopic; }

 private:
  std::string zmqAddress, zmqTopic;
  int Fs;
  bool zmqBind;
  static ZmqPublisher bind_publisher;   // shared PUB socket when binding (vfo.cpp:4,128-131)
  ZmqPublisher connect_publisher;
  int decimateCount;
  uint32_t outputRate;
  float gain;
  double mixer_freq;
  bool demodUSB;
  int cstyle;
  int filterbw, offsetbw;
  int scalecomp;
  int samplesPerBuffer_, lateDecimate_;
  bool inited_, socketsReady_;
  std::shared_ptr<aero::DdcBank> bank_;
  int index_;
};
