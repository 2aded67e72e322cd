// Host-side mirror of the reference class `Publisher` (/root/reference/publish/publisher.h:12-66):
// same constructor arguments, isRunning(), run(), completed() notification and interrupt handlers,
// same .ini semantics (loadSettings, publisher.cpp:55-227). The SoapySDR device is replaced by an IQ
// file / synthetic source selected by the device string, and every VFO of the settings file is
// batched into ONE GPU bank (pinned host ring + async H2D inside libaeroddc.so).
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "iqsource.h"
#include "vfo.h"

class Publisher {
 public:
  Publisher(const std::string& deviceStr, bool enableBiast, bool enableDcc, const std::string& settingsPath);
  Publisher(const Publisher&) = delete;
  Publisher& operator=(const Publisher&) = delete;
  ~Publisher();

  bool isRunning() const { return running; }
  void run();                                   // starts the reader thread (QtConcurrent::run in the reference)
  void wait();                                  // joins it
  std::function<void()> completed;              // the reference's `completed()` signal

  void handleHup() {}
  void handleInterrupt() { running = false; }
  void handleTerminate() { running = false; }

  // inspection (tests, CLI): the VFO tree loadSettings built
  const std::vector<vfo*>& mainVfos() const { return VFOmain; }
  const std::vector<vfo*>& subVfos(int main_idx) const { return VFOsub[main_idx]; }
  const std::vector<vfo*>& flatVfos() const { return VFOflat; }
  int sampleRate() const { return Fs; }
  int blockLen() const { return buflen / 2; }
  bool dcc() const { return enableDcc; }
  const std::string& lastError() const { return error; }
  long long blocksProcessed() const { return blocks; }

  // only parse the settings (no device, no GPU): what the constructor does first
  static bool parseOnly(const std::string& settingsPath, Publisher** out, std::string* err);

 private:
  Publisher() {}
  bool loadSettings(const std::string& settingsPath);
  int matchMainVfo(int vfo_freq) const;   // index of the main VFO a [vfos] entry hangs under, or -1
  void readerThread();
  void demodData(void* block);
  void transmitData();

  // the reference accepts {288000, 1536000, 1920000} (publisher.h:32); 2.4 and 61.44 MS/s are the
  // BASELINE.json benchmark rates the GPU bank adds
  const std::vector<int> validSampleRates = {288000, 1536000, 1920000, 2400000, 61440000};

  std::thread mainReader;
  std::atomic<bool> running{false};
  bool enableBiast = false, enableDcc = false;
  int Fs = 0, center_frequency = 0, tuner_gain = 496, tuner_gain_idx = 0, tuner_idx = 0;
  int buflen = 0;
  int nVFO = 0;
  std::vector<vfo*> VFOsub[3];
  std::vector<vfo*> VFOmain;
  std::vector<vfo*> VFOflat;      // VFOs of a file without [main_vfos]: fed by the raw stream (see loadSettings)
  std::unique_ptr<aero::IqSource> source;
  std::shared_ptr<aero::DdcBank> bank;
  std::string error;
  long long blocks = 0;
};
