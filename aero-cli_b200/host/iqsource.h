// IQ block sources that replace the SoapySDR device of Publisher::readerThread
// (/root/reference/publish/publisher.cpp:234-283) for benchmarking and replay (BASELINE.json north_star:
// "a new IQ file source replacing SoapySDR").
//   file=<path>,format=cu8|cs16|cf32[,repeat=N]     raw interleaved I,Q file
//   synthetic=<seed>[,format=...][,blocks=N]        deterministic noise + carriers
// Either takes [,throttle=X] (deliver blocks no faster than X times real time; 1 = like an SDR, default 0 = as fast
// as the bank goes) and [,delay=SECONDS] (wait before the first block so that ZeroMQ subscribers can join).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>

namespace aero {

int formatBytes(int fmt);                       // bytes per complex sample
int parseFormat(const std::string& name);       // AERODDC_CU8/CS16/CF32 or -1

class IqSource {
 public:
  virtual ~IqSource() {}
  virtual int format() const = 0;
  // fill `dst` with exactly n_complex samples; false at end of stream (a partial block is dropped)
  virtual bool read(void* dst, size_t n_complex) = 0;
  // read() behind the delay= / throttle= pacing; `sample_rate` turns block lengths into time
  bool next(void* dst, size_t n_complex, int sample_rate);
  // "file=...,..." / "synthetic=..." ; nullptr + message on error
  static std::unique_ptr<IqSource> open(const std::string& deviceStr, std::string* err);

 private:
  static std::unique_ptr<IqSource> openSource(const std::string& deviceStr, std::string* err);
  double throttle_ = 0.0, delay_ = 0.0;
  bool started_ = false;
  std::chrono::steady_clock::time_point t0_;
  double delivered_ = 0.0;   // seconds of signal handed out so far
};

}  // namespace aero
