// IQ block sources that replace the SoapySDR device of Publisher::readerThread
// (/root/reference/publish/publisher.cpp:234-283) for benchmarking and replay (BASELINE.json north_star:
// "a new IQ file source replacing SoapySDR").
//   file=<path>,format=cu8|cs16|cf32[,repeat=N][,throttle=1]   raw interleaved I,Q file
//   synthetic=<seed>[,format=...][,blocks=N]                   deterministic noise + carriers
#pragma once
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
  // "file=...,..." / "synthetic=..." ; nullptr + message on error
  static std::unique_ptr<IqSource> open(const std::string& deviceStr, std::string* err);
};

}  // namespace aero
