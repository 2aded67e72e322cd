#include "iqsource.h"

#include <cmath>
#include <cstring>
#include <map>
#include <sstream>
#include <thread>
#include <vector>

#include "../../include/aeroddc.h"

namespace aero {

int formatBytes(int fmt) { return fmt == AERODDC_CU8 ? 2 : (fmt == AERODDC_CS16 ? 4 : 8); }
int parseFormat(const std::string& n) {
  if (n == "cu8" || n == "CU8") return AERODDC_CU8;
  if (n == "cs16" || n == "CS16") return AERODDC_CS16;
  if (n == "cf32" || n == "CF32") return AERODDC_CF32;
  return -1;
}

namespace {
std::map<std::string, std::string> parseArgs(const std::string& s) {
  std::map<std::string, std::string> kv;
  std::istringstream in(s);
  std::string item;
  while (std::getline(in, item, ',')) {
    const size_t eq = item.find('=');
    if (eq == std::string::npos) kv[item] = "";
    else kv[item.substr(0, eq)] = item.substr(eq + 1);
  }
  return kv;
}

class FileSource : public IqSource {
 public:
  FileSource(FILE* f, int fmt, long repeat) : f_(f), fmt_(fmt), repeat_(repeat) {}
  ~FileSource() override { if (f_) fclose(f_); }
  int format() const override { return fmt_; }
  bool read(void* dst, size_t n) override {
    const size_t want = n * formatBytes(fmt_);
    size_t got = fread(dst, 1, want, f_);
    while (got < want) {
      if (repeat_ == 0) return false;
      if (repeat_ > 0) --repeat_;
      rewind(f_);
      const size_t more = fread((char*)dst + got, 1, want - got, f_);
      if (more == 0) return false;
      got += more;
    }
    return true;
  }
 private:
  FILE* f_;
  int fmt_;
  long repeat_;   // remaining rewinds; -1 = forever
};

// counter-based generator: sample n depends only on (seed, n), so any block can be regenerated
class SyntheticSource : public IqSource {
 public:
  SyntheticSource(uint64_t seed, int fmt, long blocks) : seed_(seed), fmt_(fmt), blocks_(blocks), n_(0) {}
  int format() const override { return fmt_; }
  bool read(void* dst, size_t n) override {
    if (blocks_ == 0) return false;
    if (blocks_ > 0) --blocks_;
    for (size_t i = 0; i < n; ++i, ++n_) {
      const double u1 = unit(n_, 0), u2 = unit(n_, 7777);
      const double ph = 2 * M_PI * std::fmod(n_ * 0.0173, 1.0);
      const double re = 0.6 * ((u1 - 0.5) + 0.4 * std::cos(ph)), im = 0.6 * ((u2 - 0.5) + 0.4 * std::sin(ph));
      if (fmt_ == AERODDC_CF32) { ((float*)dst)[2 * i] = (float)re; ((float*)dst)[2 * i + 1] = (float)im; }
      else if (fmt_ == AERODDC_CS16) { ((int16_t*)dst)[2 * i] = (int16_t)std::lrint(re * 32767); ((int16_t*)dst)[2 * i + 1] = (int16_t)std::lrint(im * 32767); }
      else { ((uint8_t*)dst)[2 * i] = (uint8_t)std::lrint(re * 127 + 127.4); ((uint8_t*)dst)[2 * i + 1] = (uint8_t)std::lrint(im * 127 + 127.4); }
    }
    return true;
  }
 private:
  double unit(uint64_t n, uint64_t salt) const {
    uint64_t x = (n + seed_ + salt) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return (double)(x >> 40) / (double)(1 << 24);
  }
  uint64_t seed_;
  int fmt_;
  long blocks_;
  uint64_t n_;
};
}  // namespace

bool IqSource::next(void* dst, size_t n, int sample_rate) {
  using clock = std::chrono::steady_clock;
  if (!started_) {
    started_ = true;
    if (delay_ > 0) std::this_thread::sleep_for(std::chrono::duration<double>(delay_));
    t0_ = clock::now();
  }
  if (throttle_ > 0 && sample_rate > 0) {
    // block k becomes available when an SDR running at throttle_ x real time would have finished capturing it
    delivered_ += (double)n / (double)sample_rate;
    std::this_thread::sleep_until(t0_ + std::chrono::duration_cast<clock::duration>(std::chrono::duration<double>(delivered_ / throttle_)));
  }
  return read(dst, n);
}

std::unique_ptr<IqSource> IqSource::open(const std::string& deviceStr, std::string* err) {
  std::unique_ptr<IqSource> src = openSource(deviceStr, err);
  if (src) {
    auto kv = parseArgs(deviceStr);
    if (kv.count("throttle")) src->throttle_ = atof(kv["throttle"].c_str());
    if (kv.count("delay")) src->delay_ = atof(kv["delay"].c_str());
    if (src->throttle_ < 0 || src->delay_ < 0) { if (err) *err = "throttle= and delay= must not be negative"; return nullptr; }
  }
  return src;
}

std::unique_ptr<IqSource> IqSource::openSource(const std::string& deviceStr, std::string* err) {
  auto kv = parseArgs(deviceStr);
  int fmt = AERODDC_CF32;
  if (kv.count("format")) {
    fmt = parseFormat(kv["format"]);
    if (fmt < 0) { if (err) *err = "unknown IQ format: " + kv["format"]; return nullptr; }
  }
  if (kv.count("file")) {
    FILE* f = fopen(kv["file"].c_str(), "rb");
    if (!f) { if (err) *err = "cannot open IQ file: " + kv["file"]; return nullptr; }
    const long repeat = kv.count("repeat") ? atol(kv["repeat"].c_str()) : 0;
    return std::unique_ptr<IqSource>(new FileSource(f, fmt, repeat));
  }
  if (kv.count("synthetic")) {
    const long blocks = kv.count("blocks") ? atol(kv["blocks"].c_str()) : -1;
    return std::unique_ptr<IqSource>(new SyntheticSource(strtoull(kv["synthetic"].c_str(), nullptr, 10), fmt, blocks));
  }
  if (err) *err = "failed to find device: " + deviceStr + " (expected file=<path>,format=cu8|cs16|cf32 or synthetic=<seed>)";
  return nullptr;
}

}  // namespace aero
