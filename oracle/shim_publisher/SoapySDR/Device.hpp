// TEST INFRASTRUCTURE (oracle/): the slice of SoapySDR::Device that /root/reference/publish/publisher.cpp:27-51,
// 241-272 calls, backed by an IQ file (implemented in oracle/ref_publisher_harness.cpp). Device::make takes the
// same "file=<path>,format=cu8|cs16|cf32" string as aero-publish-b200 -d; readStream delivers CF32 like every
// SoapySDR driver does, exactly bufflen/2 samples per call (the reference sizes its VFO blocks on that,
// publisher.cpp:93-100,241-242,267-268).
#ifndef AERODDC_SOAPY_DEVICE_HPP
#define AERODDC_SOAPY_DEVICE_HPP
#include <map>
#include <string>
#include <vector>
namespace SoapySDR {
typedef std::map<std::string, std::string> Kwargs;
class Stream;
class Device {
public:
  static Device* make(const std::string& args);
  static void unmake(Device* d);
  void setGainMode(int, size_t, bool) {}
  void setGain(int, size_t, double) {}
  void setFrequency(int, size_t, double) {}
  void setSampleRate(int, size_t, double) {}
  void setDCOffsetMode(int, size_t, bool) {}
  void writeSetting(const std::string&, const std::string&) {}
  Stream* setupStream(int direction, const std::string& format, const std::vector<size_t>& channels, const Kwargs& args);
  int activateStream(Stream*) { return 0; }
  int deactivateStream(Stream*) { return 0; }
  void closeStream(Stream*) {}
  int readStream(Stream* s, void* const* buffs, size_t numElems, int& flags, long long& timeNs, long timeoutUs);
  void* impl = nullptr;
};
}
#endif
