#include "vfo.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace aero {

DdcBank::DdcBank(int sample_rate, int block_len, int in_format, const std::vector<int>& devices) : block_len_(block_len), format_(in_format) {
  if (aeroddc_fleet_create(&bank_, sample_rate, block_len, in_format, devices.data(), (int)devices.size()) != AERODDC_OK)
    throw std::runtime_error(std::string("aeroddc_fleet_create: ") + aeroddc_last_error());
}
DdcBank::~DdcBank() { aeroddc_fleet_destroy(bank_); }
void DdcBank::finalize() {
  if (finalized_) return;
  if (aeroddc_fleet_finalize(bank_) != AERODDC_OK) throw std::runtime_error(std::string("aeroddc_fleet_finalize: ") + aeroddc_last_error());
  finalized_ = true;
}
void DdcBank::process(const void* host_iq, size_t n_complex) {
  if (aeroddc_fleet_process(bank_, host_iq, n_complex) != AERODDC_OK)
    throw std::runtime_error(std::string("aeroddc_fleet_process: ") + aeroddc_last_error());
}
void DdcBank::submit(const void* host_iq, size_t n_complex) {
  if (aeroddc_fleet_submit(bank_, host_iq, n_complex) != AERODDC_OK)
    throw std::runtime_error(std::string("aeroddc_fleet_submit: ") + aeroddc_last_error());
}
void DdcBank::wait() {
  if (aeroddc_fleet_wait(bank_) != AERODDC_OK) throw std::runtime_error(std::string("aeroddc_fleet_wait: ") + aeroddc_last_error());
}

}  // namespace aero

ZmqPublisher vfo::bind_publisher;

vfo::vfo() {   // defaults of vfo::vfo (vfo.cpp:5-29); Fs / decimateCount / mixer_freq are uninitialised there
  gain = 0.01f;
  demodUSB = true;
  filterbw = 0;
  offsetbw = 0;
  mpVFOs = nullptr;
  scalecomp = 1;
  cstyle = 0;
  Fs = 0;
  decimateCount = 0;
  mixer_freq = 0;
  outputRate = 0;
  zmqBind = false;
  samplesPerBuffer_ = 0;
  lateDecimate_ = 0;
  inited_ = false;
  socketsReady_ = false;
  index_ = -1;
}

vfo::~vfo() {   // a vfo owns its sub-VFOs (vfo.cpp:50-55)
  if (mpVFOs)
    for (vfo* s : *mpVFOs) delete s;
}

void vfo::init(int samplesPerBuffer, bool bind, int lateDecimate) {
  samplesPerBuffer_ = samplesPerBuffer;
  lateDecimate_ = lateDecimate;
  // outputRate as vfo::init derives it (vfo.cpp:62-88)
  int targetRate = (int)(Fs / std::pow(2.0, decimateCount));
  if (demodUSB && lateDecimate > 0) targetRate = targetRate / lateDecimate;
  outputRate = (uint32_t)targetRate;
  zmqBind = bind;   // the sockets themselves are opened by connectSockets(), not while settings are being parsed
  inited_ = true;
}

// The socket part of vfo::init (vfo.cpp:128-136), deferred until the VFO is about to run: parsing a settings file
// (aero-publish-b200 --plan) must not bind the publisher's address. Only VFOs that can publish open a socket: a main VFO
// that feeds sub-VFOs never sends (vfo.cpp:167-172), and IQ output needs a topic (vfo.cpp:300).
void vfo::connectSockets() {
  if (mpVFOs && !mpVFOs->empty()) {
    for (vfo* s : *mpVFOs) s->connectSockets();
    return;
  }
  if (socketsReady_ || (!demodUSB && zmqTopic.empty())) return;
  if (!ZmqPublisher::available())
    throw std::runtime_error("ZeroMQ output unavailable: libzmq could not be loaded (set AERODDC_LIBZMQ to its path) and no sink is installed");
  if (zmqBind) {
    if (!vfo::bind_publisher.connected) {
      vfo::bind_publisher.setAddress(zmqAddress);
      vfo::bind_publisher.setBind(true);
      vfo::bind_publisher.connect();
    }
  } else {
    connect_publisher.setBind(false);
    connect_publisher.setAddress(zmqAddress);
    connect_publisher.connect();
  }
  socketsReady_ = true;
}

void vfo::setZmqAddress(const std::string& address) { zmqAddress = address; }
void vfo::setZmqTopic(const std::string& top) { zmqTopic = top; }
void vfo::setFs(int samplerate) { Fs = samplerate; }
void vfo::setDecimationCount(int count) { decimateCount = count; }
void vfo::setMixerFreq(double freq) { mixer_freq = freq; }
double vfo::getMixerFreq() { return mixer_freq; }
int vfo::getOutRate() { return (int)(Fs / std::pow(2.0, decimateCount)); }
void vfo::setOffsetBandwidth(double bw) { offsetbw = (int)bw; }   // never called by the reference
void vfo::setFilterBandwidth(double bw) { filterbw = (int)bw; }
void vfo::setGain(float g) { gain = g; }
void vfo::setCompressonStyle(int st) { cstyle = st; }
void vfo::setScaleComp(int scale) { scalecomp = scale; }
void vfo::setDemodUSB(bool usb) { demodUSB = usb; }
bool vfo::getDemodUSB() { return demodUSB; }
void vfo::setVFOs(std::vector<vfo*>* vfos) { mpVFOs = vfos; }

void vfo::addToBank(const std::shared_ptr<aero::DdcBank>& bank, int parent) {
  if (!inited_) throw std::runtime_error("vfo::init was not called");
  if (offsetbw > 1) throw std::runtime_error("offset bandwidth (osc_bfo) is dead code in the reference and is not supported");
  aeroddc_vfo_desc d;
  memset(&d, 0, sizeof d);
  d.mixer_freq = mixer_freq;
  d.decim_count = decimateCount;
  d.late_decimate = lateDecimate_;
  d.filter_bw = filterbw;
  d.gain = gain;
  d.demod_usb = demodUSB ? 1 : 0;
  d.compress_style = cstyle;
  d.scale_comp = scalecomp;
  strncpy(d.topic, zmqTopic.c_str(), sizeof d.topic - 1);
  d.parent = parent;
  const int idx = aeroddc_fleet_add_vfo(bank->handle(), &d);
  if (idx < 0) throw std::runtime_error(std::string("aeroddc_fleet_add_vfo(") + zmqTopic + "): " + aeroddc_last_error());
  bank_ = bank;
  index_ = idx;
  if (mpVFOs)
    for (vfo* s : *mpVFOs) s->addToBank(bank, idx);
}

void vfo::process(const std::vector<cpx_typef>& samples) {
  if (!bank_) {   // driven on its own, like the reference's object: private bank with this VFO and its sub-VFOs
    auto bank = std::make_shared<aero::DdcBank>(Fs, samplesPerBuffer_, AERODDC_CF32);
    addToBank(bank, -1);
    bank->finalize();
  }
  if (!bank_->finalized()) bank_->finalize();
  // std::complex<float> is layout-compatible with interleaved float I,Q (publisher.cpp:288-299 builds exactly this)
  bank_->process(samples.data(), samples.size());
  transmitData();
}

void vfo::transmitData() {
  if (mpVFOs && !mpVFOs->empty()) {   // a main VFO only recurses (vfo.cpp:167-172)
    for (vfo* s : *mpVFOs) s->transmitData();
    return;
  }
  const void* payload = nullptr;
  size_t n = 0;
  uint32_t rate = 0;
  if (aeroddc_fleet_output(bank_->handle(), index_, &payload, &n, &rate) != AERODDC_OK)
    throw std::runtime_error(std::string("aeroddc_fleet_output: ") + aeroddc_last_error());
  if (!demodUSB && zmqTopic.empty()) return;   // vfo.cpp:300: IQ is only sent when a topic is set
  connectSockets();                             // first message of a VFO driven on its own; a no-op afterwards
  ZmqPublisher& pub = zmqBind ? vfo::bind_publisher : connect_publisher;
  pub.publish((unsigned char*)payload, (uint32_t)n, zmqTopic, rate);
}
