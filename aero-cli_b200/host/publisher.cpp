#include "publisher.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "ini.h"

#define CRIT(...) do { fprintf(stderr, "[CRIT] " __VA_ARGS__); fputc('\n', stderr); } while (0)

Publisher::Publisher(const std::string& deviceStr, bool enableBiast_, bool enableDcc_, const std::string& settingsPath) {
  enableBiast = enableBiast_;
  enableDcc = enableDcc_;
  running = false;
  if (!loadSettings(settingsPath)) {   // publisher.cpp:22-25
    CRIT("[ERROR] failed to parse and load settings: %s", error.c_str());
    return;
  }
  std::string err;
  source = aero::IqSource::open(deviceStr, &err);   // SoapySDR::Device::make in the reference (publisher.cpp:27-31)
  if (!source) {
    error = err;
    CRIT("[ERROR] %s", err.c_str());
    return;
  }
  try {
    const int fmt = source->format();
    // GPUs: "gpus=N" in the device string (or AERODDC_GPUS) shards the VFOs over devices 0..N-1 (NCCL broadcast of each block)
    int ngpu = 1;
    const size_t gp = deviceStr.find("gpus=");
    if (gp != std::string::npos) ngpu = atoi(deviceStr.c_str() + gp + 5);
    else if (const char* e = getenv("AERODDC_GPUS")) ngpu = atoi(e);
    std::vector<int> devs;
    for (int i = 0; i < std::max(1, ngpu); ++i) devs.push_back(i);
    bank = std::make_shared<aero::DdcBank>(Fs, buflen / 2, fmt, devs);
    // "mode=fast" in the device string selects the tolerance mode (fused arithmetic, not bit-identical; include/aeroddc.h)
    if (deviceStr.find("mode=fast") != std::string::npos && aeroddc_fleet_set_mode(bank->handle(), AERODDC_MODE_FAST) != AERODDC_OK)
      throw std::runtime_error(aeroddc_last_error());
    // "mode=tensor": the tensor-core formulation (same tolerance class; raw-fed VFOs with decim_count >= 6 on cf32 input)
    if (deviceStr.find("mode=tensor") != std::string::npos && aeroddc_fleet_set_mode(bank->handle(), AERODDC_MODE_TENSOR) != AERODDC_OK)
      throw std::runtime_error(aeroddc_last_error());
    if (enableDcc && aeroddc_fleet_set_dc_correction(bank->handle(), 1) != AERODDC_OK) throw std::runtime_error(aeroddc_last_error());   // publisher.cpp:292-296, on the GPU
    for (vfo* m : VFOmain) m->addToBank(bank, -1);
    for (vfo* f : VFOflat) f->addToBank(bank, -1);
    bank->finalize();
    for (vfo* m : VFOmain) m->connectSockets();   // the socket half of vfo::init; throws when there is no way to publish
    for (vfo* f : VFOflat) f->connectSockets();
  } catch (const std::exception& e) {
    error = e.what();
    CRIT("[ERROR] %s", e.what());
    return;
  }
  running = true;
}

Publisher::~Publisher() {
  running = false;
  if (mainReader.joinable()) mainReader.join();
  for (vfo* m : VFOmain) delete m;   // deletes its sub-VFOs (vfo.cpp:50-55)
  for (vfo* f : VFOflat) delete f;
}

bool Publisher::parseOnly(const std::string& settingsPath, Publisher** out, std::string* err) {
  Publisher* p = new Publisher();
  const bool ok = p->loadSettings(settingsPath);
  if (!ok && err) *err = p->error;
  if (ok && out) *out = p; else delete p;
  return ok;
}

// ---- settings-file rules (Publisher::loadSettings, publisher.cpp:55-227), one function per rule -----------------------
namespace {

// publisher.cpp:93-100: four blocks per second of interleaved floats, five when a quarter is not a multiple of 512
struct BlockSplit { int parts, floats; };
BlockSplit splitBlocks(int fs) {
  const long long twice = 2LL * fs;
  if ((int)(twice / 4) % 512 > 0) return {5, (int)(twice / 5)};
  return {4, (int)(twice / 4)};
}

// publisher.cpp:140-141: integer ratio, log2 truncated; a ratio of 1 is no decimation
int mainDecimation(int fs, int out_rate) {
  const int ratio = fs / out_rate;
  return ratio == 1 ? 0 : (int)std::log2(ratio);
}

// publisher.cpp:164-176: an explicit out_rate wins, else the channel's bit rate picks the audio rate
int audioRate(int out_rate, int data_rate) {
  if (out_rate != 0 || data_rate <= 0) return out_rate;
  if (data_rate == 600) return 12000;
  if (data_rate == 1200) return 24000;
  return 48000;
}

// publisher.cpp:198-210: a parent at 5 x or 6 x 48 kHz uses the late /5 or /6 stage; otherwise plain halvings
struct Chain { int halvings, late; };
Chain subChain(int fs, int parent_rate, int out_rate) {
  const int k = parent_rate / 48000;
  if (k == 5 || k == 6) return {(int)std::log2(parent_rate / (k * out_rate)), k};
  return {(int)std::log2(fs / out_rate) - (int)std::log2(fs / parent_rate), 0};
}

}  // namespace

// publisher.cpp:183-193: the first main VFO that only feeds sub-VFOs and lies within ONE output rate (not half) wins
int Publisher::matchMainVfo(int vfo_freq) const {
  for (size_t a = 0; a < VFOmain.size(); ++a) {
    vfo* m = VFOmain[a];   // the getters are non-const like the reference's (vfo.h:31-38)
    const int offset = std::abs((center_frequency - m->getMixerFreq()) - vfo_freq);
    if (offset < m->getOutRate() && !m->getDemodUSB()) return (int)a;
  }
  return -1;
}

bool Publisher::loadSettings(const std::string& settingsPath) {
  aero::IniSettings ini;
  if (!ini.load(settingsPath)) {
    error = "Provided settings file path either doesn't exist or isn't a file: " + settingsPath;
    return false;
  }
  Fs = ini.toInt("sample_rate");
  if (Fs == 0) { error = "Provided sample rate in settings file either doesn't exist or isn't an integer"; return false; }
  if (std::find(validSampleRates.begin(), validSampleRates.end(), Fs) == validSampleRates.end()) {
    error = "Provided sample rate is not supported: " + std::to_string(Fs);
    return false;
  }
  center_frequency = ini.toInt("center_frequency");
  tuner_idx = ini.toInt("auto_start_tuner_idx");
  if (ini.toInt("auto_start_biast") == 1) enableBiast = true;
  if (ini.toInt("tuner_gain") > 0) tuner_gain = ini.toInt("tuner_gain");
  if (ini.toInt("remote_rtl_gain_idx") > 0) tuner_gain_idx = ini.toInt("remote_rtl_gain_idx");
  if (ini.value("correct_dc_bias") == "1") enableDcc = true;
  const int mix_offset = ini.toInt("mix_offset");
  const std::string pub_address = ini.value("zmq_address");
  const BlockSplit split = splitBlocks(Fs);
  buflen = split.floats;

  // [main_vfos] (publisher.cpp:118-148): mix + decimate only; their stage-D stream feeds the sub-VFOs
  const int n_main = ini.beginReadArray("main_vfos");
  if (n_main > 3) { error = "more than 3 main VFOs (VFOsub[3], publisher.h:50)"; return false; }
  for (int i = 0; i < n_main; ++i) {
    ini.setArrayIndex(i);
    const int rate = ini.toInt("out_rate");
    if (rate <= 0) { error = "main VFO without out_rate"; return false; }
    vfo* m = new vfo();
    if (ini.toInt("compress_scale") > 0) m->setScaleComp(ini.toInt("compress_scale"));
    if (!ini.value("zmq_address").empty() && !ini.value("zmq_topic").empty()) {
      m->setZmqAddress(ini.value("zmq_address"));
      m->setZmqTopic(ini.value("zmq_topic"));
    }
    m->setFs(Fs);
    m->setDecimationCount(mainDecimation(Fs, rate));
    m->setMixerFreq(center_frequency - ini.toInt("frequency"));
    m->setDemodUSB(false);
    m->setCompressonStyle(1);
    m->init(buflen / 2, false);
    m->setVFOs(&VFOsub[i]);
    VFOmain.push_back(m);
  }
  ini.endArray();

  // [vfos] (publisher.cpp:156-222): USB-demodulating leaves, each behind the main VFO it matches
  nVFO = ini.beginReadArray("vfos");
  for (int i = 0; i < nVFO; ++i) {
    ini.setArrayIndex(i);
    const std::string label = "VFO " + std::to_string(i + 1);
    const int vfo_freq = ini.toInt("frequency") + mix_offset;
    const int out_rate = audioRate(ini.toInt("out_rate"), ini.toInt("data_rate"));
    if (out_rate <= 0) { error = label + " has neither out_rate nor data_rate"; return false; }
    const int parent = matchMainVfo(vfo_freq);
    if (parent < 0 && !VFOmain.empty()) {
      // the reference would park it in VFOsub[0] and feed it a stream of the wrong rate (publisher.cpp:219)
      error = label + " matches no main VFO";
      return false;
    }
    const int parent_mixer = parent < 0 ? 0 : (int)VFOmain[parent]->getMixerFreq();
    const int parent_rate = parent < 0 ? Fs : VFOmain[parent]->getOutRate();
    const Chain chain = subChain(Fs, parent_rate, out_rate);
    vfo* v = new vfo();
    v->setZmqTopic(ini.value("topic"));
    v->setZmqAddress(pub_address);
    v->setDecimationCount(chain.halvings);
    v->setFilterBandwidth(ini.toInt("filter_bandwidth"));
    v->setGain((float)ini.toFloat("gain") / 100);
    v->setMixerFreq((center_frequency - parent_mixer) - vfo_freq);
    v->setFs(parent_rate);
    v->setCompressonStyle(1);
    v->init(parent_rate / split.parts, true, chain.late);
    // Without [main_vfos] the reference builds these VFOs but never runs them (demodData walks the main VFOs only,
    // publisher.cpp:301-305); here such a file describes a flat bank on the raw stream.
    if (parent >= 0) VFOsub[parent].push_back(v);
    else VFOflat.push_back(v);
  }
  ini.endArray();
  return true;
}

void Publisher::run() { mainReader = std::thread([this] { readerThread(); }); }
void Publisher::wait() { if (mainReader.joinable()) mainReader.join(); }

void Publisher::readerThread() {
  // The reference reads a block and runs every VFO on it before reading the next (publisher.cpp:266-276). Here the
  // two halves of the pinned host ring alternate: while the GPUs work on block k (async H2D, kernels, payload D2H)
  // the source fills the other half with block k+1, and block k+1 is submitted before block k is published.
  void* slot[2] = {nullptr, nullptr};
  size_t bytes = 0;
  const size_t n = (size_t)buflen / 2;
  if (!running) goto Exit;
  for (int i = 0; i < 2; ++i)
    if (aeroddc_fleet_host_slot(bank->handle(), i, &slot[i], &bytes) != AERODDC_OK) { error = aeroddc_last_error(); goto Exit; }
  try {
    // a failed read is "SoapySDR could not read stream from SDR" in the reference (publisher.cpp:269-272): end of stream
    bool have = source->next(slot[0], n, Fs);
    if (have) demodData(slot[0]);
    for (long long k = 0; have; ++k) {
      const bool more = running && source->next(slot[(k + 1) & 1], n, Fs);
      bank->wait();                                  // block k; its payloads stay valid until the next wait()
      if (more) demodData(slot[(k + 1) & 1]);
      transmitData();
      have = more;
    }
  } catch (const std::exception& e) {
    error = e.what();
    CRIT("%s", e.what());
  }
Exit:
  running = false;
  if (completed) completed();
}

// hands one raw block to every main VFO, sub-VFO and flat VFO: one asynchronous GPU pass (publisher.cpp:285-306)
void Publisher::demodData(void* block) { bank->submit(block, (size_t)buflen / 2); }

// one ZeroMQ message per publishing VFO for the block that wait() just retired (publisher.cpp:301-305 -> vfo::transmitData)
void Publisher::transmitData() {
  for (vfo* m : VFOmain) m->transmitData();
  for (vfo* f : VFOflat) f->transmitData();
  ++blocks;
}
