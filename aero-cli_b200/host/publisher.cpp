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
    if (enableDcc && aeroddc_fleet_set_dc_correction(bank->handle(), 1) != AERODDC_OK) throw std::runtime_error(aeroddc_last_error());   // publisher.cpp:292-296, on the GPU
    for (vfo* m : VFOmain) m->addToBank(bank, -1);
    for (vfo* f : VFOflat) f->addToBank(bank, -1);
    bank->finalize();
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

bool Publisher::loadSettings(const std::string& settingsPath) {
  aero::IniSettings settings;
  if (!settings.load(settingsPath)) {
    error = "Provided settings file path either doesn't exist or isn't a file: " + settingsPath;
    return false;
  }
  Fs = settings.toInt("sample_rate");
  if (Fs == 0) { error = "Provided sample rate in settings file either doesn't exist or isn't an integer"; return false; }
  bool valid = false;
  for (int r : validSampleRates) valid = valid || r == Fs;
  if (!valid) { error = "Provided sample rate is not supported: " + std::to_string(Fs); return false; }

  center_frequency = settings.toInt("center_frequency");
  tuner_idx = settings.toInt("auto_start_tuner_idx");
  enableBiast = enableBiast || (settings.toInt("auto_start_biast") == 1);
  const int gain = settings.toInt("tuner_gain");
  const int remote_gain_idx = settings.toInt("remote_rtl_gain_idx");
  const int mix_offset = settings.toInt("mix_offset");

  // usually 4 buffers per Fs but in some cases 5 due to multiple of 512 (publisher.cpp:93-100)
  int bufsplit = 4;
  if (double((int((2 * (long long)Fs) / 4)) % 512) > 0) {
    buflen = int((2 * (long long)Fs) / 5);
    bufsplit = 5;
  } else {
    buflen = int((2 * (long long)Fs) / 4);
  }
  if (gain > 0) tuner_gain = gain;
  if (remote_gain_idx > 0) tuner_gain_idx = remote_gain_idx;
  const std::string zmq_address = settings.value("zmq_address");
  enableDcc = enableDcc || settings.value("correct_dc_bias") == "1";

  const int msize = settings.beginReadArray("main_vfos");
  if (msize > 3) { error = "more than 3 main VFOs (VFOsub[3], publisher.h:50)"; return false; }
  for (int i = 0; i < msize; ++i) {   // publisher.cpp:118-148
    settings.setArrayIndex(i);
    vfo* pVFO = new vfo();
    const int vfo_freq = settings.toInt("frequency");
    const int vfo_out_rate = settings.toInt("out_rate");
    const std::string output_connect = settings.value("zmq_address");
    const std::string out_topic = settings.value("zmq_topic");
    const int compscale = settings.toInt("compress_scale");
    if (vfo_out_rate <= 0) { delete pVFO; error = "main VFO without out_rate"; return false; }
    if (compscale > 0) pVFO->setScaleComp(compscale);
    if (output_connect != "" && out_topic != "") {
      pVFO->setZmqAddress(output_connect);
      pVFO->setZmqTopic(out_topic);
    }
    pVFO->setFs(Fs);
    pVFO->setDecimationCount(Fs / vfo_out_rate == 1 ? 0 : int(log2(Fs / vfo_out_rate)));
    pVFO->setMixerFreq(center_frequency - vfo_freq);
    pVFO->setDemodUSB(false);
    pVFO->setCompressonStyle(1);
    pVFO->init(buflen / 2, false);
    pVFO->setVFOs(&VFOsub[i]);
    VFOmain.push_back(pVFO);
  }
  settings.endArray();

  const int size = settings.beginReadArray("vfos");
  nVFO = size;
  for (int i = 0; i < size; ++i) {   // publisher.cpp:156-222
    settings.setArrayIndex(i);
    vfo* pVFO = new vfo();
    const int vfo_freq = settings.toInt("frequency") + mix_offset;
    const int data_rate = settings.toInt("data_rate");
    int out_rate = settings.toInt("out_rate");
    if (out_rate == 0 && data_rate > 0) {
      switch (data_rate) {
        case 600: out_rate = 12000; break;
        case 1200: out_rate = 24000; break;
        default: out_rate = 48000; break;
      }
    }
    if (out_rate <= 0) { delete pVFO; error = "VFO " + std::to_string(i + 1) + " has neither out_rate nor data_rate"; return false; }
    const int filterbw = settings.toInt("filter_bandwidth");
    int main_vfo_freq = 0;
    int main_vfo_out_rate = Fs;
    int main_idx = -1;
    for (size_t a = 0; a < VFOmain.size(); a++) {   // first main VFO within one output rate wins (not half)
      const int diff = std::abs((center_frequency - VFOmain[a]->getMixerFreq()) - vfo_freq);
      if (diff < VFOmain[a]->getOutRate() && !VFOmain[a]->getDemodUSB()) {
        main_idx = (int)a;
        main_vfo_freq = VFOmain[a]->getMixerFreq();
        main_vfo_out_rate = VFOmain[a]->getOutRate();
        break;
      }
    }
    pVFO->setZmqTopic(settings.value("topic"));
    pVFO->setZmqAddress(zmq_address);
    int lateDecimate = 0;
    if ((main_vfo_out_rate / 48000) == 5) {
      pVFO->setDecimationCount(int(log2(main_vfo_out_rate / (5 * out_rate))));
      lateDecimate = 5;
    } else if ((main_vfo_out_rate / 48000) == 6) {
      pVFO->setDecimationCount(int(log2(main_vfo_out_rate / (6 * out_rate))));
      lateDecimate = 6;
    } else {
      pVFO->setDecimationCount(int(log2(Fs / out_rate)) - int(log2(Fs / main_vfo_out_rate)));
    }
    pVFO->setFilterBandwidth(filterbw);
    pVFO->setGain((float)settings.toFloat("gain") / 100);
    pVFO->setMixerFreq((center_frequency - main_vfo_freq) - vfo_freq);
    pVFO->setFs(main_vfo_out_rate);
    pVFO->setCompressonStyle(1);
    pVFO->init(main_vfo_out_rate / bufsplit, true, lateDecimate);
    if (main_idx >= 0) {
      VFOsub[main_idx].push_back(pVFO);
    } else if (VFOmain.empty()) {
      // The reference parks such a VFO in VFOsub[0] and never processes it (demodData only walks the main
      // VFOs, publisher.cpp:301-305). Here a settings file without [main_vfos] describes a flat bank on the raw stream.
      VFOflat.push_back(pVFO);
    } else {
      // with main VFOs present the reference would feed this VFO a stream of the wrong rate (publisher.cpp:219)
      delete pVFO;
      error = "VFO " + std::to_string(i + 1) + " matches no main VFO";
      return false;
    }
  }
  settings.endArray();
  return true;
}

void Publisher::run() { mainReader = std::thread([this] { readerThread(); }); }
void Publisher::wait() { if (mainReader.joinable()) mainReader.join(); }

void Publisher::readerThread() {
  void* slot[2] = {nullptr, nullptr};
  size_t bytes = 0;
  if (!running) goto Exit;
  for (int i = 0; i < 2; ++i)
    if (aeroddc_fleet_host_slot(bank->handle(), i, &slot[i], &bytes) != AERODDC_OK) { error = aeroddc_last_error(); goto Exit; }
  while (running) {
    void* dst = slot[blocks & 1];   // the source writes straight into the pinned ring
    if (!source->next(dst, (size_t)buflen / 2, Fs)) {
      // "SoapySDR could not read stream from SDR" in the reference (publisher.cpp:269-272): end of stream
      break;
    }
    try {
      demodData(dst);
    } catch (const std::exception& e) {
      error = e.what();
      CRIT("%s", e.what());
      break;
    }
  }
Exit:
  running = false;
  if (completed) completed();
}

void Publisher::demodData(void* block) {
  bank->process(block, (size_t)buflen / 2);       // every main VFO, sub-VFO and flat VFO in one GPU pass
  for (vfo* m : VFOmain) m->transmitData();       // publisher.cpp:301-305 + vfo::transmitData
  for (vfo* f : VFOflat) f->transmitData();
  ++blocks;
}
