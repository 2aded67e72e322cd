// TEST INFRASTRUCTURE (oracle/) - never linked into the product.
//
// C entry points around the reference's UNMODIFIED aero-decode signal chain, compiled from /root/reference/decode
// into oracle/_ref/libref_decode.so (recipe: oracle/Makefile). This is the "unchanged aero-decode" of the
// end-to-end check (BASELINE.json configs[4]): int16 payloads exactly as they travel in the
// third ZeroMQ frame go in, the decoded ACARS / ISU records come out, and the test compares the record set obtained
// from the GPU bank's payloads with the one obtained from the CPU chain's payloads.
//
// What is and is not the reference here:
//   reference, unmodified: aerol.cpp (frame sync, de-interleaver, descrambler, CRC, SU/ISU/ACARS parsing),
//     mskdemodulator.cpp, oqpskdemodulator.cpp, DSP.cpp, coarsefreqestimate.cpp, jconvolutionalcodec.cpp, jfft.cpp,
//     fftwrapper.cpp, hunter.cpp, databasetext.cpp;
//   stand-ins: Qt (oracle/shim_decode/qt_decode_shim.h), libcorrect's Viterbi (oracle/viterbi_restated.c, parity
//     unpinned), and this file, which plays the part of (1) moc - the bodies of the signal functions - and (2)
//     Decoder::Decoder's wiring (decode/decode.cpp:104-243): audioReceived -> demodulator.dataReceived ->
//     processDemodulatedSoftBits -> AeroL; SignalStatus -> SignalHunter; newFreqCenter -> CenterFreqChangedSlot;
//     DataCarrierDetect -> SignalHunter::handleDcd; ACARSsignal / ACARSfragmentsignal -> output.
//   Not compiled: decode.cpp / output.cpp / forwarder.cpp (ZeroMQ, QJson, libacars, sockets). Records are therefore
//     printed in this harness's own one-line format, not in the reference's output formats.
// Wall-clock pieces are tied to the audio clock so that runs are reproducible: AeroL's 1 s DCD timer
// (aerol.cpp:900-902) is driven once per second of audio fed.
#include "qt_decode_shim.h"

#include <string>

// The harness calls two private slots (AeroL::updateDCD, ParserISU::acarslookupresult) and reaches the
// demodulators' coarse frequency estimator the way Qt's meta-object calls would. Access specifiers do not change
// the class layout, so this affects only what this translation unit may name.
#define private public
#define protected public
#include "aerol.h"
#include "hunter.h"
#include "mskdemodulator.h"
#include "oqpskdemodulator.h"
#undef private
#undef protected

namespace {

struct Chain;

class ChainRoot : public QObject {
public:
  explicit ChainRoot(Chain* c) : QObject(nullptr), chain(c) {}
  Chain* chain;
};

struct Chain {
  ChainRoot* root = nullptr;
  AeroL* aerol = nullptr;
  MskDemodulator* msk = nullptr;
  OqpskDemodulator* oqpsk = nullptr;
  SignalHunter* hunter = nullptr;
  int bitrate = 600;
  bool reassembly = true;
  double audio_clock = 0.0;   // seconds of audio since the last DCD timer tick
  std::string out;            // newline-terminated records
  std::string log;            // hunter / DCD events (diagnostics, not compared)
};

Chain* chain_of(const QObject* o) {
  while (o) {
    if (const ChainRoot* r = dynamic_cast<const ChainRoot*>(o)) return r->chain;
    o = o->parent();
  }
  return nullptr;
}

std::string hex_escape(const std::string& s) {
  std::string r;
  char b[8];
  for (unsigned char c : s) {
    if (c >= 0x20 && c < 0x7F && c != '\\' && c != '|') r += (char)c;
    else { std::snprintf(b, sizeof b, "\\x%02X", c); r += b; }
  }
  return r;
}

void record_acars(Chain* c, const char* kind, ACARSItem& it) {
  char head[256];
  std::snprintf(head, sizeof head, "%s|AES=%06X|GES=%02X|QNO=%02X|REFNO=%02X|MODE=%02X|TAK=%02X|BI=%02X|DL=%d|MORE=%d|NONACARS=%d|", kind,
                it.isuitem.AESID, it.isuitem.GESID, it.isuitem.QNO, it.isuitem.REFNO, (unsigned char)it.MODE, it.TAK, it.BI, (int)it.downlink,
                (int)it.moretocome, (int)it.nonacars);
  c->out += head;
  c->out += "LABEL=" + hex_escape(it.LABEL.toStdString()) + "|REG=" + hex_escape(it.PLANEREG.toStdString()) + "|TEXT=" +
            hex_escape(it.message.toStdString()) + "\n";
}

}  // namespace

// ---- what moc would generate: signal bodies, routed as decode.cpp / the constructors connect them --------------------

void SignalHunter::newFreqCenter(double f) {
  Chain* c = chain_of(this);
  if (!c) return;
  char b[64];
  std::snprintf(b, sizeof b, "hunt %.1f\n", f);
  c->log += b;
  if (c->msk) c->msk->CenterFreqChangedSlot(f);       // decode.cpp:223-224
  if (c->oqpsk) c->oqpsk->CenterFreqChangedSlot(f);   // decode.cpp:196-197
}
void SignalHunter::noSignalAfterScan() {
  if (Chain* c = chain_of(this)) c->log += "full scan without signal\n";
}
void SignalHunter::dcdChange(bool, bool now) {
  if (Chain* c = chain_of(this)) c->log += now ? "dcd 1\n" : "dcd 0\n";
}

void CoarseFreqEstimate::FreqOffsetEstimate(double est) {
  // mskdemodulator.cpp:75-76, oqpskdemodulator.cpp:57-58: the estimator's parent is its demodulator
  if (MskDemodulator* m = dynamic_cast<MskDemodulator*>(parent())) m->FreqOffsetEstimateSlot(est);
  else if (OqpskDemodulator* o = dynamic_cast<OqpskDemodulator*>(parent())) o->FreqOffsetEstimateSlot(est);
}

void MskDemodulator::ScatterPoints(const QVector<cpx_type>&) {}
void MskDemodulator::SymbolPhase(double) {}
void MskDemodulator::BBOverlapedBuffer(const QVector<cpx_type>& b) { coarsefreqestimate->ProcessBasebandData(b); }   // mskdemodulator.cpp:72-74
void MskDemodulator::OrgOverlapedBuffer(const QVector<double>&) {}
void MskDemodulator::Plottables(double, double, double) {}
void MskDemodulator::PeakVolume(double) {}
void MskDemodulator::processDemodulatedSoftBits(const QVector<short>& bits) {
  if (Chain* c = chain_of(this)) c->aerol->processDemodulatedSoftBits(bits);   // decode.cpp:216-218
}
void MskDemodulator::RxData(const QByteArray&) {}
void MskDemodulator::MSESignal(double) {}
void MskDemodulator::SignalStatus(bool got) {
  if (Chain* c = chain_of(this)) c->hunter->updatedSignalStatus(got);          // decode.cpp:219-220
}
void MskDemodulator::WarningTextSignal(const QString&) {}
void MskDemodulator::EbNoMeasurmentSignal(double) {}
void MskDemodulator::SampleRateChanged(double) {}
void MskDemodulator::BitRateChanged(double, bool) {}

void OqpskDemodulator::ScatterPoints(const QVector<cpx_type>&) {}
void OqpskDemodulator::OrgOverlapedBuffer(const QVector<double>&) {}
void OqpskDemodulator::PeakVolume(double) {}
void OqpskDemodulator::SampleRateChanged(double) {}
void OqpskDemodulator::BitRateChanged(double, bool) {}
void OqpskDemodulator::Plottables(double, double, double) {}
void OqpskDemodulator::BBOverlapedBuffer(const QVector<cpx_type>& b) { coarsefreqestimate->ProcessBasebandData(b); }  // oqpskdemodulator.cpp:54-56
void OqpskDemodulator::MSESignal(double) {}
void OqpskDemodulator::SignalStatus(bool got) {
  if (Chain* c = chain_of(this)) c->hunter->updatedSignalStatus(got);          // decode.cpp:192-193
}
void OqpskDemodulator::WarningTextSignal(const QString&) {}
void OqpskDemodulator::EbNoMeasurmentSignal(double) {}
void OqpskDemodulator::processDemodulatedSoftBits(const QVector<short>& bits) {
  if (Chain* c = chain_of(this)) c->aerol->processDemodulatedSoftBits(bits);   // decode.cpp:189-191
}

void AeroL::DataCarrierDetect(bool status) {
  Chain* c = chain_of(this);
  if (c && c->hunter) c->hunter->handleDcd(status);                             // decode.cpp:228-229
}
void AeroL::ACARSfragmentsignal(ACARSItem& it) {
  if (Chain* c = chain_of(this)) record_acars(c, "FRAGMENT", it);               // decode.cpp:233-236 (--disable-reassembly)
}
void AeroL::ACARSsignal(ACARSItem& it) {
  if (Chain* c = chain_of(this)) record_acars(c, "ACARS", it);                  // decode.cpp:238-241
}
void AeroL::Errorsignal(QString& e) {
  if (Chain* c = chain_of(this)) c->out += "ERROR|" + hex_escape(e.toStdString()) + "\n";
}
void AeroL::Voicesignal(QByteArray&, QString&) {}
void AeroL::Voicesignal(const QByteArray&) {}
void AeroL::CChannelAssignmentSignal(CChannelAssignmentItem& it) {
  if (Chain* c = chain_of(this)) {
    char b[200];
    std::snprintf(b, sizeof b, "CASSIGN|AES=%06X|GES=%02X|TYPE=%02X|RX=%.4f|TX=%.4f\n", it.AESID, it.GESID, it.type, it.receive_freq, it.transmit_freq);
    c->out += b;
  }
}
void AeroL::Call_progress_Signal(QByteArray) {}

// aerol.cpp:886-891 connects the parser's signals to AeroL's signals of the same name (its parent)
void ParserISU::ACARSsignal(ACARSItem& it) {
  if (AeroL* a = dynamic_cast<AeroL*>(parent())) a->ACARSsignal(it);
}
void ParserISU::ACARSfragmentsignal(ACARSItem& it) {
  if (AeroL* a = dynamic_cast<AeroL*>(parent())) a->ACARSfragmentsignal(it);
}
void ParserISU::Errorsignal(QString& e) {
  if (AeroL* a = dynamic_cast<AeroL*>(parent())) a->Errorsignal(e);
}

// aerol.cpp:329-331: the lookup result comes back to the parser that owns the DataBaseTextUser
void DataBaseTextUser::result(bool ok, int ref, const QStringList& values) {
  if (ParserISU* p = dynamic_cast<ParserISU*>(parent())) p->acarslookupresult(ok, ref, values);
}
void DataBaseText::asyncDbLookupFromAES(const QString&, const QString&, int, QObject*, const char*) {}

// ---- C entry points -----------------------------------------------------------------------------------------------

extern "C" {

// decode.cpp:117-160,199-226 for the continuous (non-burst) channel types: 600 / 1200 bps MSK, 10500 bps OQPSK.
void* refdec_create(int bitrate) {
  if (bitrate != 600 && bitrate != 1200 && bitrate != 10500) return nullptr;
  Chain* c = new Chain;
  c->bitrate = bitrate;
  c->root = new ChainRoot(c);
  c->aerol = new AeroL(c->root);
  c->aerol->setBitRate(bitrate);
  c->aerol->setBurstmode(false);
  c->hunter = new SignalHunter(15, c->root);
  if (bitrate > 1200) {
    OqpskDemodulator::Settings s;
    s.zmqAudio = true;
    s.freq_center = 0;
    c->oqpsk = new OqpskDemodulator(c->root);
    c->oqpsk->setAFC(true);
    c->oqpsk->setCPUReduce(false);
    c->oqpsk->setSettings(s);
    c->hunter->setParams(0, 25000, 10500);
  } else {
    MskDemodulator::Settings s;
    s.zmqAudio = true;
    s.freq_center = 0;
    s.Fs = bitrate == 600 ? 12000 : 24000;
    c->msk = new MskDemodulator(c->root);
    c->msk->setAFC(true);
    c->msk->setCPUReduce(false);
    c->msk->setSettings(s);
    c->hunter->setParams(0, 6000, 900);
  }
  return c;
}

// One ZeroMQ message (decode.cpp:341-347): frame 2 = rate, frame 3 = payload.
void refdec_feed(void* h, const int16_t* audio, size_t n, uint32_t rate) {
  Chain* c = (Chain*)h;
  QByteArray qdata((const char*)audio, (int)(n * sizeof(int16_t)));
  if (c->msk) c->msk->dataReceived(qdata, rate);
  if (c->oqpsk) c->oqpsk->dataReceived(qdata, rate);
  c->audio_clock += (double)n / (double)rate;
  while (c->audio_clock >= 1.0) {   // the 1 s QTimer of aerol.cpp:900-902, on the audio clock
    c->audio_clock -= 1.0;
    c->aerol->updateDCD();
  }
}

// Soft bits straight into the frame decoder (what a demodulator emits: 0..255, >= 128 is a one).
void refdec_feed_softbits(void* h, const int16_t* bits, size_t n) {
  Chain* c = (Chain*)h;
  QVector<short> v;
  v.reserve(n);
  for (size_t i = 0; i < n; i++) v.push_back(bits[i]);
  c->aerol->processDemodulatedSoftBits(v);
}

static size_t drain(std::string& s, char* buf, size_t cap) {
  const size_t need = s.size();
  if (buf && cap > need) {
    std::memcpy(buf, s.data(), need);
    buf[need] = 0;
    s.clear();
  }
  return need;
}
// Returns the size of the pending text; copies and clears it when cap > size.
size_t refdec_output(void* h, char* buf, size_t cap) { return drain(((Chain*)h)->out, buf, cap); }
size_t refdec_log(void* h, char* buf, size_t cap) { return drain(((Chain*)h)->log, buf, cap); }
double refdec_demod_freq(void* h) {
  Chain* c = (Chain*)h;
  return c->msk ? c->msk->getCurrentFreq() : c->oqpsk ? c->oqpsk->getCurrentFreq() : 0.0;
}

void refdec_destroy(void* h) {
  Chain* c = (Chain*)h;
  if (!c) return;
  delete c->msk;
  delete c->oqpsk;
  delete c->hunter;
  delete c->aerol;
  delete c->root;
  delete c;
}

}  // extern "C"
