// Host-side (init-time) coefficient generation for the DDC bank. Must agree bit for bit with the
// reference's designers because the taps are float32 inputs of the device arithmetic:
//   firfilter::low_pass / compute_ntaps / hamming   /root/reference/publish/firfilter.cpp:46-99,186-193
//   FIRHilbert::FIRHilbert                          /root/reference/publish/dsp.cpp:181-215
//   Oscillator rotation                             /root/reference/publish/oscillator.cpp:8-10
#pragma once
#include <vector>

namespace aeroddc {

// Hamming-windowed sinc low-pass with DC gain `gain`. Empty result = parameters rejected
// (the reference throws std::out_of_range, firfilter.cpp:100-112).
std::vector<float> design_lowpass(double gain, double fs, double cutoff, double transition);

// 125-tap (len) Hilbert transformer, energy-normalised and time-reversed as the reference stores it.
std::vector<float> design_hilbert(int len, int fs_param);

// (float)cos, (float)sin of the per-sample phase step.
void design_rotation(double fs, double freq, float* c, float* s);

// Parameters derived in vfo::init (vfo.cpp:62-79,88-102,111-112).
struct TailPlan {
  int out_rate;       // outputRate sent in ZMQ frame 2
  int n_stage;        // stage-D samples per block
  int n_out;          // output samples per block
  int late;           // 0 or lateDecimate
  std::vector<float> late_taps, usb_taps, hilbert_taps;
};
// returns false when a design is rejected
bool plan_tail(int fs, int block_len, int decim, int late, int filter_bw, bool demod_usb, TailPlan* out);

}  // namespace aeroddc
