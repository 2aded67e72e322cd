#include "design.h"

#include <cmath>

namespace aeroddc {

static const double kPi = 3.14159265358979323846264338327950288;

std::vector<float> design_lowpass(double gain, double fs, double cutoff, double transition) {
  std::vector<float> h;
  if (!(fs > 0.0) || !(cutoff > 0.0) || cutoff > fs / 2 || !(transition > 0.0)) return h;
  // tap count from the window's stop-band attenuation (Hamming: 53 dB), forced odd
  int n = (int)(53.0 * fs / (22.0 * transition));
  if (!(n & 1)) ++n;
  const int half = (n - 1) / 2;
  // the window is evaluated in double against a float-typed (n-1) and stored as float
  std::vector<float> win(n);
  const float span = static_cast<float>(n - 1);
  for (int i = 0; i < n; ++i) win[i] = static_cast<float>(0.54 - 0.46 * std::cos((2 * kPi * i) / span));
  // truncated ideal response, each tap rounded to float as it is produced
  h.resize(n);
  const double w0 = 2 * kPi * cutoff / fs;
  for (int k = -half; k <= half; ++k) {
    const double ideal = (k == 0) ? w0 / kPi : std::sin(k * w0) / (k * kPi);
    h[k + half] = static_cast<float>(ideal * win[k + half]);
  }
  // DC gain of the float taps, summed in double from the centre outwards
  double dc = h[half];
  for (int k = 1; k <= half; ++k) dc += 2 * h[k + half];
  const double scale = gain / dc;
  for (int i = 0; i < n; ++i) h[i] = static_cast<float>(h[i] * scale);
  return h;
}

std::vector<float> design_hilbert(int len, int fs_param) {
  std::vector<float> raw(len), h(len);
  const int mid = len / 2;
  float energy = 0.0f;   // float accumulation and float sqrt, as the reference ends up doing
  for (int i = 0; i < len; ++i) {
    const int k = i - mid;
    raw[i] = (k == 0) ? 0.0f : static_cast<float>(fs_param / (kPi * k) * (1 - std::cos(kPi * k)));
    const float sq = raw[i] * raw[i];
    energy = energy + sq;
  }
  const double norm = static_cast<double>(std::sqrt(energy));
  for (int i = 0; i < len; ++i) h[i] = static_cast<float>(raw[len - 1 - i] / norm);
  return h;
}

void design_rotation(double fs, double freq, float* c, float* s) {
  const double step = 2.0 * kPi * freq / fs;
  *c = static_cast<float>(std::cos(step));
  *s = static_cast<float>(std::sin(step));
}

bool plan_tail(int fs, int block_len, int decim, int late, int filter_bw, bool demod_usb, TailPlan* out) {
  TailPlan t;
  int rate = (int)(fs / std::pow(2.0, decim));
  int n_stage = (int)(block_len / std::pow(2.0, decim));
  t.n_stage = n_stage;
  int n_out = n_stage;
  t.late = 0;
  if (demod_usb && late > 0) {
    t.late = late;
    rate = rate / late;
    n_out = n_out / late;
    t.late_taps = design_lowpass(2, rate * late, rate / 2, (double)rate / (late - 1));
    if (t.late_taps.empty()) return false;
  }
  t.out_rate = rate;
  t.n_out = n_out;
  if (filter_bw > 0) {   // designed even when the VFO is not USB-demodulated; only used when it is
    t.usb_taps = design_lowpass(2, rate, filter_bw, (double)filter_bw / 4);
    if (t.usb_taps.empty()) return false;
    if (!demod_usb) t.usb_taps.clear();
  }
  t.hilbert_taps = design_hilbert(125, n_out);
  *out = t;
  return true;
}

}  // namespace aeroddc
