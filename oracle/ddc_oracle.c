/* TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, never linked or imported by the product path.
 *
 * CPU restatement, in plain C, of the reference's per-VFO digital down-converter chain
 * (aero-publish, /root/reference/publish). It is written from the algorithm, not from the
 * code: the reference is a set of stateful objects (circular buffers, linear queues, a table
 * pointer); this file states the same arithmetic as functions of sample indices:
 *
 *   nco(n)          = q[L-1] if n == 0 else q[n mod L]            (oscillator.cpp:4-39, vfo.cpp:155-161)
 *   x0[n]           = nco(n) * iq[n]                               (vfo.cpp:157)
 *   x_{s+1}[j]      = half-band(x_s window 2j-10 .. 2j)            (halfbanddecimator.cpp:35-60, dsp.cpp:102-150)
 *                     with the block-boundary rule of dsp.cpp:163-172 (see hb_stage below)
 *   tail            = late FIR / delay - Hilbert / fir_usb / int16 (vfo.cpp:188-258, dsp.cpp:64-78,216-231)
 *
 * Pinned against: the SURVEY.md section-8c golden anchors and oracle/_ref/libref_vfo.so (the
 * unmodified reference compiled here) by tests/test_oracle_golden.py and the fixtures in
 * tests/golden/. Build: gcc -std=c11 -O2 -ffp-contract=off (never -ffast-math; FMA contraction
 * changes the output bits, SURVEY.md finding 4).
 *
 * Every float operation below is written as a separate statement on `float` lvalues so that
 * no excess precision or contraction can creep in.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

#define HB_TAPS 11
#define HILBERT_LEN 125
#define DELAY_LEN 62 /* (125-1)/2, vfo.cpp:111 */
#define MAX_STAGES 8 /* hdecimator[8], vfo.h:63 */

/* ------------------------------------------------------------------------------------------
 * Host-side designs (init time)
 * ---------------------------------------------------------------------------------------- */

/* firfilter::low_pass + compute_ntaps + hamming, Hamming window only
 * (firfilter.cpp:46-99 and :186-193). Returns the tap count; writes at most cap taps. */
int ddc_lowpass_taps(double gain, double fs, double fc, double tw, float *out, int cap) {
  if (!(fs > 0.0) || !(fc > 0.0) || fc > fs / 2 || !(tw > 0)) /* firfilter.cpp:100-112 */
    return -1;
  int ntaps = (int)(53.0 * fs / (22.0 * tw)); /* Hamming: 53 dB (firfilter.cpp:91-99,118-120) */
  if ((ntaps & 1) == 0)
    ntaps++;
  if (ntaps > cap)
    return ntaps;
  float *w = (float *)malloc(sizeof(float) * (size_t)ntaps);
  float Mw = (float)(ntaps - 1);
  for (int n = 0; n < ntaps; n++)
    w[n] = (float)(0.54 - 0.46 * cos((2 * M_PI * n) / Mw)); /* firfilter.cpp:186-193 */
  int M = (ntaps - 1) / 2;
  double fwT0 = 2 * M_PI * fc / fs;
  for (int n = -M; n <= M; n++) {
    if (n == 0)
      out[n + M] = (float)(fwT0 / M_PI * w[n + M]);
    else
      out[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
  }
  double fmax = out[M];
  for (int n = 1; n <= M; n++)
    fmax += 2 * out[n + M];
  gain /= fmax;
  for (int i = 0; i < ntaps; i++)
    out[i] = (float)(out[i] * gain);
  free(w);
  return ntaps;
}

/* FIRHilbert::FIRHilbert coefficient table (dsp.cpp:181-215). `fs` is the integer the
 * reference passes (samplesOut, vfo.cpp:112); it cancels up to rounding in the normalisation.
 * Note the float sum of squares and the float square root (std::sqrt(float) overload). */
void ddc_hilbert_taps(int len, int fs, float *out) {
  float *tmp = (float *)malloc(sizeof(float) * (size_t)len);
  float sumsq = 0;
  for (int n = 0; n < len; n++) {
    if (n == len / 2)
      tmp[n] = 0;
    else
      tmp[n] = (float)(fs / (M_PI * (n - len / 2)) * (1 - cos(M_PI * (n - len / 2))));
    float sq = tmp[n] * tmp[n];
    sumsq = sumsq + sq;
  }
  double g = (double)sqrtf(sumsq);
  for (int i = 0; i < len; i++)
    out[i] = (float)(tmp[len - i - 1] / g);
  free(tmp);
}

/* Oscillator::Oscillator table (oscillator.cpp:4-28): q[i], i < L = (int)Fs, interleaved re,im.
 * Complex multiply as GCC emits it for std::complex<float>: (a*c - b*d, a*d + b*c). */
void ddc_nco_rotation(double fs, double f, float *rot_re, float *rot_im) {
  double ang = 2.0 * M_PI * f / fs;
  *rot_re = (float)cos(ang);
  *rot_im = (float)sin(ang);
}
static inline void nco_step(float *re, float *im, float c, float d) {
  float a = *re, b = *im;
  float ac = a * c, bd = b * d, ad = a * d, bc = b * c;
  float nr = ac - bd;
  float ni = ad + bc;
  float r2 = nr * nr, i2 = ni * ni;
  float s = r2 + i2;
  float norm = 1.95f - s;
  *re = nr * norm;
  *im = ni * norm;
}
void ddc_nco_table(double fs, double f, float *q /* 2*L */) {
  float c, d;
  ddc_nco_rotation(fs, f, &c, &d);
  int L = (int)fs;
  float re = 1.0f, im = 0.0f;
  for (int i = 0; i < L; i++) {
    nco_step(&re, &im, c, d);
    q[2 * i] = re;
    q[2 * i + 1] = im;
  }
}

/* Raw sample formats -> interleaved float I,Q. cf32 is what the reference receives from
 * SoapySDR (publisher.cpp:254); cu8 and cs16 are this project's file-source additions
 * (SURVEY.md section 8d): (u8 - 127.4f) / 128.0f and s16 / 32768.0f. */
void ddc_unpack(int fmt, const void *raw, long long n_complex, float *out) {
  long long n = 2 * n_complex;
  if (fmt == 0) {
    const uint8_t *p = (const uint8_t *)raw;
    for (long long i = 0; i < n; i++) {
      float v = (float)p[i] - 127.4f;
      out[i] = v / 128.0f;
    }
  } else if (fmt == 1) {
    const int16_t *p = (const int16_t *)raw;
    for (long long i = 0; i < n; i++)
      out[i] = (float)p[i] / 32768.0f;
  } else {
    memcpy(out, raw, sizeof(float) * (size_t)n);
  }
}

/* Publisher::demodData's optional DC removal (publisher.cpp:292-296), in place on interleaved float I,Q:
 *   avept = avept * (1.0f - 0.000001f) + 0.000001f * curr;  curr -= avept;   (std::complex<float>: per rail)
 * state[2] is the running average (function-static in the reference). Pinned: the unmodified publisher.cpp
 * (oracle/_ref/ref_publish) with --enable-dcc yields byte-identical payloads to this function followed by the
 * chain (tests/test_reference_publisher.py). */
void ddc_dc_correct(float *iq, long long n_complex, float *state) {
  const float k = 1.0f - 0.000001f, c = 0.000001f;
  for (long long i = 0; i < n_complex; i++)
    for (int r = 0; r < 2; r++) {
      float x = iq[2 * i + r];
      float t0 = state[r] * k;
      float t1 = c * x;
      state[r] = t0 + t1;
      iq[2 * i + r] = x - state[r];
    }
}

/* ------------------------------------------------------------------------------------------
 * One VFO
 * ---------------------------------------------------------------------------------------- */
typedef struct ddc_oracle {
  int Fs, B, D, L; /* input rate, block length, half-band stages, late decimation (0,5,6) */
  int demod_usb, cstyle, scalecomp;
  float gain;
  int filter_bw;
  float rot_re, rot_im;
  int nco_len;
  float *nco; /* q table, 2*L floats */
  long long n_abs; /* absolute index of the next input sample */
  long long blocks;
  /* stage streams of the current block and the last HB_TAPS samples of each from the
   * previous block (all interleaved re,im) */
  float *stage[MAX_STAGES + 1];
  int stage_len[MAX_STAGES + 1];
  float prev_tail[MAX_STAGES][2 * HB_TAPS];
  /* tail: continuous streams with enough history in front */
  int ntl;
  float *tl; /* late FIR taps */
  int nu;
  float *tu; /* fir_usb taps */
  float hil[HILBERT_LEN];
  float *hist_d;
  int hist_d_len; /* stage-D history (late FIR): ntl complex samples */
  float *hist_m;
  int hist_m_len; /* post-late-FIR stream history: HILBERT_LEN-1 complex samples */
  float *hist_u;
  int hist_u_len; /* usb stream history: nu real samples */
  int out_rate, n_out;
} ddc_oracle;

static const float hb_p0 = 0.0060431029837374152f; /* halfbanddecimator.h:84-87 */
static const float hb_p2 = -0.049372515458761493f;
static const float hb_p4 = 0.29332944952052842f;
static const float hb_p5 = 0.5f;

void ddc_oracle_destroy(ddc_oracle *o) {
  if (!o)
    return;
  free(o->nco);
  for (int s = 0; s <= MAX_STAGES; s++)
    free(o->stage[s]);
  free(o->tl);
  free(o->tu);
  free(o->hist_d);
  free(o->hist_m);
  free(o->hist_u);
  free(o);
}

/* Parameters follow the vfo setters + init (vfo.h:16-40, vfo.cpp:57-139). Returns NULL if the
 * block contract of SURVEY.md section 8b is violated. */
ddc_oracle *ddc_oracle_create(int Fs, int B, int D, int L, double mixer_freq, float gain,
                              int filter_bw, int demod_usb, int cstyle, int scalecomp) {
  if (D < 0 || D > MAX_STAGES || B <= 0 || (B % (1 << D)) != 0 || L < 0 || L == 1)
    return NULL;
  int nD = B >> D;
  if (demod_usb && L > 0 && (nD % L) != 0)
    return NULL;
  if (D > 0 && (B >> (D - 1)) < HB_TAPS) /* the boundary rule needs 11 samples of the previous block */
    return NULL;
  ddc_oracle *o = (ddc_oracle *)calloc(1, sizeof(ddc_oracle));
  o->Fs = Fs, o->B = B, o->D = D, o->L = (demod_usb ? L : 0);
  o->demod_usb = demod_usb, o->cstyle = cstyle, o->scalecomp = scalecomp > 0 ? scalecomp : 1;
  o->gain = gain;
  o->filter_bw = filter_bw;
  o->nco_len = (int)(double)Fs;
  o->nco = (float *)malloc(sizeof(float) * 2 * (size_t)o->nco_len);
  ddc_nco_rotation((double)Fs, mixer_freq, &o->rot_re, &o->rot_im);
  ddc_nco_table((double)Fs, mixer_freq, o->nco);
  for (int s = 0; s <= D; s++) {
    o->stage_len[s] = B >> s;
    o->stage[s] = (float *)calloc(2 * (size_t)(B >> s), sizeof(float));
  }
  /* vfo.cpp:62-79 */
  int targetRate = (int)(Fs / pow(2, D));
  int samplesOut = (int)(B / pow(2, D));
  if (o->L > 0) {
    targetRate = targetRate / o->L;
    samplesOut = samplesOut / o->L;
    float tmp[4096];
    int n = ddc_lowpass_taps(2, targetRate * o->L, targetRate / 2,
                             (double)targetRate / (o->L - 1), tmp, 4096);
    if (n <= 0 || n > 4096) {
      ddc_oracle_destroy(o);
      return NULL;
    }
    o->ntl = n;
    o->tl = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(o->tl, tmp, sizeof(float) * (size_t)n);
  }
  o->out_rate = targetRate;
  o->n_out = samplesOut;
  if (filter_bw > 0 && demod_usb) { /* vfo.cpp:92-102 */
    float tmp[8192];
    int n = ddc_lowpass_taps(2, targetRate, filter_bw, (double)filter_bw / 4, tmp, 8192);
    if (n <= 0 || n > 8192) {
      ddc_oracle_destroy(o);
      return NULL;
    }
    o->nu = n;
    o->tu = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(o->tu, tmp, sizeof(float) * (size_t)n);
  }
  ddc_hilbert_taps(HILBERT_LEN, samplesOut, o->hil); /* vfo.cpp:112 */
  o->hist_d_len = o->ntl;
  o->hist_d = (float *)calloc(2 * (size_t)(o->hist_d_len + 1), sizeof(float));
  o->hist_m_len = HILBERT_LEN - 1;
  o->hist_m = (float *)calloc(2 * (size_t)o->hist_m_len, sizeof(float));
  o->hist_u_len = o->nu;
  o->hist_u = (float *)calloc((size_t)o->hist_u_len + 1, sizeof(float));
  return o;
}

int ddc_oracle_out_rate(const ddc_oracle *o) { return o->out_rate; }
/* payload bytes per block (vfo.cpp:114-120,289-313) */
int ddc_oracle_out_bytes(const ddc_oracle *o) {
  if (o->demod_usb)
    return o->n_out * 2;
  return o->cstyle == 1 ? o->n_out : o->n_out * 2;
}
int ddc_oracle_stage(const ddc_oracle *o, int s, float *out, int cap_complex) {
  if (s < 0 || s > o->D)
    return -1;
  int n = o->stage_len[s] < cap_complex ? o->stage_len[s] : cap_complex;
  memcpy(out, o->stage[s], sizeof(float) * 2 * (size_t)n);
  return o->stage_len[s];
}

/* Half-band stage on one block.
 *   out[j] = ((p0*(w0+w10) + p2*(w2+w8)) + p4*(w4+w6)) + p5*w5,   w[t] = in[2j-10+t]
 * (dsp.cpp:141-147; the leading "0 +" of `outsum += ...` cannot change a finite value).
 * Boundary rule, from how FIRQueueBackToFront re-seeds the queue (dsp.cpp:163-172): for a
 * window index l = 2j-10+t < 0 the value is 0 in the first block (dsp.cpp:48-52) and, in later
 * blocks, the PREVIOUS block's sample at n_in - 1 + l, one older than the true predecessor:
 * the newest sample of each block never reaches the next block's history. */
static void hb_stage(const float *in, int n_in, const float *prev_tail /* last 11 of prev */,
                     int have_prev, float *out) {
  int n_out = n_in / 2;
  for (int j = 0; j < n_out; j++) {
    float wr[HB_TAPS], wi[HB_TAPS];
    for (int t = 0; t < HB_TAPS; t++) {
      int l = 2 * j - 10 + t;
      if (l >= 0) {
        wr[t] = in[2 * l];
        wi[t] = in[2 * l + 1];
      } else if (!have_prev) {
        wr[t] = 0.0f;
        wi[t] = 0.0f;
      } else {
        int k = HB_TAPS - 1 + l; /* prev index n_in-1+l, as an index into the last 11 */
        wr[t] = prev_tail[2 * k];
        wi[t] = prev_tail[2 * k + 1];
      }
    }
    for (int c = 0; c < 2; c++) {
      const float *w = c ? wi : wr;
      float s0 = w[0] + w[10];
      float s2 = w[2] + w[8];
      float s4 = w[4] + w[6];
      float m0 = hb_p0 * s0;
      float m2 = hb_p2 * s2;
      float m4 = hb_p4 * s4;
      float m5 = hb_p5 * w[5];
      float a = m0 + m2;
      a = a + m4;
      a = a + m5;
      float y = 0.0f;
      y = y + a;
      out[2 * j + c] = y;
    }
  }
}

/* double -> short the way GCC/x86-64 does it for in-range values (truncate toward zero); the
 * out-of-range case is undefined in C++ and is pinned here, for both the
 * oracle and the GPU, to what cvttsd2si + 16-bit store produces. */
static inline int16_t to_short(double v) {
  int32_t i;
  if (!(v > -2147483649.0 && v < 2147483648.0))
    i = INT32_MIN;
  else
    i = (int32_t)v;
  return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}
static inline int8_t to_schar(float v) { /* float -> signed char via cvttss2si + 8-bit store */
  int32_t i;
  if (!(v > -2147483904.0f && v < 2147483648.0f))
    i = INT32_MIN;
  else
    i = (int32_t)v;
  return (int8_t)(uint8_t)((uint32_t)i & 0xFFu);
}

/* y = sum_{i<N} p[i] * x[i], accumulated left to right from 0 in float (dsp.cpp:64-78). */
static inline float dot_seq(const float *p, const float *x, int stride, int N) {
  float acc = 0.0f;
  for (int i = 0; i < N; i++) {
    float m = p[i] * x[(size_t)i * stride];
    acc = acc + m;
  }
  return acc;
}

/* One vfo::process call on exactly B complex samples (interleaved float). Writes the ZMQ
 * frame-3 payload to out and returns its length in bytes. */
int ddc_oracle_process(ddc_oracle *o, const float *iq, int n_complex, unsigned char *out) {
  if (n_complex != o->B)
    return -1;
  const int B = o->B, D = o->D;
  /* mix (vfo.cpp:155-161): sample n uses q[L-1] when n == 0, else q[n mod L] */
  float *x0 = o->stage[0];
  for (int i = 0; i < B; i++) {
    long long n = o->n_abs + i;
    long long idx = (n == 0) ? (o->nco_len - 1) : (n % o->nco_len);
    float a = o->nco[2 * idx], b = o->nco[2 * idx + 1];
    float c = iq[2 * i], d = iq[2 * i + 1];
    float ac = a * c, bd = b * d, ad = a * d, bc = b * c;
    x0[2 * i] = ac - bd;
    x0[2 * i + 1] = ad + bc;
  }
  o->n_abs += B;
  /* half-band cascade (vfo.cpp:163-165) */
  for (int s = 0; s < D; s++) {
    int n_in = o->stage_len[s];
    hb_stage(o->stage[s], n_in, o->prev_tail[s], o->blocks > 0, o->stage[s + 1]);
    /* keep this block's last 11 stage-s samples for the next block's boundary rule */
    memcpy(o->prev_tail[s], o->stage[s] + 2 * (size_t)(n_in - HB_TAPS), sizeof(float) * 2 * HB_TAPS);
  }
  o->blocks++;
  const float *xd = o->stage[D];
  const int nD = o->stage_len[D];

  if (!o->demod_usb) { /* vfo::compress (vfo.cpp:260-287) */
    if (o->cstyle == 1) {
      for (int i = 0; i < nD; i++) {
        float r = xd[2 * i] / (float)o->scalecomp;
        r = r * 128.0f;
        float q = xd[2 * i + 1] / (float)o->scalecomp;
        q = q * 128.0f;
        int re = to_schar(r), im = to_schar(q);
        out[i] = (unsigned char)((re & 0xF0) | ((im & 0xF0) >> 4));
      }
      return nD;
    }
    for (int i = 0; i < nD; i++) {
      float r = xd[2 * i] * 128.0f;
      float q = xd[2 * i + 1] * 128.0f;
      out[2 * i] = (unsigned char)to_schar(r);
      out[2 * i + 1] = (unsigned char)to_schar(q);
    }
    return 2 * nD;
  }

  /* ---- USB demodulation tail (vfo.cpp:188-258) ---- */
  /* (1) late decimation: every L-th stage-D sample (phase restarts per block, vfo.cpp:218-220;
   * nD % L == 0 is enforced so it is continuous) goes through fir_decI/Q:
   *   m[k] = sum_{i<T} tl[i] * xD[n - T + i],  n = k*L  -- newest sample excluded (dsp.cpp:64-78) */
  int n_m = o->L > 0 ? nD / o->L : nD;
  float *m = (float *)malloc(sizeof(float) * 2 * (size_t)(o->hist_m_len + n_m));
  memcpy(m, o->hist_m, sizeof(float) * 2 * (size_t)o->hist_m_len);
  float *mm = m + 2 * (size_t)o->hist_m_len;
  if (o->L > 0) {
    int T = o->ntl;
    float *ext = (float *)malloc(sizeof(float) * 2 * (size_t)(T + nD));
    memcpy(ext, o->hist_d, sizeof(float) * 2 * (size_t)T);
    memcpy(ext + 2 * (size_t)T, xd, sizeof(float) * 2 * (size_t)nD);
    for (int k = 0; k < n_m; k++) {
      const float *w = ext + 2 * (size_t)(k * o->L); /* = xD[n-T], n = k*L, in ext coordinates */
      mm[2 * k] = dot_seq(o->tl, w, 2, T);
      mm[2 * k + 1] = dot_seq(o->tl, w + 1, 2, T);
    }
    memcpy(o->hist_d, ext + 2 * (size_t)nD, sizeof(float) * 2 * (size_t)T);
    free(ext);
  } else {
    memcpy(mm, xd, sizeof(float) * 2 * (size_t)nD);
  }
  /* (2) usb[k] = I[k-62] - sum_{i<125} hil[i] * Q[k-124+i]  (newest included, dsp.cpp:216-231;
   * DelayThing of length 62, dsp.h:74-96); float - double -> float equals the float subtraction. */
  float *u = (float *)malloc(sizeof(float) * (size_t)(o->hist_u_len + n_m));
  memcpy(u, o->hist_u, sizeof(float) * (size_t)o->hist_u_len);
  float *uu = u + o->hist_u_len;
  for (int k = 0; k < n_m; k++) {
    const float *wq = mm + 2 * (ptrdiff_t)(k - (HILBERT_LEN - 1)) + 1;
    float h = dot_seq(o->hil, wq, 2, HILBERT_LEN);
    float di = mm[2 * (ptrdiff_t)(k - DELAY_LEN)];
    uu[k] = (float)((double)di - (double)h);
  }
  memcpy(o->hist_m, m + 2 * (size_t)n_m, sizeof(float) * 2 * (size_t)o->hist_m_len);
  /* (3) optional fir_usb (newest excluded again), then gain and conversion */
  int16_t *o16 = (int16_t *)out;
  for (int k = 0; k < n_m; k++) {
    float v;
    if (o->nu > 0) {
      if (o->L > 0) /* vfo.cpp:236-238: filter the usb sample */
        v = dot_seq(o->tu, uu + (ptrdiff_t)(k - o->nu), 1, o->nu);
      else /* vfo.cpp:203-207: same thing */
        v = dot_seq(o->tu, uu + (ptrdiff_t)(k - o->nu), 1, o->nu);
    } else {
      v = uu[k];
    }
    float g = v * o->gain;
    o16[k] = to_short((double)g * 32768.0);
  }
  if (o->hist_u_len > 0)
    memcpy(o->hist_u, u + n_m, sizeof(float) * (size_t)o->hist_u_len);
  free(m);
  free(u);
  return 2 * n_m;
}

/* Convenience: run the same block n_blocks times, discarding output (CPU-baseline timing). */
void ddc_oracle_process_repeat(ddc_oracle *o, const float *iq, int n_complex, int n_blocks,
                               unsigned char *scratch) {
  for (int b = 0; b < n_blocks; b++)
    ddc_oracle_process(o, iq, n_complex, scratch);
}
