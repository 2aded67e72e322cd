// aero-ddc-b200: the bank object behind the C ABI of include/aeroddc.h.
//
// Host orchestration only: VFO bookkeeping, host-side filter design (design.cpp), HBM layout,
// streams/events, and the launches of the kernels in ddc_kernels.cuh / aux_kernels.cuh.
// Replaces the per-VFO objects of /root/reference/publish/vfo.cpp:57-139 (init) and the block pump
// of /root/reference/publish/publisher.cpp:285-306 (demodData -> vfo::process for every VFO).
#include "../include/aeroddc.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "ddc_kernels.cuh"
#include "design.h"

namespace {

thread_local std::string t_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}
}  // namespace
// shared with fleet.cu
int aeroddc_set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}
namespace {
#define CU(x)                                                                                            \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) return fail(AERODDC_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

using namespace aeroddc;

struct VfoRec {
  aeroddc_vfo_desc d;
  TailPlan plan;
  int fs_in, blk_in; // rate and block length of the stream this VFO mixes (the raw IQ, or its parent's output)
  int children;      // number of VFOs fed by this one (> 0: publishes nothing itself, vfo.cpp:167-172)
  int slot;          // column in the device tables (VFOs are grouped by input stream and decimation count)
  int hist;          // stage-D history samples kept in front of each block
  size_t xd_off;     // float2 offset of this VFO's row (history first) in d_xd
  size_t out_bytes;  // payload bytes per block
  size_t out_off;    // offset of the payload row in the output buffers
  size_t taps_off[3];
  std::vector<float> hil_nz;   // non-zero Hilbert taps ...
  std::vector<int> hil_nz_idx; // ... and their indices
  size_t hil_idx_off;
};

struct Group {   // VFOs sharing input stream and D: one launch of the main kernel per block
  int parent;      // -1: raw IQ of the bank; else the VFO whose stage-D stream is the input
  int D, base, count;
  int fs_in, blk_in;
  int S, W, Wb, nseg;
  int Q, P;          // parts per segment and part length (chained CTAs, see ddc_kernels.cuh)
  int* d_flags = nullptr;
  float2* d_hand = nullptr;
};

int raw_bytes(int fmt) { return fmt == AERODDC_CU8 ? 2 : (fmt == AERODDC_CS16 ? 4 : 8); }

int ilcm(int a, int b) {
  int x = a, y = b;
  while (y) { int t = x % y; x = y; y = t; }
  return a / x * b;
}

}  // namespace

struct aeroddc_bank {
  int fs = 0, B = 0, fmt = 0, device = 0;
  bool finalized = false;
  int mode = AERODDC_MODE_EXACT;
  bool dcc = false;
  float* d_dcc_out = nullptr;    // DC-corrected cf32 block
  float* d_dcc_state = nullptr;  // running average per rail
  std::vector<VfoRec> vfos;
  std::vector<Group> groups;
  int vfo_pitch = 0;
  int n_sm = 148;
  int nck = 0;

  // device memory
  float2 *d_rot = nullptr, *d_qlast = nullptr, *d_ckpt = nullptr;
  float2* d_state[2] = {nullptr, nullptr};
  float2* d_xd = nullptr;       // all stage-D rows, per VFO: [hist][n_stage]
  float2** d_xd_rows = nullptr; // [vfo_pitch] pointer to stage-D index 0 of each column's row
  int* d_nco_len = nullptr;     // [vfo_pitch]
  int* d_err = nullptr;         // chained-CTA watchdog flag: device view of h_err (zero-copy pinned host memory)
  volatile int* h_err = nullptr;
  int nck_max = 0;
  float* d_taps = nullptr;
  int* d_hil_idx = nullptr;
  TailVfo* d_tail = nullptr;
  unsigned char* d_out = nullptr;
  size_t out_total = 0;
  unsigned char* d_in[2] = {nullptr, nullptr};
  size_t in_bytes = 0;
  size_t dev_bytes = 0;

  // host memory
  unsigned char* h_in[2] = {nullptr, nullptr};
  unsigned char* h_out[3] = {nullptr, nullptr, nullptr};   // three slots: a payload stays valid until the next wait()

  cudaStream_t s_compute = nullptr, s_copy = nullptr, s_d2h = nullptr;   // kernels | H2D of raw blocks | D2H of payloads
  cudaEvent_t ev_tail[2] = {};   // payload rows of parity p written
  cudaEvent_t ev_d2h[2] = {};    // payload rows of parity p copied out (may be overwritten)
  cudaEvent_t ev_h2d[2] = {};
  cudaEvent_t ev_sw0 = nullptr, ev_sw1 = nullptr;
  cudaEvent_t ev_done[3] = {}, ev_k0[3] = {}, ev_k1[3] = {}, ev_m0[3] = {}, ev_m1[3] = {};

  long long blocks_submitted = 0, blocks_done = 0;
  int cur_out = -1;   // h_out slot returned by output()
  float last_kernel_ms = 0, last_main_ms = 0;
  int last_launches = 0;
  int launches_per_block = 0;
  size_t tail_smem = 0;
  int tail_chunks = 0;
  bool any_hist = false;
};

namespace {

template <int NF, int FMT, bool FAST>
cudaError_t launch_main_m(const MainParams& p, dim3 grid, cudaStream_t s) {
  const int smem = TileSmem<FMT>::kTotal;
  cudaError_t e = cudaFuncSetAttribute(ddc_main_kernel<NF, FMT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  ddc_main_kernel<NF, FMT, FAST><<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}
template <int NF, int FMT>
cudaError_t launch_main_t(const MainParams& p, dim3 grid, cudaStream_t s, bool fast) {
  return fast ? launch_main_m<NF, FMT, true>(p, grid, s) : launch_main_m<NF, FMT, false>(p, grid, s);
}
template <int FMT>
cudaError_t launch_main_f(int nf, const MainParams& p, dim3 grid, cudaStream_t s, bool fast) {
  switch (nf) {
    case 0: return launch_main_t<0, FMT>(p, grid, s, fast);
    case 1: return launch_main_t<1, FMT>(p, grid, s, fast);
    case 2: return launch_main_t<2, FMT>(p, grid, s, fast);
    case 3: return launch_main_t<3, FMT>(p, grid, s, fast);
    case 4: return launch_main_t<4, FMT>(p, grid, s, fast);
    default: return launch_main_t<5, FMT>(p, grid, s, fast);
  }
}
cudaError_t launch_main(int fmt, int nf, const MainParams& p, dim3 grid, cudaStream_t s, bool fast) {
  if (fmt == AERODDC_CU8) return launch_main_f<FMT_CU8>(nf, p, grid, s, fast);
  if (fmt == AERODDC_CS16) return launch_main_f<FMT_CS16>(nf, p, grid, s, fast);
  return launch_main_f<FMT_CF32>(nf, p, grid, s, fast);
}

void free_all(aeroddc_bank* b) {
  cudaSetDevice(b->device);
  if (b->s_compute) cudaStreamSynchronize(b->s_compute);
  if (b->s_copy) cudaStreamSynchronize(b->s_copy);
  if (b->s_d2h) cudaStreamSynchronize(b->s_d2h);
  cudaFree(b->d_rot); cudaFree(b->d_qlast); cudaFree(b->d_ckpt);
  cudaFree(b->d_state[0]); cudaFree(b->d_state[1]);
  for (Group& g : b->groups) { cudaFree(g.d_flags); cudaFree(g.d_hand); }
  if (b->h_err) cudaFreeHost((void*)b->h_err);
  cudaFree(b->d_dcc_out); cudaFree(b->d_dcc_state);
  cudaFree(b->d_xd); cudaFree(b->d_xd_rows); cudaFree(b->d_nco_len); cudaFree(b->d_taps); cudaFree(b->d_hil_idx); cudaFree(b->d_tail); cudaFree(b->d_out);
  cudaFree(b->d_in[0]); cudaFree(b->d_in[1]);
  for (int i = 0; i < 2; ++i) {
    if (b->h_in[i]) cudaFreeHost(b->h_in[i]);
    if (b->ev_h2d[i]) cudaEventDestroy(b->ev_h2d[i]);
  }
  for (int i = 0; i < 3; ++i) {
    if (b->h_out[i]) cudaFreeHost(b->h_out[i]);
    if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
    if (b->ev_k0[i]) cudaEventDestroy(b->ev_k0[i]);
    if (b->ev_k1[i]) cudaEventDestroy(b->ev_k1[i]);
    if (b->ev_m0[i]) cudaEventDestroy(b->ev_m0[i]);
    if (b->ev_m1[i]) cudaEventDestroy(b->ev_m1[i]);
  }
  if (b->ev_sw0) cudaEventDestroy(b->ev_sw0);
  if (b->ev_sw1) cudaEventDestroy(b->ev_sw1);
  if (b->s_compute) cudaStreamDestroy(b->s_compute);
  if (b->s_copy) cudaStreamDestroy(b->s_copy);
  if (b->s_d2h) cudaStreamDestroy(b->s_d2h);
  for (int i = 0; i < 2; ++i) { if (b->ev_tail[i]) cudaEventDestroy(b->ev_tail[i]); if (b->ev_d2h[i]) cudaEventDestroy(b->ev_d2h[i]); }
}

// enqueue everything that follows the arrival of the raw block in device memory
int enqueue_block(aeroddc_bank* b, const void* dev_iq) {
  const int slot = (int)(b->blocks_submitted % 3);   // payload/event slot
  cudaStream_t s = b->s_compute;
  const int par = (int)(b->blocks_submitted & 1);
  int launches = 0;
  CU(cudaEventRecord(b->ev_k0[slot], s));
  if (b->dcc) {   // sequential DC removal of the raw stream; the VFOs then read the corrected cf32 block
    if (b->fmt == AERODDC_CU8) dcc_kernel<0><<<1, 32, 0, s>>>(dev_iq, b->d_dcc_out, b->d_dcc_state, b->B);
    else if (b->fmt == AERODDC_CS16) dcc_kernel<1><<<1, 32, 0, s>>>(dev_iq, b->d_dcc_out, b->d_dcc_state, b->B);
    else dcc_kernel<2><<<1, 32, 0, s>>>(dev_iq, b->d_dcc_out, b->d_dcc_state, b->B);
    CU(cudaGetLastError());
    ++launches;
    dev_iq = b->d_dcc_out;
  }
  CU(cudaEventRecord(b->ev_m0[slot], s));
  for (const Group& g : b->groups) {
    MainParams p;
    // a sub-VFO group reads its parent's stage-D stream of this block (already enqueued on this stream)
    p.iq = g.parent < 0 ? dev_iq : (const void*)(b->d_xd + b->vfos[g.parent].xd_off + b->vfos[g.parent].hist);
    p.ckpt = b->d_ckpt;
    p.rot = b->d_rot;
    p.qlast = b->d_qlast;
    p.state_in = b->d_state[par];
    p.state_out = b->d_state[par ^ 1];
    p.xd_rows = b->d_xd_rows;
    p.block_abs = b->blocks_submitted * (long long)g.blk_in;
    p.vfo_pitch = b->vfo_pitch;
    p.vfo_base = g.base;
    p.vfo_count = g.count;
    p.D = g.D;
    p.B = g.blk_in;
    p.S = g.S;
    p.W = g.W;
    p.nseg = g.nseg;
    p.Wb = g.Wb;
    p.nco_len = g.fs_in;
    p.one = 1.0f;
    p.transient = 4 * kNcoStride;
    p.nck = b->nck_max;
    p.Q = g.Q;
    p.P = g.P;
    p.ngroups = (g.count + kVfoPerCta - 1) / kVfoPerCta;
    p.flags = g.d_flags;
    p.ticket = g.d_flags + (size_t)p.ngroups * g.nseg;   // the counter lives behind the flags
    p.hand = g.d_hand;
    p.err = b->d_err;
    CU(cudaMemsetAsync(g.d_flags, 0, sizeof(int) * ((size_t)p.ngroups * g.nseg + 1), s));
    dim3 grid((unsigned)(p.ngroups * (1 + g.Q * g.nseg)));
    CU(launch_main((g.parent < 0 && !b->dcc) ? b->fmt : AERODDC_CF32, std::min(g.D, kFastStages), p, grid, s, b->mode == AERODDC_MODE_FAST));
    ++launches;
  }
  CU(cudaEventRecord(b->ev_m1[slot], s));
  {
    // payload rows are double-buffered by block parity; wait until the copy-out of this parity (two blocks ago) is done
    if (b->blocks_submitted >= 2) CU(cudaStreamWaitEvent(s, b->ev_d2h[par], 0));
    dim3 grid(b->tail_chunks, (unsigned)b->vfos.size());
    tail_kernel<<<grid, kTailThreads, b->tail_smem, s>>>(b->d_tail, 1.0f, (size_t)par * b->out_total);
    CU(cudaGetLastError());
    ++launches;
  }
  CU(cudaEventRecord(b->ev_k1[slot], s));
  CU(cudaEventRecord(b->ev_tail[par], s));
  // keep the last `hist` stage-D samples of every VFO in front of the next block
  if (b->any_hist) {
    xd_shift_kernel<<<(unsigned)b->vfos.size(), 256, 0, s>>>(b->d_tail);
    CU(cudaGetLastError());
    ++launches;
  }
  // payloads leave on their own stream, overlapping the next block's kernels
  CU(cudaStreamWaitEvent(b->s_d2h, b->ev_tail[par], 0));
  CU(cudaMemcpyAsync(b->h_out[slot], b->d_out + (size_t)par * b->out_total, b->out_total, cudaMemcpyDeviceToHost, b->s_d2h));
  CU(cudaEventRecord(b->ev_d2h[par], b->s_d2h));
  CU(cudaEventRecord(b->ev_done[slot], b->s_d2h));
  b->launches_per_block = launches;
  b->blocks_submitted++;
  return AERODDC_OK;
}

}  // namespace

extern "C" {

int aeroddc_abi_version(void) { return AERODDC_ABI_VERSION; }

int aeroddc_design_lowpass(double gain, double fs, double cutoff, double transition, float* taps, int cap) {
  const std::vector<float> h = design_lowpass(gain, fs, cutoff, transition);
  if (h.empty()) return fail(AERODDC_ERR_DESIGN, "low_pass rejected fs=%g cutoff=%g transition=%g (firfilter.cpp:100-112)", fs, cutoff, transition);
  if (taps) std::copy(h.begin(), h.begin() + std::min<size_t>(h.size(), (size_t)std::max(cap, 0)), taps);
  return (int)h.size();
}
int aeroddc_design_hilbert(int len, int fs_param, float* taps, int cap) {
  if (len < 1) return fail(AERODDC_ERR_ARG, "len must be positive");
  const std::vector<float> h = design_hilbert(len, fs_param);
  if (taps) std::copy(h.begin(), h.begin() + std::min<size_t>(h.size(), (size_t)std::max(cap, 0)), taps);
  return (int)h.size();
}
int aeroddc_design_rotation(double fs, double freq, float* cos_out, float* sin_out) {
  if (!cos_out || !sin_out) return fail(AERODDC_ERR_ARG, "NULL argument");
  design_rotation(fs, freq, cos_out, sin_out);
  return AERODDC_OK;
}
const char* aeroddc_last_error(void) { return t_err.c_str(); }

int aeroddc_bank_create(aeroddc_bank** out, int sample_rate, int block_len, int in_format, int device) {
  if (!out) return fail(AERODDC_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (sample_rate <= 0 || block_len <= 0) return fail(AERODDC_ERR_ARG, "sample_rate and block_len must be positive");
  if (block_len > sample_rate)
    return fail(AERODDC_ERR_ARG, "block_len %d > sample_rate %d: the reference's half-band queue holds inlen+11 samples (dsp.cpp:43)",
                block_len, sample_rate);
  if (in_format < AERODDC_CU8 || in_format > AERODDC_CF32) return fail(AERODDC_ERR_ARG, "unknown input format %d", in_format);
  if (block_len % kChunk) return fail(AERODDC_ERR_ARG, "block_len must be a multiple of %d", kChunk);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(AERODDC_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(AERODDC_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  aeroddc_bank* b = new (std::nothrow) aeroddc_bank();
  if (!b) return fail(AERODDC_ERR_NOMEM, "out of host memory");
  b->fs = sample_rate;
  b->B = block_len;
  b->fmt = in_format;
  b->device = device;
  *out = b;
  return AERODDC_OK;
}

int aeroddc_plan_segments(int block_len, int decim_count, int n_vfos, int n_sm, double waves, int parts, aeroddc_segment_plan* out) {
  if (!out || block_len <= 0 || decim_count < 0 || decim_count > kMaxStages || n_vfos < 1 || n_sm < 1 || !(waves > 0))
    return fail(AERODDC_ERR_ARG, "bad planning arguments");
  const int D = decim_count;
  const int cstep = ilcm(kChunk, 1 << D);
  if (block_len % cstep) return fail(AERODDC_ERR_ARG, "block of %d samples is not a multiple of %d", block_len, cstep);
  const int align = ilcm(kNcoStride, 1 << D);
  out->warmup = D == 0 ? 0 : ((10 << D) + cstep - 1) / cstep * cstep;   // 10*(2^D - 1) samples reach the deepest stage's history
  out->boundary_warmup = D == 0 ? 0 : (11 << D);                        // the next block's shifted history needs 11 samples per stage
  const int groups = (n_vfos + kVfoPerCta - 1) / kVfoPerCta;
  // One wave of CTAs (kCtasPerSm per SM) cuts the block into segments, each paying one warm-up of W samples.
  // The warp scheduler favours some resident warps, so equal CTAs of a single wave finish at different times;
  // each segment is therefore processed as Q chained parts by Q short CTAs (state handed over through HBM),
  // which evens the load without further warm-ups.
  const int target = std::max(1, (int)std::floor((double)kCtasPerSm * n_sm * waves / groups) - 1);   // -1: the boundary CTA
  int S = (block_len + target - 1) / target;
  S = std::max(S, std::max(4 * out->warmup, 4096));
  S = (S + align - 1) / align * align;
  out->segment_len = S;
  out->n_segments = (block_len + S - 1) / S;
  int Q = parts > 0 ? parts : 16;
  const int pal = ilcm(kTile, 1 << D);
  while (Q > 1 && S / Q < 8 * pal) --Q;   // keep parts long enough that the hand-over stays negligible
  Q = std::min(Q, 60);
  out->parts = Q;
  out->part_len = ((S + Q - 1) / Q + pal - 1) / pal * pal;
  out->vfo_groups = groups;
  out->ctas = groups * (1 + Q * out->n_segments);
  return AERODDC_OK;
}

int aeroddc_bank_set_mode(aeroddc_bank* b, int mode) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->blocks_submitted > 0) return fail(AERODDC_ERR_STATE, "the arithmetic mode cannot change once blocks were processed");
  if (mode != AERODDC_MODE_EXACT && mode != AERODDC_MODE_FAST) return fail(AERODDC_ERR_ARG, "unknown mode %d", mode);
  b->mode = mode;
  return AERODDC_OK;
}

int aeroddc_bank_set_dc_correction(aeroddc_bank* b, int enable) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->finalized) return fail(AERODDC_ERR_STATE, "set DC correction before finalize");
  b->dcc = enable != 0;
  return AERODDC_OK;
}

int aeroddc_bank_add_vfo(aeroddc_bank* b, const aeroddc_vfo_desc* d) {
  if (!b || !d) return fail(AERODDC_ERR_ARG, "NULL argument");
  if (b->finalized) return fail(AERODDC_ERR_STATE, "bank already finalized");
  if (d->decim_count < 0 || d->decim_count > kMaxStages)
    return fail(AERODDC_ERR_ARG, "decim_count %d outside 0..8 (vfo.h:63)", d->decim_count);
  VfoRec r;
  r.d = *d;
  r.d.topic[sizeof r.d.topic - 1] = 0;
  r.children = 0;
  if (d->parent >= (int)b->vfos.size() || d->parent < -1) return fail(AERODDC_ERR_ARG, "parent %d is not an earlier VFO", d->parent);
  if (d->parent >= 0) {
    const VfoRec& pr = b->vfos[d->parent];
    if (pr.d.parent >= 0) return fail(AERODDC_ERR_ARG, "only one level of main -> sub VFOs (publisher.cpp:118-219)");
    r.fs_in = (int)(pr.fs_in / std::pow(2.0, pr.d.decim_count));   // vfo::getOutRate (vfo.cpp:147)
    r.blk_in = pr.plan.n_stage;
  } else {
    r.fs_in = b->fs;
    r.blk_in = b->B;
  }
  const int step = ilcm(kChunk, 1 << d->decim_count);
  if (r.blk_in % step) return fail(AERODDC_ERR_ARG, "input block %d not a multiple of %d for D=%d", r.blk_in, step, d->decim_count);
  if (d->decim_count > 0 && r.blk_in < 20 * (1 << d->decim_count))
    return fail(AERODDC_ERR_ARG, "input block %d too short for D=%d (need >= %d)", r.blk_in, d->decim_count, 20 << d->decim_count);
  if (r.blk_in > r.fs_in) return fail(AERODDC_ERR_ARG, "input block %d longer than its rate %d (dsp.cpp:43)", r.blk_in, r.fs_in);
  const int late = d->demod_usb ? d->late_decimate : 0;
  if (late < 0 || late == 1) return fail(AERODDC_ERR_ARG, "late_decimate must be 0 or >= 2");
  const int n_stage = r.blk_in >> d->decim_count;
  if (late > 0 && n_stage % late) return fail(AERODDC_ERR_ARG, "stage-D block %d not divisible by late_decimate %d", n_stage, late);
  if (r.d.scale_comp <= 0) r.d.scale_comp = 1;
  if (!plan_tail(r.fs_in, r.blk_in, d->decim_count, late, d->filter_bw, d->demod_usb != 0, &r.plan))
    return fail(AERODDC_ERR_DESIGN, "filter design rejected (fs=%d D=%d late=%d bw=%d)", r.fs_in, d->decim_count, late, d->filter_bw);
  if (r.plan.n_out < 1) return fail(AERODDC_ERR_ARG, "no output samples per block");
  const int T = (int)r.plan.late_taps.size(), U = (int)r.plan.usb_taps.size();
  if (U > (late > 0 ? 1024 : 4096)) return fail(AERODDC_ERR_ARG, "fir_usb with %d taps is not supported (max %d)", U, late > 0 ? 1024 : 4096);
  r.hist = d->demod_usb ? (U + kHilbert - 1) * std::max(late, 1) + T : 0;
  r.hist = (r.hist + 1) & ~1;
  r.out_bytes = d->demod_usb ? (size_t)r.plan.n_out * 2 : (d->compress_style == 1 ? (size_t)r.plan.n_out : (size_t)r.plan.n_out * 2);
  b->vfos.push_back(r);
  if (d->parent >= 0) b->vfos[d->parent].children++;
  return (int)b->vfos.size() - 1;
}

int aeroddc_bank_finalize(aeroddc_bank* b) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->finalized) return fail(AERODDC_ERR_STATE, "already finalized");
  if (b->vfos.empty()) return fail(AERODDC_ERR_STATE, "no VFOs");
  CU(cudaSetDevice(b->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, b->device));
  if (prop.major < 10) return fail(AERODDC_ERR_CUDA, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  b->n_sm = prop.multiProcessorCount;

  // ---- group VFOs by (input stream, D); raw-fed groups first so that parents run before children ----
  const int nv = (int)b->vfos.size();
  int col = 0;
  for (int parent = -1; parent < nv; ++parent) {
    if (parent >= 0 && b->vfos[parent].children == 0) continue;
    for (int D = 0; D <= kMaxStages; ++D) {
      Group g;
      g.parent = parent; g.D = D; g.base = col; g.count = 0;
      for (int i = 0; i < nv; ++i)
        if (b->vfos[i].d.parent == parent && b->vfos[i].d.decim_count == D) {
          b->vfos[i].slot = col++; g.count++;
          g.fs_in = b->vfos[i].fs_in; g.blk_in = b->vfos[i].blk_in;
        }
      if (g.count) b->groups.push_back(g);
    }
  }
  b->vfo_pitch = (col + 3) & ~3;
  const char* env_waves = getenv("AERODDC_WAVES");
  const double waves = env_waves ? std::max(0.05, atof(env_waves)) : 1.0;
  const char* env_parts = getenv("AERODDC_PARTS");
  const int env_parts_n = env_parts ? std::max(1, atoi(env_parts)) : 0;
  for (Group& g : b->groups) {
    aeroddc_segment_plan pl;
    const int rc = aeroddc_plan_segments(g.blk_in, g.D, g.count, b->n_sm, waves, env_parts_n, &pl);
    if (rc != AERODDC_OK) return rc;
    g.W = pl.warmup; g.Wb = pl.boundary_warmup; g.S = pl.segment_len; g.nseg = pl.n_segments; g.Q = pl.parts; g.P = pl.part_len;
  }

  for (Group& g : b->groups) {
    const size_t nk = (size_t)((g.count + kVfoPerCta - 1) / kVfoPerCta) * g.nseg;
    CU(cudaMalloc((void**)&g.d_flags, sizeof(int) * (nk + 1)));   // + the ticket counter
    CU(cudaMemset(g.d_flags, 0, sizeof(int) * (nk + 1)));
    CU(cudaMalloc((void**)&g.d_hand, sizeof(float2) * nk * kHandSlots * kThreads));
  }
  CU(cudaHostAlloc((void**)&b->h_err, sizeof(int), cudaHostAllocMapped));
  *b->h_err = 0;
  CU(cudaHostGetDevicePointer((void**)&b->d_err, (void*)b->h_err, 0));

  // ---- constant tables ----
  std::vector<float2> h_rot(b->vfo_pitch, make_float2(1.0f, 0.0f));
  std::vector<int> h_len(b->vfo_pitch, 1);
  b->nck_max = 1;
  for (const VfoRec& r : b->vfos) {
    design_rotation((double)r.fs_in, r.d.mixer_freq, &h_rot[r.slot].x, &h_rot[r.slot].y);
    h_len[r.slot] = r.fs_in;
    b->nck_max = std::max(b->nck_max, (r.fs_in + kNcoStride - 1) / kNcoStride);
  }
  size_t bytes = 0;
  auto dmalloc = [&](void** p, size_t n) { bytes += n; return cudaMalloc(p, n); };
  CU(dmalloc((void**)&b->d_rot, sizeof(float2) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_qlast, sizeof(float2) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_nco_len, sizeof(int) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_ckpt, sizeof(float2) * (size_t)b->nck_max * b->vfo_pitch));
  CU(cudaMemset(b->d_ckpt, 0, sizeof(float2) * (size_t)b->nck_max * b->vfo_pitch));
  CU(cudaMemset(b->d_qlast, 0, sizeof(float2) * b->vfo_pitch));
  CU(cudaMemcpy(b->d_rot, h_rot.data(), sizeof(float2) * b->vfo_pitch, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b->d_nco_len, h_len.data(), sizeof(int) * b->vfo_pitch, cudaMemcpyHostToDevice));
  for (int i = 0; i < 2; ++i) {
    const size_t n = sizeof(float2) * (size_t)kMaxStages * kStateSlots * b->vfo_pitch;
    CU(dmalloc((void**)&b->d_state[i], n));
    CU(cudaMemset(b->d_state[i], 0, n));   // first block: all-zero history (dsp.cpp:48-52)
  }

  // ---- stage-D rows and payload rows ----
  size_t taps_total = 0, out_off = 0, xd_total = 0, idx_total = 0;
  for (VfoRec& r : b->vfos) {
    r.xd_off = xd_total;
    xd_total += ((size_t)r.hist + r.plan.n_stage + 1) & ~(size_t)1;
    if (r.hist > 0) b->any_hist = true;
    r.taps_off[0] = taps_total; taps_total += r.plan.late_taps.size();
    r.taps_off[1] = taps_total; taps_total += r.plan.usb_taps.size();
    r.hil_nz.clear(); r.hil_nz_idx.clear();
    for (size_t i = 0; i < r.plan.hilbert_taps.size(); ++i)
      if (r.plan.hilbert_taps[i] != 0.0f) { r.hil_nz.push_back(r.plan.hilbert_taps[i]); r.hil_nz_idx.push_back((int)i); }
    r.taps_off[2] = taps_total; taps_total += r.hil_nz.size();
    r.hil_idx_off = idx_total; idx_total += r.hil_nz.size();
    r.out_off = out_off;
    if (r.children) r.out_bytes = 0;   // a main VFO with sub-VFOs only feeds them (vfo.cpp:167-172)
    out_off += (r.out_bytes + 15) & ~(size_t)15;
  }
  b->out_total = std::max<size_t>(out_off, 16);
  CU(dmalloc((void**)&b->d_xd, sizeof(float2) * xd_total));
  CU(cudaMemset(b->d_xd, 0, sizeof(float2) * xd_total));
  std::vector<float2*> h_rows(b->vfo_pitch, b->d_xd);
  for (const VfoRec& r : b->vfos) h_rows[r.slot] = b->d_xd + r.xd_off + r.hist;
  CU(dmalloc((void**)&b->d_xd_rows, sizeof(float2*) * b->vfo_pitch));
  CU(cudaMemcpy(b->d_xd_rows, h_rows.data(), sizeof(float2*) * b->vfo_pitch, cudaMemcpyHostToDevice));
  std::vector<float> h_taps(std::max<size_t>(taps_total, 1));
  std::vector<int> h_idx(std::max<size_t>(idx_total, 1));
  for (const VfoRec& r : b->vfos) {
    std::copy(r.plan.late_taps.begin(), r.plan.late_taps.end(), h_taps.begin() + r.taps_off[0]);
    std::copy(r.plan.usb_taps.begin(), r.plan.usb_taps.end(), h_taps.begin() + r.taps_off[1]);
    std::copy(r.hil_nz.begin(), r.hil_nz.end(), h_taps.begin() + r.taps_off[2]);
    std::copy(r.hil_nz_idx.begin(), r.hil_nz_idx.end(), h_idx.begin() + r.hil_idx_off);
  }
  CU(dmalloc((void**)&b->d_hil_idx, sizeof(int) * h_idx.size()));
  CU(cudaMemcpy(b->d_hil_idx, h_idx.data(), sizeof(int) * h_idx.size(), cudaMemcpyHostToDevice));
  CU(dmalloc((void**)&b->d_taps, sizeof(float) * h_taps.size()));
  CU(cudaMemcpy(b->d_taps, h_taps.data(), sizeof(float) * h_taps.size(), cudaMemcpyHostToDevice));
  CU(dmalloc((void**)&b->d_out, 2 * b->out_total));   // two parities: the tail of block k+1 runs while block k's payloads copy out
  CU(cudaMemset(b->d_out, 0, 2 * b->out_total));
  std::vector<TailVfo> h_tail(nv);
  size_t tail_smem = 0;
  int max_out = 0;
  for (int i = 0; i < nv; ++i) {
    const VfoRec& r = b->vfos[i];
    TailVfo& t = h_tail[i];
    t.xd = b->d_xd + r.xd_off + r.hist;
    t.out = b->d_out + r.out_off;
    t.late_taps = b->d_taps + r.taps_off[0];
    t.usb_taps = b->d_taps + r.taps_off[1];
    t.hil_taps = b->d_taps + r.taps_off[2];
    t.hil_idx = b->d_hil_idx + r.hil_idx_off;
    t.n_hil = (int)r.hil_nz.size();
    t.hil_regular = 1;
    for (size_t j = 0; j < r.hil_nz_idx.size(); ++j) if (r.hil_nz_idx[j] != (int)(2 * j + 1)) t.hil_regular = 0;
    t.n_stage = r.plan.n_stage;
    t.n_out = r.children ? 0 : r.plan.n_out;
    t.hist = r.hist;
    t.late = r.plan.late;
    t.T = (int)r.plan.late_taps.size();
    t.U = (int)r.plan.usb_taps.size();
    t.demod_usb = r.d.demod_usb != 0;
    t.cstyle = r.d.compress_style;
    t.scalecomp = r.d.scale_comp;
    t.gain = r.d.gain;
    max_out = std::max(max_out, t.n_out);
    const size_t n_m = kHilbert - 1 + t.U + kTailChunk;
    const size_t sm = sizeof(float) * (2 * n_m + t.U + kTailChunk + t.T + t.U + 2 * kHilbert + 2) +
                      (t.late > 0 ? sizeof(float2) * ((n_m - 1) * t.late + t.T) : 0) + 16;
    tail_smem = std::max(tail_smem, sm);
  }
  b->tail_smem = tail_smem;
  b->tail_chunks = std::max(1, (max_out + kTailChunk - 1) / kTailChunk);
  CU(cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem));
  CU(dmalloc((void**)&b->d_tail, sizeof(TailVfo) * nv));
  CU(cudaMemcpy(b->d_tail, h_tail.data(), sizeof(TailVfo) * nv, cudaMemcpyHostToDevice));

  if (b->dcc) {
    CU(dmalloc((void**)&b->d_dcc_out, sizeof(float) * 2 * (size_t)b->B));
    CU(dmalloc((void**)&b->d_dcc_state, sizeof(float) * 2));
    CU(cudaMemset(b->d_dcc_state, 0, sizeof(float) * 2));
  }

  // ---- input staging ----
  b->in_bytes = (size_t)b->B * raw_bytes(b->fmt);
  for (int i = 0; i < 2; ++i) {
    CU(dmalloc((void**)&b->d_in[i], b->in_bytes));
    CU(cudaHostAlloc((void**)&b->h_in[i], b->in_bytes, cudaHostAllocDefault));
    CU(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 3; ++i) {
    CU(cudaHostAlloc((void**)&b->h_out[i], std::max<size_t>(b->out_total, 16), cudaHostAllocDefault));
    memset(b->h_out[i], 0, std::max<size_t>(b->out_total, 16));
    CU(cudaEventCreate(&b->ev_k0[i])); CU(cudaEventCreate(&b->ev_k1[i]));
    CU(cudaEventCreate(&b->ev_m0[i])); CU(cudaEventCreate(&b->ev_m1[i]));
    CU(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreate(&b->ev_sw0));
  CU(cudaEventCreate(&b->ev_sw1));
  CU(cudaStreamCreateWithFlags(&b->s_compute, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&b->s_copy, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&b->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU(cudaEventCreateWithFlags(&b->ev_tail[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_d2h[i], cudaEventDisableTiming));
  }

  // ---- NCO checkpoints: the exact sequential recurrence, one thread per VFO ----
  {
    const int threads = 32;
    nco_checkpoint_kernel<<<(b->vfo_pitch + threads - 1) / threads, threads, 0, b->s_compute>>>(
        b->d_rot, b->d_nco_len, b->d_ckpt, b->d_qlast, b->vfo_pitch, b->vfo_pitch, kNcoStride);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(b->s_compute));
  }
  b->dev_bytes = bytes;
  b->finalized = true;
  return AERODDC_OK;
}

int aeroddc_bank_submit_device(aeroddc_bank* b, const void* dev_iq, size_t n_complex, void* ready_event) {
  if (!b || !dev_iq) return fail(AERODDC_ERR_ARG, "NULL argument");
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  if (n_complex != (size_t)b->B)
    return fail(AERODDC_ERR_ARG, "block of %zu samples, bank was created for %d (vfo.cpp:155,164 block contract)", n_complex, b->B);
  if (b->blocks_submitted - b->blocks_done >= 2) return fail(AERODDC_ERR_STATE, "two blocks already in flight; call wait()");
  if ((uintptr_t)dev_iq & 15) return fail(AERODDC_ERR_ARG, "device block must be 16-byte aligned");
  CU(cudaSetDevice(b->device));
  if (ready_event) CU(cudaStreamWaitEvent(b->s_compute, (cudaEvent_t)ready_event, 0));
  return enqueue_block(b, dev_iq);
}

int aeroddc_bank_submit(aeroddc_bank* b, const void* host_iq, size_t n_complex) {
  if (!b || !host_iq) return fail(AERODDC_ERR_ARG, "NULL argument");
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  if (n_complex != (size_t)b->B)
    return fail(AERODDC_ERR_ARG, "block of %zu samples, bank was created for %d (vfo.cpp:155,164 block contract)", n_complex, b->B);
  if (b->blocks_submitted - b->blocks_done >= 2) return fail(AERODDC_ERR_STATE, "two blocks already in flight; call wait()");
  CU(cudaSetDevice(b->device));
  const int slot = (int)(b->blocks_submitted & 1);
  const void* src = host_iq;
  if (host_iq != b->h_in[0] && host_iq != b->h_in[1]) {   // pageable caller memory: stage through the pinned ring
    memcpy(b->h_in[slot], host_iq, b->in_bytes);
    src = b->h_in[slot];
  }
  // d_in[slot] was last read by the block submitted two calls ago, which wait() has retired
  CU(cudaMemcpyAsync(b->d_in[slot], src, b->in_bytes, cudaMemcpyHostToDevice, b->s_copy));
  CU(cudaEventRecord(b->ev_h2d[slot], b->s_copy));
  CU(cudaStreamWaitEvent(b->s_compute, b->ev_h2d[slot], 0));
  return enqueue_block(b, b->d_in[slot]);
}

int aeroddc_bank_wait(aeroddc_bank* b) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->blocks_done >= b->blocks_submitted) return fail(AERODDC_ERR_STATE, "nothing in flight");
  CU(cudaSetDevice(b->device));
  const int slot = (int)(b->blocks_done % 3);
  CU(cudaEventSynchronize(b->ev_done[slot]));
  CU(cudaEventElapsedTime(&b->last_kernel_ms, b->ev_k0[slot], b->ev_k1[slot]));
  CU(cudaEventElapsedTime(&b->last_main_ms, b->ev_m0[slot], b->ev_m1[slot]));
  b->last_launches = b->launches_per_block;
  b->cur_out = slot;
  b->blocks_done++;
  if (*b->h_err) return fail(AERODDC_ERR_CUDA, "a chained segment CTA timed out waiting for its predecessor (internal error)");
  return AERODDC_OK;
}

int aeroddc_bank_process(aeroddc_bank* b, const void* host_iq, size_t n_complex) {
  int rc = aeroddc_bank_submit(b, host_iq, n_complex);
  if (rc != AERODDC_OK) return rc;
  while (b->blocks_done < b->blocks_submitted) {
    rc = aeroddc_bank_wait(b);
    if (rc != AERODDC_OK) return rc;
  }
  return AERODDC_OK;
}

int aeroddc_bank_host_slot(aeroddc_bank* b, int slot, void** ptr, size_t* bytes) {
  if (!b || !ptr) return fail(AERODDC_ERR_ARG, "NULL argument");
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  if (slot < 0 || slot > 1) return fail(AERODDC_ERR_ARG, "slot must be 0 or 1");
  *ptr = b->h_in[slot];
  if (bytes) *bytes = b->in_bytes;
  return AERODDC_OK;
}

int aeroddc_bank_output(aeroddc_bank* b, int vfo, const void** payload, size_t* nbytes, uint32_t* rate) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (vfo < 0 || vfo >= (int)b->vfos.size()) return fail(AERODDC_ERR_ARG, "vfo index %d out of range", vfo);
  if (b->cur_out < 0) return fail(AERODDC_ERR_STATE, "no completed block yet");
  const VfoRec& r = b->vfos[vfo];
  if (payload) *payload = b->h_out[b->cur_out] + r.out_off;
  if (nbytes) *nbytes = r.out_bytes;
  if (rate) *rate = (uint32_t)r.plan.out_rate;
  return AERODDC_OK;
}

const char* aeroddc_bank_topic(aeroddc_bank* b, int vfo) {
  if (!b || vfo < 0 || vfo >= (int)b->vfos.size()) return "";
  return b->vfos[vfo].d.topic;
}

int aeroddc_bank_stage_d(aeroddc_bank* b, int vfo, float* host_out, size_t cap_complex) {
  if (!b || !host_out) return fail(AERODDC_ERR_ARG, "NULL argument");
  if (vfo < 0 || vfo >= (int)b->vfos.size()) return fail(AERODDC_ERR_ARG, "vfo index %d out of range", vfo);
  if (b->blocks_done < 1 || b->blocks_done != b->blocks_submitted) return fail(AERODDC_ERR_STATE, "needs a completed block and nothing in flight");
  CU(cudaSetDevice(b->device));
  const VfoRec& r = b->vfos[vfo];
  const size_t n = std::min<size_t>(cap_complex, (size_t)r.plan.n_stage);
  // the block stays at [hist, hist + n_stage) of the row until the next block overwrites it; the
  // history shift only rewrites [0, hist)
  CU(cudaStreamSynchronize(b->s_compute));
  CU(cudaMemcpy(host_out, b->d_xd + r.xd_off + r.hist, n * sizeof(float2), cudaMemcpyDeviceToHost));
  return (int)r.plan.n_stage;
}

int aeroddc_bank_num_vfos(aeroddc_bank* b) { return b ? (int)b->vfos.size() : 0; }

int aeroddc_bank_last_timing(aeroddc_bank* b, float* kernel_ms, int* launches) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (kernel_ms) *kernel_ms = b->last_kernel_ms;
  if (launches) *launches = b->last_launches;
  return AERODDC_OK;
}
int aeroddc_bank_last_main_ms(aeroddc_bank* b, float* main_ms) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (main_ms) *main_ms = b->last_main_ms;
  return AERODDC_OK;
}
int aeroddc_bank_stopwatch(aeroddc_bank* b, int which, float* ms) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  CU(cudaSetDevice(b->device));
  if (which == 0) { CU(cudaEventRecord(b->ev_sw0, b->s_compute)); return AERODDC_OK; }
  if (which == 1) { CU(cudaEventRecord(b->ev_sw0, b->s_copy)); return AERODDC_OK; }
  if (which != 2 || !ms) return fail(AERODDC_ERR_ARG, "which must be 0, 1 or 2 (with ms)");
  // the region ends when the last payload copy has finished: make the compute stream wait for it first
  if (b->blocks_submitted > 0) CU(cudaStreamWaitEvent(b->s_compute, b->ev_d2h[(b->blocks_submitted - 1) & 1], 0));
  CU(cudaEventRecord(b->ev_sw1, b->s_compute));
  CU(cudaEventSynchronize(b->ev_sw1));
  CU(cudaEventElapsedTime(ms, b->ev_sw0, b->ev_sw1));
  return AERODDC_OK;
}
int aeroddc_bank_device_bytes(aeroddc_bank* b, size_t* bytes) {
  if (!b || !bytes) return fail(AERODDC_ERR_ARG, "NULL argument");
  *bytes = b->dev_bytes;
  return AERODDC_OK;
}

void aeroddc_bank_destroy(aeroddc_bank* b) {
  if (!b) return;
  free_all(b);
  delete b;
}

int aeroddc_measure_fp32_peak(int device, double* tflops, double* sm_clock_mhz) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(AERODDC_ERR_CUDA, "no CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  float* out = nullptr;
  long long* clk = nullptr;
  CU(cudaMalloc(&out, sizeof(float) * blocks * threads));
  CU(cudaMalloc(&clk, sizeof(long long)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  fp32_peak_kernel<<<blocks, threads>>>(out, 2000, 1.0000001f, 1e-9f, clk);
  CU(cudaDeviceSynchronize());
  float best = 1e30f;
  long long cyc = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CU(cudaEventRecord(e0));
    fp32_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f, clk);
    CU(cudaEventRecord(e1));
    CU(cudaDeviceSynchronize());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CU(cudaMemcpy(&cyc, clk, sizeof cyc, cudaMemcpyDeviceToHost)); }
  }
  const double fma_lanes = (double)iters * 64 * 2 * (double)blocks * threads;   // packed: 2 lanes per instruction
  if (tflops) *tflops = 2.0 * fma_lanes / (best * 1e-3) / 1e12;
  if (sm_clock_mhz) *sm_clock_mhz = (double)cyc / 1000.0;   // the kernel reports cycles per microsecond x1000
  cudaFree(out); cudaFree(clk); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return AERODDC_OK;
}

}  // extern "C"
