// aero-ddc-b200: the bank object behind the C ABI of include/aeroddc.h.
//
// Host orchestration only: VFO bookkeeping, host-side filter design (design.cpp), HBM layout,
// streams/events, and the launches of the kernels in ddc_kernels.cuh / aux_kernels.cuh.
// Replaces the per-VFO objects of /root/reference/publish/vfo.cpp:57-139 (init) and the block pump
// of /root/reference/publish/publisher.cpp:285-306 (demodData -> vfo::process for every VFO).
#include "../../include/aeroddc.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ddc_kernels.cuh"
#include "aux_kernels.cuh"
#include "tc_kernels.cuh"
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

struct Group {   // VFOs sharing input stream and DA = min(D, 5): one launch of the main and of the deep kernel per block
  int parent;      // -1: raw IQ of the bank; else the VFO whose stage-D stream is the input
  int DA, Dmax, base, count;
  int fs_in, blk_in;
  int S, W, Wb, nseg;
  int Q, P;          // parts per full segment and part length (chained CTAs, see ddc_kernels.cuh)
  int nchains, nparts; // chains = 32-VFO groups x segments; parts of all chains together (the last segment may be shorter)
  bool direct;       // no VFO of the group has a deep stage and the bank runs in order: the main kernel writes the stage-D rows itself
  int mid_pitch;     // lanes of the stage-DA stream: 32 per VFO group
  int n_mid;         // stage-DA samples per block
  int deep_T, deep_ranges;
  int* d_sched = nullptr;      // [2 + nchains + (nparts - nchains)] scheduler words of the main kernel
  size_t sched_words = 0;
  float2* d_hand = nullptr;
  float2* d_mid[2] = {nullptr, nullptr};   // [32-VFO group][n_mid][32], by block parity
  uint4* d_filt = nullptr;     // tensor mode: per-VFO modulated composite filters, [tiles of 64 VFOs][40 k-steps][8 KB]
  int tc_ntiles = 0;           // > 0: this group runs on ddc_tc_kernel (all but the zones, which stay on the FP32 kernel)
  TcSeg* d_segs = nullptr;     // host-planned stretches of the (VFO tile, time) plane, one run of them per CTA
  int* d_cta_seg = nullptr;
  int tc_grid = 0;
};

// cudaFuncSetAttribute is state of the (function, device) pair, shared by every bank of the process: a second bank with
// shorter filters must not shrink the limit under a first bank that needs more (the reference's own Publisher builds one
// private bank per main VFO when it drives the product's vfo class, INTEGRATION.md option C). Keep the maximum ever asked.
cudaError_t raise_tail_smem_limit(int device, size_t bytes) {
  static std::mutex mu;
  static size_t limit[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  const int d = device & 63;
  if (bytes <= limit[d]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) limit[d] = bytes;
  return e;
}

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
  float* d_dcc_out[2] = {nullptr, nullptr};   // DC-corrected cf32 block, by block parity (block k+1 is corrected while block k computes)
  float* d_dcc_state = nullptr;  // running average per rail
  size_t dcc_smem = 0;           // dynamic shared memory the DC kernel asks for (a whole SM's worth, see enqueue_block)
  bool nested = false;           // some VFO feeds sub-VFOs, or some group needs no deep kernel: the kernels of a block then run
                                 // strictly in order on one stream (no overlap with the next block's main kernel)
  bool poisoned = false;         // a chained CTA gave up waiting: every later call fails
  std::vector<VfoRec> vfos;
  std::vector<Group> groups;
  int vfo_pitch = 0;
  int n_sm = 148;
  int nck = 0;

  // device memory
  float2 *d_rot = nullptr, *d_qlast = nullptr, *d_ckpt = nullptr;
  float2* d_state[3] = {nullptr, nullptr, nullptr};   // boundary history, rotating by block (the deep kernel of block k still reads
                                                      // state[k % 3] while the main kernel of block k+1 writes state[(k+2) % 3])
  unsigned char* d_vfo_D = nullptr;                   // [vfo_pitch] half-band stages per column
  float2* d_xd = nullptr;       // all stage-D rows, per VFO: [hist][n_stage]; two copies, by block parity, so that block k+1's kernels
                                // write their rows while block k's tail still reads its own (the history shift copies across)
  float2** d_xd_rows[2] = {nullptr, nullptr};   // [vfo_pitch] pointer to stage-D index 0 of each column's row, per parity
  int tc_fstages = 10;
  bool gather = false;           // sliced blocks are first gathered into local HBM by the copy engines (tensor mode, see submit_device_sliced)
  float2* d_pw = nullptr;       // tensor mode: [kTcPwRows][vfo_pitch] unit rotation powers u^r
  int* d_nco_len = nullptr;     // [vfo_pitch]
  int* d_post_ctr = nullptr;    // work-item counters of the persistent post-processing kernels: [0] tail, [1 + g] deep kernel of group g
  int post_ctas_deep = 0, post_ctas_tail = 0;   // their grid sizes
  int* d_err = nullptr;         // chained-CTA watchdog flag: device view of h_err (zero-copy pinned host memory)
  volatile int* h_err = nullptr;
  int nck_max = 0;
  float* d_taps = nullptr;
  int* d_hil_idx = nullptr;
  TailVfo* d_tail[2] = {nullptr, nullptr};   // per block parity (they differ in the stage-D row pointers)
  unsigned char* d_out = nullptr;
  size_t out_total = 0;
  unsigned char* d_in[2] = {nullptr, nullptr};
  size_t in_bytes = 0;
  size_t dev_bytes = 0;
  size_t xd_total = 0;

  // host memory
  unsigned char* h_in[2] = {nullptr, nullptr};
  unsigned char* h_out[3] = {nullptr, nullptr, nullptr};   // three slots: a payload stays valid until the next wait()

  // s_compute: main kernels. s_post: deep + tail + history shift of block k, overlapping the main kernel of block k+1
  // (== s_compute for nested banks). s_dcc: DC removal one block ahead. s_copy: H2D of raw blocks. s_d2h: payloads out.
  cudaStream_t s_compute = nullptr, s_post = nullptr, s_dcc = nullptr, s_copy = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_post[2] = {};   // tail + history shift of parity p finished (the stage-D rows may be overwritten)
  cudaEvent_t ev_main[2] = {};   // main kernels of parity p finished
  cudaEvent_t ev_dcc[2] = {};    // corrected block of parity p ready
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
  cudaDeviceSynchronize();
  cudaFree(b->d_post_ctr);
  cudaFree(b->d_rot); cudaFree(b->d_qlast); cudaFree(b->d_ckpt); cudaFree(b->d_vfo_D);
  for (int i = 0; i < 3; ++i) cudaFree(b->d_state[i]);
  for (Group& g : b->groups) { cudaFree(g.d_sched); cudaFree(g.d_hand); cudaFree(g.d_mid[0]); cudaFree(g.d_mid[1]); cudaFree(g.d_filt); cudaFree(g.d_segs); cudaFree(g.d_cta_seg); }
  cudaFree(b->d_pw);
  if (b->h_err) cudaFreeHost((void*)b->h_err);
  cudaFree(b->d_dcc_out[0]); cudaFree(b->d_dcc_out[1]); cudaFree(b->d_dcc_state);
  cudaFree(b->d_xd); cudaFree(b->d_xd_rows[0]); cudaFree(b->d_xd_rows[1]); cudaFree(b->d_nco_len); cudaFree(b->d_taps); cudaFree(b->d_hil_idx); cudaFree(b->d_tail[0]); cudaFree(b->d_tail[1]); cudaFree(b->d_out);
  cudaFree(b->d_in[0]); cudaFree(b->d_in[1]);
  for (int i = 0; i < 2; ++i) {
    if (b->h_in[i]) cudaFreeHost(b->h_in[i]);
    for (cudaEvent_t e : {b->ev_h2d[i], b->ev_tail[i], b->ev_d2h[i], b->ev_main[i], b->ev_dcc[i], b->ev_post[i]}) if (e) cudaEventDestroy(e);
  }
  for (int i = 0; i < 3; ++i) {
    if (b->h_out[i]) cudaFreeHost(b->h_out[i]);
    for (cudaEvent_t e : {b->ev_done[i], b->ev_k0[i], b->ev_k1[i], b->ev_m0[i], b->ev_m1[i]}) if (e) cudaEventDestroy(e);
  }
  if (b->ev_sw0) cudaEventDestroy(b->ev_sw0);
  if (b->ev_sw1) cudaEventDestroy(b->ev_sw1);
  if (b->s_post && b->s_post != b->s_compute) cudaStreamDestroy(b->s_post);
  for (cudaStream_t st : {b->s_compute, b->s_dcc, b->s_copy, b->s_d2h}) if (st) cudaStreamDestroy(st);
}

// Enqueue everything that follows the arrival of the raw block in device memory. The caller has already made the
// first stream of the chain (s_dcc with DC correction, else s_compute) wait for the block.
//   s_compute : [memset flags, main kernel] per group                       -> ev_main
//   s_post    : deep kernel per group, tail, history shift (after ev_main)  -> ev_tail     (block k's post-processing
//               overlaps the main kernel of block k+1; a nested bank runs both on one stream, group by group)
//   s_d2h     : payload copy-out (after ev_tail)                            -> ev_done
int enqueue_block(aeroddc_bank* b, const RawBlock& raw_in) {
  const long long k = b->blocks_submitted;
  const int slot = (int)(k % 3);   // payload/event slot
  const int par = (int)(k & 1);
  cudaStream_t sA = b->s_compute, sB = b->s_post;
  const bool fast = b->mode != AERODDC_MODE_EXACT;
  int launches = 0;
  RawBlock raw = raw_in;
  int raw_fmt = b->fmt;
  if (b->dcc) {
    // Sequential DC removal of the raw stream, one block ahead of the VFOs, which then read the corrected cf32 block.
    // The recurrence is one dependent FMUL + FADD per sample; next to 16 warps that saturate the FP32 pipe its warp
    // would be starved (measured: 19x slower), so the CTA (recurrence warp + a loading and a storing warp, dcc_kernel)
    // asks for a whole SM's shared memory and thereby keeps the SM to itself: 1/148 of the machine for the rate-limiting
    // step of a DC-corrected stream.
    const size_t ex = b->dcc_smem;
    if (b->fmt == AERODDC_CU8) dcc_kernel<0><<<1, kDccThreads, ex, b->s_dcc>>>(raw, b->d_dcc_out[par], b->d_dcc_state, b->B);
    else if (b->fmt == AERODDC_CS16) dcc_kernel<1><<<1, kDccThreads, ex, b->s_dcc>>>(raw, b->d_dcc_out[par], b->d_dcc_state, b->B);
    else dcc_kernel<2><<<1, kDccThreads, ex, b->s_dcc>>>(raw, b->d_dcc_out[par], b->d_dcc_state, b->B);
    CU(cudaGetLastError());
    ++launches;
    CU(cudaEventRecord(b->ev_dcc[par], b->s_dcc));
    CU(cudaStreamWaitEvent(sA, b->ev_dcc[par], 0));
    raw.slice[0] = b->d_dcc_out[par];
    raw.n_slices = 1;
    raw.slice_len = b->B;
    raw_fmt = AERODDC_CF32;
  }
  CU(cudaEventRecord(b->ev_k0[slot], sA));
  if (b->nested) CU(cudaMemsetAsync(b->d_post_ctr, 0, sizeof(int) * (1 + b->groups.size()), sA));
  CU(cudaEventRecord(b->ev_m0[slot], sA));
  const float2* state_in = b->d_state[k % 3];
  float2* state_out = b->d_state[(k + 1) % 3];

  auto launch_deep = [&](const Group& g, cudaStream_t st) -> int {
    DeepParams d;
    d.mid = g.d_mid[par];
    d.state_in = state_in;
    d.state_out = state_out;
    d.xd_rows = b->d_xd_rows[par];
    d.vfo_D = b->d_vfo_D;
    d.counter = b->d_post_ctr + 1 + (&g - &b->groups[0]);
    d.vfo_pitch = b->vfo_pitch;
    d.vfo_base = g.base;
    d.vfo_count = g.count;
    d.DA = g.DA;
    d.n_mid = g.n_mid;
    d.T = g.deep_T;
    d.Wd = kDeepWarm;
    d.nranges = g.deep_ranges;
    d.ngroups = (g.count + 31) / 32;
    d.one = 1.0f;
    const int items = d.nranges * d.ngroups;
    const unsigned grid = (unsigned)std::min(items, b->post_ctas_deep);
    const size_t ring = sizeof(float2) * 2 * kDeepStep * 32;
    if (fast) ddc_deep_kernel<true><<<grid, 32, ring, st>>>(d);
    else ddc_deep_kernel<false><<<grid, 32, ring, st>>>(d);
    CU(cudaGetLastError());
    ++launches;
    return AERODDC_OK;
  };

  for (const Group& g : b->groups) {
    MainParams p;
    if (g.parent < 0) {
      p.raw = raw;
    } else {   // a sub-VFO group reads its parent's stage-D stream of this block (already enqueued on this stream)
      p.raw.slice[0] = b->d_xd + (size_t)par * b->xd_total + b->vfos[g.parent].xd_off + b->vfos[g.parent].hist;
      p.raw.n_slices = 1;
      p.raw.slice_len = g.blk_in;
    }
    p.ckpt = b->d_ckpt;
    p.rot = b->d_rot;
    p.qlast = b->d_qlast;
    p.state_in = state_in;
    p.state_out = state_out;
    p.mid = g.direct ? nullptr : g.d_mid[par];
    p.n_mid = g.n_mid;
    p.xd_rows = b->d_xd_rows[par];
    p.block_abs = k * (long long)g.blk_in;
    p.vfo_pitch = b->vfo_pitch;
    p.vfo_base = g.base;
    p.vfo_count = g.count;
    p.DA = g.DA;
    p.B = g.blk_in;
    p.S = g.S;
    p.W = g.W;
    p.nseg = g.nseg;
    p.nco_len = g.fs_in;
    p.one = 1.0f;
    p.transient = 4 * kNcoStride;
    p.nck = b->nck_max;
    p.P = g.P;
    p.ngroups = (g.count + kVfoPerCta - 1) / kVfoPerCta;
    p.nchains = g.nchains;
    p.sched = g.d_sched;
    p.hand = g.d_hand;
    p.err = b->d_err;
    p.seg_off = 0;
    p.cold0 = 0;
    p.nbound = p.ngroups;
    const int fmt_in = g.parent < 0 ? raw_fmt : AERODDC_CF32;
    if (g.tc_ntiles > 0) {
      // Tensor mode. The FP32 kernel keeps what is not a clean FIR window: the head of the block (the half-band queues
      // are re-seeded with a one-sample shift at every block start, dsp.cpp:163-172) and the zone after an oscillator
      // restart (amplitude transient of the recurrence, oscillator.cpp:19-24); its stage-5 samples go to the mid stream,
      // where ddc_tc_kernel - which does every other output and all the later stages - picks them up.
      const int zone = kTcHead * 32;                                     // 2048 samples
      const long long idx0 = p.block_abs % g.fs_in;
      const long long wrap = idx0 == 0 ? 0 : (long long)g.fs_in - idx0;   // in-block sample at which the table restarts
      int head = zone, fix_at = -1;
      if (wrap > 0 && wrap < g.blk_in) {
        const int a = (int)(wrap / kNcoStride) * kNcoStride;
        if (a < zone) head = a + zone; else fix_at = a;
      }
      head = std::min(head, g.blk_in);
      auto launch_range = [&](int off, int len, bool with_boundary) -> int {
        MainParams r = p;
        r.mid = g.d_mid[par];
        r.seg_off = off; r.S = len; r.P = len; r.nseg = 1; r.nchains = r.ngroups;
        r.cold0 = off > 0; r.nbound = with_boundary ? r.ngroups : 0;
        CU(cudaMemsetAsync(g.d_sched, 0, sizeof(int) * g.sched_words, sA));
        CU(launch_main(AERODDC_CF32, g.DA, r, dim3((unsigned)(r.nbound + r.nchains)), sA, true));
        ++launches;
        return AERODDC_OK;
      };
      { const int rc = launch_range(0, head, true); if (rc != AERODDC_OK) return rc; }
      if (fix_at >= 0) { const int rc = launch_range(fix_at, zone, false); if (rc != AERODDC_OK) return rc; }
      TcParams t;
      t.raw = p.raw;
      t.filt = g.d_filt;
      t.ckpt = b->d_ckpt;
      t.pw = b->d_pw;
      t.zone = g.d_mid[par];
      t.state_in = state_in;
      t.state_out = state_out;
      t.xd_rows = b->d_xd_rows[par];
      t.vfo_D = b->d_vfo_D;
      t.segs = g.d_segs;
      t.cta_seg = g.d_cta_seg;
      t.block_abs = p.block_abs;
      t.nco_len = g.fs_in;
      t.nck = (g.fs_in + kNcoStride - 1) / kNcoStride;
      t.vfo_pitch = b->vfo_pitch;
      t.vfo_base = g.base;
      t.vfo_count = g.count;
      t.n_mid = g.n_mid;
      t.z0_end = head / 32;
      t.z1_lo = fix_at >= 0 ? fix_at / 32 : 0;
      t.z1_hi = fix_at >= 0 ? std::min((fix_at + zone) / 32, g.n_mid) : 0;
      // the stage-D rows of this parity were last read by the tail of block k-2
      if (k >= 2) CU(cudaStreamWaitEvent(sA, b->ev_post[par], 0));
      if (b->tc_fstages == 6) ddc_tc_kernel<6><<<(unsigned)g.tc_grid, kTcThreads, TcSmem<6>::kTotal, sA>>>(t);
      else ddc_tc_kernel<10><<<(unsigned)g.tc_grid, kTcThreads, TcSmem<10>::kTotal, sA>>>(t);
      CU(cudaGetLastError());
      ++launches;
    } else {
      CU(cudaMemsetAsync(g.d_sched, 0, sizeof(int) * g.sched_words, sA));
      dim3 grid((unsigned)(p.ngroups + g.nparts));
      CU(launch_main(fmt_in, g.DA, p, grid, sA, fast));
      ++launches;
    }
    if (b->nested && !g.direct && g.tc_ntiles == 0) { const int rc = launch_deep(g, sA); if (rc != AERODDC_OK) return rc; }
  }
  CU(cudaEventRecord(b->ev_m1[slot], sA));
  if (!b->nested) {
    CU(cudaEventRecord(b->ev_main[par], sA));
    CU(cudaStreamWaitEvent(sB, b->ev_main[par], 0));
    CU(cudaMemsetAsync(b->d_post_ctr, 0, sizeof(int) * (1 + b->groups.size()), sB));
    for (const Group& g : b->groups) if (g.tc_ntiles == 0) { const int rc = launch_deep(g, sB); if (rc != AERODDC_OK) return rc; }
  }
  {
    // payload rows are double-buffered by block parity; wait until the copy-out of this parity (two blocks ago) is done
    if (k >= 2) CU(cudaStreamWaitEvent(sB, b->ev_d2h[par], 0));
    const int items = b->tail_chunks * (int)b->vfos.size();
    tail_kernel<<<(unsigned)std::min(items, b->post_ctas_tail), kTailThreads, b->tail_smem, sB>>>(
        b->d_tail[par], 1.0f, (size_t)par * b->out_total, b->d_post_ctr, (int)b->vfos.size(), b->tail_chunks);
    CU(cudaGetLastError());
    ++launches;
  }
  CU(cudaEventRecord(b->ev_k1[slot], sB));
  CU(cudaEventRecord(b->ev_tail[par], sB));
  // keep the last `hist` stage-D samples of every VFO in front of the next block
  if (b->any_hist) {
    xd_shift_kernel<<<(unsigned)b->vfos.size(), 256, 0, sB>>>(b->d_tail[par], b->d_tail[par ^ 1]);
    CU(cudaGetLastError());
    ++launches;
  }
  CU(cudaEventRecord(b->ev_post[par], sB));
  // payloads leave on their own stream, overlapping the next block's kernels
  CU(cudaStreamWaitEvent(b->s_d2h, b->ev_tail[par], 0));
  CU(cudaMemcpyAsync(b->h_out[slot], b->d_out + (size_t)par * b->out_total, b->out_total, cudaMemcpyDeviceToHost, b->s_d2h));
  CU(cudaEventRecord(b->ev_d2h[par], b->s_d2h));
  CU(cudaEventRecord(b->ev_done[slot], b->s_d2h));
  b->launches_per_block = launches;
  b->blocks_submitted++;
  return AERODDC_OK;
}

int check_submit(aeroddc_bank* b, size_t n_complex) {
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  if (b->poisoned) return fail(AERODDC_ERR_CUDA, "the bank stopped after an internal error (a chained segment CTA timed out); destroy it");
  if (n_complex != (size_t)b->B)
    return fail(AERODDC_ERR_ARG, "block of %zu samples, bank was created for %d (vfo.cpp:155,164 block contract)", n_complex, b->B);
  if (b->blocks_submitted - b->blocks_done >= 2) return fail(AERODDC_ERR_STATE, "two blocks already in flight; call wait()");
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
  const int DA = std::min(decim_count, kFastStages);   // stages of the main kernel; the deep kernel needs no segmentation plan
  const int cstep = ilcm(kChunk, 1 << DA);
  if (block_len % cstep) return fail(AERODDC_ERR_ARG, "block of %d samples is not a multiple of %d", block_len, cstep);
  const int align = ilcm(kNcoStride, 1 << DA);
  out->warmup = DA == 0 ? 0 : ((10 << DA) + cstep - 1) / cstep * cstep;     // 10*(2^DA - 1) samples reach the last register stage's history
  out->boundary_warmup = decim_count == 0 ? 0 : (11 << decim_count);        // the next block's shifted history needs 11 samples per stage
  const int groups = (n_vfos + kVfoPerCta - 1) / kVfoPerCta;
  // One wave of one-warp CTAs (kCtasPerSm per SM), plus a few percent, cuts the block into segments, each paying one
  // warm-up of W samples. The warp scheduler favours some resident warps, so equal CTAs of a single wave would finish
  // at different times; each segment is therefore a chain of Q parts handed from CTA to CTA through HBM, and the
  // kernel's FIFO of ready chains (a few more chains than SM slots keep it non-empty) lets every chain advance at the
  // average pace of all slots.
  const int target = std::max(1, (int)std::ceil((double)kCtasPerSm * n_sm * waves * 1.04 / groups));
  int S = (block_len + target - 1) / target;
  S = std::max(S, std::max(2 * out->warmup, 512));   // a small bank: short segments, for latency rather than efficiency
  S = (S + align - 1) / align * align;
  out->segment_len = S;
  out->n_segments = (block_len + S - 1) / S;
  int Q = parts > 0 ? parts : 16;
  const int pal = ilcm(kTile, 1 << DA);
  while (Q > 1 && S / Q < 8 * pal) --Q;   // keep parts long enough (2048 samples) that the hand-over stays negligible
  Q = std::min(Q, 60);
  out->parts = Q;
  out->part_len = ((S + Q - 1) / Q + pal - 1) / pal * pal;
  out->vfo_groups = groups;
  const int last = block_len - (out->n_segments - 1) * S;                      // the last segment may be shorter
  const int parts_total = (out->n_segments - 1) * ((S + out->part_len - 1) / out->part_len) + (last + out->part_len - 1) / out->part_len;
  out->ctas = groups * (1 + parts_total);
  return AERODDC_OK;
}

int aeroddc_plan_tensor_stretches(int n_tiles, int n_mid, int n_sm, int* cta_first, int* stretches, int cap) {
  if (n_tiles < 1 || n_mid < 1 || n_sm < 1) return fail(AERODDC_ERR_ARG, "bad planning arguments");
  const long long U = (long long)n_tiles * n_mid;
  const int P = (int)std::max<long long>(1, std::min<long long>(n_sm, U / 1024));
  if (!cta_first && !stretches) return P;
  auto bound = [&](int c) -> long long {   // first output (tile-major) of CTA c, cut at a multiple of 256 inside its tile
    if (c >= P) return U;
    const long long gg = U * c / P;
    return gg / n_mid * n_mid + (gg % n_mid) / kTcCols * kTcCols;
  };
  int n = 0;
  for (int c = 0; c < P; ++c) {
    if (cta_first) cta_first[c] = n;
    long long gg = bound(c);
    const long long g1 = bound(c + 1);
    while (gg < g1) {
      const int nt = (int)(gg / n_mid), m_lo = (int)(gg % n_mid);
      const int m_hi = (int)std::min<long long>(n_mid, m_lo + (g1 - gg));
      if (stretches) {
        if (n >= cap) return fail(AERODDC_ERR_ARG, "stretch buffer too small");
        stretches[3 * n] = nt; stretches[3 * n + 1] = m_lo; stretches[3 * n + 2] = m_hi;
      }
      ++n;
      gg += m_hi - m_lo;
    }
  }
  if (cta_first) cta_first[P] = n;
  return n;
}

int aeroddc_bank_set_mode(aeroddc_bank* b, int mode) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->blocks_submitted > 0) return fail(AERODDC_ERR_STATE, "the arithmetic mode cannot change once blocks were processed");
  if (mode != AERODDC_MODE_EXACT && mode != AERODDC_MODE_FAST && mode != AERODDC_MODE_TENSOR) return fail(AERODDC_ERR_ARG, "unknown mode %d", mode);
  if (mode == AERODDC_MODE_TENSOR && b->finalized) return fail(AERODDC_ERR_STATE, "AERODDC_MODE_TENSOR must be chosen before finalize (its filter tables are built there)");
  if (b->mode == AERODDC_MODE_TENSOR && mode != AERODDC_MODE_TENSOR && b->finalized) return fail(AERODDC_ERR_STATE, "the bank was finalized for AERODDC_MODE_TENSOR");
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

  // ---- group VFOs by (input stream, min(D, 5)); raw-fed groups first so that parents run before children ----
  const int nv = (int)b->vfos.size();
  int col = 0;
  for (int parent = -1; parent < nv; ++parent) {
    if (parent >= 0 && b->vfos[parent].children == 0) continue;
    if (parent >= 0) b->nested = true;
    for (int DA = 0; DA <= kFastStages; ++DA) {
      Group g;
      g.parent = parent; g.DA = DA; g.Dmax = 0; g.base = col; g.count = 0;
      for (int i = 0; i < nv; ++i)
        if (b->vfos[i].d.parent == parent && std::min(b->vfos[i].d.decim_count, kFastStages) == DA) {
          b->vfos[i].slot = col++; g.count++;
          g.Dmax = std::max(g.Dmax, b->vfos[i].d.decim_count);
          g.fs_in = b->vfos[i].fs_in; g.blk_in = b->vfos[i].blk_in;
        }
      if (g.count) b->groups.push_back(g);
    }
  }
  b->vfo_pitch = (col + 3) & ~3;
  for (const Group& g : b->groups) if (g.Dmax <= kFastStages) b->nested = true;   // such a group writes its stage-D rows directly: keep order
  const char* env_waves = getenv("AERODDC_WAVES");
  const double waves = env_waves ? std::max(0.05, atof(env_waves)) : 1.0;
  const char* env_parts = getenv("AERODDC_PARTS");
  const int env_parts_n = env_parts ? std::max(1, atoi(env_parts)) : 0;
  size_t bytes = 0;
  auto dmalloc = [&](void** p, size_t n) { bytes += n; return cudaMalloc(p, n); };
  for (Group& g : b->groups) {
    aeroddc_segment_plan pl;
    const int rc = aeroddc_plan_segments(g.blk_in, g.DA, g.count, b->n_sm, waves, env_parts_n, &pl);
    if (rc != AERODDC_OK) return rc;
    g.W = pl.warmup; g.S = pl.segment_len; g.nseg = pl.n_segments; g.Q = pl.parts; g.P = pl.part_len;
    g.direct = g.Dmax <= kFastStages;
    g.mid_pitch = 32 * ((g.count + 31) / 32);   // whole 32-VFO groups: [group][time][32 lanes]
    g.n_mid = g.blk_in >> g.DA;
    // time ranges of the deep kernel: long where the stream is long (each range re-reads 72 samples of run-in), short
    // where a small bank would otherwise leave the machine idle
    g.deep_T = g.n_mid >= 65536 ? 1024 : (g.n_mid >= 8192 ? 256 : 64);   // multiples of the deep kernel's 32-sample step
    g.deep_ranges = (g.n_mid + g.deep_T - 1) / g.deep_T;
    const int ng = (g.count + kVfoPerCta - 1) / kVfoPerCta;
    g.nchains = ng * g.nseg;
    g.nparts = pl.ctas - ng;
    g.sched_words = 2 + (size_t)g.nparts;   // counters + parts done per chain + queue (one entry per part after a chain's first)
    CU(dmalloc((void**)&g.d_sched, sizeof(int) * g.sched_words));
    CU(cudaMemset(g.d_sched, 0, sizeof(int) * g.sched_words));
    CU(dmalloc((void**)&g.d_hand, sizeof(float2) * (size_t)g.nchains * kHandSlots * kThreads));
    if (!g.direct)
      for (int i = 0; i < 2; ++i) CU(dmalloc((void**)&g.d_mid[i], sizeof(float2) * (size_t)g.n_mid * g.mid_pitch));
  }
  CU(dmalloc((void**)&b->d_post_ctr, sizeof(int) * (1 + b->groups.size())));
  // Persistent grids of the post-processing kernels, in CTAs per SM. A flat bank overlaps block k's post-processing with
  // block k+1's main kernel. Measured on B200: too few light warps per SM beside the FP32-saturating main kernel are
  // starved by the warp scheduler (2 + 1 per SM: the deep kernel then takes longer than the main kernel), too many take
  // register-file slots from main-kernel CTAs for longer than their work is worth.
  const char* env_post = getenv("AERODDC_POST_CTAS");   // experiments: "deep,tail" CTAs per SM
  // Flat bank in the FP32 modes (post-processing beside the next block's main kernel): 4 + 4 per SM measured best once the
  // deep kernel fetched by bulk copies (805 Gsps at 1024 VFOs against 788 with 12 + 6). Kernels that run alone - nested
  // banks, whose kernels go in order, and tensor mode, whose tail cannot share an SM with the tensor kernel - take the SMs' fill.
  const bool beside_main = !b->nested && b->mode != AERODDC_MODE_TENSOR;
  int per_sm_deep = beside_main ? 4 : kDeepCtasPerSm, per_sm_tail = beside_main ? 4 : 6;
  if (env_post) sscanf(env_post, "%d,%d", &per_sm_deep, &per_sm_tail);
  b->post_ctas_deep = std::max(1, per_sm_deep) * b->n_sm;
  b->post_ctas_tail = std::max(1, per_sm_tail) * b->n_sm;
  CU(cudaHostAlloc((void**)&b->h_err, sizeof(int), cudaHostAllocMapped));
  *b->h_err = 0;
  CU(cudaHostGetDevicePointer((void**)&b->d_err, (void*)b->h_err, 0));

  // ---- constant tables ----
  std::vector<float2> h_rot(b->vfo_pitch, make_float2(1.0f, 0.0f));
  std::vector<int> h_len(b->vfo_pitch, 1);
  std::vector<unsigned char> h_D(b->vfo_pitch, 0);
  b->nck_max = 1;
  for (const VfoRec& r : b->vfos) {
    design_rotation((double)r.fs_in, r.d.mixer_freq, &h_rot[r.slot].x, &h_rot[r.slot].y);
    h_len[r.slot] = r.fs_in;
    h_D[r.slot] = (unsigned char)r.d.decim_count;
    b->nck_max = std::max(b->nck_max, (r.fs_in + kNcoStride - 1) / kNcoStride);
  }
  CU(dmalloc((void**)&b->d_rot, sizeof(float2) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_qlast, sizeof(float2) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_nco_len, sizeof(int) * b->vfo_pitch));
  CU(dmalloc((void**)&b->d_vfo_D, b->vfo_pitch));
  CU(dmalloc((void**)&b->d_ckpt, sizeof(float2) * (size_t)b->nck_max * b->vfo_pitch));
  CU(cudaMemset(b->d_ckpt, 0, sizeof(float2) * (size_t)b->nck_max * b->vfo_pitch));
  CU(cudaMemset(b->d_qlast, 0, sizeof(float2) * b->vfo_pitch));
  CU(cudaMemcpy(b->d_rot, h_rot.data(), sizeof(float2) * b->vfo_pitch, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b->d_nco_len, h_len.data(), sizeof(int) * b->vfo_pitch, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b->d_vfo_D, h_D.data(), b->vfo_pitch, cudaMemcpyHostToDevice));
  for (int i = 0; i < 3; ++i) {
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
  b->xd_total = xd_total;
  CU(dmalloc((void**)&b->d_xd, sizeof(float2) * 2 * xd_total));
  CU(cudaMemset(b->d_xd, 0, sizeof(float2) * 2 * xd_total));
  for (int par = 0; par < 2; ++par) {
    std::vector<float2*> h_rows(b->vfo_pitch, b->d_xd);
    for (const VfoRec& r : b->vfos) h_rows[r.slot] = b->d_xd + (size_t)par * xd_total + r.xd_off + r.hist;
    CU(dmalloc((void**)&b->d_xd_rows[par], sizeof(float2*) * b->vfo_pitch));
    CU(cudaMemcpy(b->d_xd_rows[par], h_rows.data(), sizeof(float2*) * b->vfo_pitch, cudaMemcpyHostToDevice));
  }
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
  CU(raise_tail_smem_limit(b->device, tail_smem));
  for (int par = 0; par < 2; ++par) {
    for (int i = 0; i < nv; ++i) h_tail[i].xd = b->d_xd + (size_t)par * xd_total + b->vfos[i].xd_off + b->vfos[i].hist;
    CU(dmalloc((void**)&b->d_tail[par], sizeof(TailVfo) * nv));
    CU(cudaMemcpy(b->d_tail[par], h_tail.data(), sizeof(TailVfo) * nv, cudaMemcpyHostToDevice));
  }

  if (b->dcc) {
    b->dcc_smem = (size_t)prop.sharedMemPerBlockOptin - 1024;   // (almost) everything an SM can give one CTA: no main-kernel CTA fits beside it
    CU(cudaFuncSetAttribute(dcc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->dcc_smem));
    CU(cudaFuncSetAttribute(dcc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->dcc_smem));
    CU(cudaFuncSetAttribute(dcc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->dcc_smem));
    for (int i = 0; i < 2; ++i) CU(dmalloc((void**)&b->d_dcc_out[i], sizeof(float) * 2 * (size_t)b->B));
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
  // the post-processing stream outranks the main stream: its small kernels take the next SM slots that come free while
  // the following block's main kernel is running
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CU(cudaStreamCreateWithPriority(&b->s_compute, cudaStreamNonBlocking, prio_lo));
  if (b->nested) b->s_post = b->s_compute;
  else CU(cudaStreamCreateWithPriority(&b->s_post, cudaStreamNonBlocking, prio_hi));
  CU(cudaStreamCreateWithPriority(&b->s_dcc, cudaStreamNonBlocking, prio_hi));
  CU(cudaStreamCreateWithFlags(&b->s_copy, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&b->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU(cudaEventCreateWithFlags(&b->ev_tail[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_d2h[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_main[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_post[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_dcc[i], cudaEventDisableTiming));
  }
  // every instance of the main kernel may be launched by this bank: raise their dynamic shared-memory limits once
  // (the values are compile-time constants, so banks never disagree about them)
  // (all are below the 48 KB default; nothing to do)

  // ---- NCO checkpoints: the exact sequential recurrence, one thread per VFO ----
  {
    const int threads = 32;
    nco_checkpoint_kernel<<<(b->vfo_pitch + threads - 1) / threads, threads, 0, b->s_compute>>>(
        b->d_rot, b->d_nco_len, b->d_ckpt, b->d_qlast, b->vfo_pitch, b->vfo_pitch, kNcoStride);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(b->s_compute));
  }
  // ---- tensor mode: per-VFO modulated composite filters and rotation powers ----
  if (b->mode == AERODDC_MODE_TENSOR) {
    // g = the five half-band stages as one FIR decimating by 32: h (*) up2(h) (*) up4(h) (*) up8(h) (*) up16(h), 311 taps
    const double h11[11] = {HB_P0, 0, HB_P2, 0, HB_P4, HB_P5, HB_P4, 0, HB_P2, 0, HB_P0};
    std::vector<double> g5(h11, h11 + 11);
    for (int s = 1; s < kFastStages; ++s) {
      std::vector<double> up((10 << s) + 1, 0.0), nx(g5.size() + (10 << s), 0.0);
      for (int i = 0; i < 11; ++i) up[(size_t)i << s] = h11[i];
      for (size_t i = 0; i < g5.size(); ++i)
        for (size_t j = 0; j < up.size(); ++j) nx[i + j] += g5[i] * up[j];
      g5.swap(nx);
    }
    if ((int)g5.size() != kTcTaps) return fail(AERODDC_ERR_DESIGN, "composite half-band response has %zu taps", g5.size());
    double* d_g = nullptr;
    bool any = false;
    for (Group& g : b->groups) {
      if (g.parent >= 0 || g.DA != kFastStages || g.direct) continue;          // raw-fed groups with deep stages only
      if (!(b->fmt == AERODDC_CF32 || b->dcc)) continue;                       // the X tiles are staged from cf32
      if (g.blk_in < 8 * kTcHead * 32 || g.n_mid % 32) continue;                // short or ragged blocks: not worth it
      if (!d_g) {
        CU(cudaMalloc((void**)&d_g, sizeof(double) * kTcTaps));
        CU(cudaMemcpy(d_g, g5.data(), sizeof(double) * kTcTaps, cudaMemcpyHostToDevice));
      }
      g.tc_ntiles = (g.count + kTcVfos - 1) / kTcVfos;
      CU(dmalloc((void**)&g.d_filt, (size_t)g.tc_ntiles * kTcKSteps * kTcFSlab));
      const int n = g.tc_ntiles * kTcKSteps * kTcRails;
      tc_build_filters_kernel<<<(n + 127) / 128, 128, 0, b->s_compute>>>(b->d_rot, d_g, g.base, g.count, g.tc_ntiles, g.d_filt);
      CU(cudaGetLastError());
      // one contiguous stretch of the (VFO tile, time) plane per CTA, cut at multiples of 256 outputs and at tile ends
      const int P = aeroddc_plan_tensor_stretches(g.tc_ntiles, g.n_mid, b->n_sm, nullptr, nullptr, 0);
      std::vector<int> cta_seg(P + 1, 0), flat(3 * (size_t)(P + g.tc_ntiles) + 3, 0);
      const int nseg = aeroddc_plan_tensor_stretches(g.tc_ntiles, g.n_mid, b->n_sm, cta_seg.data(), flat.data(), (int)flat.size() / 3);
      if (nseg < 0) return nseg;
      std::vector<TcSeg> segs((size_t)nseg);
      for (int i = 0; i < nseg; ++i) { segs[i].nt = flat[3 * i]; segs[i].m_lo = flat[3 * i + 1]; segs[i].m_hi = flat[3 * i + 2]; }
      g.tc_grid = P;
      CU(dmalloc((void**)&g.d_segs, sizeof(TcSeg) * segs.size()));
      CU(cudaMemcpy(g.d_segs, segs.data(), sizeof(TcSeg) * segs.size(), cudaMemcpyHostToDevice));
      CU(dmalloc((void**)&g.d_cta_seg, sizeof(int) * cta_seg.size()));
      CU(cudaMemcpy(g.d_cta_seg, cta_seg.data(), sizeof(int) * cta_seg.size(), cudaMemcpyHostToDevice));
      any = true;
    }
    if (any) {
      CU(dmalloc((void**)&b->d_pw, sizeof(float2) * (size_t)kTcPwRows * b->vfo_pitch));
      tc_build_pw_kernel<<<dim3((b->vfo_pitch + 127) / 128, kTcPwRows), 128, 0, b->s_compute>>>(b->d_rot, b->vfo_pitch, b->d_pw);
      CU(cudaGetLastError());
      const char* env_fs = getenv("AERODDC_TC_FSTAGES");   // experiments: depth of the filter-slab ring (6 or 10)
      b->tc_fstages = env_fs && atoi(env_fs) == 6 ? 6 : 10;
      CU(cudaFuncSetAttribute(ddc_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<6>::kTotal));
      CU(cudaFuncSetAttribute(ddc_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<10>::kTotal));
      CU(cudaStreamSynchronize(b->s_compute));
    }
    if (d_g) cudaFree(d_g);
  }
  {
    bool any_tc = false;
    for (const Group& g : b->groups) any_tc = any_tc || g.tc_ntiles > 0;
    const char* env_g = getenv("AERODDC_GATHER");   // experiments: 0 = always read slices in place, 1 = always gather
    b->gather = env_g ? atoi(env_g) != 0 : any_tc;
  }
  b->dev_bytes = bytes;
  b->finalized = true;
  return AERODDC_OK;
}

int aeroddc_bank_submit_device_sliced(aeroddc_bank* b, const void* const* slices, int n_slices, size_t slice_len, size_t n_complex,
                                      void* const* ready_events, int n_events) {
  if (!b || !slices) return fail(AERODDC_ERR_ARG, "NULL argument");
  int rc = check_submit(b, n_complex);
  if (rc != AERODDC_OK) return rc;
  if (n_slices < 1 || n_slices > kMaxSlices) return fail(AERODDC_ERR_ARG, "n_slices must be 1..%d", kMaxSlices);
  if (n_slices > 1 && (slice_len < (size_t)kTile || slice_len % kChunk))
    return fail(AERODDC_ERR_ARG, "slice_len must be a multiple of %d and at least %d complex samples", kChunk, kTile);
  if (n_slices > 1 && (slice_len * (size_t)(n_slices - 1) >= n_complex || slice_len * (size_t)n_slices < n_complex))
    return fail(AERODDC_ERR_ARG, "%d slices of %zu samples do not tile a block of %zu", n_slices, slice_len, n_complex);
  if (n_events < 0 || (n_events > 0 && !ready_events)) return fail(AERODDC_ERR_ARG, "bad ready_events");
  RawBlock raw;
  for (int i = 0; i < kMaxSlices; ++i) raw.slice[i] = nullptr;
  for (int i = 0; i < n_slices; ++i) {
    if (!slices[i]) return fail(AERODDC_ERR_ARG, "slice %d is NULL", i);
    if ((uintptr_t)slices[i] & 15) return fail(AERODDC_ERR_ARG, "device block must be 16-byte aligned");
    raw.slice[i] = slices[i];
  }
  raw.n_slices = n_slices;
  raw.slice_len = n_slices > 1 ? (int)slice_len : b->B;
  CU(cudaSetDevice(b->device));
  cudaStream_t first = b->dcc ? b->s_dcc : b->s_compute;
  if (n_slices > 1 && b->gather) {
    // Tensor mode reads every raw sample once per 64-VFO tile and finishes a block in a fraction of a millisecond per
    // GPU: pulled in place, the tiles of one GPU would cross NVLink (V / N / 64) times and the links, not the tensor
    // pipe, would set the pace (measured: 0.89 ms per block at 8 GPUs against 0.45 ms for the same shard fed from local
    // HBM). So the copy engines gather the slices into the bank's own input buffer first - each byte crosses NVLink once,
    // on the copy stream, while the previous block computes - and the kernels read local memory.
    const int par = (int)(b->blocks_submitted & 1);   // d_in[par] was last read by the block submitted two calls ago (retired)
    const size_t bps = b->in_bytes / (size_t)b->B;
    for (int i = 0; i < n_events; ++i)
      if (ready_events[i]) CU(cudaStreamWaitEvent(b->s_copy, (cudaEvent_t)ready_events[i], 0));
    for (int i = 0; i < n_slices; ++i) {
      const size_t off = (size_t)i * slice_len;
      const size_t cnt = std::min(slice_len, (size_t)b->B - off);
      CU(cudaMemcpyAsync(b->d_in[par] + off * bps, slices[i], cnt * bps, cudaMemcpyDefault, b->s_copy));
    }
    CU(cudaEventRecord(b->ev_h2d[par], b->s_copy));
    CU(cudaStreamWaitEvent(first, b->ev_h2d[par], 0));
    for (int i = 0; i < kMaxSlices; ++i) raw.slice[i] = nullptr;
    raw.slice[0] = b->d_in[par];
    raw.n_slices = 1;
    raw.slice_len = b->B;
    return enqueue_block(b, raw);
  }
  for (int i = 0; i < n_events; ++i)
    if (ready_events[i]) CU(cudaStreamWaitEvent(first, (cudaEvent_t)ready_events[i], 0));
  return enqueue_block(b, raw);
}

int aeroddc_bank_submit_device(aeroddc_bank* b, const void* dev_iq, size_t n_complex, void* ready_event) {
  if (!b || !dev_iq) return fail(AERODDC_ERR_ARG, "NULL argument");
  return aeroddc_bank_submit_device_sliced(b, &dev_iq, 1, n_complex, n_complex, &ready_event, ready_event ? 1 : 0);
}

int aeroddc_bank_submit(aeroddc_bank* b, const void* host_iq, size_t n_complex) {
  if (!b || !host_iq) return fail(AERODDC_ERR_ARG, "NULL argument");
  int rc = check_submit(b, n_complex);
  if (rc != AERODDC_OK) return rc;
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
  CU(cudaStreamWaitEvent(b->dcc ? b->s_dcc : b->s_compute, b->ev_h2d[slot], 0));
  RawBlock raw;
  for (int i = 0; i < kMaxSlices; ++i) raw.slice[i] = nullptr;
  raw.slice[0] = b->d_in[slot];
  raw.n_slices = 1;
  raw.slice_len = b->B;
  return enqueue_block(b, raw);
}

int aeroddc_bank_wait(aeroddc_bank* b) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (b->blocks_done >= b->blocks_submitted) return fail(AERODDC_ERR_STATE, "nothing in flight");
  CU(cudaSetDevice(b->device));
  const int slot = (int)(b->blocks_done % 3);
  CU(cudaEventSynchronize(b->ev_done[slot]));
  b->blocks_done++;
  if (b->poisoned || *b->h_err) {
    // A chained CTA gave up waiting for its predecessor (never observed; the watchdog exists so that a scheduling
    // surprise cannot hang the GPU). Whatever that block produced is garbage: wipe it and stop the bank.
    b->poisoned = true;
    for (int i = 0; i < 3; ++i) memset(b->h_out[i], 0, b->out_total);
    return fail(AERODDC_ERR_CUDA, "a chained segment CTA timed out waiting for its predecessor (internal error); payloads were zeroed, the bank is stopped");
  }
  CU(cudaEventElapsedTime(&b->last_kernel_ms, b->ev_k0[slot], b->ev_k1[slot]));
  CU(cudaEventElapsedTime(&b->last_main_ms, b->ev_m0[slot], b->ev_m1[slot]));
  b->last_launches = b->launches_per_block;
  b->cur_out = slot;
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

int aeroddc_bank_reset(aeroddc_bank* b) {
  if (!b) return fail(AERODDC_ERR_ARG, "NULL bank");
  if (!b->finalized) return fail(AERODDC_ERR_STATE, "bank not finalized");
  if (b->blocks_done != b->blocks_submitted) return fail(AERODDC_ERR_STATE, "blocks in flight; call wait() first");
  CU(cudaSetDevice(b->device));
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < 3; ++i) CU(cudaMemset(b->d_state[i], 0, sizeof(float2) * (size_t)kMaxStages * kStateSlots * b->vfo_pitch));
  CU(cudaMemset(b->d_xd, 0, sizeof(float2) * 2 * b->xd_total));
  if (b->d_dcc_state) CU(cudaMemset(b->d_dcc_state, 0, sizeof(float) * 2));
  b->blocks_submitted = b->blocks_done = 0;
  b->cur_out = -1;
  if (b->poisoned) return fail(AERODDC_ERR_CUDA, "the bank stopped after an internal error; destroy it");
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
  // the block stays at [hist, hist + n_stage) of its parity's row until the block after next overwrites it
  CU(cudaStreamSynchronize(b->s_post));
  CU(cudaMemcpy(host_out, b->d_xd + (size_t)((b->blocks_done - 1) & 1) * b->xd_total + r.xd_off + r.hist, n * sizeof(float2), cudaMemcpyDeviceToHost));
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
