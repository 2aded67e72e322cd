// aero-ddc-b200: tensor-core formulation of the NCO mix + the first five half-band stages (AERODDC_MODE_TENSOR),
// with the remaining half-band stages fused into the epilogue.
//
// Same path as ddc_main_kernel + ddc_deep_kernel (vfo.cpp:155-161 mix, halfbanddecimator.cpp:35-60 x D), recast as the
// complex GEMM north_star names. NOT bit-identical: a tolerance mode (max |err| <= 1e-4 FS, error SNR >= 80 dB), checked
// against the exact mode in tests/test_gpu_parity.py and by decoded-frame identity.
//
// Algebra (scratch/tc_model.py checks it on the CPU). Five cascaded 11-tap half-band decimators are one 311-tap FIR g
// decimating by 32: z[m] = sum_t g[t] q[n_s + t] x[n_s + t], n_s = 32 m - 310. The reference oscillator q is a
// drifting float32 recurrence, but inside one window it is a pure rotation of the value at the window start:
// q[n_s + t] = q[n_s] u^t, u = rot / |rot| (amplitude and phase noise of the recurrence over 320 steps: ~1e-7).
// So  z[m] = q[n_s] * sum_t (g[t] u^t) x[n_s + t]: a per-VFO complex FIR with constant taps G_v[t] = g[t] u_v^t,
// followed by ONE complex multiply per output with the exact recurrence value q[n_s] (exact checkpoint x rotation).
// q[n_s] restarts from (1, 0) with an amplitude transient at every table wrap and the half-band queues are re-seeded
// with a one-sample shift at every block start (dsp.cpp:163-172): windows touching either are not a clean FIR and
// stay on the FP32 kernel (the head of every block and the zone after a wrap; bank.cu launches it over those
// stretches first and this kernel picks its stage-5 samples up from the `zone` stream).
//
// GEMM shape. Windows are padded to 320 samples starting at 32 m - 312 (a multiple of 8). Rows of X are the block cut
// into 32-sample rows (64 floats re/im interleaved); output m needs rows m-10 .. m.
//   D[rail][m] = sum over 40 k-steps of 8 samples (K = 16 bf16) of F[rail][k-step] * X[m + j][chunk]
// rail = (VFO, re/im): rail 0 taps (Gr, -Gi), rail 1 (Gi, Gr). A operand = the filter slab of the k-step (128 rails = 64
// VFOs), streamed from L2 with bulk copies. B operand = X tile in shared memory, K-major, no swizzle, laid out
// [16-byte chunk][row][16 B]: rows are 16 bytes apart, so the j-row shift of a k-step is just a start-address offset of
// the matrix descriptor - every raw sample is staged ONCE per tile instead of ten times (no im2col). Accumulators:
// M=128 rails x N=256 outputs fp32 in TMEM, double-buffered (512 columns). With rails on the TMEM lanes, an epilogue
// thread owns ONE rail and reads its time series along the columns: it rotates by q (partner rail by shuffle), then
// runs the remaining D-5 half-band stages on its rail in registers (the half-band taps are real, so the rails are
// independent) and writes stage-D samples. No intermediate stream, no second kernel.
// Precision: operands are split in bf16 hi + mid (x = hi + mid + O(2^-17 x)); three products hi*hi + mid*hi + hi*mid
// per k-step (the dropped terms are ~2^-17 relative: 107 dB SNR in the model), fp32 accumulation. The two correction
// products skip the k-steps whose taps are below 1e-6 of full scale at their 2^-9 weight (92 instead of 120 MMAs per tile).
//
// Work split: the (VFO tile, time) plane is cut on the host into one contiguous stretch per CTA (equal work, one wave
// of persistent CTAs); a stretch that does not begin at a block start runs the fused stages in over 96 stage-5 samples.
// Warp roles (320 threads, one CTA per SM): warps 0-3 epilogue, warps 4-7 stage X (global -> bf16 hi/mid -> shared),
// warp 8 streams filter slabs (bulk copy), warp 9 issues tcgen05.mma.
#pragma once
#include <cuda_bf16.h>

#include "ddc_kernels.cuh"

namespace aeroddc {

constexpr int kTcRails = 128;             // rails per tile (MMA M)
constexpr int kTcVfos = kTcRails / 2;     // VFOs per tile
constexpr int kTcCols = 256;              // outputs (stage-5 samples) per tile (MMA N)
constexpr int kTcTaps = 311;              // 1 + 10 * (2^5 - 1)
constexpr int kTcWin = 320;               // padded window, samples
constexpr int kTcLead = 312;              // window start = 32 m - kTcLead
constexpr int kTcKSteps = kTcWin / 8;     // 40 MMA k-steps of 8 samples
constexpr int kTcBack = 10;               // rows before the output's own row
constexpr int kTcRows = 272;              // rows per X tile (266 used)
constexpr int kTcXPart = 8 * kTcRows * 16;            // one bf16 part of an X tile: [8 chunks][rows][16 B]
constexpr int kTcXStage = 2 * kTcXPart;               // hi | mid
constexpr int kTcXStages = 2;
constexpr int kTcFPart = 2 * kTcRails * 16;           // one bf16 part of a filter slab: [2 chunks][128 rails][16 B]
constexpr int kTcFSlab = 2 * kTcFPart;                // hi | mid
template <int FS> struct TcSmem {   // FS = depth of the filter-slab ring
  static constexpr int kBars = 2 * kTcXStages + 2 * FS + 4;
  static constexpr int kTotal = kTcXStages * kTcXStage + FS * kTcFSlab + 8 * kBars + 16;
};
constexpr int kTcThreads = 320;
constexpr int kTcPwRows = 512;            // rotation table rows: u^r, r = 0 .. 511
constexpr int kTcHead = 64;               // outputs [0, kTcHead) of a block (2048 samples) stay on the FP32 kernel
// The correction products (hi x mid) are 2^-9 of the main product, so they only need the taps that matter at that scale:
// outside the central 209 taps sum |g| = 4e-4, i.e. 8e-7 of full scale after the 2^-9 (-119 dB in energy).
constexpr int kTcCorrLo = 7, kTcCorrHi = 33;          // k-steps [7, 33): padded taps 56 .. 263, centre at 157
constexpr int kTcRunIn = 96;              // stage-5 samples of run-in for the fused stages (10 * (2^3 - 1), rounded to 32)

struct TcSeg { int nt, m_lo, m_hi; };     // a stretch of one VFO tile: stage-5 outputs [m_lo, m_hi), multiples of 32

struct TcParams {
  RawBlock raw;               // the block (cf32), possibly in slices on several GPUs (a 32-sample row never straddles slices)
  const uint4* filt;          // [n_ntiles][kTcKSteps][kTcFSlab / 16]
  const float2* ckpt;         // [nck][vfo_pitch] exact NCO checkpoints (state after 256 k steps)
  const float2* pw;           // [kTcPwRows][vfo_pitch] u^r
  const float2* zone;         // [32-VFO group][n_mid][32] stage-5 samples the FP32 kernel computed for the zones
  const float2* state_in;     // [kMaxStages][kStateSlots][vfo_pitch] block-start history (stages 5.. are used here)
  float2* state_out;
  float2* const* xd_rows;     // [vfo_pitch] per VFO: where stage-D sample 0 of this block goes
  const unsigned char* vfo_D; // [vfo_pitch]
  const TcSeg* segs;          // host-planned stretches
  const int* cta_seg;         // [gridDim.x + 1] stretches of CTA c: [cta_seg[c], cta_seg[c + 1])
  long long block_abs;
  int nco_len, nck, vfo_pitch, vfo_base, vfo_count;
  int n_mid;                  // B / 32
  int z0_end;                 // zone 0 = outputs [0, z0_end)
  int z1_lo, z1_hi;           // zone 1 = outputs [z1_lo, z1_hi) (empty when z1_hi == 0)
};

// ---- tcgen05 helpers ----
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8-row core matrices of 16-byte rows; LBO = distance between the two core matrices of one K step,
// SBO = distance between 8-row groups (cute/arch/mma_sm100_desc.hpp SmemDescriptor; version 1 = sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16, A = B = bf16 (K-major), D = f32, M = 128, N = 256 (InstrDescriptor bit layout of the same header)
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcCols >> 3) << 17) | ((uint32_t)(kTcRails >> 4) << 24);
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
        "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
        "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x -> bf16 hi and bf16 mid = bf16(x - hi), two values packed per word (element order = memory order)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& mid) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 m = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  mid = *reinterpret_cast<const uint32_t*>(&m);
}

// One rail of one half-band stage, fused arithmetic (the tolerance modes' hb_out_fast, one lane of it).
// e[k] = x[2(j-5+k)], o[k] = x[2(j-3+k)+1], ox = the three odd-phase samples before o[0] (for the block-end history).
struct RailHb { float e[5], o[3], ox[3]; };
__device__ __forceinline__ float rail_pair(RailHb& h, float xe, float xo) {
  const float s0 = h.e[0] + xe, s2 = h.e[1] + h.e[4], s4 = h.e[2] + h.e[3];
  const float y = fmaf(s4, HB_P4, fmaf(h.o[0], HB_P5, fmaf(s2, HB_P2, s0 * HB_P0)));
  h.ox[0] = h.ox[1]; h.ox[1] = h.ox[2]; h.ox[2] = h.o[0];
  h.e[0] = h.e[1]; h.e[1] = h.e[2]; h.e[2] = h.e[3]; h.e[3] = h.e[4]; h.e[4] = xe;
  h.o[0] = h.o[1]; h.o[1] = h.o[2]; h.o[2] = xo;
  return y;
}

// the tile sequence of this CTA, walked identically by every warp role
struct TcWalk {
  const TcSeg* segs;
  int si, si_end, nt, m_lo, m_hi, m0;
  bool seg_first;   // m0 is the first tile of its stretch
  __device__ __forceinline__ void open() {
    const TcSeg s = segs[si];
    nt = s.nt; m_lo = s.m_lo; m_hi = s.m_hi;
    m0 = m_lo > 0 ? m_lo - kTcRunIn : 0;
    seg_first = true;
  }
  __device__ __forceinline__ bool start(const TcParams& p) {
    segs = p.segs;
    si = p.cta_seg[blockIdx.x];
    si_end = p.cta_seg[blockIdx.x + 1];
    if (si >= si_end) return false;
    open();
    return true;
  }
  __device__ __forceinline__ bool next() {
    m0 += kTcCols;
    seg_first = false;
    if (m0 < m_hi) return true;
    if (++si >= si_end) return false;
    open();
    return true;
  }
};

template <int kTcFStages>
__global__ void __launch_bounds__(kTcThreads, 1) ddc_tc_kernel(const TcParams p) {
  constexpr int kTcBars = TcSmem<kTcFStages>::kBars;
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned char* xs = tsm;
  unsigned char* fsl = tsm + kTcXStages * kTcXStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fsl + kTcFStages * kTcFSlab);
  uint64_t* xfull = bars;
  uint64_t* xempty = xfull + kTcXStages;
  uint64_t* ffull = xempty + kTcXStages;
  uint64_t* fempty = ffull + kTcFStages;
  uint64_t* tfull = fempty + kTcFStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcXStages; ++i) { mbar_init(&xfull[i], 128); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < kTcFStages; ++i) { mbar_init(&ffull[i], 1); mbar_init(&fempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {   // the whole TMEM: two 256-column accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  TcWalk w;
  const bool any = w.start(p);

  if (warp < 4) {
    // ===== epilogue: one rail per thread: D (TMEM) x q[n_s] -> fused half-band stages -> stage-D rows =====
    const int rail = lane & 1;
    int slot = 0, col = 0, nd = 0;
    bool active = false;
    float2* xd = nullptr;
    RailHb hb[kDeepStages];
    // q for output mc + i: the oscillator state after idx1 + 32 i steps, idx1 = table index of the window's first sample
    // plus one: exact checkpoint (every 256 steps) times the unit rotation over the remainder. The remainder cycles
    // through 8 values per 32 outputs (the same 8 all along a stretch: a chunk advances the index by 1024; taken as
    // r0 + 32 k without wrapping at 256, so the checkpoint row is simply c0 + i / 8) and the checkpoint row advances
    // every 8 outputs; the four rows of the NEXT chunk are fetched while this one is processed.
    // (Indices past the table end only occur in the zones, whose outputs are replaced: clamped, never wrapped.)
    int idx1 = 1, r0 = -1;
    P2 PX[4], PY[4], SPX[4], SPY[4];   // the 8 rotation powers of a chunk, two consecutive outputs per packed register (S* = signed for this rail)
    float2 ckn[4];
    const float2* ck_col = p.ckpt;
    const float2* pw_col = p.pw;
    int tl = 0;
    if (any) do {
      const int a = tl & 1;
      if (w.seg_first) {
        slot = w.nt * kTcVfos + warp * 16 + (lane >> 1);
        active = slot < p.vfo_count;
        col = min(p.vfo_base + slot, p.vfo_pitch - 1);
        nd = max((int)p.vfo_D[col] - kFastStages, 0);
        xd = p.xd_rows[col];
        ck_col = p.ckpt + col;
        pw_col = p.pw + col;
        long long n_s = (p.block_abs + 32ll * w.m0 - kTcLead) % p.nco_len;
        if (n_s < 0) n_s += p.nco_len;
        idx1 = (int)n_s + 1;
        r0 = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) ckn[j] = __ldg(ck_col + (size_t)min((idx1 >> 8) + j, p.nck - 1) * p.vfo_pitch);
#pragma unroll
        for (int s = 0; s < kDeepStages; ++s) {
#pragma unroll
          for (int k = 0; k < 5; ++k) hb[s].e[k] = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) { hb[s].o[k] = 0.f; hb[s].ox[k] = 0.f; }
          if (w.m_lo == 0 && s < nd) {   // block start: the saved (shifted) history of this stage, this rail
            const float2* st = p.state_in + (size_t)(kFastStages + s) * kStateSlots * p.vfo_pitch + col;
#pragma unroll
            for (int k = 0; k < 5; ++k) { const float2 v = st[(size_t)k * p.vfo_pitch]; hb[s].e[k] = rail ? v.y : v.x; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float2 v = st[(size_t)(5 + k) * p.vfo_pitch]; hb[s].o[k] = rail ? v.y : v.x; }
          }
        }
      }
      mbar_wait(&tfull[a], (tl >> 1) & 1);
      tc_fence_after();
      const int zs = active ? slot : 0;          // lanes beyond the last VFO read VFO 0's zone samples (never stored)
      const float2* zone_col = p.zone + (size_t)(zs >> 5) * p.n_mid * 32 + (zs & 31);
#pragma unroll 1
      for (int cb = 0; cb < kTcCols; cb += 32) {
        const int mc = w.m0 + cb;                  // 32 consecutive stage-5 outputs, mc a multiple of 32
        if (mc >= w.m_hi) break;                   // uniform: nothing of this stretch beyond
        float d[32];
        tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kTcCols + cb), d);
        tc_ld_wait();
        if ((idx1 & 255) != r0) {                   // first chunk of a stretch, or the table restarted inside it
          r0 = idx1 & 255;
          const float sg = rail ? 1.f : -1.f;       // rail 0: z = dr fr - di fi; rail 1: z = di fr + dr fi
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float2 u0 = __ldg(pw_col + (size_t)(r0 + 64 * kk) * p.vfo_pitch), u1 = __ldg(pw_col + (size_t)(r0 + 64 * kk + 32) * p.vfo_pitch);
            PX[kk] = pack2(u0.x, u1.x); PY[kk] = pack2(u0.y, u1.y);
            SPX[kk] = pack2(sg * u0.x, sg * u1.x); SPY[kk] = pack2(sg * u0.y, sg * u1.y);
          }
        }
        float2 ckc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ckc[j] = ckn[j];
        idx1 += 1024;
        if (idx1 > p.nco_len) idx1 -= p.nco_len;
#pragma unroll
        for (int j = 0; j < 4; ++j) ckn[j] = __ldg(ck_col + (size_t)min((idx1 >> 8) + j, p.nck - 1) * p.vfo_pitch);
        const bool in_zone = mc < p.z0_end || (mc < p.z1_hi && mc + 32 > p.z1_lo);
        float z[32];
        // two consecutive outputs per packed instruction: f = ck * u^r (complex), z = own * Re f -+ partner * Im f
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float cx = ckc[j >> 2].x, cy = ckc[j >> 2].y;
          const P2 fr = fma2(bcast2(-cy), PY[j & 3], mul2s(PX[j & 3], cx));
          const P2 sf = fma2(bcast2(cy), SPX[j & 3], mul2s(SPY[j & 3], cx));
          const P2 own = pack2(d[2 * j], d[2 * j + 1]);
          const P2 oth = pack2(__shfl_xor_sync(0xffffffffu, d[2 * j], 1), __shfl_xor_sync(0xffffffffu, d[2 * j + 1], 1));
          unpack2(fma2(own, fr, mul2(oth, sf)), z[2 * j], z[2 * j + 1]);
        }
        if (in_zone) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = mc + i;
            if (m < p.z0_end || (m >= p.z1_lo && m < p.z1_hi)) { const float2 v = zone_col[(size_t)m * 32]; z[i] = rail ? v.y : v.x; }
          }
        }
        // fused half-band stages on this rail: 32 -> 16 -> 8 -> 4
        float y0[16], y1[8], y2[4];
#pragma unroll
        for (int j = 0; j < 16; ++j) y0[j] = rail_pair(hb[0], z[2 * j], z[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j) y1[j] = rail_pair(hb[1], y0[2 * j], y0[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) y2[j] = rail_pair(hb[2], y1[2 * j], y1[2 * j + 1]);
        const bool keep = mc >= w.m_lo;            // run-in outputs are dropped (m_lo is a multiple of 32)
        // the even lane of a pair writes (re, im); lanes of a warp may differ in their stage count
        const bool wr = keep && active && rail == 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float o = __shfl_xor_sync(0xffffffffu, y2[j], 1); if (wr && nd == 3) xd[(mc >> 3) + j] = make_float2(y2[j], o); }
        if (__any_sync(0xffffffffu, nd < 3)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float o = __shfl_xor_sync(0xffffffffu, y1[j], 1); if (wr && nd == 2) xd[(mc >> 2) + j] = make_float2(y1[j], o); }
#pragma unroll
          for (int j = 0; j < 16; ++j) { const float o = __shfl_xor_sync(0xffffffffu, y0[j], 1); if (wr && nd == 1) xd[(mc >> 1) + j] = make_float2(y0[j], o); }
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float o = __shfl_xor_sync(0xffffffffu, z[j], 1); if (wr && nd == 0) xd[mc + j] = make_float2(z[j], o); }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[a]);
      // end of the block's stream: the shifted history the fused stages start the next block from (see boundary_role)
      if (w.m0 + kTcCols >= w.m_hi && w.m_hi == p.n_mid && active) {
#pragma unroll
        for (int s = 0; s < kDeepStages; ++s) {
          if (s < nd) {
            float* st = reinterpret_cast<float*>(p.state_out + (size_t)(kFastStages + s) * kStateSlots * p.vfo_pitch + col) + rail;
            const size_t pitch = (size_t)p.vfo_pitch * 2;
            st[0] = hb[s].ox[0];               // x[n-11]
            st[pitch] = hb[s].ox[1];           // x[n-9]
            st[2 * pitch] = hb[s].ox[2];       // x[n-7]
            st[3 * pitch] = hb[s].o[0];        // x[n-5]
            st[4 * pitch] = hb[s].o[1];        // x[n-3]
#pragma unroll
            for (int k = 0; k < 3; ++k) st[(size_t)(5 + k) * pitch] = hb[s].e[2 + k];   // x[n-6], x[n-4], x[n-2]
          }
        }
      }
      ++tl;
    } while (w.next());
  } else if (warp < 8) {
    // ===== X producer: rows of the raw block -> bf16 hi / mid, [chunk][row][16 B] =====
    // A tile needs rows m0-10 .. m0+255. The 256 body rows come from global memory, two rows per thread with all 32
    // 16-byte loads of both in flight at once (the block may live on another GPU: one NVLink round trip per tile, not
    // three); the 10 halo rows in front are the last 10 body rows of the previous tile of the stretch, which still sit,
    // converted, in the other stage.
    const int ptid = threadIdx.x - 128;
    int tl = 0;
    auto fetch_row = [&](int rho, float4 (&r)[16]) {
      const bool valid = rho >= 0 && rho < p.n_mid;
      const int g0 = (valid ? rho : 0) * 32;
      const int sl = p.raw.n_slices > 1 ? g0 / p.raw.slice_len : 0;
      const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(p.raw.slice[sl]) + (g0 - sl * p.raw.slice_len));
#pragma unroll
      for (int c = 0; c < 16; ++c) r[c] = valid ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_row = [&](unsigned char* hi_base, int i, const float4 (&r)[16]) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 h, mdl;
        split_pair(r[2 * c].x, r[2 * c].y, h.x, mdl.x);
        split_pair(r[2 * c].z, r[2 * c].w, h.y, mdl.y);
        split_pair(r[2 * c + 1].x, r[2 * c + 1].y, h.z, mdl.z);
        split_pair(r[2 * c + 1].z, r[2 * c + 1].w, h.w, mdl.w);
        *reinterpret_cast<uint4*>(hi_base + (c * kTcRows + i) * 16) = h;
        *reinterpret_cast<uint4*>(hi_base + kTcXPart + (c * kTcRows + i) * 16) = mdl;
      }
    };
    if (any) do {
      const int xsi = tl & 1;
      const int row0 = w.m0 - kTcBack;
      asm volatile("bar.sync 1, 128;" ::: "memory");   // every producer thread is through with the previous tile (its rows are read below)
      mbar_wait(&xempty[xsi], ((tl >> 1) & 1) ^ 1);
      unsigned char* hi_base = xs + xsi * kTcXStage;
      {
        float4 ra[16], rb[16];
        fetch_row(row0 + kTcBack + ptid, ra);
        fetch_row(row0 + kTcBack + 128 + ptid, rb);
        store_row(hi_base, kTcBack + ptid, ra);
        store_row(hi_base, kTcBack + 128 + ptid, rb);
      }
      if (w.seg_first) {
        if (ptid < kTcBack) {
          float4 r[16];
          fetch_row(row0 + ptid, r);
          store_row(hi_base, ptid, r);
        }
      } else if (ptid < 8 * kTcBack) {
        const unsigned char* prev = xs + (xsi ^ 1) * kTcXStage;
        const int row = ptid >> 3, c = ptid & 7;
        *reinterpret_cast<uint4*>(hi_base + (c * kTcRows + row) * 16) = *reinterpret_cast<const uint4*>(prev + (c * kTcRows + kTcCols + row) * 16);
        *reinterpret_cast<uint4*>(hi_base + kTcXPart + (c * kTcRows + row) * 16) = *reinterpret_cast<const uint4*>(prev + kTcXPart + (c * kTcRows + kTcCols + row) * 16);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
      mbar_arrive(&xfull[xsi]);
      ++tl;
    } while (w.next());
  } else if (warp == 8) {
    // ===== filter slabs: one 8 KB bulk copy per k-step =====
    if (lane == 0 && any) {
      int fsi = 0;
      uint32_t fph = 1;                                   // parity to wait for on the empty barrier of slot fsi
      do {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.filt) + (size_t)w.nt * kTcKSteps * kTcFSlab;
#pragma unroll 4
        for (int ks = 0; ks < kTcKSteps; ++ks) {
          mbar_wait(&fempty[fsi], fph);
          mbar_expect_tx(&ffull[fsi], kTcFSlab);
          tma_bulk_g2s(fsl + fsi * kTcFSlab, src + (size_t)ks * kTcFSlab, kTcFSlab, &ffull[fsi]);
          if (++fsi == kTcFStages) { fsi = 0; fph ^= 1; }
        }
      } while (w.next());
    }
  } else {
    // ===== MMA issuer =====
    // One thread issues; its loop is the critical path of the kernel (a dependent instruction costs it ~5 cycles, an MMA
    // runs 128), so everything per k-step is a compile-time constant (the 40 steps are unrolled) or an increment.
    if (lane == 0 && any) {
      int fsi = 0;
      uint32_t fph = 0;                                   // parity to wait for on the full barrier of slot fsi
      int tl = 0;
      const uint64_t fdesc0 = tc_desc(smem_u32(fsl), kTcRails * 16, 128);
      do {
        const int a = tl & 1, xsi = tl & 1;
        mbar_wait(&tempty[a], ((tl >> 1) & 1) ^ 1);
        mbar_wait(&xfull[xsi], (tl >> 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(a * kTcCols);
        const uint64_t xdesc_hi = tc_desc(smem_u32(xs + xsi * kTcXStage), kTcRows * 16, 128);
        const uint64_t xdesc_mid = xdesc_hi + (uint64_t)(kTcXPart >> 4);
#pragma unroll
        for (int ks = 0; ks < kTcKSteps; ++ks) {
          mbar_wait(&ffull[fsi], fph);
          tc_fence_after();
          constexpr int kDummy = 0; (void)kDummy;
          const int s = 8 + 8 * ks;                       // first sample of the k-step, counted from row (m - 10)
          const uint64_t xo = (uint64_t)(((s & 31) >> 2) * kTcRows + (s >> 5));   // in 16-byte units: added to the address field
          const uint64_t ah = fdesc0 + (uint64_t)(fsi * (kTcFSlab >> 4));
          tc_mma(d, ah, xdesc_hi + xo, ks > 0);
          if (ks >= kTcCorrLo && ks < kTcCorrHi) {   // the two correction products, where the taps are large enough to matter
            tc_mma(d, ah, xdesc_mid + xo, 1);
            tc_mma(d, ah + (uint64_t)(kTcFPart >> 4), xdesc_hi + xo, 1);
          }
          tc_commit(&fempty[fsi]);
          if (++fsi == kTcFStages) { fsi = 0; fph ^= 1; }
        }
        tc_commit(&xempty[xsi]);
        tc_commit(&tfull[a]);
        ++tl;
      } while (w.next());
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- finalize-time tables ----
// filt[nt][ks][part][chunk][rail][8 bf16]: rail = 2 * (VFO within the tile) + re/im; element k of the k-step: sample
// 8 ks + k / 2 of the padded window, component k & 1. Rail 0 (real): (Gr, -Gi); rail 1 (imaginary): (Gi, Gr);
// G = g[t' - 2] u^t'.
__global__ void tc_build_filters_kernel(const float2* __restrict__ rot, const double* __restrict__ g, int vfo_base, int vfo_count,
                                        int n_ntiles, uint4* __restrict__ filt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;   // over n_ntiles * kTcKSteps * kTcRails
  if (n >= n_ntiles * kTcKSteps * kTcRails) return;
  const int rail_n = n % kTcRails, ks = (n / kTcRails) % kTcKSteps, nt = n / (kTcRails * kTcKSteps);
  const int slot = nt * kTcVfos + (rail_n >> 1), rail = rail_n & 1;
  float val[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) val[k] = 0.f;
  if (slot < vfo_count) {
    const float2 rt = rot[vfo_base + slot];
    const double theta = atan2((double)rt.y, (double)rt.x);
    for (int k = 0; k < 16; ++k) {
      const int tp = 8 * ks + (k >> 1), tg = tp - 2;
      if (tg < 0 || tg >= kTcTaps) continue;
      double sn, cs;
      sincos(theta * (double)tp, &sn, &cs);
      const double gr = g[tg] * cs, gi = g[tg] * sn;
      val[k] = (float)(rail == 0 ? ((k & 1) ? -gi : gr) : ((k & 1) ? gr : gi));
    }
  }
  uint32_t hi[8], mid[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) split_pair(val[2 * k], val[2 * k + 1], hi[k], mid[k]);
  uint4* slab = filt + (size_t)(nt * kTcKSteps + ks) * (kTcFSlab / 16);
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    slab[ch * kTcRails + rail_n] = make_uint4(hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
    slab[(kTcFPart / 16) + ch * kTcRails + rail_n] = make_uint4(mid[4 * ch], mid[4 * ch + 1], mid[4 * ch + 2], mid[4 * ch + 3]);
  }
}

// pw[r][col] = u^r, u = rot / |rot|
__global__ void tc_build_pw_kernel(const float2* __restrict__ rot, int vfo_pitch, float2* __restrict__ pw) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (col >= vfo_pitch) return;
  const float2 rt = rot[col];
  const double theta = atan2((double)rt.y, (double)rt.x);
  double sn, cs;
  sincos(theta * (double)r, &sn, &cs);
  pw[(size_t)r * vfo_pitch + col] = make_float2((float)cs, (float)sn);
}

}  // namespace aeroddc
