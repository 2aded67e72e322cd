// aero-ddc-b200: device code for the multi-VFO digital down-converter bank (sm_100a).
//
// Replaces, for every VFO at once, the reference's per-VFO chain
//   vfo::process mix loop        /root/reference/publish/vfo.cpp:155-161
//   Oscillator (NCO recurrence)  /root/reference/publish/oscillator.cpp:4-39
//   HalfBandDecimator::decimate  /root/reference/publish/halfbanddecimator.cpp:35-60
//   FIR half-band queue kernels  /root/reference/publish/dsp.cpp:102-172
//
// Design: one thread carries ONE VFO through time; its I and Q rails sit in the two
// lanes of the Blackwell packed-FP32 instructions (FMUL2/FFMA2, PTX mul/fma.rn.f32x2). Every multiply
// and add of the reference is issued un-fused and in the reference's order, so results are
// bit-identical to the CPU chain; packing halves the issue slots per flop, which leaves room for the
// shared-memory broadcast loads and register moves next to a saturated FP32 pipe.
//
// Two kernels share the half-band cascade:
//   ddc_main_kernel  unpack + NCO + mix + the first min(D, 5) half-band stages, all in registers. A CTA is ONE warp
//                    (32 VFOs) with its own TMA ring of raw tiles; time is cut into segments (each warmed up over the
//                    10*(2^5-1) samples before it) and every segment into chained parts whose state is handed from
//                    CTA to CTA through HBM. Output: the stage-min(D,5) stream, [time][VFO] (coalesced).
//   ddc_deep_kernel  the remaining D-5 stages (rate <= Fs/32) from that stream: one warp = 32 VFOs x one time
//                    range, history in registers, warmed up over 10*(2^(D-5)-1) of its own input samples.
// The raw block may be given as up to 8 equal slices living in different GPUs' HBM (peer memory over NVLink): the TMA
// tile loads pull each tile from the slice that holds it, so the exchange is fused with the compute tile by tile.
// Only bank.cu includes this header (compute-only translation unit; no host-side state here).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aeroddc {

constexpr int kThreads = 32;           // threads per CTA of the main kernel: one warp, each thread carries one VFO
constexpr int kVfoPerCta = kThreads;
constexpr int kChunk = 32;             // input samples per unrolled inner step
constexpr int kTile = 256;             // input samples per shared-memory tile
constexpr int kMaxStages = 8;          // hdecimator[8], vfo.h:63
constexpr int kFastStages = 5;         // half-band stages of the main kernel (registers); the rest run in the deep kernel
constexpr int kDeepStages = kMaxStages - kFastStages;
constexpr int kStateSlots = 8;         // per stage: 5 even-phase + 3 odd-phase history samples
constexpr int kNcoStride = 256;        // NCO checkpoint spacing (samples)
constexpr int kHandSlots = kFastStages * kStateSlots + 1;   // half-band history of the register stages + oscillator state
constexpr int kCtasPerSm = 16;         // 16 resident warps per SM at <= 128 registers per thread
constexpr int kMaxSlices = 8;          // a raw block may be spread over up to 8 GPUs

enum { FMT_CU8 = 0, FMT_CS16 = 1, FMT_CF32 = 2 };

// ---------------------------------------------------------------------------------------------
// packed 2 x fp32 arithmetic: lane 0 = I (real), lane 1 = Q (imaginary) of ONE VFO.
// All round-to-nearest, never fused.
// ---------------------------------------------------------------------------------------------
struct P2 { unsigned long long v; };

__device__ __forceinline__ P2 pack2(float a, float b) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(P2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ P2 bcast2(float s) { return pack2(s, s); }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { P2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d; }
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with -fmad=false (the scalar
// .rn forms are left alone), which would change the reference's bits. So additions are issued as
// fma(a, 1.0f, b), with 1.0f arriving as a kernel parameter (a uniform register) so that ptxas can
// neither simplify it back to an add nor fuse across it. The result is exactly round(a + b).
// FFMA2, FMUL2 and FADD2 share one pipe and one rate.
struct Ones { P2 one; };
__device__ __forceinline__ P2 add2(const Ones& k, P2 a, P2 b) { return fma2(a, k.one, b); }
// A true FADD2, for sums whose operands are never products (the tap-pair sums of the half-band filter): there is no
// multiply ptxas could fuse it with, and it reads two register pairs instead of three operands.
__device__ __forceinline__ P2 addp(P2 a, P2 b) { P2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ P2 mul2s(P2 a, float s) { return mul2(a, bcast2(s)); }   // FMUL2 R, R.F32x2, R.F32
__device__ __forceinline__ P2 pzero() { P2 z; z.v = 0ull; return z; }

// half-band taps actually used by the reference (halfbanddecimator.h:84-87, 11-tap set)
#define HB_P0 0.0060431029837374152f
#define HB_P2 (-0.049372515458761493f)
#define HB_P4 0.29332944952052842f
#define HB_P5 0.5f

// Per-VFO oscillator constants: rotA = (c, d), rotB = (-d, c) with (c, d) = ((float)cos, (float)sin).
struct Rot { P2 a, b; };

// One NCO step: v *= rot; v *= 1.95f - |v|^2   (oscillator.cpp:19-24), complex product as GCC
// evaluates std::complex<float>: (a*c - b*d, a*d + b*c). Here (a*c, a*d) + (b*(-d), b*c): negating d
// is exact, so lane 0 is round(round(a*c) - round(b*d)) as in the reference.
__device__ __forceinline__ void nco_step(const Ones& k, float& a, float& b, const Rot& rot) {
  const P2 n = add2(k, mul2s(rot.a, a), mul2s(rot.b, b));
  float r2, i2;
  unpack2(mul2(n, n), r2, i2);
  const float nm = __fsub_rn(1.95f, __fadd_rn(r2, i2));
  unpack2(mul2s(n, nm), a, b);
}

// mix: osc * sample (vfo.cpp:157) = (a*c - b*d, a*d + b*c), a,b = oscillator, c,d = raw sample:
// (c, d)*a + (-d, c)*b. ptxas folds the swap and the sign into FMUL2 operand modifiers (.LO_HI.NP).
__device__ __forceinline__ P2 mix(const Ones& k, float a, float b, const float2& s) {
  return add2(k, mul2s(pack2(s.x, s.y), a), mul2s(pack2(-s.y, s.x), b));
}

// Per-stage history in polyphase form: e[k] = x[2(j-5+k)], k<5 (even-phase samples) and
// o[k] = x[2(j-3+k)+1], k<3 (odd-phase), j = index of the next output. Each entry is (I, Q).
struct HbState { P2 e[5]; P2 o[3]; };

// y[j] = ((p0*(w0+w10) + p2*(w2+w8)) + p4*(w4+w6)) + p5*w5, w[t] = x[2j-10+t]  (dsp.cpp:141-147),
// both rails at once.
// The centre tap is exactly 0.5: fl(0.5*w5) is exact for every normal w5 (a halving changes only the exponent), so
// fl(acc + fl(0.5*w5)) == fma(0.5, w5, acc) bit for bit. The two forms can differ only when w5 is so small that its
// half is not representable (|w5| < 2^-125) AND acc is subnormal too, i.e. on signals below 1e-37 of full scale; such
// values end as 0 in every int16 / int8 payload either way (tests/test_gpu_parity.py::test_subnormal_inputs_...).
// -DAERODDC_HB_CENTER_MUL restores the separate multiply (one more FMUL2 per half-band output).
__device__ __forceinline__ P2 hb_out(const Ones& k, P2 e0, P2 e1, P2 e2, P2 e3, P2 e4, P2 xe, P2 o0) {
  P2 s0 = addp(e0, xe), s2 = addp(e1, e4), s4 = addp(e2, e3);
  P2 m0 = mul2s(s0, HB_P0), m2 = mul2s(s2, HB_P2), m4 = mul2s(s4, HB_P4);
#ifndef AERODDC_HB_CENTER_MUL
  return fma2(o0, bcast2(HB_P5), add2(k, add2(k, m0, m2), m4));
#else
  return add2(k, add2(k, add2(k, m0, m2), m4), mul2s(o0, HB_P5));
#endif
}

// ---------------------------------------------------------------------------------------------
// Tolerance mode (AERODDC_MODE_FAST): the same chain with fused multiply-adds. It is NOT bit-identical
// to the reference; it stays within BASELINE.json's tolerance (max |err| <= 1e-4 of full scale, error
// SNR >= 80 dB on signals of normal level) at ~55 % of the exact mode's FP32 work:
//  * oscillator: exact checkpoints every kNcoStride samples (the same table as the exact mode); in
//    between, one fused complex rotation per sample (the recurrence's amplitude factor is 1 +- 1e-7 in
//    steady state). The ~200-sample amplitude transient after each table restart and the restart
//    itself run the full recurrence (fused).
//  * mix and half-band taps: fma chains instead of separately rounded products.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void nco_rotate_fast(float& a, float& b, const Rot& rot) {
  unpack2(fma2(rot.b, bcast2(b), mul2s(rot.a, a)), a, b);
}
__device__ __forceinline__ void nco_step_fused(float& a, float& b, const Rot& rot) {
  const P2 n = fma2(rot.b, bcast2(b), mul2s(rot.a, a));
  float nr, ni;
  unpack2(n, nr, ni);
  const float nm = 1.95f - fmaf(ni, ni, nr * nr);
  unpack2(mul2s(n, nm), a, b);
}
__device__ __forceinline__ P2 mix_fast(float a, float b, const float2& s) {
  return fma2(pack2(-s.y, s.x), bcast2(b), mul2s(pack2(s.x, s.y), a));
}
__device__ __forceinline__ P2 hb_out_fast(const Ones& k, P2 e0, P2 e1, P2 e2, P2 e3, P2 e4, P2 xe, P2 o0) {
  const P2 s0 = addp(e0, xe), s2 = addp(e1, e4), s4 = addp(e2, e3);
  return fma2(s4, bcast2(HB_P4), fma2(o0, bcast2(HB_P5), fma2(s2, bcast2(HB_P2), mul2s(s0, HB_P0))));
}

// feed the even-phase sample x[2j] and return output j; then feed the odd-phase sample x[2j+1]
template <bool FAST>
__device__ __forceinline__ P2 hb_even(const Ones& k, HbState& h, P2 xe) {
  const P2 y = FAST ? hb_out_fast(k, h.e[0], h.e[1], h.e[2], h.e[3], h.e[4], xe, h.o[0])
                    : hb_out(k, h.e[0], h.e[1], h.e[2], h.e[3], h.e[4], xe, h.o[0]);
  h.e[0] = h.e[1]; h.e[1] = h.e[2]; h.e[2] = h.e[3]; h.e[3] = h.e[4]; h.e[4] = xe;
  return y;
}
__device__ __forceinline__ void hb_odd(HbState& h, P2 xo) { h.o[0] = h.o[1]; h.o[1] = h.o[2]; h.o[2] = xo; }
// the same, also remembering the three odd-phase samples that fall out of h.o (deep kernel: saved block-end history)
template <bool FAST>
__device__ __forceinline__ P2 deep_pair(const Ones& k, HbState& h, P2 (&ox)[3], P2 xe, P2 xo);
// consume the pair (x[2j], x[2j+1]) and return output j
template <bool FAST>
__device__ __forceinline__ P2 hb_pair(const Ones& k, HbState& h, P2 xe, P2 xo) {
  const P2 y = hb_even<FAST>(k, h, xe);
  hb_odd(h, xo);
  return y;
}
template <bool FAST>
__device__ __forceinline__ P2 deep_pair(const Ones& k, HbState& h, P2 (&ox)[3], P2 xe, P2 xo) {
  ox[0] = ox[1]; ox[1] = ox[2]; ox[2] = h.o[0];
  return hb_pair<FAST>(k, h, xe, xo);
}

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct RawBlock {
  const void* slice[kMaxSlices];   // slice i holds samples [i*slice_len, (i+1)*slice_len) of the block, format FMT
  int n_slices;
  int slice_len;                   // complex samples per slice (even; the last slice may be shorter)
};

struct MainParams {
  RawBlock raw;              // the block on the device(s), B complex samples
  const float2* ckpt;        // [nck][vfo_pitch] NCO state after k*kNcoStride steps from (1,0)
  const float2* rot;         // [vfo_pitch] per-VFO rotation (cos, sin) as floats
  const float2* qlast;       // [vfo_pitch] q[L-1], the value sample 0 is mixed with
  const float2* state_in;    // [kMaxStages][kStateSlots][vfo_pitch] history at the start of this block
  float2* state_out;         // same layout, history for the start of the next block
  float2* mid;               // [ngroups][B >> DA][32] output: the stage-DA stream of this block, one contiguous run of 256-byte
  int n_mid;                 // rows per 32-VFO group (n_mid = B >> DA rows), so that writer and reader both stream sequentially.
  float2* const* xd_rows;    // mid == NULL (no VFO of the launch has a deep stage and the bank runs its kernels in order):
                             // the samples go straight into the per-VFO stage-D rows ([vfo_pitch] pointers)
  long long block_abs;       // absolute index of the block's first sample
  int vfo_pitch;             // padded VFO count (row pitch of ckpt/rot/state/mid)
  int vfo_base, vfo_count;   // VFO slice handled by this launch (all share DA = min(D, 5))
  int DA;                    // half-band stages of this kernel
  int B, S, W, nseg;         // block length, segment length, warm-up length, segments per block
  int nco_len;               // L = (int)Fs, the oscillator table length
  float one;                 // 1.0f (see add2)
  int transient;             // tolerance mode: table indices below this use the full recurrence
  int nck;                   // rows of ckpt
  // A segment of one VFO group is a CHAIN of consecutive parts of P samples, processed by different CTAs that hand the
  // filter and oscillator state on through `hand`. Which CTA runs which part is decided at run time by a FIFO of ready
  // chains (see ddc_main_kernel): a CTA that finishes a part appends its chain to the queue, a CTA that starts takes
  // the chain that has waited longest. Every chain therefore advances at the average pace of all SM slots, whatever the
  // speed of the individual slot (the warp scheduler favours some resident warps), and all chains end together.
  int P, ngroups, nchains;   // part length; VFO groups of 32; chains = ngroups * nseg
  int* sched;                // [0] items taken, [1] queue tail, [2 .. 2+nchains) parts completed per chain, then the queue
                             // (chain + 1 per entry); zeroed by the host before every launch
  float2* hand;              // [nchains][kHandSlots][kThreads]
  int* err;                  // set to 1 if a CTA gave up waiting for its queue entry (should never happen)
  // A launch may cover only a stretch of the block (tensor mode keeps the head of every block and the zone after an
  // oscillator restart on this kernel): segment k starts at seg_off + k*S; cold0 = segment 0 does not continue the saved
  // block history but runs in over W samples like the others; nbound = boundary CTAs of this launch (ngroups or 0).
  int seg_off, cold0, nbound;
};

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int FMT> struct RawBytes { static constexpr int v = FMT == FMT_CU8 ? 2 : (FMT == FMT_CS16 ? 4 : 8); };

// raw -> float, exactly as the CPU side of this project defines it:
// cu8: (u8 - 127.4f) / 128.0f ; cs16: s16 / 32768.0f ; cf32: as is.
__device__ __forceinline__ float cvt_u8(unsigned v) { return __fdiv_rn(__fsub_rn((float)v, 127.4f), 128.0f); }
__device__ __forceinline__ float cvt_s16(int v) { return __fdiv_rn((float)v, 32768.0f); }
template <int FMT> __device__ __forceinline__ float2 load_raw(const void* base, size_t n) {
  if (FMT == FMT_CU8) { const uchar2 v = reinterpret_cast<const uchar2*>(base)[n]; return make_float2(cvt_u8(v.x), cvt_u8(v.y)); }
  if (FMT == FMT_CS16) { const short2 v = reinterpret_cast<const short2*>(base)[n]; return make_float2(cvt_s16(v.x), cvt_s16(v.y)); }
  return reinterpret_cast<const float2*>(base)[n];
}
// sample n of a (possibly sliced) block
template <int FMT> __device__ __forceinline__ float2 load_raw_block(const RawBlock& rb, int n) {
  const int s = rb.n_slices > 1 ? n / rb.slice_len : 0;
  return load_raw<FMT>(rb.slice[s], (size_t)(n - s * rb.slice_len));
}

// ---------------------------------------------------------------------------------------------
// The register-resident cascade: kChunk input samples -> kChunk >> NF outputs.
// SPECIAL handles the rare events inside a chunk: oscillator table wrap (index reaches L: restart from (1,0),
// oscillator.cpp:31-39), absolute sample 0 (mixed with q[L-1], because the constructor leaves _vector at the last
// table entry, oscillator.cpp:12-27) and, in tolerance mode, a checkpoint stride boundary that does not fall on a
// chunk start (possible after a table wrap when L is not a multiple of the chunk).
// ---------------------------------------------------------------------------------------------
template <int NF, bool SPECIAL, bool FAST>
__device__ __forceinline__ void fast_chunk(const Ones& k1, float& oa, float& ob, const Rot& rot,
                                           HbState (&hb)[kFastStages > 0 ? kFastStages : 1],
                                           const float2* __restrict__ tile, P2 (&out)[kChunk >> NF],
                                           int& idx, int nco_len, long long n_abs, float qa, float qb,
                                           const float2* __restrict__ ckpt_col, int vfo_pitch, int transient) {
  P2 x0[kChunk];
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const float2 s = tile[i];   // one raw sample (c, d), broadcast to the warp
    if (SPECIAL) {
      if (idx == nco_len) { idx = 0; oa = 1.0f; ob = 0.0f; }
      if (FAST && idx >= transient && (idx % kNcoStride) == 0) {   // snap back to the exact table
        const float2 c0 = ckpt_col[(size_t)(idx / kNcoStride) * vfo_pitch];
        oa = c0.x; ob = c0.y;
      }
    }
    if (!FAST) nco_step(k1, oa, ob, rot);
    else if (SPECIAL) { if (idx < transient) nco_step_fused(oa, ob, rot); else nco_rotate_fast(oa, ob, rot); }
    else nco_rotate_fast(oa, ob, rot);
    float a = oa, b = ob;
    if (SPECIAL) { if (n_abs + i == 0) { a = qa; b = qb; } idx++; }
    x0[i] = FAST ? mix_fast(a, b, s) : mix(k1, a, b, s);
  }
  if (!SPECIAL) idx += kChunk;
  // NF half-band stages, compacting in place (output j overwrites slot j after slots 2j, 2j+1 were read)
#pragma unroll
  for (int s = 0; s < NF; ++s) {
#pragma unroll
    for (int j = 0; j < (kChunk >> (s + 1)); ++j) x0[j] = hb_pair<FAST>(k1, hb[s], x0[2 * j], x0[2 * j + 1]);
  }
#pragma unroll
  for (int i = 0; i < (kChunk >> NF); ++i) out[i] = x0[i];
}

__device__ __forceinline__ P2 load_p2(const float2* p) { const float2 v = *p; return pack2(v.x, v.y); }
__device__ __forceinline__ P2 load_p2_cg(const float2* p) { const float2 v = __ldcg(p); return pack2(v.x, v.y); }   // L2 only: data another CTA wrote during this launch
__device__ __forceinline__ void store_p2(float2* p, P2 c) { float a, b; unpack2(c, a, b); *p = make_float2(a, b); }

// ---------------------------------------------------------------------------------------------
// Boundary role: recompute, from the last Wb samples of this block, the history the register stages of every VFO will
// see at the start of the NEXT block. The reference re-seeds each half-band queue with queue[n-1 .. n+9] instead of the
// last 11 samples (FIR::FIRQueueBackToFront, dsp.cpp:163-172): history index l<0 of the next block is this block's
// sample n-1+l, so the newest sample x[n-1] is dropped and the window is one sample older than the true one. In
// polyphase terms the next block's even-phase history is this block's odd samples x[n-11], x[n-9] .. x[n-3] and its
// odd-phase history is this block's even samples x[n-6], x[n-4], x[n-2] (n = the stage's input count, even).
// One extra CTA per 32 VFOs runs the ordinary register cascade over the last 11*2^NF samples (zero history in front:
// 10*(2^NF - 1) samples of run-in, then 12 valid inputs of the last stage) and keeps the last 12 inputs of every stage.
// The stages beyond the register cascade get their history from the deep kernel, which sees the end of their streams.
// ---------------------------------------------------------------------------------------------
template <int M> __device__ __forceinline__ void keep_last12(P2 (&h)[12], const P2* in) {   // M new inputs, oldest first
  if (M >= 12) {
#pragma unroll
    for (int k = 0; k < 12; ++k) h[k] = in[M - 12 + k];
  } else {
#pragma unroll
    for (int k = 0; k < 12 - M; ++k) h[k] = h[k + M];
#pragma unroll
    for (int k = 0; k < M; ++k) h[12 - M + k] = in[k];
  }
}

template <int NF, int FMT, bool FAST>
__device__ void boundary_role(const MainParams& p, int vfo, bool active, float2* tile) {
  if (NF == 0) return;                                  // no half-band stage in this kernel: nothing to save
  constexpr int NS = NF > 0 ? NF : 1;
  P2 last[NS][12];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int k = 0; k < 12; ++k) last[s][k] = pzero();
  HbState hb[kFastStages > 0 ? kFastStages : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int k = 0; k < 5; ++k) hb[s].e[k] = pzero();
#pragma unroll
    for (int k = 0; k < 3; ++k) hb[s].o[k] = pzero();
  }
  Ones k1; k1.one = bcast2(p.one);
  const float2 r = p.rot[vfo];
  Rot rot; rot.a = pack2(r.x, r.y); rot.b = pack2(-r.y, r.x);
  const float2 ql = p.qlast[vfo];
  const float2* ckpt_col = p.ckpt + vfo;
  const int Wb = ((11 << NF) + kChunk - 1) / kChunk * kChunk;
  const int start = p.B - Wb;                           // in-block index of the first run-in sample (>= 0: blocks hold >= 20*2^D samples)
  long long n_abs = p.block_abs + start;
  int idx = (int)(n_abs % p.nco_len);
  float oa, ob;
  {
    const int ck = idx / kNcoStride, rem = idx % kNcoStride;
    const float2 c = ckpt_col[(size_t)ck * p.vfo_pitch];
    oa = c.x; ob = c.y;
    for (int i = 0; i < rem; ++i) nco_step(k1, oa, ob, rot);
  }
#pragma unroll 1
  for (int c = 0; c < Wb; c += kChunk) {
    __syncwarp();
    tile[threadIdx.x] = load_raw_block<FMT>(p.raw, start + c + (int)threadIdx.x);   // kThreads == kChunk: one sample per lane
    __syncwarp();
    // the chunk, as fast_chunk<NF, true, FAST> runs it, with every stage's inputs kept
    P2 x0[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const float2 sm = tile[i];
      if (idx == p.nco_len) { idx = 0; oa = 1.0f; ob = 0.0f; }
      if (FAST && idx >= p.transient && (idx % kNcoStride) == 0) {
        const float2 c0 = ckpt_col[(size_t)(idx / kNcoStride) * p.vfo_pitch];
        oa = c0.x; ob = c0.y;
      }
      if (!FAST) nco_step(k1, oa, ob, rot);
      else if (idx < p.transient) nco_step_fused(oa, ob, rot);
      else nco_rotate_fast(oa, ob, rot);
      float a = oa, b = ob;
      if (n_abs + i == 0) { a = ql.x; b = ql.y; }
      idx++;
      x0[i] = FAST ? mix_fast(a, b, sm) : mix(k1, a, b, sm);
    }
    n_abs += kChunk;
    keep_last12<kChunk>(last[0], x0);
    if (NF > 1) {
      P2 y0[kChunk / 2];
#pragma unroll
      for (int j = 0; j < kChunk / 2; ++j) y0[j] = hb_pair<FAST>(k1, hb[0], x0[2 * j], x0[2 * j + 1]);
      keep_last12<kChunk / 2>(last[NS > 1 ? 1 : 0], y0);
      if (NF > 2) {
        P2 y1[kChunk / 4];
#pragma unroll
        for (int j = 0; j < kChunk / 4; ++j) y1[j] = hb_pair<FAST>(k1, hb[1], y0[2 * j], y0[2 * j + 1]);
        keep_last12<kChunk / 4>(last[NS > 2 ? 2 : 0], y1);
        if (NF > 3) {
          P2 y2[kChunk / 8];
#pragma unroll
          for (int j = 0; j < kChunk / 8; ++j) y2[j] = hb_pair<FAST>(k1, hb[2], y1[2 * j], y1[2 * j + 1]);
          keep_last12<kChunk / 8>(last[NS > 3 ? 3 : 0], y2);
          if (NF > 4) {
            P2 y3[kChunk / 16];
#pragma unroll
            for (int j = 0; j < kChunk / 16; ++j) y3[j] = hb_pair<FAST>(k1, hb[3], y2[2 * j], y2[2 * j + 1]);
            keep_last12<kChunk / 16>(last[NS > 4 ? 4 : 0], y3);
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      float2* st = p.state_out + (size_t)s * kStateSlots * p.vfo_pitch + vfo;
#pragma unroll
      for (int k = 0; k < 5; ++k) store_p2(st + (size_t)k * p.vfo_pitch, last[s][1 + 2 * k]);        // x[n-11], x[n-9] .. x[n-3]
#pragma unroll
      for (int k = 0; k < 3; ++k) store_p2(st + (size_t)(5 + k) * p.vfo_pitch, last[s][6 + 2 * k]);  // x[n-6], x[n-4], x[n-2]
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Main kernel. 1-D grid of one-warp CTAs: ngroups boundary CTAs, then Q * nseg * ngroups segment-part CTAs.
// Shared memory: raw tile ring (TMA bulk copies) | one converted float tile (cu8 / cs16 only; cf32 tiles are read
// where the TMA put them) | mbarriers.
// ---------------------------------------------------------------------------------------------
template <int FMT> struct TileSmem {
  static constexpr int kRawStages = 3;
  static constexpr int kRawBytes = kTile * RawBytes<FMT>::v;
  static constexpr int kCvtBytes = FMT == FMT_CF32 ? 0 : kTile * 8;
  static constexpr int kTotal = kRawStages * kRawBytes + kCvtBytes + 64;
};

template <int NF, int FMT, bool FAST>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) ddc_main_kernel(const MainParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  using TS = TileSmem<FMT>;
  unsigned char* raw = smem;
  float2* cvt = reinterpret_cast<float2*>(smem + TS::kRawStages * TS::kRawBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TS::kRawStages * TS::kRawBytes + TS::kCvtBytes);

  const int tid = threadIdx.x;
  // 1-D grid of anonymous CTAs; a CTA finds out what it is when it STARTS. Item t (an atomic counter) is: the boundary
  // role of VFO group t for t < nbound (= ngroups); part 0 of chain t - ngroups for the next nchains items; after that, entry
  // t - ngroups - nchains of the ready queue, i.e. the next part of whichever chain was handed on at that position.
  // Every entry below t has been taken by a CTA that started earlier and is resident or finished, so the entry this
  // CTA waits for is always on its way - whatever order the hardware dispatches blocks in.
  // The item travels from lane 0 to the warp through shared memory: a shared-memory load from a uniform address is
  // uniform to the compiler, so the loop counters and branches below stay in the uniform datapath.
  __shared__ int s_item[2];
  if (tid == 0) {
    int chain = 0, q = 0;
    const int t = atomicAdd(p.sched, 1);
    if (t < p.nbound) {
      chain = -1 - t;
    } else if (t < p.nbound + p.nchains) {
      chain = t - p.nbound;
    } else {
      const int* entry = p.sched + 2 + p.nchains + (t - p.nbound - p.nchains);
      int v, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(entry) : "memory");
        if (v == 0) { __nanosleep(100); if (++spins > (1 << 25)) { *p.err = 1; break; } }
      } while (v == 0);
      chain = v - 1;                                     // -1 if the watchdog gave up: handled below
      if (v != 0) q = __ldcg(p.sched + 2 + chain);       // parts of this chain already done (written before the entry)
      else chain = -1 - p.ngroups;                       // marker: leave
    }
    s_item[0] = chain;
    s_item[1] = q;
  }
  __syncwarp();
  const int chain = s_item[0], q = s_item[1];
  if (chain < -p.ngroups) return;
  const bool is_boundary = chain < 0;
  const int gy = is_boundary ? -1 - chain : chain / p.nseg;
  const int seg = is_boundary ? 0 : chain - gy * p.nseg;
  const int slot = gy * kVfoPerCta + tid;               // this thread's VFO within the slice
  const bool active = slot < p.vfo_count;
  const int vfo = p.vfo_base + (active ? slot : 0);     // inactive lanes shadow VFO 0 of the slice, never store

  if (is_boundary) {
    boundary_role<NF, FMT, FAST>(p, vfo, active, reinterpret_cast<float2*>(raw));
    return;
  }

  const int seg_start = p.seg_off + seg * p.S;
  const int seg_end = min(seg_start + p.S, p.B);
  const int part_start = seg_start + q * p.P;
  const int part_end = min(part_start + p.P, seg_end);
  const bool from_history = seg == 0 && !p.cold0;         // part 0 of segment 0 starts from the saved block history
  const int warm = (q > 0 || from_history) ? 0 : p.W;
  const int first = part_start - warm;                    // in-block index of the first sample processed
  const int total = warm + (part_end - part_start);       // multiple of max(kChunk, 2^DA)
  const int ntiles = (total + kTile - 1) / kTile;
  float2* hand = p.hand + ((size_t)chain * kHandSlots) * kThreads + tid;

  if (tid == 0) {
    for (int i = 0; i < TS::kRawStages; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // tile t = samples [first + t*kTile, ...) of the block; a tile that straddles two slices takes two bulk copies
  auto issue = [&](int t) {
    const int n = min(kTile, total - t * kTile);
    const int g = first + t * kTile;
    uint64_t* bar = &bars[t % TS::kRawStages];
    unsigned char* dst = raw + (t % TS::kRawStages) * TS::kRawBytes;
    mbar_expect_tx(bar, (unsigned)n * RawBytes<FMT>::v);
    const int s = p.raw.n_slices > 1 ? g / p.raw.slice_len : 0;
    const int off = g - s * p.raw.slice_len;
    const int n1 = p.raw.n_slices > 1 ? min(n, p.raw.slice_len - off) : n;
    tma_bulk_g2s(dst, reinterpret_cast<const unsigned char*>(p.raw.slice[s]) + (size_t)off * RawBytes<FMT>::v, (unsigned)n1 * RawBytes<FMT>::v, bar);
    if (n1 < n)
      tma_bulk_g2s(dst + (size_t)n1 * RawBytes<FMT>::v, reinterpret_cast<const unsigned char*>(p.raw.slice[s + 1]), (unsigned)(n - n1) * RawBytes<FMT>::v, bar);
  };
  if (tid == 0) {
    for (int t = 0; t < TS::kRawStages && t < ntiles; ++t) issue(t);
  }

  // ---- per-thread state ----
  Ones k1; k1.one = bcast2(p.one);
  const float2 r = p.rot[vfo];
  Rot rot; rot.a = pack2(r.x, r.y); rot.b = pack2(-r.y, r.x);
  const float2 ql = p.qlast[vfo];
  const float2* ckpt_col = p.ckpt + vfo;
  HbState hb[kFastStages > 0 ? kFastStages : 1];
  if (q > 0) {
    // take over the state the previous part of this chain left behind (lane 0's acquire above + this fence order the loads)
    __threadfence();
#pragma unroll
    for (int s = 0; s < NF; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) hb[s].e[k] = load_p2_cg(hand + (size_t)(s * kStateSlots + k) * kThreads);
#pragma unroll
      for (int k = 0; k < 3; ++k) hb[s].o[k] = load_p2_cg(hand + (size_t)(s * kStateSlots + 5 + k) * kThreads);
    }
  } else if (from_history) {
#pragma unroll
    for (int s = 0; s < NF; ++s) {
      const float2* st = p.state_in + (size_t)s * kStateSlots * p.vfo_pitch + vfo;
#pragma unroll
      for (int k = 0; k < 5; ++k) hb[s].e[k] = load_p2(st + (size_t)k * p.vfo_pitch);
#pragma unroll
      for (int k = 0; k < 3; ++k) hb[s].o[k] = load_p2(st + (size_t)(5 + k) * p.vfo_pitch);
    }
  } else {
#pragma unroll
    for (int s = 0; s < NF; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) hb[s].e[k] = pzero();
#pragma unroll
      for (int k = 0; k < 3; ++k) hb[s].o[k] = pzero();
    }
  }
  long long n_abs = p.block_abs + first;
  int idx = (int)(n_abs % p.nco_len);
  float oa, ob;
  if (q > 0) {   // the oscillator continues exactly where the previous part stopped
    const float2 o = __ldcg(hand + (size_t)(kHandSlots - 1) * kThreads);
    oa = o.x; ob = o.y;
  } else {       // nearest exact checkpoint, then the recurrence itself up to the first sample
    const int ck = idx / kNcoStride, rem = idx % kNcoStride;
    const float2 c = ckpt_col[(size_t)ck * p.vfo_pitch];
    oa = c.x; ob = c.y;
    for (int i = 0; i < rem; ++i) nco_step(k1, oa, ob, rot);
  }

  // output cursor in the stage-DA stream: outputs before the part's own start (warm-up) are discarded
  float2* mid = p.mid ? p.mid + (size_t)gy * p.n_mid * 32 + tid : p.xd_rows[vfo];
  const size_t mid_stride = p.mid ? (size_t)32 : (size_t)1;
  const int out_first = part_start >> NF;                 // first stage-DA index this part owns
  int out_pos = first >> NF;                              // stage-DA index of the next output produced
  float2 nxt = make_float2(0.f, 0.f);                     // tolerance mode: prefetched checkpoint of stride nxt_k
  int nxt_k = -1;

  for (int t = 0; t < ntiles; ++t) {
    const int n = min(kTile, total - t * kTile);
    mbar_wait(&bars[t % TS::kRawStages], (t / TS::kRawStages) & 1);
    const float2* tile;
    if (FMT == FMT_CF32) {
      tile = reinterpret_cast<const float2*>(raw + (t % TS::kRawStages) * TS::kRawBytes);
    } else {
      // unpack once per CTA: raw -> float (c, d), shared by the 32 VFOs of the warp
      const unsigned char* src = raw + (t % TS::kRawStages) * TS::kRawBytes;
      __syncwarp();   // every lane has finished reading the previous converted tile
      for (int i = tid; i < n; i += kThreads) cvt[i] = load_raw<FMT>(src, i);
      __syncwarp();
      if (tid == 0 && t + TS::kRawStages < ntiles) issue(t + TS::kRawStages);   // the raw slot is free again
      tile = cvt;
    }

#pragma unroll 2
    for (int c = 0; c < n; c += kChunk) {
      P2 out[kChunk >> NF];
      // rare variant: oscillator table wrap inside the chunk, absolute sample 0, and (tolerance mode) the restart
      // transient or a checkpoint stride boundary strictly inside the chunk
      const int in_stride = idx % kNcoStride;
      const bool special = (idx + kChunk > p.nco_len) || (n_abs == 0) ||
                           (FAST && (idx < p.transient || (in_stride != 0 && in_stride + kChunk > kNcoStride)));
      if (FAST && !special && in_stride == 0) {   // tolerance mode: snap back to the exact table every stride
        const int kk = idx / kNcoStride;
        const float2 c0 = (kk == nxt_k) ? nxt : ckpt_col[(size_t)kk * p.vfo_pitch];
        oa = c0.x; ob = c0.y;
        nxt_k = min(kk + 1, p.nck - 1);                     // prefetch the next stride's checkpoint (used 8 chunks later)
        nxt = ckpt_col[(size_t)nxt_k * p.vfo_pitch];
      }
      if (special) fast_chunk<NF, true, FAST>(k1, oa, ob, rot, hb, tile + c, out, idx, p.nco_len, n_abs, ql.x, ql.y, ckpt_col, p.vfo_pitch, p.transient);
      else         fast_chunk<NF, false, FAST>(k1, oa, ob, rot, hb, tile + c, out, idx, p.nco_len, n_abs, ql.x, ql.y, ckpt_col, p.vfo_pitch, p.transient);
      n_abs += kChunk;
#pragma unroll
      for (int i = 0; i < (kChunk >> NF); ++i) {
        if (out_pos >= out_first && active) store_p2(mid + (size_t)out_pos * mid_stride, out[i]);   // mid stream: 32 lanes write one 256-byte row
        out_pos++;
      }
    }
    if (FMT == FMT_CF32) {
      __syncwarp();   // every lane has finished reading this raw slot
      if (tid == 0 && t + TS::kRawStages < ntiles) issue(t + TS::kRawStages);
    }
  }

  // hand the chain on: state to HBM, then the chain joins the tail of the ready queue
  if (part_end < seg_end) {
#pragma unroll
    for (int s = 0; s < NF; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) store_p2(hand + (size_t)(s * kStateSlots + k) * kThreads, hb[s].e[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) store_p2(hand + (size_t)(s * kStateSlots + 5 + k) * kThreads, hb[s].o[k]);
    }
    hand[(size_t)(kHandSlots - 1) * kThreads] = make_float2(oa, ob);
    __threadfence();
    __syncwarp();
    if (tid == 0) {
      p.sched[2 + chain] = q + 1;
      const int pos = atomicAdd(p.sched + 1, 1);
      int* entry = p.sched + 2 + p.nchains + pos;
      const int v = chain + 1;
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(entry), "r"(v) : "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Deep kernel: half-band stages DA .. D-1 of every VFO of one launch group, from the stage-DA stream the main kernel
// wrote ([time][VFO], so a warp's 32 VFOs read one 256-byte row per time step), to the per-VFO stage-D rows the
// tail kernel reads. One warp = 32 VFOs x one range of T input samples, history in registers; ranges after the first
// run in over Wd = 10*(2^3 - 1) (rounded up to 96) earlier samples with zero history and discard those outputs; range 0
// starts from the block's saved (shifted, see boundary_role) history. A VFO with D <= 5 has no deep stage: its
// samples are only moved into its row. Lanes of a warp may differ in their stage count.
// ---------------------------------------------------------------------------------------------
struct DeepParams {
  const float2* mid;          // [ngroups][n_mid][32]
  const float2* state_in;     // [kMaxStages][kStateSlots][vfo_pitch] history at the start of this block
  float2* state_out;          // same layout: the warp that reaches the end of the stream saves the deep stages' history
  float2* const* xd_rows;     // [vfo_pitch] per VFO: where stage-D sample 0 of this block goes
  const unsigned char* vfo_D; // [vfo_pitch] half-band stages of each VFO
  int* counter;               // work-item counter, zeroed by the host before every launch
  int vfo_pitch, vfo_base, vfo_count;
  int DA;                     // stages already done by the main kernel
  int n_mid;                  // B >> DA
  int T, Wd, nranges, ngroups;
  float one;
};

constexpr int kDeepStep = 32;          // input samples per unrolled step of the deep kernel; its bulk-copy ring holds two steps (16 KB per warp)
constexpr int kDeepWarm = 96;          // 10*(2^3 - 1) input samples reach the last deep stage's history; rounded up to a step
constexpr int kDeepCtasPerSm = 12;     // 12 x 16 KB of ring per SM; up to 168 registers per thread

// A persistent grid of one-warp CTAs takes (time range, 32-VFO group) items from a counter. The work is a stream of
// 8 B per VFO per input sample of this kernel - bound by memory, next to nothing for the FP32 pipe - so the bank launches
// only a few of these warps per SM (4 beside the following block's FP32-bound main kernel, on a high-priority stream;
// kDeepCtasPerSm when the kernel runs alone): they sit beside the main kernel instead of displacing it.
template <bool FAST>
__global__ void __launch_bounds__(32, kDeepCtasPerSm) ddc_deep_kernel(const DeepParams p) {
  extern __shared__ __align__(128) float2 ring[];   // [2][kDeepStep][32]
  const int lane = threadIdx.x;
  Ones k1; k1.one = bcast2(p.one);
  __shared__ int s_next;
  __shared__ __align__(8) uint64_t bars[2];
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  int gdone = 0;                                           // steps this CTA has consumed so far: ring slot and barrier phase go on across items
  for (;;) {
  __syncwarp();
  if (lane == 0) s_next = atomicAdd(p.counter, 1);
  __syncwarp();
  const int item = s_next;   // through shared memory: uniform to the compiler (see ddc_main_kernel)
  if (item >= p.nranges * p.ngroups) break;
  const int range = item / p.ngroups;
  const int slot = (item - range * p.ngroups) * 32 + lane;
  const bool active = slot < p.vfo_count;
  const int vfo = p.vfo_base + (active ? slot : 0);
  const int D = p.vfo_D[vfo];
  const int nd = max(D - p.DA, 0);                         // 0..3 deep stages for this VFO
  HbState hb[kDeepStages];
  P2 ox[kDeepStages][3];                                   // the three odd-phase samples before hb[s].o[0] (for the saved history)
#pragma unroll
  for (int s = 0; s < kDeepStages; ++s)
#pragma unroll
    for (int k = 0; k < 3; ++k) ox[s][k] = pzero();
  const int t0 = range * p.T;
  const int t1 = min(t0 + p.T, p.n_mid);
  const bool from_start = t0 <= p.Wd;                      // the run-in would reach the block start: begin there, with the saved history
  const int ts = from_start ? 0 : t0 - p.Wd;
  if (from_start) {
#pragma unroll
    for (int s = 0; s < kDeepStages; ++s) {
      const int st_i = min(p.DA + s, kMaxStages - 1);
      const float2* st = p.state_in + (size_t)st_i * kStateSlots * p.vfo_pitch + vfo;
#pragma unroll
      for (int k = 0; k < 5; ++k) hb[s].e[k] = s < nd ? load_p2(st + (size_t)k * p.vfo_pitch) : pzero();
#pragma unroll
      for (int k = 0; k < 3; ++k) hb[s].o[k] = s < nd ? load_p2(st + (size_t)(5 + k) * p.vfo_pitch) : pzero();
    }
  } else {
#pragma unroll
    for (int s = 0; s < kDeepStages; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) hb[s].e[k] = pzero();
#pragma unroll
      for (int k = 0; k < 3; ++k) hb[s].o[k] = pzero();
    }
  }
  float2* xd = p.xd_rows[vfo];
  const float2* src = p.mid + (size_t)(item - range * p.ngroups) * p.n_mid * 32 + lane;
  // kDeepStep (32) input samples at a time while they last (ts and t0 are multiples of 32, so every stage starts a step
  // on its even phase): 16 / 8 / 4 outputs of deep stage 1 / 2 / 3, fully unrolled so that the filter histories are
  // renamed rather than moved. The kernel is a stream from HBM and the bytes in flight set its speed: the 32 rows of a
  // step are one contiguous 8 KB run of the stream, which lane 0 fetches one step ahead with ONE bulk copy into a
  // two-step shared-memory ring (cp.async.bulk + mbarrier; 32 per-lane cp.async per step kept the MIO queue full:
  // mio_throttle was 45 % of this kernel's stall samples).
  int t = ts;
  const int nstep = (t1 - ts) / kDeepStep;
  const float2* rows = src - lane;
  auto fetch = [&](int g) {
    if (g < nstep && lane == 0) {
      const int q = gdone + g;
      mbar_expect_tx(&bars[q & 1], kDeepStep * 32 * (unsigned)sizeof(float2));
      tma_bulk_g2s(ring + ((q & 1) * kDeepStep) * 32, rows + (size_t)(ts + kDeepStep * g) * 32, kDeepStep * 32 * (unsigned)sizeof(float2), &bars[q & 1]);
    }
  };
  fetch(0);
  for (int g = 0; g < nstep; ++g, t += kDeepStep) {
    const int q = gdone + g;
    __syncwarp();                                                  // every lane is through with step g-1, whose slot the next copy overwrites
    fetch(g + 1);
    mbar_wait(&bars[q & 1], (unsigned)(q >> 1) & 1u);              // step g has landed
    const float2* got = ring + ((q & 1) * kDeepStep) * 32 + lane;
    const bool keep = active && t >= t0;
    if (nd == 0) {
      if (keep) {
#pragma unroll
        for (int i = 0; i < kDeepStep; ++i) xd[t + i] = got[i * 32];
      }
      continue;
    }
    P2 y0[kDeepStep / 2];
#pragma unroll
    for (int i = 0; i < kDeepStep / 2; ++i) {
      const float2 a = got[(2 * i) * 32], c = got[(2 * i + 1) * 32];
      y0[i] = deep_pair<FAST>(k1, hb[0], ox[0], pack2(a.x, a.y), pack2(c.x, c.y));
    }
    if (nd == 1) {
      if (keep) {
#pragma unroll
        for (int i = 0; i < kDeepStep / 2; ++i) store_p2(xd + (t >> 1) + i, y0[i]);
      }
      continue;
    }
    P2 y1[kDeepStep / 4];
#pragma unroll
    for (int i = 0; i < kDeepStep / 4; ++i) y1[i] = deep_pair<FAST>(k1, hb[1], ox[1], y0[2 * i], y0[2 * i + 1]);
    if (nd == 2) {
      if (keep) {
#pragma unroll
        for (int i = 0; i < kDeepStep / 4; ++i) store_p2(xd + (t >> 2) + i, y1[i]);
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < kDeepStep / 8; ++i) {
      const P2 y2 = deep_pair<FAST>(k1, hb[2], ox[2], y1[2 * i], y1[2 * i + 1]);
      if (keep) store_p2(xd + (t >> 3) + i, y2);
    }
  }
  gdone += nstep;
  // the last range of a stream whose length is not a multiple of 32: sample by sample
  for (; t < t1; ++t) {
    P2 x = load_p2(src + (size_t)t * 32);
    int cnt = t;
    bool stop = false;
#pragma unroll
    for (int s = 0; s < kDeepStages; ++s) {
      if (!stop && s < nd) {
        if (cnt & 1) { ox[s][0] = ox[s][1]; ox[s][1] = ox[s][2]; ox[s][2] = hb[s].o[0]; hb_odd(hb[s], x); stop = true; }
        else { x = hb_even<FAST>(k1, hb[s], x); cnt >>= 1; }
      }
    }
    if (!stop && active && t >= t0) store_p2(xd + (t >> nd), x);
  }
  // end of the stream: the shifted history the deep stages start the next block from (see boundary_role)
  if (t1 == p.n_mid && active) {
#pragma unroll
    for (int s = 0; s < kDeepStages; ++s) {
      if (s < nd) {
        float2* st = p.state_out + (size_t)(p.DA + s) * kStateSlots * p.vfo_pitch + vfo;
        store_p2(st, ox[s][0]);                                                      // x[n-11]
        store_p2(st + (size_t)1 * p.vfo_pitch, ox[s][1]);                            // x[n-9]
        store_p2(st + (size_t)2 * p.vfo_pitch, ox[s][2]);                            // x[n-7]
        store_p2(st + (size_t)3 * p.vfo_pitch, hb[s].o[0]);                          // x[n-5]
        store_p2(st + (size_t)4 * p.vfo_pitch, hb[s].o[1]);                          // x[n-3]
#pragma unroll
        for (int k = 0; k < 3; ++k) store_p2(st + (size_t)(5 + k) * p.vfo_pitch, hb[s].e[2 + k]);   // x[n-6], x[n-4], x[n-2]
      }
    }
  }
  }   // next item
}

}  // namespace aeroddc
