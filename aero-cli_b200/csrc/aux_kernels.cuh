// aero-ddc-b200: the small kernels around the main cascade kernel.
//   nco_checkpoint_kernel : Oscillator::Oscillator table recurrence   oscillator.cpp:12-27
//   tail_kernel           : usb_demod / usb_decimdemod / compress     vfo.cpp:188-287
//                           FIR::FIRUpdateAndProcess, FIRHilbert, DelayThing   dsp.cpp:64-78,216-231, dsp.h:74-96
//   fp32_peak_kernel      : register-only FFMA issue-rate probe (roofline denominator)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aeroddc {

// ---------------------------------------------------------------------------------------------
// NCO checkpoints. One thread per VFO walks the reference recurrence v *= rot; v *= 1.95f - |v|^2
// for L steps in un-fused float arithmetic (scalar __fmul_rn/__fadd_rn are never contracted) and
// stores the state after every `stride` steps: ckpt[k][v] = state after k*stride steps (k = 0 is
// (1,0)), i.e. the value from which one more step yields table entry q[k*stride]. qlast = q[L-1].
// ---------------------------------------------------------------------------------------------
__global__ void nco_checkpoint_kernel(const float2* __restrict__ rot, const int* __restrict__ nco_len, float2* __restrict__ ckpt,
                                      float2* __restrict__ qlast, int n_vfo, int vfo_pitch, int stride) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vfo) return;
  const int L = nco_len[v];   // (int)Fs of the stream this VFO mixes (a sub-VFO runs at its parent's output rate)
  const float c = rot[v].x, d = rot[v].y;
  float a = 1.0f, b = 0.0f;
  int k = 0;
  for (int i = 0; i < L; i += stride) {
    ckpt[(size_t)k * vfo_pitch + v] = make_float2(a, b);
    ++k;
    const int n = min(stride, L - i);
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const float nr = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, d));
      const float ni = __fadd_rn(__fmul_rn(a, d), __fmul_rn(b, c));
      const float nm = __fsub_rn(1.95f, __fadd_rn(__fmul_rn(nr, nr), __fmul_rn(ni, ni)));
      a = __fmul_rn(nr, nm);
      b = __fmul_rn(ni, nm);
    }
  }
  qlast[v] = make_float2(a, b);
}

// ---------------------------------------------------------------------------------------------
// Tail. Stateless given the stage-D stream with enough history in front of the block:
//   m[k] = late > 0 ? sum_{i<T} tl[i] * xD[k*late - T + i] : xD[k]        (newest sample excluded)
//   u[k] = Re m[k-62] - sum_{i<125} hil[i] * Im m[k-124+i]                (newest included)
//   y[k] = U > 0 ? sum_{i<U} tu[i] * u[k-U+i] : u[k]                      (newest excluded)
//   out[k] = (short) trunc(double(y[k] * gain) * 32768.0)
// every sum accumulated left to right from 0.0f with separately rounded products.
// ---------------------------------------------------------------------------------------------
struct TailVfo {
  const float2* xd;      // points at stage-D index 0 of this block; negative indices are history
  unsigned char* out;    // payload row
  const float* late_taps;
  const float* usb_taps;
  const float* hil_taps;   // the non-zero Hilbert taps, in ascending tap order ...
  const int* hil_idx;      // ... and their tap indices (every other tap of the 125 is exactly +-0.0f and cannot change a sum)
  int n_hil;
  int hil_regular;         // 1 when the non-zero taps are exactly the odd indices 1, 3, ..., 2*n_hil-1 (always, for the 125-tap design)
  int n_stage, n_out;
  int hist;              // stage-D samples kept in front of the block
  int late, T, U;
  int demod_usb, cstyle, scalecomp;
  float gain;
};

constexpr int kTailChunk = 512;
constexpr int kTailThreads = 256;
constexpr int kHilbert = 125;
constexpr int kDelay = 62;

__device__ __forceinline__ short to_short_x86(float g) {
  // double(g) * 32768.0 is exact; conversion truncates toward zero; out of int32 range x86 yields
  // INT_MIN whose low 16 bits are 0 (the C++ conversion is undefined there; the oracle pins the same)
  const double v = (double)g * 32768.0;
  int i;
  if (!(v > -2147483649.0 && v < 2147483648.0)) i = (int)0x80000000;
  else i = __double2int_rz(v);
  return (short)(unsigned short)((unsigned)i & 0xFFFFu);
}
__device__ __forceinline__ int to_schar_x86(float v) {
  int i;
  if (!(v > -2147483904.0f && v < 2147483648.0f)) i = (int)0x80000000;
  else i = __float2int_rz(v);
  return (int)(signed char)(unsigned char)((unsigned)i & 0xFFu);
}

// packed (I,Q) helpers for the late FIR; additions as fma(a, 1.0f, b) with 1.0f a kernel parameter, because
// ptxas would contract mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (see ddc_kernels.cuh)
__device__ __forceinline__ unsigned long long tl_pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long tl_mul(unsigned long long a, unsigned long long b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long tl_fma(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// one (VFO, chunk of kTailChunk outputs) item; every thread of the CTA takes the same path through it
__device__ __forceinline__ void tail_item(const TailVfo* __restrict__ vfos, int vfo_i, int chunk, float one, size_t out_offset, unsigned char* tsm) {
  TailVfo v = vfos[vfo_i];
  v.out += out_offset;   // payload rows are double-buffered by block parity
  const int k0 = chunk * kTailChunk;
  if (k0 >= v.n_out) return;
  const int kc = min(kTailChunk, v.n_out - k0);
  const int tid = threadIdx.x;

  if (!v.demod_usb) {   // vfo::compress (vfo.cpp:260-287)
    for (int k = tid; k < kc; k += kTailThreads) {
      const float2 s = v.xd[k0 + k];
      if (v.cstyle == 1) {
        const float sc = (float)v.scalecomp;
        const int re = to_schar_x86(__fmul_rn(__fdiv_rn(s.x, sc), 128.0f));
        const int im = to_schar_x86(__fmul_rn(__fdiv_rn(s.y, sc), 128.0f));
        v.out[k0 + k] = (unsigned char)((re & 0xF0) | ((im & 0xF0) >> 4));
      } else {
        v.out[2 * (k0 + k)] = (unsigned char)to_schar_x86(__fmul_rn(s.x, 128.0f));
        v.out[2 * (k0 + k) + 1] = (unsigned char)to_schar_x86(__fmul_rn(s.y, 128.0f));
      }
    }
    return;
  }

  // shared layout: mI[n_m] | mQ[n_m] | u[n_u] | taps | staged stage-D window (late > 0 only)
  const int n_m = (kHilbert - 1) + v.U + kc;
  const int n_u = v.U + kc;
  const int n_w = v.late > 0 ? (n_m - 1) * v.late + v.T : 0;   // stage-D samples the late FIR of this chunk reads
  float* mI = reinterpret_cast<float*>(tsm);
  float* mQ = mI + n_m;
  float* u = mQ + n_m;
  float* tl = u + n_u;
  float* tu = tl + v.T;
  float* th = tu + v.U;
  float2* win = reinterpret_cast<float2*>(th + 2 * kHilbert + ((n_m * 2 + n_u + v.T + v.U) & 1));
  for (int i = tid; i < v.T; i += kTailThreads) tl[i] = v.late_taps[i];
  for (int i = tid; i < v.U; i += kTailThreads) tu[i] = v.usb_taps[i];
  int* thi = reinterpret_cast<int*>(th + kHilbert);
  for (int i = tid; i < v.n_hil; i += kTailThreads) { th[i] = v.hil_taps[i]; thi[i] = v.hil_idx[i]; }

  // phase 0/1: m[j] for k = kbase + j
  const int kbase = k0 - v.U - (kHilbert - 1);
  if (v.late > 0) {
    const float2* w0 = v.xd + ((long long)kbase * v.late - v.T);   // coalesced copy of the window
    for (int i = tid; i < n_w; i += kTailThreads) win[i] = w0[i];
    __syncthreads();
    const unsigned long long ONE = tl_pack(one, one);
    for (int j = tid; j < n_m; j += kTailThreads) {
      const float2* w = win + j * v.late;
      unsigned long long acc = 0ull;   // (0.0f, 0.0f)
#pragma unroll 7
      for (int i = 0; i < v.T; ++i) {
        const float2 sx = w[i];
        const float t = tl[i];
        acc = tl_fma(acc, ONE, tl_mul(tl_pack(sx.x, sx.y), tl_pack(t, t)));   // acc + tl[i]*x, both rails, un-fused
      }
      float ar, ai;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(ar), "=f"(ai) : "l"(acc));
      mI[j] = ar;
      mQ[j] = ai;
    }
  } else {
    for (int j = tid; j < n_m; j += kTailThreads) {
      const float2 sx = v.xd[kbase + j];
      mI[j] = sx.x;
      mQ[j] = sx.y;
    }
  }
  __syncthreads();
  // phase 2: u[j] for k = k0 - U + j ; m index of k is j + 124
  for (int j = tid; j < n_u; j += kTailThreads) {
    float h = 0.0f;
    const float* w = mQ + j;
    if (v.hil_regular) {
#pragma unroll 31
      for (int i = 0; i < v.n_hil; ++i) h = __fadd_rn(h, __fmul_rn(th[i], w[2 * i + 1]));
    } else {
#pragma unroll 7
      for (int i = 0; i < v.n_hil; ++i) h = __fadd_rn(h, __fmul_rn(th[i], w[thi[i]]));
    }
    u[j] = __fsub_rn(mI[j + (kHilbert - 1) - kDelay], h);
  }
  __syncthreads();
  // phase 3
  short* o16 = reinterpret_cast<short*>(v.out);
  for (int k = tid; k < kc; k += kTailThreads) {
    float y;
    if (v.U > 0) {
      const float* w = u + k;   // u index of output k0+k is k + U; window starts U earlier
      y = 0.0f;
      for (int i = 0; i < v.U; ++i) y = __fadd_rn(y, __fmul_rn(tu[i], w[i]));
    } else {
      y = u[k];
    }
    o16[k0 + k] = to_short_x86(__fmul_rn(y, v.gain));
  }
}

// Persistent grid: CTAs take (VFO, chunk) items from a counter (zeroed by the host before every launch). The bank
// launches about one CTA per SM on its high-priority post-processing stream, so the tail of block k runs in a few SM
// slots beside the main kernel of block k+1 instead of displacing it.
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const TailVfo* __restrict__ vfos, float one, size_t out_offset,
                                                             int* counter, int n_vfos, int chunks_per_vfo) {
  extern __shared__ __align__(16) unsigned char tsm[];
  __shared__ int s_item;
  const int n_items = n_vfos * chunks_per_vfo;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    tail_item(vfos, item / chunks_per_vfo, item % chunks_per_vfo, one, out_offset, tsm);
    __syncthreads();   // shared memory and s_item are reused by the next item
  }
}

// DC removal of Publisher::demodData (publisher.cpp:292-296), exact and therefore sequential:
//   avept = avept * (1.0f - 0.000001f) + 0.000001f * x;  x -= avept      (std::complex<float> ops = per rail)
// The only serial part is one dependent FMUL + FADD per sample and rail (8 cycles); everything else is taken off that
// chain by a three-warp pipeline over batches of kDccBatch samples in shared memory (mbarrier ring, kDccStages deep):
//   warp 1  loads the raw batch (coalesced 16-byte loads, from whichever slice of the block holds it), converts it,
//           keeps x and the products t = 0.000001f * x, the latter de-interleaved per rail;
//   warp 0  walks the recurrence: lane 0 the I rail, lane 1 the Q rail; four samples per LDS.128 of t (loaded one group
//           ahead), four FMUL + FADD pairs, one STS.128 of the four running averages - 2.5 instructions per sample
//           beside the 8-cycle dependency, on a scheduler it has to itself;
//   warp 2  subtracts the averages from x and stores the corrected cf32 block (coalesced 16-byte stores).
// The running average persists in `state` across blocks. Before (one warp doing all three jobs, values handed round by
// shuffles): 163 ms per 15.36 M-sample block, now 80 ms (the dependent pair costs 10.2 cycles per sample at the nominal clock;
// FFMA forms of the two operations and deeper prefetch of t were measured and change nothing); round 1 (one thread per rail
// chasing pointers): 1.2 s.
constexpr int kDccBatch = 1024;    // complex samples per batch
constexpr int kDccStages = 4;
constexpr int kDccThreads = 96;
constexpr int kDccStageBytes = 3 * 2 * kDccBatch * (int)sizeof(float);   // x (interleaved) | t (per rail) | averages (per rail)
constexpr int kDccSmemMin = kDccStages * kDccStageBytes + 3 * kDccStages * 8;

__device__ __forceinline__ void dcc_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// two consecutive complex samples (s, s + 1) of the block as floats; s is even, slices hold an even number of samples
template <int FMT> __device__ __forceinline__ float4 dcc_load2(const RawBlock& rb, int s, int n) {
  const void* base = rb.slice[0];
  int s_in = s;
#pragma unroll
  for (int k = 1; k < kMaxSlices; ++k)                  // static indices into the parameter block
    if (k < rb.n_slices && s >= k * rb.slice_len) { base = rb.slice[k]; s_in = s - k * rb.slice_len; }
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s + 1 < n) {
    if (FMT == 2) {
      v = __ldg(reinterpret_cast<const float4*>(base) + (s_in >> 1));
    } else if (FMT == 1) {
      const short4 r = __ldg(reinterpret_cast<const short4*>(base) + (s_in >> 1));
      v = make_float4(cvt_s16(r.x), cvt_s16(r.y), cvt_s16(r.z), cvt_s16(r.w));
    } else {
      const uchar4 r = __ldg(reinterpret_cast<const uchar4*>(base) + (s_in >> 1));
      v = make_float4(cvt_u8(r.x), cvt_u8(r.y), cvt_u8(r.z), cvt_u8(r.w));
    }
  } else if (s < n) {
    const float2 h = load_raw<FMT>(base, (size_t)s_in);
    v.x = h.x; v.y = h.y;
  }
  return v;
}

template <int FMT>
__global__ void __launch_bounds__(kDccThreads) dcc_kernel(const RawBlock rb, float* __restrict__ out, float* __restrict__ state, int n) {
  extern __shared__ __align__(16) unsigned char dsm[];
  uint64_t* full_t = reinterpret_cast<uint64_t*>(dsm + kDccStages * kDccStageBytes);
  uint64_t* full_a = full_t + kDccStages;
  uint64_t* empty = full_a + kDccStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3 * kDccStages; ++i) mbar_init(&full_t[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nbatch = (n + kDccBatch - 1) / kDccBatch;
  constexpr int kPairs = kDccBatch / 64;                 // float4 (two samples) per lane and batch

  if (warp == 1) {
    // ---- loads, conversion, t = c * x ----
    const float cc = 0.000001f;
    for (int bt = 0; bt < nbatch; ++bt) {
      const int st = bt % kDccStages;
      float* xs = reinterpret_cast<float*>(dsm + st * kDccStageBytes);
      float* ts = xs + 2 * kDccBatch;
      float4 v[kPairs];
#pragma unroll
      for (int i = 0; i < kPairs; ++i) v[i] = dcc_load2<FMT>(rb, bt * kDccBatch + 2 * (lane + 32 * i), n);
      mbar_wait(&empty[st], ((unsigned)(bt / kDccStages) & 1u) ^ 1u);
#pragma unroll
      for (int i = 0; i < kPairs; ++i) {
        const int q = lane + 32 * i;                      // samples 2q, 2q + 1 of the batch
        reinterpret_cast<float4*>(xs)[q] = v[i];
        reinterpret_cast<float2*>(ts)[q] = make_float2(__fmul_rn(cc, v[i].x), __fmul_rn(cc, v[i].z));
        reinterpret_cast<float2*>(ts + kDccBatch)[q] = make_float2(__fmul_rn(cc, v[i].y), __fmul_rn(cc, v[i].w));
      }
      __syncwarp();
      if (lane == 0) dcc_arrive(&full_t[st]);
    }
  } else if (warp == 0) {
    // ---- the recurrence ----
    const int rail = lane & 1;
    const float kk = 1.0f - 0.000001f;
    float a = state[rail];
    for (int bt = 0; bt < nbatch; ++bt) {
      const int st = bt % kDccStages;
      const unsigned ph = (unsigned)(bt / kDccStages) & 1u;
      float* xs = reinterpret_cast<float*>(dsm + st * kDccStageBytes);
      const float4* tp = reinterpret_cast<const float4*>(xs + 2 * kDccBatch + rail * kDccBatch);
      float4* ap = reinterpret_cast<float4*>(xs + 4 * kDccBatch + rail * kDccBatch);
      const int m = min(kDccBatch, n - bt * kDccBatch);
      const int g8 = m >> 3;
      mbar_wait(&full_t[st], ph);
      // eight samples per turn; their products are loaded TWO turns (130 cycles) ahead: the warp issues in order, so a
      // shared-memory load still in flight when its first use comes up would stall the chain itself
      float4 b0 = tp[0], b1 = tp[1], c0 = tp[2], c1 = tp[3];
#pragma unroll 3
      for (int g = 0; g < g8; ++g) {
        const float4 t0 = b0, t1 = b1;
        b0 = c0; b1 = c1;
        const int nx = min(2 * g + 4, kDccBatch / 4 - 2);
        c0 = tp[nx]; c1 = tp[nx + 1];
        float4 r0, r1;
        a = __fadd_rn(__fmul_rn(a, kk), t0.x); r0.x = a;
        a = __fadd_rn(__fmul_rn(a, kk), t0.y); r0.y = a;
        a = __fadd_rn(__fmul_rn(a, kk), t0.z); r0.z = a;
        a = __fadd_rn(__fmul_rn(a, kk), t0.w); r0.w = a;
        if (lane < 2) ap[2 * g] = r0;
        a = __fadd_rn(__fmul_rn(a, kk), t1.x); r1.x = a;
        a = __fadd_rn(__fmul_rn(a, kk), t1.y); r1.y = a;
        a = __fadd_rn(__fmul_rn(a, kk), t1.z); r1.z = a;
        a = __fadd_rn(__fmul_rn(a, kk), t1.w); r1.w = a;
        if (lane < 2) ap[2 * g + 1] = r1;
      }
      if (m & 7) {                                        // ragged end of the block
        const float* t1 = reinterpret_cast<const float*>(tp);
        float* a1 = reinterpret_cast<float*>(ap);
        for (int i = g8 * 8; i < m; ++i) {
          a = __fadd_rn(__fmul_rn(a, kk), t1[i]);
          if (lane < 2) a1[i] = a;
        }
      }
      __syncwarp();
      if (lane == 0) dcc_arrive(&full_a[st]);
    }
    if (lane < 2) state[lane] = a;
  } else {
    // ---- x - avept, stores ----
    for (int bt = 0; bt < nbatch; ++bt) {
      const int st = bt % kDccStages;
      const unsigned ph = (unsigned)(bt / kDccStages) & 1u;
      const float* xs = reinterpret_cast<const float*>(dsm + st * kDccStageBytes);
      const float* as = xs + 4 * kDccBatch;
      mbar_wait(&full_a[st], ph);
#pragma unroll
      for (int i = 0; i < kPairs; ++i) {
        const int q = lane + 32 * i;
        const int s = bt * kDccBatch + 2 * q;
        const float4 x = reinterpret_cast<const float4*>(xs)[q];
        const float2 ar = reinterpret_cast<const float2*>(as)[q];
        const float2 ai = reinterpret_cast<const float2*>(as + kDccBatch)[q];
        if (s + 1 < n)
          reinterpret_cast<float4*>(out)[(size_t)s >> 1] = make_float4(__fsub_rn(x.x, ar.x), __fsub_rn(x.y, ai.x), __fsub_rn(x.z, ar.y), __fsub_rn(x.w, ai.y));
        else if (s < n)
          reinterpret_cast<float2*>(out)[s] = make_float2(__fsub_rn(x.x, ar.x), __fsub_rn(x.y, ai.x));
      }
      __syncwarp();
      if (lane == 0) dcc_arrive(&empty[st]);
    }
  }
}

// Keep the last `hist` stage-D samples of every VFO in front of its next block. The rows are double-buffered by block
// parity: the history goes from this block's row (cur) to the front of the other parity's row (nxt).
__global__ void __launch_bounds__(256) xd_shift_kernel(const TailVfo* __restrict__ cur, const TailVfo* __restrict__ nxt) {
  const TailVfo v = cur[blockIdx.x];
  const int hist = v.hist;
  const float2* src = v.xd + v.n_stage - hist;
  float2* dst = const_cast<float2*>(nxt[blockIdx.x].xd) - hist;
  for (int i = threadIdx.x; i < hist; i += 256) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// FP32 peak probe: 8 independent FFMA2 chains per thread, 64 resident warps per SM.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float m, float c, long long* clk) {
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a), "f"(a + 0.5f));
  }
  unsigned long long pm, pc;
  asm("mov.b64 %0, {%1, %1};" : "=l"(pm) : "f"(m));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pm), "l"(pc));
    }
  }
  const long long t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i]));
    r += a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  // SM clock in kHz: cycles per nanosecond x 1e6
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = (long long)((double)(t1 - t0) / (double)(g1 - g0 ? g1 - g0 : 1) * 1e6);
}

}  // namespace aeroddc
