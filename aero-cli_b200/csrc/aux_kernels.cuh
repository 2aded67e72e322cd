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
// Lane 0 walks the I rail and lane 1 the Q rail (one dependent FMUL + FADD per sample); the running average persists in
// `state` across blocks. All 32 lanes feed them: a batch is 32 consecutive raw values (16 complex samples, one coalesced
// load), kDccAhead batches are in flight in registers, the two rails pick their values out of a batch with shuffles one
// batch ahead of the recurrence, and every lane corrects and stores its own value (one coalesced store per batch) with
// the averages the two rails leave in shared memory. Measured: 165 ms per 15.36 M-sample block (the pointer-chasing
// one-thread-per-rail form of this loop waited for memory at every sample: 1.17 s).
// The block may come in slices (RawBlock, ddc_kernels.cuh; slice lengths are multiples of 32 samples).
constexpr int kDccAhead = 16;

template <int FMT> __device__ __forceinline__ float dcc_load(const void* raw, size_t i) {
  if (FMT == 0) return __fdiv_rn(__fsub_rn((float)reinterpret_cast<const unsigned char*>(raw)[i], 127.4f), 128.0f);
  if (FMT == 1) return __fdiv_rn((float)reinterpret_cast<const short*>(raw)[i], 32768.0f);
  return reinterpret_cast<const float*>(raw)[i];
}

template <int FMT>
__device__ __forceinline__ float dcc_walk(const void* raw, float* __restrict__ o, int m, int lane, float a, float (*avg)[32]) {
  const int rail = lane & 1;
  const float k = 1.0f - 0.000001f, c = 0.000001f;
  const int nb = m / 16;                               // batches of 32 raw values (m is a multiple of 16)
  float buf[kDccAhead];
#pragma unroll
  for (int j = 0; j < kDccAhead; ++j) buf[j] = j < nb ? dcc_load<FMT>(raw, (size_t)j * 32 + lane) : 0.0f;
  float xn[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) xn[i] = __shfl_sync(0xffffffffu, buf[0], 2 * i + rail);
  for (int b0 = 0; b0 < nb; b0 += kDccAhead) {
#pragma unroll
    for (int j = 0; j < kDccAhead; ++j) {
      const int bt = b0 + j;
      float xc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) xc[i] = xn[i];
      const float mine = buf[j];                         // this lane's own raw value of batch bt
      if (bt + kDccAhead < nb) buf[j] = dcc_load<FMT>(raw, (size_t)(bt + kDccAhead) * 32 + lane);   // uniform
#pragma unroll
      for (int i = 0; i < 16; ++i) xn[i] = __shfl_sync(0xffffffffu, buf[(j + 1) % kDccAhead], 2 * i + rail);
      if (bt < nb) {                                     // uniform
        float* av = avg[j & 1];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          a = __fadd_rn(__fmul_rn(a, k), __fmul_rn(c, xc[i]));
          if (lane < 2) av[2 * i + rail] = a;            // the running average after sample i of the batch, per rail
        }
        __syncwarp();
        o[(size_t)bt * 32 + lane] = __fsub_rn(mine, av[lane]);
      }
    }
  }
  return a;
}

template <int FMT>
__global__ void __launch_bounds__(32) dcc_kernel(const RawBlock rb, float* __restrict__ out, float* __restrict__ state, int n) {
  __shared__ float avg[2][32];
  const int lane = threadIdx.x;
  float a = state[lane & 1];
  int done = 0;
#pragma unroll
  for (int s = 0; s < kMaxSlices; ++s) {               // static index into the parameter block
    if (s < rb.n_slices && done < n) {
      const int m = min(rb.slice_len, n - done);
      a = dcc_walk<FMT>(rb.slice[s], out + 2 * (size_t)done, m, lane, a, avg);
      done += m;
    }
  }
  if (lane < 2) state[lane] = a;
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
