// aero-ddc-b200: tensor-core formulation of the NCO mix + the first five half-band stages (AERODDC_MODE_TENSOR).
//
// Same path as ddc_main_kernel (vfo.cpp:155-161 mix, halfbanddecimator.cpp:35-60 x 5), recast as the complex GEMM
// north_star names. NOT bit-identical: a tolerance mode (max |err| <= 1e-4 FS, error SNR >= 80 dB), checked against
// the exact mode in tests/test_gpu_parity.py and by decoded-frame identity.
//
// Algebra (scratch/tc_model.py checks it on the CPU). Five cascaded 11-tap half-band decimators are one 311-tap FIR g
// decimating by 32: z[m] = sum_t g[t] q[n_s + t] x[n_s + t], n_s = 32 m - 310. The reference oscillator q is a
// drifting float32 recurrence, but inside one window it is a pure rotation of the value at the window start:
// q[n_s + t] = q[n_s] u^t, u = rot / |rot| (amplitude and phase noise of the recurrence over 320 steps: ~1e-7).
// So  z[m] = q[n_s] * sum_t (g[t] u^t) x[n_s + t]: a per-VFO complex FIR with constant taps G_v[t] = g[t] u_v^t,
// followed by ONE complex multiply per output with the exact recurrence value q[n_s] (exact checkpoint x rotation).
// q[n_s] restarts from (1, 0) with an amplitude transient at every table wrap and the half-band queues are re-seeded
// with a one-sample shift at every block start (dsp.cpp:163-172): windows touching either are not a clean FIR and
// stay on the FP32 kernel (the head of every block and the zone after a wrap; see bank.cu).
//
// GEMM shape. Windows are padded to 320 samples starting at 32 m - 312 (a multiple of 8). Rows of X are the block cut
// into 32-sample rows (64 floats re/im interleaved); output m needs rows m-10 .. m. D[m][n] = sum over 40 k-steps of
// 8 samples (K = 16 bf16) of X[m + j][chunk] * F[n][k-step], n = (VFO, rail): rail 0 taps (Gr, -Gi), rail 1 (Gi, Gr).
// A operand = X tile in shared memory, K-major, no swizzle, laid out [16-byte chunk][row][16 B]: rows are 16 bytes
// apart, so the j-row shift of a k-step is just a start-address offset of the matrix descriptor - every raw sample is
// staged ONCE per tile instead of ten times (no im2col). B operand = the filter slab of the k-step, streamed from L2
// with bulk copies. Accumulators: M=128 outputs x N=256 rails fp32 in TMEM, double-buffered (512 columns).
// Precision: operands are split in bf16 hi + mid (x = hi + mid + O(2^-17 x)); three products hi*hi + mid*hi + hi*mid
// per k-step (the dropped terms are ~2^-17 relative: 107 dB SNR in the model), fp32 accumulation.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles): warps 0-3 epilogue (TMEM -> registers -> rotate ->
// stage-5 stream), warps 4-7 stage X (global -> bf16 hi/mid -> shared), warp 8 streams filter slabs (bulk copy),
// warp 9 issues tcgen05.mma.
#pragma once
#include <cuda_bf16.h>

#include "ddc_kernels.cuh"

namespace aeroddc {

constexpr int kTcM = 128;                 // outputs (stage-5 samples) per tile
constexpr int kTcN = 256;                 // rails per tile
constexpr int kTcVfos = kTcN / 2;         // VFOs per tile
constexpr int kTcTaps = 311;              // 1 + 10 * (2^5 - 1)
constexpr int kTcWin = 320;               // padded window, samples
constexpr int kTcLead = 312;              // window start = 32 m - kTcLead
constexpr int kTcKSteps = kTcWin / 8;     // 40 MMA k-steps of 8 samples
constexpr int kTcBack = 10;               // rows before the output's own row
constexpr int kTcRows = 144;              // rows per X tile (138 used)
constexpr int kTcXPart = 8 * kTcRows * 16;            // one bf16 part of an X tile: [8 chunks][rows][16 B]
constexpr int kTcXStage = 2 * kTcXPart;               // hi | mid
constexpr int kTcXStages = 2;
constexpr int kTcFPart = 2 * kTcN * 16;               // one bf16 part of a filter slab: [2 chunks][256 rails][16 B]
constexpr int kTcFSlab = 2 * kTcFPart;                // hi | mid
template <int FS> struct TcSmem {   // FS = filter-slab ring depth
  static constexpr int kBars = 2 * kTcXStages + 2 * FS + 4;
  static constexpr int kTotal = kTcXStages * kTcXStage + FS * kTcFSlab + 8 * kBars + 16;
};
constexpr int kTcThreads = 320;
constexpr int kTcPwRows = 512;            // rotation table rows: u^r, r = 0 .. 511
constexpr int kTcHead = 64;               // outputs [0, kTcHead) of a block (2048 samples) stay on the FP32 kernel

struct TcParams {
  RawBlock raw;               // the block (cf32), possibly in slices on several GPUs (a 32-sample row never straddles slices)
  const uint4* filt;          // [n_ntiles][kTcKSteps][kTcFSlab / 16]
  const float2* ckpt;         // [nck][vfo_pitch] exact NCO checkpoints (state after 256 k steps)
  const float2* pw;           // [kTcPwRows][vfo_pitch] u^r
  float2* mid;                // [32-VFO group][n_mid][32] stage-5 stream of this block
  long long block_abs;
  int nco_len, nck, vfo_pitch, vfo_base, vfo_count, mid_groups;
  int n_mid;                  // B / 32
  int m_first, m_end;         // outputs [m_first, m_end)
  int n_ntiles, n_mtiles;
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
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
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
  const int ntiles = p.n_ntiles * p.n_mtiles;

  if (warp < 4) {
    // ===== epilogue: D (TMEM) x q[n_s] -> stage-5 stream =====
    int tl = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const int a = tl & 1;
      const int nt = t / p.n_mtiles, mt = t - nt * p.n_mtiles;
      const int m = p.m_first + mt * kTcM + warp * 32 + lane;
      mbar_wait(&tfull[a], (tl >> 1) & 1);
      tc_fence_after();
      // the oscillator value the window's first sample is mixed with: S(idx + 1) = ckpt[c] * u^r
      const long long n_s = p.block_abs + 32ll * m - kTcLead;
      const int idx1 = (int)(n_s % p.nco_len) + 1;
      const int c = min(idx1 >> 8, p.nck - 1);
      const int r = idx1 - (c << 8);
      const float2* ck_row = p.ckpt + (size_t)c * p.vfo_pitch;
      const float2* pw_row = p.pw + (size_t)r * p.vfo_pitch;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int slot0 = nt * kTcVfos + q * 32;          // first VFO slot (within the launch group) of this 32-group
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kTcN + q * 64);
        uint32_t d0[32], d1[32];
        tc_ld32(taddr, d0);
        tc_ld32(taddr + 32, d1);
        tc_ld_wait();
        if (slot0 < p.mid_groups * 32 && m < p.m_end) {
          float4* out = reinterpret_cast<float4*>(p.mid + ((size_t)(slot0 >> 5) * p.n_mid + m) * 32);
#pragma unroll
          for (int v = 0; v < 32; v += 2) {
            float z[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = min(p.vfo_base + slot0 + v + e, p.vfo_pitch - 1);
              const float2 ck = __ldg(ck_row + col), pw = __ldg(pw_row + col);
              const float fr = ck.x * pw.x - ck.y * pw.y, fi = ck.x * pw.y + ck.y * pw.x;
              const int w = 2 * (v + e);
              const float dr = __uint_as_float(w < 32 ? d0[w & 31] : d1[w & 31]);
              const float di = __uint_as_float(w < 32 ? d0[(w + 1) & 31] : d1[(w + 1) & 31]);
              z[2 * e] = dr * fr - di * fi;
              z[2 * e + 1] = dr * fi + di * fr;
            }
            out[v >> 1] = make_float4(z[0], z[1], z[2], z[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[a]);
    }
  } else if (warp < 8) {
    // ===== X producer: rows of the raw block -> bf16 hi / mid, [chunk][row][16 B] =====
    const int ptid = threadIdx.x - 128;
    int tl = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const int xsi = tl & 1;
      const int mt = t % p.n_mtiles;
      const int row0 = p.m_first + mt * kTcM - kTcBack;
      mbar_wait(&xempty[xsi], ((tl >> 1) & 1) ^ 1);
      unsigned char* hi_base = xs + xsi * kTcXStage;
      for (int i = ptid; i < kTcM + kTcBack; i += 128) {
        const int rho = row0 + i;
        const bool valid = rho >= 0 && rho < p.n_mid;
        const int g0 = (valid ? rho : 0) * 32;
        const int sl = p.raw.n_slices > 1 ? g0 / p.raw.slice_len : 0;
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(p.raw.slice[sl]) + (g0 - sl * p.raw.slice_len));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;
          if (valid) { u = __ldg(src + 2 * c); w = __ldg(src + 2 * c + 1); }
          uint4 h, mdl;
          split_pair(u.x, u.y, h.x, mdl.x);
          split_pair(u.z, u.w, h.y, mdl.y);
          split_pair(w.x, w.y, h.z, mdl.z);
          split_pair(w.z, w.w, h.w, mdl.w);
          *reinterpret_cast<uint4*>(hi_base + (c * kTcRows + i) * 16) = h;
          *reinterpret_cast<uint4*>(hi_base + kTcXPart + (c * kTcRows + i) * 16) = mdl;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
      mbar_arrive(&xfull[xsi]);
    }
  } else if (warp == 8) {
    // ===== filter slabs: one 16 KB bulk copy per k-step =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int nt = t / p.n_mtiles;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.filt) + (size_t)nt * kTcKSteps * kTcFSlab;
        for (int ks = 0; ks < kTcKSteps; ++ks, ++it) {
          const int fsi = it % kTcFStages;
          mbar_wait(&fempty[fsi], ((it / kTcFStages) & 1) ^ 1);
          mbar_expect_tx(&ffull[fsi], kTcFSlab);
          tma_bulk_g2s(fsl + fsi * kTcFSlab, src + (size_t)ks * kTcFSlab, kTcFSlab, &ffull[fsi]);
        }
      }
    }
  } else {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0;
      int tl = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
        const int a = tl & 1, xsi = tl & 1;
        mbar_wait(&tempty[a], ((tl >> 1) & 1) ^ 1);
        mbar_wait(&xfull[xsi], (tl >> 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(a * kTcN);
        const uint32_t xhi = smem_u32(xs + xsi * kTcXStage), xmid = xhi + kTcXPart;
        for (int ks = 0; ks < kTcKSteps; ++ks, ++it) {
          const int fsi = it % kTcFStages;
          mbar_wait(&ffull[fsi], (it / kTcFStages) & 1);
          tc_fence_after();
          const int s = 8 + 8 * ks;                       // first sample of the k-step, counted from row (m - 10)
          const uint32_t xo = (uint32_t)((((s & 31) >> 2) * kTcRows + (s >> 5)) * 16);
          const uint64_t ah = tc_desc(xhi + xo, kTcRows * 16, 128), am = tc_desc(xmid + xo, kTcRows * 16, 128);
          const uint32_t fb = smem_u32(fsl + fsi * kTcFSlab);
          const uint64_t bh = tc_desc(fb, kTcN * 16, 128), bm = tc_desc(fb + kTcFPart, kTcN * 16, 128);
          tc_mma(d, ah, bh, ks > 0);
          tc_mma(d, am, bh, 1);
          tc_mma(d, ah, bm, 1);
          tc_commit(&fempty[fsi]);
        }
        tc_commit(&xempty[xsi]);
        tc_commit(&tfull[a]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- finalize-time tables ----
// filt[nt][ks][part][chunk][n][8 bf16]: n = 2 * (VFO within the tile) + rail; element k of the k-step: sample 8 ks + k / 2
// of the padded window, component k & 1. Rail 0 (real): (Gr, -Gi); rail 1 (imaginary): (Gi, Gr); G = g[t' - 2] u^t'.
__global__ void tc_build_filters_kernel(const float2* __restrict__ rot, const double* __restrict__ g, int vfo_base, int vfo_count,
                                        int n_ntiles, uint4* __restrict__ filt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;   // over n_ntiles * kTcKSteps * kTcN
  if (n >= n_ntiles * kTcKSteps * kTcN) return;
  const int rail_n = n % kTcN, ks = (n / kTcN) % kTcKSteps, nt = n / (kTcN * kTcKSteps);
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
    slab[ch * kTcN + rail_n] = make_uint4(hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
    slab[(kTcFPart / 16) + ch * kTcN + rail_n] = make_uint4(mid[4 * ch], mid[4 * ch + 1], mid[4 * ch + 2], mid[4 * ch + 3]);
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
