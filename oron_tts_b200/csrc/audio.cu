// Log-mel STFT front end, Vocos iSTFT head and peak normalisation (n_fft 1024, hop 256).
// Both transforms pack TWO real frames into one 1024-point complex FFT (frame A -> real part,
// frame B -> imaginary part) that a single warp runs in registers (fft_warp.cuh), so a CTA of 8 warps
// transforms 16 frames between two block barriers and the complex spectra / frame buffers never touch
// HBM: log-mel reads the waveform once and writes [n_mels, T]; the iSTFT reads the head activations
// once and writes the waveform once.
//
// These kernels are bound by fp32 instruction issue, not by HBM: one packed transform is ~1.2 k
// instructions per lane against 4 KB (log-mel) / 8.4 KB (iSTFT) of HBM traffic per frame pair
// ("audio kernels").
#include "../../include/oron_b200.h"
#include "fft_warp.cuh"
#include "host_util.h"
#include "ptx.cuh"

using namespace oron;

namespace {

constexpr int NFFT = 1024;
constexpr int HOP = 256;
constexpr int NBIN = 513;
constexpr int AUD_WARPS = 8;                   // one frame pair per warp; two CTAs per SM overlap load and math phases
constexpr int AUD_THREADS = AUD_WARPS * 32;
constexpr int AUD_FR = 2 * AUD_WARPS;          // frames transformed per CTA pass
constexpr int XB_BYTES = fw::XB_ELEMS * 8;     // per-warp exchange tile

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------
// log-mel
// ---------------------------------------------------------------------------
constexpr int MEL_MAXBAND = 32;  // widest triangular filter (bins) kept in smem
constexpr int MEL_MAXM = 128;
// Filter bands (oron_logmel_bands): int lo[128] | int n[128] | float w[32][128]  (w[i][m] = fb[lo[m] + i][m])
constexpr int MEL_BANDS_BYTES = 2 * MEL_MAXM * 4 + MEL_MAXBAND * MEL_MAXM * 4;
constexpr int MEL_SMEM = AUD_WARPS * XB_BYTES + NFFT * 8 + NFFT * 4 + MEL_MAXBAND * MEL_MAXM * 4 +
                         AUD_FR * (MEL_MAXM + 1) * 4 + 2 * MEL_MAXM * 4;

// The non-zero band of every triangular mel filter, once per filterbank (it used to be rediscovered by every CTA of
// every launch: 51 k dependent global loads per CTA, 16 % of the kernel's stall samples).
__global__ void __launch_bounds__(256)
logmel_bands_kernel(const float* __restrict__ fb, int n_mels, uint8_t* __restrict__ bands) {
  __shared__ int lo_s[MEL_MAXM], hi_s[MEL_MAXM];
  int* lo_g = reinterpret_cast<int*>(bands);
  int* n_g = lo_g + MEL_MAXM;
  float* w_g = reinterpret_cast<float*>(n_g + MEL_MAXM);
  for (int m = threadIdx.x; m < MEL_MAXM; m += blockDim.x) { lo_s[m] = NBIN; hi_s[m] = -1; }
  __syncthreads();
  for (int i = threadIdx.x; i < NBIN * n_mels; i += blockDim.x) {
    if (fb[i] != 0.f) {
      const int k = i / n_mels, m = i - k * n_mels;
      atomicMin(&lo_s[m], k);
      atomicMax(&hi_s[m], k);
    }
  }
  __syncthreads();
  for (int m = threadIdx.x; m < MEL_MAXM; m += blockDim.x) {
    const bool any = m < n_mels && lo_s[m] <= hi_s[m];
    const int lo = any ? lo_s[m] : 0;
    const int n = any ? hi_s[m] - lo_s[m] + 1 : 0;
    lo_g[m] = lo;
    n_g[m] = n;  // n > MEL_MAXBAND: the kernel reads this filter from fb directly
    for (int i = 0; i < MEL_MAXBAND; ++i) w_g[i * MEL_MAXM + m] = (i < n && n <= MEL_MAXBAND) ? fb[(long long)(lo + i) * n_mels + m] : 0.f;
  }
}

__global__ void __launch_bounds__(AUD_THREADS, 2)
logmel_kernel(const float* __restrict__ wav, long long ld_wav, int nb, int n_samples, int n_frames,
              const float* __restrict__ window, const float* __restrict__ fb, const uint8_t* __restrict__ bands, int n_mels,
              float clip, float* __restrict__ out) {
  using fw::c64;
  extern __shared__ __align__(16) uint8_t dsm[];
  float2* xb_all = reinterpret_cast<float2*>(dsm);
  float2* tw = xb_all + AUD_WARPS * fw::XB_ELEMS;
  float* win = reinterpret_cast<float*>(tw + NFFT);
  float (*fbs)[MEL_MAXM] = reinterpret_cast<float (*)[MEL_MAXM]>(win + NFFT);  // [band bin][filter]
  float (*mel_s)[MEL_MAXM + 1] = reinterpret_cast<float (*)[MEL_MAXM + 1]>(&fbs[MEL_MAXBAND][0]);
  int* band_lo = reinterpret_cast<int*>(&mel_s[AUD_FR][0]);
  int* band_n = band_lo + MEL_MAXM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* xb = xb_all + warp * fw::XB_ELEMS;
  c64* mag2 = reinterpret_cast<c64*>(xb);  // after the transform: (|A[k]|, |B[k]|), k = 0..512, of this warp's frame pair

  // ---- once per (persistent) CTA: twiddles, window, filter bands ----
  fw::fill_twiddle_table(tw);
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) win[n] = window[n];
  {
    const int* bl = reinterpret_cast<const int*>(bands);
    const float4* bw = reinterpret_cast<const float4*>(bands + 2 * MEL_MAXM * 4);
    for (int m = threadIdx.x; m < 2 * MEL_MAXM; m += blockDim.x) band_lo[m] = bl[m];  // lo[128] | n[128], contiguous here too
    float4* dst = reinterpret_cast<float4*>(&fbs[0][0]);
    for (int i = threadIdx.x; i < MEL_MAXBAND * MEL_MAXM / 4; i += blockDim.x) {  // * 1/2: |A| = |p| / 2 (below)
      const float4 q = bw[i];
      dst[i] = make_float4(0.5f * q.x, 0.5f * q.y, 0.5f * q.z, 0.5f * q.w);
    }
  }
  __syncthreads();

  // Every warp is on its own from here on: it walks frame pairs (two real frames per complex transform) with no block
  // barrier -- the transform, the spectra, the projection and the store of a pair all stay inside the warp.
  const int pairs_per_clip = (n_frames + 1) / 2;
  const long long n_pairs = (long long)nb * pairs_per_clip;
  const long long pair_step = (long long)gridDim.x * AUD_WARPS;
  for (long long pair = (long long)blockIdx.x * AUD_WARPS + warp; pair < n_pairs; pair += pair_step) {
    const int b = int(pair / pairs_per_clip);
    const int fA = 2 * int(pair - (long long)b * pairs_per_clip);
    const float* x = wav + (long long)b * ld_wav;
    {  // the next pair's samples (NFFT + HOP floats = 40 lines of 128 bytes) into L2 while this one is transformed
      const long long np = pair + pair_step;
      if (np < n_pairs) {
        const int nb_ = int(np / pairs_per_clip);
        const long long s0 = 2LL * (np - (long long)nb_ * pairs_per_clip) * HOP - NFFT / 2;
        const float* px = wav + (long long)nb_ * ld_wav;
        const long long o0 = s0 + 32LL * lane, o1 = s0 + 32LL * (lane + 32);
        if (o0 >= 0 && o0 < n_samples) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o0));
        if (lane < 8 && o1 >= 0 && o1 < n_samples) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o1));
      }
    }
    const bool hasB = fA + 1 < n_frames;
    // frame pair (fA, fA+1): reflect-padded (center=True) windowed samples, lane holds n = lane + 32 n1
    c64 v[32];
    const int base = fA * HOP - NFFT / 2 + lane;
    if (base - lane >= 0 && base - lane + NFFT + HOP <= n_samples) {  // interior: no reflection
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const float w = win[lane + 32 * n1];
        v[n1] = fw::cmul2(fw::cpack(x[base + 32 * n1], x[base + HOP + 32 * n1]), fw::cpack(w, w));
      }
    } else {
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const float w = win[lane + 32 * n1];
        int nA = base + 32 * n1;
        if (nA < 0) nA = -nA;
        if (nA >= n_samples) nA = 2 * (n_samples - 1) - nA;
        float vb = 0.f;
        if (hasB) {
          int nB = base + HOP + 32 * n1;
          if (nB < 0) nB = -nB;
          if (nB >= n_samples) nB = 2 * (n_samples - 1) - nB;
          vb = x[nB] * w;
        }
        v[n1] = fw::cpack(x[nA] * w, vb);
      }
    }
    fw::fft1024_warp<false>(v, xb, tw, lane);
    // split the two real spectra: bin k = lane + 32 k2 needs Z[k] and Z[1024 - k]; the latter sits in lane
    // (32 - lane) & 31, register 31 - k2 (lane 0: own register (32 - k2) & 31).
    // A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i: |A| = |p| / 2, |B| = |q| / 2 with
    // p = (z.x + zp.x, z.y - zp.y), q = (z.x - zp.x, z.y + zp.y); the tile holds |p|, |q| and the filter weights in
    // shared memory carry the factor 1/2
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const c64 z = v[k2];
      c64 zp = __shfl_sync(0xffffffffu, v[31 - k2], src);
      if (lane == 0) zp = v[(32 - k2) & 31];
      const c64 pq = fw::cfma2(zp, fw::cpack(1.0f, -1.0f), z), qq = fw::cfma2(zp, fw::cpack(-1.0f, 1.0f), z);
      const float2 p2 = fw::cunpack(fw::cmul2(pq, pq)), q2 = fw::cunpack(fw::cmul2(qq, qq));
      mag2[lane + 32 * k2] = fw::cpack(sqrt_approx(p2.x + p2.y), sqrt_approx(q2.x + q2.y));
    }
    if (lane == 0) {  // Nyquist bin: Z[512] = A[512] + i B[512], both real
      const float2 ny = fw::cunpack(v[16]);
      mag2[512] = fw::cpack(2.0f * fabsf(ny.x), 2.0f * fabsf(ny.y));
    }
    __syncwarp();
    // banded mel projection, both frames at once: lane -> filters lane, lane + 32, ...; rows of fbs beyond a filter's
    // band hold zeros, so the band loop runs in steps of four independent multiply-adds (bins clamped to the last one)
    for (int m = lane; m < n_mels; m += 32) {
      const int lo = band_lo[m], n = band_n[m];
      c64 acc0 = fw::cpack(0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
      if (n <= MEL_MAXBAND) {
        for (int i = 0; i < n; i += 4) {
          const float f0 = fbs[i][m], f1 = fbs[i + 1][m], f2 = fbs[i + 2][m], f3 = fbs[i + 3][m];
          const int k = lo + i;
          acc0 = fw::cfma2(mag2[min(k, 512)], fw::cpack(f0, f0), acc0);
          acc1 = fw::cfma2(mag2[min(k + 1, 512)], fw::cpack(f1, f1), acc1);
          acc2 = fw::cfma2(mag2[min(k + 2, 512)], fw::cpack(f2, f2), acc2);
          acc3 = fw::cfma2(mag2[min(k + 3, 512)], fw::cpack(f3, f3), acc3);
        }
      } else {
        for (int i = 0; i < n; ++i) {
          const float f = 0.5f * fb[(long long)(lo + i) * n_mels + m];
          acc0 = fw::cfma2(mag2[lo + i], fw::cpack(f, f), acc0);
        }
      }
      const float2 a2 = fw::cunpack(fw::cadd(fw::cadd(acc0, acc1), fw::cadd(acc2, acc3)));
      float* o = out + ((long long)b * n_mels + m) * n_frames + fA;
      o[0] = __logf(fmaxf(a2.x, clip));  // lg2.approx * ln 2: absolute error < 1e-6 on these magnitudes
      if (hasB) o[1] = __logf(fmaxf(a2.y, clip));
    }
    __syncwarp();  // the tile (magnitudes) is reused by the next pair's transform
  }
}

// ---------------------------------------------------------------------------
// iSTFT head: spectrum from the head activations -> irfft -> window -> overlap-add -> envelope
// ---------------------------------------------------------------------------
// Every warp owns a RUN of consecutive output hops of one clip and walks it two frames at a time: stage the two
// activation rows in its tile (async copies; the next pair's rows are already on their way into L2), build the packed
// spectrum, one 1024-point transform, window, and overlap-add against a lane-private carry of the three unfinished hops
// (sample n = lane + 32 k2 always belongs to the same lane, so the carry needs no synchronisation at all). No block
// barrier after the table setup, no frame is transformed twice except the four halo frames at the head of a run.
constexpr int IST_CARRY = 3 * HOP;       // floats per warp: the partial sums of the next three hops
constexpr int IST_SMEM = AUD_WARPS * (XB_BYTES + IST_CARRY * 4) + NFFT * 8 + NFFT * 4 + HOP * 4;
constexpr int IST_MIN_RUN = 8;

// exp / cos / sin of the head activations (Vocos ISTFTHead): fast-math units after an explicit
// two-constant range reduction, absolute error ~1e-6 for |phase| < 1e3.
__device__ __forceinline__ float2 polar_clip(float logmag, float phase) {
  float mg, s, c;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(mg) : "f"(logmag * 1.4426950408889634f));  // < 2^-126 flushes to 0
  mg = fminf(mg, 100.0f);
  const float q = (fmaf(phase, 0.15915494309189535f, 12582912.0f)) - 12582912.0f;  // round to nearest (|phase| < 2^22)
  float r = fmaf(q, -6.2831854820251465f, phase);   // hi part of 2 pi (fp32)
  r = fmaf(q, 1.7484555e-7f, r);                    // 2 pi - hi
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(r));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(r));
  return make_float2(mg * c, mg * s);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int mode>
__global__ void __launch_bounds__(AUD_THREADS, 2)
istft_head_kernel(const float* __restrict__ h, long long ldh, int rows_per_batch, int nb, int n_frames,
                  const float* __restrict__ window, float* __restrict__ out, long long ld_out, int run,
                  int runs_per_clip) {
  extern __shared__ __align__(16) uint8_t dsm[];
  float2* xb_all = reinterpret_cast<float2*>(dsm);
  float* carry_all = reinterpret_cast<float*>(xb_all + AUD_WARPS * fw::XB_ELEMS);
  float2* tw = reinterpret_cast<float2*>(carry_all + AUD_WARPS * IST_CARRY);
  float* wins = reinterpret_cast<float*>(tw + NFFT);  // window * transform scale
  float* ienv = wins + NFFT;                          // 1 / sum_q window[i + 256 q]^2: the envelope of an interior hop
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* xb = xb_all + warp * fw::XB_ELEMS;
  float* cw = carry_all + warp * IST_CARRY + lane;    // carry[hop 0..2][j 0..7][lane]
  constexpr int RAW_LD = fw::XB_ELEMS;                // second raw activation row inside the tile (floats)
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(h) & 15) == 0) && ((ldh & 3) == 0);

  const float scale = (mode == 1) ? (32.0f / NFFT) : (1.0f / NFFT);  // normalized=True: * sqrt(N)
  fw::fill_twiddle_table(tw);
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) wins[n] = window[n] * scale;
  for (int i = threadIdx.x; i < HOP; i += blockDim.x) {
    float e = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) e = fmaf(window[i + HOP * q], window[i + HOP * q], e);
    ienv[i] = 1.0f / e;
  }
  __syncthreads();

  const int gw = blockIdx.x * AUD_WARPS + warp, nw = gridDim.x * AUD_WARPS;
  for (int item = gw; item < nb * runs_per_clip; item += nw) {
    const int b = item / runs_per_clip;
    // padded-domain hop hh receives frames hh-3..hh and holds output samples [256 hh - 512, +256): hops 2..n_frames
    const int h0 = 2 + (item - b * runs_per_clip) * run;
    const int h1 = min(h0 + run, n_frames + 1);
    if (h0 >= h1) continue;
    const float* hb = h + (long long)b * rows_per_batch * ldh;
    float* ob = out + (long long)b * ld_out - NFFT / 2 + lane;
#pragma unroll
    for (int j = 0; j < 24; ++j) cw[32 * j] = 0.f;

    for (int A = h0 - 4; A < h1; A += 2) {  // frames A (real part) and A + 1 (imaginary part); A is even
      const bool hasA = A >= 0 && A < n_frames, hasB = A >= 0 && A + 1 < n_frames;
      fw::c64 v[32];
      if (hasA) {  // warp-uniform
        const float* ha = hb + (long long)A * ldh;
        const float* hbp = hasB ? ha + ldh : ha;
        float* rawA = reinterpret_cast<float*>(xb);
        float* rawB = rawA + RAW_LD;
        if (vec_ok) {
#pragma unroll
          for (int c = lane; c < 256; c += 32) {
            cp_async16(rawA + 4 * c, ha + 4 * c);
            cp_async16(rawB + 4 * c, hbp + 4 * c);
          }
          if (lane < 2) {
            cp_async4(rawA + 1024 + lane, ha + 1024 + lane);
            cp_async4(rawB + 1024 + lane, hbp + 1024 + lane);
          }
        } else {
          for (int c = lane; c < 2 * NBIN; c += 32) {
            cp_async4(rawA + c, ha + c);
            cp_async4(rawB + c, hbp + c);
          }
        }
        if (A + 2 < h1 && A + 2 < n_frames) {  // the next pair's rows: pull them into L2 while this pair is transformed
          const char* nx = reinterpret_cast<const char*>(ha + 2 * ldh);
          prefetch_l2(nx + 128 * lane);
          if (lane == 0) prefetch_l2(nx + 4096);
          if (A + 3 < n_frames) {
            const char* ny = reinterpret_cast<const char*>(ha + 3 * ldh);
            prefetch_l2(ny + 128 * lane);
            if (lane == 0) prefetch_l2(ny + 4096);
          }
        }
        cp_async_wait_all();
        __syncwarp();
        // lower half of Z = A + iB in registers (bin k = lane + 32 n1), mirrored half Z[N-k] = conj(A) + i conj(B)
        float2 zc[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int k = lane + 32 * n1;
          float2 P, Q;
          if (mode == 0) {
            P = polar_clip(rawA[k], rawA[NBIN + k]);
            Q = polar_clip(rawB[k], rawB[NBIN + k]);
          } else {
            P = *reinterpret_cast<const float2*>(rawA + 2 * k);
            Q = *reinterpret_cast<const float2*>(rawB + 2 * k);
          }
          if (!hasB) Q = make_float2(0.f, 0.f);
          if (n1 == 0 && lane == 0) { P.y = 0.f; Q.y = 0.f; }  // C2R ignores the imaginary part of DC
          v[n1] = fw::cpack(P.x - Q.y, P.y + Q.x);
          zc[n1] = make_float2(P.x + Q.y, Q.x - P.y);
        }
        // Nyquist bin (k = 512): real parts only, lives in lane 0 register 16
        float2 nyq = make_float2(0.f, 0.f);
        if (lane == 0) {
          if (mode == 0) {
            nyq.x = polar_clip(rawA[512], rawA[NBIN + 512]).x;
            nyq.y = hasB ? polar_clip(rawB[512], rawB[NBIN + 512]).x : 0.f;
          } else {
            nyq.x = rawA[1024];
            nyq.y = hasB ? rawB[1024] : 0.f;
          }
        }
        __syncwarp();  // the raw rows are consumed: the tile now belongs to the FFT
        // upper half: register r in [16,32) of lane L is Z[32 r + L] = mirrored value of bin 1024 - 32 r - L, which
        // lane (32 - L) & 31 computed as zc[31 - r] (L = 0: own zc[32 - r]; r = 16: the Nyquist bin)
        const int src = (32 - lane) & 31;
#pragma unroll
        for (int r = 16; r < 32; ++r) {
          float2 g;
          g.x = __shfl_sync(0xffffffffu, zc[31 - r].x, src);
          g.y = __shfl_sync(0xffffffffu, zc[31 - r].y, src);
          if (lane == 0) g = (r == 16) ? nyq : zc[(32 - r) & 15];
          v[r] = fw::cpack(g.x, g.y);
        }
        fw::fft1024_warp<true>(v, xb, tw, lane);  // real = frame A, imag = frame B, time index n = lane + 32 k2
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = 0ull;
      }
      // window, overlap-add: hop A = carry0 + A[0:256); hop A+1 = carry1 + A[256:512) + B[0:256); the rest is carried
      float o0[8], o1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = lane + 32 * j;
        const float w0 = wins[n], w1 = wins[HOP + n], w2 = wins[2 * HOP + n], w3 = wins[3 * HOP + n];
        const float2 p0 = fw::cunpack(fw::cmul2(v[j], fw::cpack(w0, w0)));
        const float2 p1 = fw::cunpack(fw::cmul2(v[8 + j], fw::cpack(w1, w1)));
        const float2 p2 = fw::cunpack(fw::cmul2(v[16 + j], fw::cpack(w2, w2)));
        const float2 p3 = fw::cunpack(fw::cmul2(v[24 + j], fw::cpack(w3, w3)));
        const float c0 = cw[32 * j], c1 = cw[32 * (8 + j)], c2 = cw[32 * (16 + j)];
        o0[j] = c0 + p0.x;
        o1[j] = c1 + p1.x + p0.y;
        cw[32 * j] = c2 + p2.x + p1.y;
        cw[32 * (8 + j)] = p3.x + p2.y;
        cw[32 * (16 + j)] = p3.y;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int hh = A + e;
        if (hh < h0 || hh >= h1) continue;  // halo pair, or the odd tail of the run
        float* dst = ob + (long long)hh * HOP;
        if (hh >= 3 && hh <= n_frames - 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[32 * j] = (e == 0 ? o0[j] : o1[j]) * ienv[lane + 32 * j];
        } else {  // first / last hop of the clip: fewer than four frames overlap
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float env = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int t = hh - 3 + q;
              const float w = window[lane + 32 * j + HOP * (3 - q)];
              if (t >= 0 && t < n_frames) env = fmaf(w, w, env);
            }
            dst[32 * j] = (e == 0 ? o0[j] : o1[j]) / env;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// peak normalisation
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, long long ldx, int n, float* __restrict__ scratch) {
  const int b = blockIdx.y;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[(long long)b * ldx + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(scratch + b), __float_as_int(m));  // m >= 0
}
__global__ void __launch_bounds__(256)
peak_scale_kernel(const float* __restrict__ x, long long ldx, int n, const float* __restrict__ scratch,
                  float* __restrict__ out, long long ldo) {
  const int b = blockIdx.y;
  const float mx = scratch[b];
  const bool silent = mx < 1e-8f;
  const float den = mx + 1e-7f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[(long long)b * ldx + i];
    out[(long long)b * ldo + i] = silent ? v : fminf(fmaxf(v / den, -1.0f), 1.0f);
  }
}

}  // namespace

extern "C" int64_t oron_logmel_bands_bytes(void) { return MEL_BANDS_BYTES; }

extern "C" int oron_logmel_bands(const float* fb, int32_t n_mels, void* bands, oron_stream_t stream) {
  if (!fb || !bands) return fail(ORON_ERR_BAD_ARG, "logmel_bands: null pointer");
  if (n_mels <= 0 || n_mels > MEL_MAXM) return fail(ORON_ERR_BAD_ARG, "logmel_bands: n_mels must be in [1,128]");
  if ((reinterpret_cast<uintptr_t>(bands) & 15) != 0) return fail(ORON_ERR_BAD_ARG, "logmel_bands: buffer must be 16-byte aligned");
  logmel_bands_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fb, n_mels, reinterpret_cast<uint8_t*>(bands));
  return check_launch("logmel_bands");
}

extern "C" int oron_logmel(const float* wav, int64_t ld_wav, int32_t nb, int32_t n_samples, const float* window,
                           const float* fb, const void* bands, int32_t n_mels, float clip, float* out, oron_stream_t stream) {
  if (!wav || !window || !fb || !bands || !out) return fail(ORON_ERR_BAD_ARG, "logmel: null pointer");
  if ((reinterpret_cast<uintptr_t>(bands) & 15) != 0) return fail(ORON_ERR_BAD_ARG, "logmel: bands must be 16-byte aligned");
  if (nb <= 0 || n_mels <= 0 || n_mels > MEL_MAXM) return fail(ORON_ERR_BAD_ARG, "logmel: n_mels must be in [1,128]");
  if (n_samples <= NFFT / 2) return fail(ORON_ERR_BAD_ARG, "logmel: reflect padding needs more than 512 samples");
  const int n_frames = 1 + n_samples / HOP;
  const long long chunks = (long long)((n_frames + AUD_FR - 1) / AUD_FR) * nb;
  const long long cap = 2LL * num_sms();  // persistent, two 8-warp CTAs per SM: the table setup is amortised over many chunks
  dim3 grid((unsigned)(chunks < cap ? chunks : cap));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEL_SMEM);
    if (e != cudaSuccess) return fail(int(e), "logmel smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  logmel_kernel<<<grid, AUD_THREADS, MEL_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      wav, ld_wav, nb, n_samples, n_frames, window, fb, reinterpret_cast<const uint8_t*>(bands), n_mels, clip, out);
  return check_launch("logmel");
}

extern "C" int oron_istft_head(const float* h, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t n_frames,
                               const float* window, int32_t mode, float* out, int64_t ld_out, oron_stream_t stream) {
  if (!h || !window || !out) return fail(ORON_ERR_BAD_ARG, "istft_head: null pointer");
  if (n_frames < 2 || n_frames > rows_per_batch || nb <= 0) return fail(ORON_ERR_BAD_ARG, "istft_head: bad frame count");
  if (ldh < 2 * NBIN) return fail(ORON_ERR_BAD_ARG, "istft_head: ldh must be >= 1026");
  // runs of output hops, one warp each: as long as possible (four halo frames per run) while every warp of the
  // persistent grid (two 8-warp CTAs per SM) still gets the same number of runs
  const long long hops = n_frames - 1;
  const long long slots = 2LL * num_sms() * AUD_WARPS;
  long long rpc = slots / nb;
  if (rpc < 1) rpc = 1;
  long long run = (hops + rpc - 1) / rpc;
  if (run < IST_MIN_RUN) run = IST_MIN_RUN;
  run = (run + 1) & ~1LL;
  rpc = (hops + run - 1) / run;
  const long long ctas = (rpc * nb + AUD_WARPS - 1) / AUD_WARPS;
  const long long cap = 2LL * num_sms();
  dim3 grid((unsigned)(ctas < cap ? ctas : cap));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(istft_head_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, IST_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(istft_head_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, IST_SMEM);
    if (e != cudaSuccess) return fail(int(e), "istft smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == 0)
    istft_head_kernel<0><<<grid, AUD_THREADS, IST_SMEM, st>>>(h, ldh, rows_per_batch, nb, n_frames, window, out, ld_out, int(run), int(rpc));
  else
    istft_head_kernel<1><<<grid, AUD_THREADS, IST_SMEM, st>>>(h, ldh, rows_per_batch, nb, n_frames, window, out, ld_out, int(run), int(rpc));
  return check_launch("istft_head");
}

extern "C" int oron_peak_normalize(const float* x, int64_t ldx, int32_t nb, int32_t n, float* out, int64_t ldo,
                                   float* scratch, oron_stream_t stream) {
  if (!x || !out || !scratch || nb <= 0 || n <= 0) return fail(ORON_ERR_BAD_ARG, "peak_normalize: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * nb, st);
  if (e != cudaSuccess) return fail(int(e), "peak_normalize memset: %s", cudaGetErrorString(e));
  int bx = (n + 256 * 8 - 1) / (256 * 8);
  if (bx > 2 * num_sms()) bx = 2 * num_sms();
  dim3 grid(bx, nb);
  absmax_kernel<<<grid, 256, 0, st>>>(x, ldx, n, scratch);
  int rc = check_launch("absmax");
  if (rc) return rc;
  peak_scale_kernel<<<grid, 256, 0, st>>>(x, ldx, n, scratch, out, ldo);
  return check_launch("peak_scale");
}
