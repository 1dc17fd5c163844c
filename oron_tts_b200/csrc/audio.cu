// Log-mel STFT front end, Vocos iSTFT head and peak normalisation (n_fft 1024, hop 256).
// Both transforms run a 1024-point radix-4 Stockham FFT in shared memory and pack TWO real
// frames into one complex transform (frame A -> real part, frame B -> imaginary part), so the
// complex spectrum / frame buffers never touch HBM: log-mel reads the waveform once and writes
// [n_mels, T]; the iSTFT reads the head activations once and writes the waveform once.
#include "../../include/oron_b200.h"
#include "host_util.h"
#include "ptx.cuh"

using namespace oron;

namespace {

constexpr int NFFT = 1024;
constexpr int HOP = 256;
constexpr int NBIN = 513;
constexpr int FFT_THREADS = 256;

// W[n] = exp(-2 pi i n / 1024)
__device__ __forceinline__ void fill_twiddles(float2* tw) {
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) {
    float s, c;
    sincospif(float(n) * (2.0f / NFFT), &s, &c);
    tw[n] = make_float2(c, -s);
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place-ish 1024-point complex FFT over two smem buffers; 256 threads, 5 radix-4 passes.
// INVERSE uses conjugated twiddles (no 1/N scaling). Result ends in `a` if the pass count is even,
// else in `b`: 5 passes -> result in b.  Caller must __syncthreads() before reading.
template <bool INVERSE>
__device__ __forceinline__ void fft1024(float2* a, float2* b, const float2* tw) {
  float2* src = a;
  float2* dst = b;
  const int j = threadIdx.x;  // 0..255
#pragma unroll
  for (int Ns = 1; Ns < NFFT; Ns <<= 2) {
    const int k = j & (Ns - 1);
    float2 v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = src[j + r * (NFFT / 4)];
    if (Ns > 1) {
      const int tstep = k * (NFFT / 4 / Ns);  // angle index for r = 1
#pragma unroll
      for (int r = 1; r < 4; ++r) {
        float2 w = tw[r * tstep];
        if (INVERSE) w.y = -w.y;
        v[r] = cmul(v[r], w);
      }
    }
    const float2 a0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
    const float2 a1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
    const float2 a2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
    const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
    // forward: multiply by -i -> (y, -x); inverse: multiply by +i -> (-y, x)
    const float2 a3 = INVERSE ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    const int base = (j / Ns) * Ns * 4 + k;
    dst[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
    dst[base + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
    dst[base + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
    dst[base + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
}

// ---------------------------------------------------------------------------
// log-mel
// ---------------------------------------------------------------------------
constexpr int MEL_FR = 32;       // frames per CTA
constexpr int MEL_MAXBAND = 32;  // widest triangular filter (bins) kept in smem
constexpr int MEL_MAXM = 128;
constexpr int MEL_SMEM = 3 * NFFT * 8 + NFFT * 4 + 2 * (NBIN + 3) * 4 + MEL_MAXM * MEL_MAXBAND * 4 +
                         MEL_FR * (MEL_MAXM + 1) * 4 + 2 * MEL_MAXM * 4;

__global__ void __launch_bounds__(FFT_THREADS)
logmel_kernel(const float* __restrict__ wav, long long ld_wav, int n_samples, int n_frames,
              const float* __restrict__ window, const float* __restrict__ fb, int n_mels, float clip,
              float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t dsm[];
  float2* bufA = reinterpret_cast<float2*>(dsm);
  float2* bufB = bufA + NFFT;
  float2* tw = bufB + NFFT;
  float* win = reinterpret_cast<float*>(tw + NFFT);
  float (*mag)[NBIN + 3] = reinterpret_cast<float (*)[NBIN + 3]>(win + NFFT);
  float (*fbs)[MEL_MAXBAND] = reinterpret_cast<float (*)[MEL_MAXBAND]>(&mag[2][0]);
  float (*mel_s)[MEL_MAXM + 1] = reinterpret_cast<float (*)[MEL_MAXM + 1]>(&fbs[MEL_MAXM][0]);
  int* band_lo = reinterpret_cast<int*>(&mel_s[MEL_FR][0]);
  int* band_n = band_lo + MEL_MAXM;

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MEL_FR;
  const float* x = wav + (long long)b * ld_wav;

  fill_twiddles(tw);
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) win[n] = window[n];
  // band limits of every mel filter (filters are contiguous triangles)
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBIN; ++k) {
      if (fb[(long long)k * n_mels + m] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    }
    if (lo < 0) { lo = 0; hi = -1; }
    band_lo[m] = lo;
    band_n[m] = hi - lo + 1;
    for (int i = 0; i < MEL_MAXBAND; ++i)
      fbs[m][i] = (i <= hi - lo && i < MEL_MAXBAND) ? fb[(long long)(lo + i) * n_mels + m] : 0.f;
  }
  __syncthreads();

  const int nfr = min(MEL_FR, n_frames - t0);
  for (int f = 0; f < nfr; f += 2) {
    // frame pair (t0+f, t0+f+1): reflect-padded (center=True) windowed samples
    const bool hasB = (f + 1 < nfr);
    for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
      int nA = (t0 + f) * HOP + i - NFFT / 2;
      if (nA < 0) nA = -nA;
      if (nA >= n_samples) nA = 2 * (n_samples - 1) - nA;
      float vb = 0.f;
      if (hasB) {
        int nB = (t0 + f + 1) * HOP + i - NFFT / 2;
        if (nB < 0) nB = -nB;
        if (nB >= n_samples) nB = 2 * (n_samples - 1) - nB;
        vb = x[nB] * win[i];
      }
      bufA[i] = make_float2(x[nA] * win[i], vb);
    }
    __syncthreads();
    fft1024<false>(bufA, bufB, tw);  // 5 passes: result in bufB (fft1024 ends with a barrier)
    // split the two real spectra and take magnitudes
    for (int k = threadIdx.x; k < NBIN; k += blockDim.x) {
      const float2 z = bufB[k];
      const float2 zn = bufB[(NFFT - k) & (NFFT - 1)];
      const float ar = 0.5f * (z.x + zn.x), ai = 0.5f * (z.y - zn.y);
      const float br = 0.5f * (z.y + zn.y), bi = -0.5f * (z.x - zn.x);
      mag[0][k] = sqrtf(ar * ar + ai * ai);
      mag[1][k] = sqrtf(br * br + bi * bi);
    }
    __syncthreads();
    // banded mel projection: threads [0,128) frame A, [128,256) frame B
    {
      const int which = threadIdx.x >> 7;
      const int m = threadIdx.x & 127;
      if (m < n_mels && (which == 0 || hasB)) {
        const int lo = band_lo[m], n = band_n[m];
        float acc = 0.f;
        if (n <= MEL_MAXBAND) {
          for (int i = 0; i < n; ++i) acc += fbs[m][i] * mag[which][lo + i];
        } else {
          for (int i = 0; i < n; ++i) acc += fb[(long long)(lo + i) * n_mels + m] * mag[which][lo + i];
        }
        mel_s[f + which][m] = logf(fmaxf(acc, clip));
      }
    }
    __syncthreads();
  }
  // out[b, m, t0 + f]: contiguous along frames
  for (int i = threadIdx.x; i < n_mels * MEL_FR; i += blockDim.x) {
    const int m = i / MEL_FR, f = i % MEL_FR;
    if (f < nfr) out[((long long)b * n_mels + m) * n_frames + t0 + f] = mel_s[f][m];
  }
}

// ---------------------------------------------------------------------------
// iSTFT head: spectrum from the head activations -> irfft -> window -> overlap-add -> envelope
// ---------------------------------------------------------------------------
constexpr int IST_HOPS = 29;             // output hops finished per CTA
constexpr int IST_FR = IST_HOPS + 3;     // frames transformed per CTA (3-frame halo recomputed)
constexpr int IST_SMEM = 3 * NFFT * 8 + NFFT * 4 + IST_HOPS * HOP * 4;

__global__ void __launch_bounds__(FFT_THREADS)
istft_head_kernel(const float* __restrict__ h, long long ldh, int rows_per_batch, int n_frames,
                  const float* __restrict__ window, int mode, float* __restrict__ out, long long ld_out) {
  extern __shared__ __align__(16) uint8_t dsm[];
  float2* bufA = reinterpret_cast<float2*>(dsm);
  float2* bufB = bufA + NFFT;
  float2* tw = bufB + NFFT;
  float* win = reinterpret_cast<float*>(tw + NFFT);
  float* ola = win + NFFT;

  const int b = blockIdx.y;
  const int hop0 = blockIdx.x * IST_HOPS;          // first padded-domain hop finished by this CTA
  const int hop1 = min(hop0 + IST_HOPS, n_frames + 3);
  const int fr_lo = max(hop0 - 3, 0);
  const int fr_hi = min(hop1 - 1, n_frames - 1);   // inclusive
  const long long base_n = (long long)hop0 * HOP;  // padded-domain sample index of ola[0]

  fill_twiddles(tw);
  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) win[n] = window[n];
  for (int n = threadIdx.x; n < IST_HOPS * HOP; n += blockDim.x) ola[n] = 0.f;
  __syncthreads();

  const float scale = (mode == 1) ? (32.0f / NFFT) : (1.0f / NFFT);  // normalized=True: * sqrt(N)
  for (int f = fr_lo; f <= fr_hi; f += 2) {
    const bool hasB = (f + 1 <= fr_hi);
    const float* ha = h + ((long long)b * rows_per_batch + f) * ldh;
    const float* hb = ha + ldh;
    // Hermitian-extend both half spectra and pack Z = A + iB
    for (int k = threadIdx.x; k < NBIN; k += blockDim.x) {
      float ar, ai, br = 0.f, bi = 0.f;
      if (mode == 0) {
        const float mg = fminf(expf(ha[k]), 100.0f);
        float s, c;
        sincosf(ha[NBIN + k], &s, &c);
        ar = mg * c; ai = mg * s;
        if (hasB) {
          const float mg2 = fminf(expf(hb[k]), 100.0f);
          sincosf(hb[NBIN + k], &s, &c);
          br = mg2 * c; bi = mg2 * s;
        }
      } else {
        ar = ha[2 * k]; ai = ha[2 * k + 1];
        if (hasB) { br = hb[2 * k]; bi = hb[2 * k + 1]; }
      }
      if (k == 0 || k == NFFT / 2) { ai = 0.f; bi = 0.f; }  // C2R ignores these imaginary parts
      // Z[k] = A[k] + i B[k] ; Z[N-k] = conj(A[k]) + i conj(B[k])
      bufA[k] = make_float2(ar - bi, ai + br);
      if (k != 0 && k != NFFT / 2) bufA[NFFT - k] = make_float2(ar + bi, br - ai);
    }
    __syncthreads();
    fft1024<true>(bufA, bufB, tw);  // result in bufB: real = frame A, imag = frame B
    // overlap-add frame A, then frame B (they overlap each other, so two phases)
    for (int which = 0; which < 2; ++which) {
      if (which == 1 && !hasB) break;
      const long long fstart = (long long)(f + which) * HOP - base_n;  // offset of sample 0 in ola
      for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
        const long long p = fstart + i;
        if (p >= 0 && p < (long long)(hop1 - hop0) * HOP) {
          const float v = (which == 0 ? bufB[i].x : bufB[i].y) * scale * win[i];
          ola[p] += v;
        }
      }
      __syncthreads();
    }
  }
  // envelope-normalise and write; output index o = n - 512, valid o in [0, 256*(n_frames-1))
  const long long out_len = (long long)HOP * (n_frames - 1);
  for (int i = threadIdx.x; i < (hop1 - hop0) * HOP; i += blockDim.x) {
    const long long n = base_n + i;
    const long long o = n - NFFT / 2;
    if (o < 0 || o >= out_len) continue;
    // frames covering n: t in [ceil((n-1023)/256), floor(n/256)] intersect [0, n_frames-1]
    int t_hi = int(n / HOP);
    int t_lo = t_hi - 3;
    if (t_lo < 0) t_lo = 0;
    if (t_hi > n_frames - 1) t_hi = n_frames - 1;
    float env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
      const float w = win[n - (long long)t * HOP];
      env += w * w;
    }
    out[(long long)b * ld_out + o] = ola[i] / env;
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

extern "C" int oron_logmel(const float* wav, int64_t ld_wav, int32_t nb, int32_t n_samples, const float* window,
                           const float* fb, int32_t n_mels, float clip, float* out, oron_stream_t stream) {
  if (!wav || !window || !fb || !out) return fail(ORON_ERR_BAD_ARG, "logmel: null pointer");
  if (nb <= 0 || n_mels <= 0 || n_mels > MEL_MAXM) return fail(ORON_ERR_BAD_ARG, "logmel: n_mels must be in [1,128]");
  if (n_samples <= NFFT / 2) return fail(ORON_ERR_BAD_ARG, "logmel: reflect padding needs more than 512 samples");
  const int n_frames = 1 + n_samples / HOP;
  dim3 grid((n_frames + MEL_FR - 1) / MEL_FR, nb);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEL_SMEM);
    if (e != cudaSuccess) return fail(int(e), "logmel smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  logmel_kernel<<<grid, FFT_THREADS, MEL_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      wav, ld_wav, n_samples, n_frames, window, fb, n_mels, clip, out);
  return check_launch("logmel");
}

extern "C" int oron_istft_head(const float* h, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t n_frames,
                               const float* window, int32_t mode, float* out, int64_t ld_out, oron_stream_t stream) {
  if (!h || !window || !out) return fail(ORON_ERR_BAD_ARG, "istft_head: null pointer");
  if (n_frames < 2 || n_frames > rows_per_batch || nb <= 0) return fail(ORON_ERR_BAD_ARG, "istft_head: bad frame count");
  if (ldh < 2 * NBIN) return fail(ORON_ERR_BAD_ARG, "istft_head: ldh must be >= 1026");
  dim3 grid((n_frames + 3 + IST_HOPS - 1) / IST_HOPS, nb);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(istft_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IST_SMEM);
    if (e != cudaSuccess) return fail(int(e), "istft smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  istft_head_kernel<<<grid, FFT_THREADS, IST_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      h, ldh, rows_per_batch, n_frames, window, mode, out, ld_out);
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
