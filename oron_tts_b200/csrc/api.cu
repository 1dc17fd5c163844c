// C ABI (include/oron_b200.h) for the tensor-core and row-wise kernels: argument checks,
// TMA descriptor construction and launches. No device memory is allocated here.
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "../../include/oron_b200.h"
#include "attn_fwd4.cuh"
#include "attn_tcgen05.cuh"
#include "ffn_tcgen05.cuh"
#include "gconv_res_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "host_util.h"
#include "rowwise.cuh"

using namespace oron;

namespace oron {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(int(e), "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ORON_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
}  // namespace oron

// ---------------------------------------------------------------------------
// TMA descriptors
// ---------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), SWIZZLE_128B, box = (64, box1, 1); OOB reads give zeros.
int oron::make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box1, int rank) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return fail(ORON_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(ORON_ERR_BAD_ARG, "TMA base not 16-byte aligned");
  if ((stride1_elems * 2) % 16 != 0 || (rank == 3 && (stride2_elems * 2) % 16 != 0))
    return fail(ORON_ERR_BAD_ARG, "TMA strides must be multiples of 16 bytes (ld %llu)", (unsigned long long)stride1_elems);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ORON_ERR_BAD_ARG, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

// ---------------------------------------------------------------------------
// GEMM
// ---------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(int(e), "gemm smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles_m = ((a.rows_per_batch + GEMM_BM - 1) / GEMM_BM) * a.nbatch;
  const int tiles_n = (a.N + BN - 1) / BN;
  int grid = tiles_m * tiles_n;
  const int cap = max_ctas > 0 ? max_ctas : num_sms();
  if (grid > cap) grid = cap;
  if (grid <= 0) return 0;
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ta, tb, a);
  if (le != cudaSuccess) return fail(int(le), "gemm launch: %s", cudaGetErrorString(le));
  return check_launch("gemm_bf16_tcgen05");
}


template <int BN, int EPI, int MNM = 0, int LNT = 0>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int max_ctas, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  auto kern = gemm2_bf16_tcgen05_kernel<BN, EPI, MNM, LNT>;
  static bool configured = false;
  static int resident_pairs = 0;  // LNT: the tail waits on other SM pairs, so every cluster of the grid must be resident at once
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(int(e), "gemm2 smem attribute: %s", cudaGetErrorString(e));
    if (LNT) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(unsigned(num_sms() & ~1));
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); return fail(ORON_ERR_UNSUPPORTED, "gemm+ln: cluster occupancy query failed"); }
      resident_pairs = n;
    }
    configured = true;
  }
  const int tiles_m = ((a.rows_per_batch + GEMM_BM - 1) / GEMM_BM) * a.nbatch;
  const int tiles = ((tiles_m + 1) / 2) * ((a.N + BN - 1) / BN);
  long long pairs = a.stream_k ? (long long)tiles * a.num_kb : tiles;  // stream-K: every pair takes an equal share of k-blocks
  int cap = (max_ctas > 0 ? max_ctas : num_sms()) / 2;
  if (LNT && cap > resident_pairs) cap = resident_pairs;
  if (pairs > cap) pairs = cap;
  if (pairs <= 0) return 0;
  cudaError_t le = launch_pdl(kern, dim3(unsigned(2 * pairs)), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ta, tb, a);
  if (le != cudaSuccess) return fail(int(le), "gemm2 launch: %s", cudaGetErrorString(le));
  return check_launch("gemm2_bf16_tcgen05");
}

// descriptor checks + kernel arguments + TMA maps of one GEMM (shared by oron_gemm_bf16 and oron_ffn_bf16)
static int gemm_prepare(const oron_gemm_desc* d, GemmArgs& a, CUtensorMap& ta, CUtensorMap& tb) {
  if (!d || !d->A || !d->W || !d->out) return fail(ORON_ERR_BAD_ARG, "gemm: null pointer");
  if (d->rows_per_batch <= 0 || d->nbatch <= 0 || d->N <= 0 || d->w_cols <= 0)
    return fail(ORON_ERR_BAD_ARG, "gemm: bad shape");
  const int taps = d->taps > 0 ? d->taps : 1;
  memset(&a, 0, sizeof(a));
  a.rows_per_batch = d->rows_per_batch;
  a.nbatch = d->nbatch;
  a.N = d->N;
  if (taps == 1) {
    a.num_kb = (d->w_cols + GEMM_BK - 1) / GEMM_BK;
    a.cpb = a.num_kb;
    a.pad = 0;
    a.grouped = 0;
  } else {
    if (d->cin_blocks <= 0 || d->w_cols != taps * d->cin_blocks * GEMM_BK)
      return fail(ORON_ERR_BAD_ARG, "conv-gemm: w_cols must equal taps*cin_blocks*64");
    if (d->grouped && (d->block_n != 64 || d->grouped != d->cin_blocks * GEMM_BK))
      return fail(ORON_ERR_BAD_ARG, "grouped conv-gemm needs block_n == 64 and grouped == cin_blocks*64");
    a.num_kb = taps * d->cin_blocks;
    a.cpb = d->cin_blocks;
    a.pad = d->pad;
    a.grouped = d->grouped;
  }
  a.act = d->act;
  a.bias = d->bias;
  a.out = d->out;
  a.ldo = d->ldo;
  a.out2 = d->out2;
  a.ldo2 = d->ldo2;
  a.addend = d->addend;
  a.ld_add = d->ld_add;
  a.gate = d->gate;
  a.gate_ld = d->gate_ld;
  a.gate_nb = d->gate_nb > 0 ? d->gate_nb : 1;
  a.gate_step_stride = d->gate_step_stride;
  a.step_ptr = d->step_ptr;
  a.rope_cos = d->rope_cos;
  a.rope_sin = d->rope_sin;
  a.rope_cols = d->rope_cols;
  a.f16_from_col = d->f16_from_col > 0 ? d->f16_from_col : 0x7fffffff;
  a.seq_lens = d->seq_lens;
  a.row_valid = d->row_valid;
  a.mask_rows = d->mask_rows;
  a.stream_k = d->stream_k != 0 ? 1 : 0;
  a.dbg = reinterpret_cast<long long*>(d->debug_stamps);

  const int epi = d->epilogue;
  // per-epilogue operand checks (vectorised epilogues assume 32-column granularity)
  const bool vec_epi = epi == EPI_QKV_ROPE || epi == EPI_GATE_RESID || epi == EPI_EMBED_DUAL ||
                       epi == EPI_MISH_MASK_BF16 || epi == EPI_MISH_MASK_RESID || epi == EPI_SCALE_RESID ||
                       epi == EPI_GELU_DROP_DUAL || epi == EPI_GELU_DROP_BWD || epi == EPI_GATE_RESID_DUAL;
  if ((epi == EPI_GELU_DROP_DUAL || epi == EPI_GELU_DROP_BWD || epi == EPI_GATE_RESID_DUAL) &&
      (!d->out2 || d->ldo2 % 8 != 0 || !d->two_sm))
    return fail(ORON_ERR_BAD_ARG, "gemm: the training epilogues need out2 (ldo2 %% 8 == 0) and two_sm");
  if (epi == EPI_GATE_RESID_DUAL && (!d->gate || !d->addend)) return fail(ORON_ERR_BAD_ARG, "gemm: GATE_RESID_DUAL needs gate and addend");
  {
    float p = d->dropout_p;
    a.drop.thresh = 0u;
    a.drop.inv_keep = 1.f;
    a.drop.key = 0u;
    if (p > 0.f) {
      if (p > 0.999f) p = 0.999f;
      uint64_t z = d->dropout_seed + 0x9E3779B97F4A7C15ull;  // splitmix64, as in train.cu
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      a.drop.key = (unsigned int)(z >> 32) ^ (unsigned int)z;
      a.drop.thresh = (unsigned int)((double)p * 4294967296.0);
      if (a.drop.thresh == 0u) a.drop.thresh = 1u;
      a.drop.inv_keep = 1.0f / (1.0f - p);
    }
  }
  if (vec_epi && (d->N % 32 != 0)) return fail(ORON_ERR_BAD_ARG, "gemm: epilogue %d needs N %% 32 == 0", epi);
  if (epi == EPI_QKV_ROPE && (!d->bias || !d->rope_cos || !d->rope_sin || d->N % 64 != 0 || d->block_n < 128))
    return fail(ORON_ERR_BAD_ARG, "gemm: QKV_ROPE needs bias, rope tables, N %% 64 == 0, block_n >= 128");
  if (epi == EPI_GATE_RESID && !d->gate) return fail(ORON_ERR_BAD_ARG, "gemm: GATE_RESID needs gate");
  if (d->stream_k && !(d->two_sm && (epi == EPI_GATE_RESID || epi == EPI_F32) && taps == 1))
    return fail(ORON_ERR_BAD_ARG, "gemm: stream_k needs two_sm, the GATE_RESID or F32 epilogue and taps == 1");
  if (d->stream_k && epi == EPI_F32 && (d->addend != nullptr || d->N % 4 != 0))
    return fail(ORON_ERR_BAD_ARG, "gemm: stream_k with the F32 epilogue adds into out: no addend, N %% 4 == 0");
  if ((epi == EPI_EMBED_DUAL || epi == EPI_MISH_MASK_RESID || epi == EPI_SCALE_RESID) && !d->addend)
    return fail(ORON_ERR_BAD_ARG, "gemm: epilogue %d needs addend", epi);
  if (epi == EPI_EMBED_DUAL && !d->out2) return fail(ORON_ERR_BAD_ARG, "gemm: EMBED_DUAL needs out2");
  if ((d->ldo % 8) != 0 && epi != EPI_F32) return fail(ORON_ERR_BAD_ARG, "gemm: ldo must be a multiple of 8");
  if (epi == EPI_F32 && (d->ldo % 4) != 0) return fail(ORON_ERR_BAD_ARG, "gemm: f32 ldo must be a multiple of 4");

  const bool two_sm = d->two_sm != 0;
  a.a_mn = d->a_mn_major != 0 ? 1 : 0;
  a.b_mn = d->b_mn_major != 0 ? 1 : 0;
  if (a.a_mn || a.b_mn) {
    if (!two_sm || taps != 1 || d->w_cols % GEMM_BK != 0 || (a.a_mn && d->nbatch != 1))
      return fail(ORON_ERR_UNSUPPORTED, "gemm: MN-major operands need two_sm, taps == 1, K %% 64 == 0 (and nbatch == 1 for A)");
  }
  int rc;
  if (a.a_mn)  // source [K, lda], columns [0, M): box = 64 MN x 64 K rows
    rc = make_tmap_bf16(&ta, d->A, uint64_t(d->rows_per_batch), uint64_t(d->w_cols), 1, uint64_t(d->lda), 0, 64, 2);
  else
    rc = make_tmap_bf16(&ta, d->A, uint64_t(d->a_cols), uint64_t(d->rows_per_batch), uint64_t(d->nbatch),
                        uint64_t(d->lda), uint64_t(d->lda) * uint64_t(d->rows_per_batch), GEMM_BM, 3);
  if (rc) return rc;
  if (a.b_mn)  // source [K, ldw], columns [0, N)
    rc = make_tmap_bf16(&tb, d->W, uint64_t(d->N), uint64_t(d->w_cols), 1, uint64_t(d->ldw), 0, 64, 2);
  else
    rc = make_tmap_bf16(&tb, d->W, uint64_t(d->w_cols), uint64_t(d->N), 1, uint64_t(d->ldw), 0,
                        uint32_t(two_sm ? d->block_n / 2 : d->block_n), 2);
  return rc;
}

// grouped conv with the activation window resident in shared memory (gconv_res_tcgen05.cuh)
template <int EPI>
static int launch_gconv_res(const oron_gemm_desc* d, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t st) {
  using Cfg = GconvResCfg;
  auto kern = gconv_res_tcgen05_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(int(e), "gconv_res smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap ta;  // same tensor as the generic kernel's A map, but one box = the whole (128 + taps - 1)-row window
  if (int rc = make_tmap_bf16(&ta, d->A, uint64_t(d->a_cols), uint64_t(d->rows_per_batch), uint64_t(d->nbatch), uint64_t(d->lda),
                              uint64_t(d->lda) * uint64_t(d->rows_per_batch), Cfg::kWinRows, 3))
    return rc;
  const int tiles_m = ((a.rows_per_batch + GEMM_BM - 1) / GEMM_BM) * a.nbatch;
  int grid = tiles_m * (a.N / Cfg::BN);
  const int cap = d->max_ctas > 0 ? d->max_ctas : num_sms();
  if (grid > cap) grid = cap;
  if (grid <= 0) return 0;
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ta, tb, a);
  if (le != cudaSuccess) return fail(int(le), "gconv_res launch: %s", cudaGetErrorString(le));
  return check_launch("gconv_res_tcgen05");
}

extern "C" int oron_gemm_bf16(const oron_gemm_desc* d, oron_stream_t stream) {
  GemmArgs a;
  CUtensorMap ta, tb;
  if (int rc = gemm_prepare(d, a, ta, tb)) return rc;
  const int epi = d->epilogue;
  const bool two_sm = d->two_sm != 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  {  // 64-channel grouped conv: resident-window kernel (ORON_GCONV_RES=0 keeps the generic per-tap fetch for A/B runs)
    const char* e = getenv("ORON_GCONV_RES");  // read per call (cheap): tests flip it inside one process
    const int res = (e && atoi(e) == 0) ? 0 : 1;
    const int taps = d->taps > 0 ? d->taps : 1;
    if (res && !two_sm && taps > 1 && taps <= GconvResCfg::kMaxTaps && d->grouped == GEMM_BK && d->cin_blocks == 1 &&
        d->block_n == GconvResCfg::BN && d->N % GconvResCfg::BN == 0 && d->pad >= 0 && d->pad < taps) {
#define ORON_GCONV_RES_CASE(EPI_) if (epi == EPI_) return launch_gconv_res<EPI_>(d, tb, a, st);
      ORON_GCONV_RES_CASE(EPI_BF16)
      ORON_GCONV_RES_CASE(EPI_F32)
      ORON_GCONV_RES_CASE(EPI_MISH_MASK_BF16)
      ORON_GCONV_RES_CASE(EPI_MISH_MASK_RESID)
      ORON_GCONV_RES_CASE(EPI_SCALE_RESID)
#undef ORON_GCONV_RES_CASE
    }
  }

  if (a.a_mn || a.b_mn) {  // backward-pass layouts: compile-time variants of the 2-SM kernel
    const int mnm = a.a_mn | (a.b_mn << 1);
#define ORON_GEMM2_MN_CASE(BN_, EPI_, MNM_) \
    if (d->block_n == BN_ && epi == EPI_ && mnm == MNM_) return launch_gemm2<BN_, EPI_, MNM_>(ta, tb, a, d->max_ctas, st);
    ORON_GEMM2_MN_CASE(128, EPI_BF16, 2)
    ORON_GEMM2_MN_CASE(256, EPI_BF16, 2)
    ORON_GEMM2_MN_CASE(128, EPI_GELU_DROP_BWD, 2)
    ORON_GEMM2_MN_CASE(256, EPI_GELU_DROP_BWD, 2)
    ORON_GEMM2_MN_CASE(128, EPI_F32, 2)
    ORON_GEMM2_MN_CASE(256, EPI_F32, 2)
    ORON_GEMM2_MN_CASE(128, EPI_F32, 3)
    ORON_GEMM2_MN_CASE(256, EPI_F32, 3)
#undef ORON_GEMM2_MN_CASE
    return fail(ORON_ERR_UNSUPPORTED, "gemm: no MN-major kernel for block_n=%d epilogue=%d a_mn=%d b_mn=%d", d->block_n, epi, a.a_mn, a.b_mn);
  }
#define ORON_GEMM2_CASE(BN_, EPI_) \
  if (two_sm && d->block_n == BN_ && epi == EPI_) return launch_gemm2<BN_, EPI_>(ta, tb, a, d->max_ctas, st);
  ORON_GEMM2_CASE(128, EPI_BF16)
  ORON_GEMM2_CASE(256, EPI_BF16)
  ORON_GEMM2_CASE(224, EPI_BF16)  // FFN up-projection at config 2: 19 x 11 = 209 tiles fill three waves of 74 SM pairs with cheaper k-blocks
  ORON_GEMM2_CASE(128, EPI_F32)
  ORON_GEMM2_CASE(256, EPI_F32)
  ORON_GEMM2_CASE(128, EPI_QKV_ROPE)
  ORON_GEMM2_CASE(256, EPI_QKV_ROPE)
  ORON_GEMM2_CASE(128, EPI_GATE_RESID)
  ORON_GEMM2_CASE(256, EPI_GATE_RESID)
  ORON_GEMM2_CASE(192, EPI_GATE_RESID)  // N = 1024 out-projection: 6 x 11 tiles = one wave of narrower (cheaper) tiles
  ORON_GEMM2_CASE(128, EPI_EMBED_DUAL)
  ORON_GEMM2_CASE(128, EPI_SCALE_RESID)
  ORON_GEMM2_CASE(256, EPI_SCALE_RESID)
  ORON_GEMM2_CASE(128, EPI_GELU_DROP_DUAL)
  ORON_GEMM2_CASE(256, EPI_GELU_DROP_DUAL)
  ORON_GEMM2_CASE(128, EPI_GATE_RESID_DUAL)
  ORON_GEMM2_CASE(256, EPI_GATE_RESID_DUAL)
#undef ORON_GEMM2_CASE
  if (two_sm) return fail(ORON_ERR_UNSUPPORTED, "gemm: no 2-SM kernel for block_n=%d epilogue=%d", d->block_n, epi);
#define ORON_GEMM_CASE(BN_, EPI_) \
  if (d->block_n == BN_ && epi == EPI_) return launch_gemm<BN_, EPI_>(ta, tb, a, d->max_ctas, st);
  ORON_GEMM_CASE(128, EPI_BF16)
  ORON_GEMM_CASE(256, EPI_BF16)
  ORON_GEMM_CASE(64, EPI_BF16)
  ORON_GEMM_CASE(128, EPI_F32)
  ORON_GEMM_CASE(256, EPI_F32)
  ORON_GEMM_CASE(64, EPI_F32)
  ORON_GEMM_CASE(128, EPI_QKV_ROPE)
  ORON_GEMM_CASE(256, EPI_QKV_ROPE)
  ORON_GEMM_CASE(128, EPI_GATE_RESID)
  ORON_GEMM_CASE(256, EPI_GATE_RESID)
  ORON_GEMM_CASE(64, EPI_GATE_RESID)
  ORON_GEMM_CASE(128, EPI_EMBED_DUAL)
  ORON_GEMM_CASE(64, EPI_MISH_MASK_BF16)
  ORON_GEMM_CASE(64, EPI_MISH_MASK_RESID)
  ORON_GEMM_CASE(128, EPI_SCALE_RESID)
  ORON_GEMM_CASE(256, EPI_SCALE_RESID)
  ORON_GEMM_CASE(64, EPI_SCALE_RESID)
#undef ORON_GEMM_CASE
  return fail(ORON_ERR_UNSUPPORTED, "gemm: no kernel for block_n=%d epilogue=%d", d->block_n, epi);
}

// ---------------------------------------------------------------------------
// GEMM + gated residual + LayerNorm / modulation of the updated rows in one launch
// ---------------------------------------------------------------------------
extern "C" int32_t oron_gemm_ln_counters(int32_t rows_per_batch, int32_t nbatch) {
  if (rows_per_batch <= 0 || nbatch <= 0) return 0;
  const int tiles_m = ((rows_per_batch + GEMM_BM - 1) / GEMM_BM) * nbatch;
  return ((tiles_m + 1) / 2 + 1) * LN_CNT_STRIDE;
}

extern "C" int oron_gemm_ln_bf16(const oron_gemm_desc* d, const oron_ln_tail* ln, oron_stream_t stream) {
  GemmArgs a;
  CUtensorMap ta, tb;
  if (int rc = gemm_prepare(d, a, ta, tb)) return rc;
  if (!ln || !ln->scale || !ln->out_bf16 || !ln->counters) return fail(ORON_ERR_BAD_ARG, "gemm+ln: null pointer");
  const int taps = d->taps > 0 ? d->taps : 1;
  if (!d->two_sm || d->epilogue != EPI_GATE_RESID || taps != 1 || a.a_mn || a.b_mn || (d->block_n != 256 && d->block_n != 192))
    return fail(ORON_ERR_UNSUPPORTED, "gemm+ln: needs two_sm, the GATE_RESID epilogue, K-major operands, taps == 1, block_n 256 or 192");
  if ((d->N != 1024 && d->N != 768 && d->N != 512 && d->N != 256 && d->N != 128) || d->ldo % 4 != 0 || ln->ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(ln->out_bf16) & 7) != 0)
    return fail(ORON_ERR_BAD_ARG, "gemm+ln: N in {128, 256, 512, 768, 1024}, 16-byte aligned f32 rows, 8-byte aligned bf16 rows");
  if (ln->n_counters < oron_gemm_ln_counters(d->rows_per_batch, d->nbatch))
    return fail(ORON_ERR_BAD_ARG, "gemm+ln: counters buffer of oron_gemm_ln_counters() int32 (zero-initialised) required");
  a.ln.counters = ln->counters;
  a.ln.scale = ln->scale;
  a.ln.shift = ln->shift;
  a.ln.mod_ld = ln->mod_ld;
  a.ln.mod_nb = ln->mod_nb > 0 ? ln->mod_nb : 1;
  a.ln.step_stride = ln->step_stride;
  a.ln.eps = ln->eps;
  a.ln.add_one = ln->add_one;
  a.ln.out = reinterpret_cast<__nv_bfloat16*>(ln->out_bf16);
  a.ln.ldo = ln->ldo;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->block_n == 256) return launch_gemm2<256, EPI_GATE_RESID, 0, 1>(ta, tb, a, d->max_ctas, st);
  return launch_gemm2<192, EPI_GATE_RESID, 0, 1>(ta, tb, a, d->max_ctas, st);
}

// ---------------------------------------------------------------------------
// FeedForward in one launch (ffn_tcgen05.cuh)
// ---------------------------------------------------------------------------
static int ffn_flag_count(int rows_per_batch, int nbatch, int ff_dim) {
  const int tiles_m = ((rows_per_batch + GEMM_BM - 1) / GEMM_BM) * nbatch;
  return 2 * ((tiles_m + 1) / 2) * ((ff_dim + 255) / 256);
}
extern "C" int64_t oron_ffn_workspace_bytes(int32_t rows_per_batch, int32_t nbatch, int32_t ff_dim) {
  if (rows_per_batch <= 0 || nbatch <= 0 || ff_dim <= 0) return 0;
  return 64 + 4ll * ffn_flag_count(rows_per_batch, nbatch, ff_dim);
}

extern "C" int oron_ffn_bf16(const oron_gemm_desc* up, const oron_gemm_desc* down, void* workspace, int64_t workspace_bytes,
                             oron_stream_t stream) {
  GemmArgs a1, a2;
  CUtensorMap ta1, tb1, ta2, tb2;
  if (int rc = gemm_prepare(up, a1, ta1, tb1)) return rc;
  if (int rc = gemm_prepare(down, a2, ta2, tb2)) return rc;
  const int taps1 = up->taps > 0 ? up->taps : 1, taps2 = down->taps > 0 ? down->taps : 1;
  if (!up->two_sm || !down->two_sm || up->block_n != 256 || down->block_n != 256 || taps1 != 1 || taps2 != 1 || a1.a_mn || a1.b_mn ||
      a2.a_mn || a2.b_mn)
    return fail(ORON_ERR_UNSUPPORTED, "ffn: both GEMMs must be plain K-major 2-SM GEMMs with block_n 256");
  if (up->epilogue != EPI_BF16 || down->epilogue != EPI_GATE_RESID)
    return fail(ORON_ERR_UNSUPPORTED, "ffn: epilogues must be BF16 (up) and GATE_RESID (down)");
  if (up->act != ACT_GELU_TANH && up->act != ACT_NONE) return fail(ORON_ERR_UNSUPPORTED, "ffn: activation must be GELU_TANH or NONE");
  if (down->A != up->out || down->lda != up->ldo || up->N != down->w_cols || up->N % 256 != 0 || up->w_cols % GEMM_BK != 0 ||
      up->rows_per_batch != down->rows_per_batch || up->nbatch != down->nbatch)
    return fail(ORON_ERR_BAD_ARG, "ffn: the down-projection must read the up-projection's output (same rows; ff_dim %% 256 == 0)");
  if (down->N % 4 != 0) return fail(ORON_ERR_BAD_ARG, "ffn: N %% 4 == 0");
  const int nflags = ffn_flag_count(up->rows_per_batch, up->nbatch, up->N);
  if (!workspace || workspace_bytes < 64 + 4ll * nflags || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return fail(ORON_ERR_BAD_ARG, "ffn: workspace of oron_ffn_workspace_bytes() bytes (zero-initialised, 16-byte aligned) required");
  a2.stream_k = 1;
  FfnSync sy;
  sy.done = reinterpret_cast<int*>(workspace);
  sy.flags = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + 64);
  sy.nflags = nflags;
  { const char* e = getenv("ORON_FFN_MODE"); sy.mode = e ? atoi(e) : 0; }
  {  // k interleave of the down-projection shares: stride ~ num_kb / golden ratio, coprime with num_kb, and its inverse
    const int n = a2.num_kb;
    auto gcd = [](int x, int y) { while (y) { int t = x % y; x = y; y = t; } return x; };
    int ks = int(n * 0.6180339887 + 0.5);
    if (ks < 1) ks = 1;
    while (gcd(ks, n) != 1) ++ks;
    ks %= n;
    if (n == 1) ks = 0;
    int inv = 0;
    for (int i = 0; i < n; ++i) if ((long long)i * ks % n == 1 % n) { inv = i; break; }
    const char* e = getenv("ORON_FFN_KSTRIDE");  // 1: contiguous k ranges (the first version)
    if (e && atoi(e) == 1) { ks = 1 % n; inv = 1 % n; }
    sy.kstride = ks;
    sy.kstride_inv = inv;
  }

  using Cfg = Gemm2Cfg<256>;
  const int tiles_m = ((a1.rows_per_batch + GEMM_BM - 1) / GEMM_BM) * a1.nbatch;
  const long long units = (long long)((tiles_m + 1) / 2) * ((a1.N / 256) * (long long)a1.num_kb + ((a2.N + 255) / 256) * (long long)a2.num_kb);
  long long pairs = units;
  const int cap = (up->max_ctas > 0 ? up->max_ctas : num_sms()) / 2;  // co-residency: never more CTAs than SMs
  if (pairs > cap) pairs = cap;
  if (pairs > num_sms() / 2) pairs = num_sms() / 2;
  if (pairs <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t le;
  if (up->act == ACT_GELU_TANH) {
    auto kern = ffn2_bf16_tcgen05_kernel<ACT_GELU_TANH>;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
      if (e != cudaSuccess) return fail(int(e), "ffn smem attribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    le = launch_pdl(kern, dim3(unsigned(2 * pairs)), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ta1, tb1, ta2, tb2, a1, a2, sy);
  } else {
    auto kern = ffn2_bf16_tcgen05_kernel<ACT_NONE>;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
      if (e != cudaSuccess) return fail(int(e), "ffn smem attribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    le = launch_pdl(kern, dim3(unsigned(2 * pairs)), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, ta1, tb1, ta2, tb2, a1, a2, sy);
  }
  if (le != cudaSuccess) return fail(int(le), "ffn launch: %s", cudaGetErrorString(le));
  return check_launch("ffn2_bf16_tcgen05");
}

// ---------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------
static long long* g_attn_dbg = nullptr;
// profiling aid (tools/attn_trace.py): int64 [n_ctas, 16] device buffer receiving per-CTA clock64 stamps, or NULL
extern "C" void oron_debug_set_attention_stamps(void* buf) { g_attn_dbg = reinterpret_cast<long long*>(buf); }
// Work decomposition: items = (batch, head, 128-query tile); two CTAs fit an SM. One CTA per item leaves a partially
// filled last wave (config 2: 352 items on 296 slots -> 2 waves for 1.19 waves of work) and pays the ~4 us prologue
// (TMEM allocation, first Q/K/V loads) once per item. The balanced schedule launches exactly the resident slots and
// gives each CTA an equal share of the flat (item, key tile) list (attn_tcgen05.cuh): it needs a workspace holding
// the plan (oron_attention_plan, once per set of sequence lengths) and the partial results of the split items.
static int g_attn_schedule = -1;
extern "C" void oron_debug_set_attention_schedule(int32_t mode) { g_attn_schedule = mode; }
static int attn_slots() { return 2 * num_sms(); }

struct AttnWsLayout {
  int grid, seg_stride;
  int64_t off_nseg, off_segs, off_merge, off_cnt, off_ml, off_o, bytes;
};
static AttnWsLayout attn_ws_layout(int nbatch, int rows_per_batch, int heads) {
  AttnWsLayout w;
  w.grid = attn_slots();
  const int64_t q_tiles = (rows_per_batch + ATT_TILE - 1) / ATT_TILE;
  const int64_t total_max = int64_t(nbatch) * heads * q_tiles * q_tiles;
  w.seg_stride = int((total_max + w.grid - 1) / w.grid) + 2;  // whole items in a share <= its units, + the two partial ends
  auto up = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  w.off_nseg = up(sizeof(AttnPlanHeader));
  w.off_segs = up(w.off_nseg + int64_t(w.grid) * 4);
  w.off_merge = up(w.off_segs + int64_t(w.grid) * w.seg_stride * int64_t(sizeof(AttnSeg)));
  w.off_cnt = up(w.off_merge + int64_t(w.grid) * int64_t(sizeof(AttnMergeEnt)));
  w.off_ml = up(w.off_cnt + int64_t(w.grid) * 4);
  w.off_o = up(w.off_ml + int64_t(2 * w.grid) * ATT_TILE * 2 * 4);
  w.bytes = up(w.off_o + int64_t(2 * w.grid) * ATT_TILE * ATT_D * 2);
  return w;
}

// ---- fourth-generation kernel (attn_fwd4.cuh): the default. ORON_ATT_VERSION=3 / oron_debug_set_attention_version(3)
// selects the round-1 kernel (kept for A/B measurements).
static int g_attn_version = 0;
static int attn_version() {
  if (g_attn_version == 0) { const char* e = getenv("ORON_ATT_VERSION"); g_attn_version = (e && atoi(e) == 3) ? 3 : 4; }
  return g_attn_version;
}
extern "C" void oron_debug_set_attention_version(int32_t v) { g_attn_version = (v == 3) ? 3 : 4; }

// Split items are merged inside the attention launch by its two combine warps, every CTA that holds a part of an item
// combining its share of the item's rows once all parts have arrived (default: 30.2 us per call at config 2), or, with
// ORON_ATT_COMBINE=kernel, by attn4_combine_kernel right behind the attention launch (33.1 us: a second launch on the
// critical path between the attention kernel and the out-projection). The first in-kernel version, where the last part
// to arrive merged the whole item alone, cost 38.1 us.
static bool attn_inline_combine() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ORON_ATT_COMBINE"); v = (e && e[0] == 'k') ? 0 : 1; }
  return v == 1;
}
// ORON_ATT_SKEW: share of a CTA's work moved from the second CTA of every SM to the first (attn4_plan_kernel). Default 0:
// measured at config 2, per call: 0 -> 33.0 us, 0.05 -> 33.4, 0.09 -> 34.5, 0.13 -> 37.0 (the asymmetry between the two CTAs
// of an SM is real -- 17.9 vs 21.5 us with equal work -- but is not a property of the block index that a static plan can use)
static float attn_skew() {
  static float v = -1.f;
  if (v < 0.f) { const char* e = getenv("ORON_ATT_SKEW"); v = e ? float(atof(e)) : 0.0f; if (v < 0.f) v = 0.f; }
  return v;
}
struct Attn4WsLayout {
  int grid, seg_stride;
  int64_t off_nseg, off_segs, off_merge, off_cnt, off_ml, off_o, bytes;
};
static Attn4WsLayout attn4_ws_layout(int nbatch, int rows_per_batch, int heads) {
  Attn4WsLayout w;
  w.grid = attn_slots();
  const int64_t q_tiles = (rows_per_batch + ATT4_TILE - 1) / ATT4_TILE;
  const int64_t total_max = int64_t(nbatch) * heads * q_tiles * q_tiles;
  w.seg_stride = int((total_max + w.grid - 1) / w.grid) + 2;  // whole items in a share <= its units, + the two partial ends
  auto up = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  w.off_nseg = up(sizeof(Attn4PlanHeader));
  w.off_segs = up(w.off_nseg + int64_t(w.grid) * 4);
  w.off_merge = up(w.off_segs + int64_t(w.grid) * w.seg_stride * int64_t(sizeof(Attn4Seg)));
  w.off_cnt = up(w.off_merge + int64_t(w.grid) * int64_t(sizeof(Attn4Merge)));
  w.off_ml = up(w.off_cnt + int64_t(w.grid) * 8);
  w.off_o = up(w.off_ml + int64_t(2 * w.grid) * ATT4_TILE * 2 * 4);
  w.bytes = up(w.off_o + int64_t(2 * w.grid) * ATT4_TILE * ATT4_D * 2);
  return w;
}

extern "C" int64_t oron_attention_workspace_bytes(int32_t nbatch, int32_t rows_per_batch, int32_t heads) {
  if (nbatch <= 0 || rows_per_batch <= 0 || heads <= 0) return 0;
  const int64_t a = attn_ws_layout(nbatch, rows_per_batch, heads).bytes, b = attn4_ws_layout(nbatch, rows_per_batch, heads).bytes;
  return a > b ? a : b;
}

extern "C" int oron_attention_plan(const int32_t* seq_lens, int32_t nbatch, int32_t rows_per_batch, int32_t heads,
                                   void* workspace, int64_t workspace_bytes, oron_stream_t stream) {
  if (!workspace || nbatch <= 0 || rows_per_batch <= 0 || heads <= 0) return fail(ORON_ERR_BAD_ARG, "attention_plan: bad argument");
  if (workspace_bytes < oron_attention_workspace_bytes(nbatch, rows_per_batch, heads) || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return fail(ORON_ERR_BAD_ARG, "attention_plan: workspace too small or not 16-byte aligned");
  char* p = reinterpret_cast<char*>(workspace);
  if (attn_version() == 4) {
    const Attn4WsLayout w = attn4_ws_layout(nbatch, rows_per_batch, heads);
    attn4_plan_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<Attn4PlanHeader*>(p), reinterpret_cast<int*>(p + w.off_nseg), reinterpret_cast<Attn4Seg*>(p + w.off_segs),
        reinterpret_cast<Attn4Merge*>(p + w.off_merge), reinterpret_cast<int*>(p + w.off_cnt), seq_lens, nbatch, rows_per_batch, heads, w.grid, w.seg_stride, g_attn_schedule == 1 ? 1 : 0, attn_skew());
    return check_launch("attn4_plan");
  }
  const AttnWsLayout w = attn_ws_layout(nbatch, rows_per_batch, heads);
  attn_plan_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<AttnPlanHeader*>(p), reinterpret_cast<int*>(p + w.off_nseg), reinterpret_cast<AttnSeg*>(p + w.off_segs),
      reinterpret_cast<AttnMergeEnt*>(p + w.off_merge), reinterpret_cast<int*>(p + w.off_cnt), seq_lens, nbatch, rows_per_batch,
      heads, w.grid, w.seg_stride);
  return check_launch("attn_plan");
}

static int attention4_impl(const CUtensorMap& tq, void* out, int64_t ldo, int32_t nbatch, int32_t rows_per_batch, int32_t heads,
                           const int32_t* seq_lens, float scale, void* workspace, int64_t workspace_bytes, float* lse,
                           cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES);
    if (e != cudaSuccess) return fail(int(e), "attention smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const Attn4WsLayout w = attn4_ws_layout(nbatch, rows_per_batch, heads);
  const bool have_ws = workspace != nullptr && workspace_bytes >= w.bytes && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0;
  Attn4Args a;
  memset(&a, 0, sizeof(a));
  a.rows_per_batch = rows_per_batch;
  a.nbatch = nbatch;
  a.heads = heads;
  a.seq_lens = seq_lens;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.dbg = g_attn_dbg;
  a.lse = lse;
  a.q_tiles = (rows_per_batch + ATT4_TILE - 1) / ATT4_TILE;
  const long long items = (long long)a.q_tiles * heads * nbatch;
  // with a planned workspace: at most two CTAs per SM, equal shares of the key-tile list; g_attn_schedule == 0 forces
  // one CTA per item
  const bool planned = have_ws && g_attn_schedule != 0;
  if (planned) {
    char* p = reinterpret_cast<char*>(workspace);
    a.plan_hdr = reinterpret_cast<const Attn4PlanHeader*>(p);
    a.plan_nseg = reinterpret_cast<const int*>(p + w.off_nseg);
    a.plan_segs = reinterpret_cast<const Attn4Seg*>(p + w.off_segs);
    a.plan_merge = reinterpret_cast<const Attn4Merge*>(p + w.off_merge);
    a.ws_cnt = attn_inline_combine() ? reinterpret_cast<int*>(p + w.off_cnt) : nullptr;
    a.ws_ml = reinterpret_cast<float*>(p + w.off_ml);
    a.ws_o = reinterpret_cast<__half*>(p + w.off_o);
  }
  dim3 grid(planned ? unsigned(w.grid) : unsigned(items));
  static int abl = -1;  // ORON_ATT_ABL: ablation variants for tools/kernel_bench.py (wrong results, timing only)
  if (abl < 0) { const char* e = getenv("ORON_ATT_ABL"); abl = e ? atoi(e) : 0; }
  cudaError_t le;
  if (abl != 0) {
    auto set = [&](auto k) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES); return launch_pdl(k, grid, dim3(ATT4_THREADS), ATT4_SMEM_BYTES, st, tq, a); };
    switch (abl) {
      case 128: le = set(attn_fwd4_kernel<false, 128>); break;
      case 256: le = set(attn_fwd4_kernel<false, 256>); break;
      case 384: le = set(attn_fwd4_kernel<false, 384>); break;
      case 7: le = set(attn_fwd4_kernel<false, 7>); break;
      case 15: le = set(attn_fwd4_kernel<false, 15>); break;
      case 23: le = set(attn_fwd4_kernel<false, 23>); break;
      case 39: le = set(attn_fwd4_kernel<false, 39>); break;
      case 71: le = set(attn_fwd4_kernel<false, 71>); break;
      case 55: le = set(attn_fwd4_kernel<false, 55>); break;
      default: le = a.dbg ? set(attn_fwd4_kernel<true, 127>) : set(attn_fwd4_kernel<false, 127>); break;
    }
  } else {
  // ORON_ATT_TRACE=lite: the production instantiation with start / end wall-clock stamps only
  static int lite = -1;
  if (lite < 0) { const char* e = getenv("ORON_ATT_TRACE"); lite = (e && e[0] == 'l') ? 1 : 0; }
  le = (a.dbg && !lite) ? launch_pdl(attn_fwd4_kernel<true>, grid, dim3(ATT4_THREADS), ATT4_SMEM_BYTES, st, tq, a)
             : launch_pdl(attn_fwd4_kernel<false>, grid, dim3(ATT4_THREADS), ATT4_SMEM_BYTES, st, tq, a);
  }
  if (le != cudaSuccess) return fail(int(le), "attention launch: %s", cudaGetErrorString(le));
  int rc = check_launch("attn_fwd4");
  if (rc) return rc;
  // items are split only when there are more of them than CTAs (or when the test aid forces shares): then the parts are
  // combined by a second, small launch right behind the first
  if (planned && a.ws_cnt == nullptr && (items > w.grid || g_attn_schedule == 1)) {
    le = launch_pdl(attn4_combine_kernel, dim3(unsigned(4 * w.grid)), dim3(256), 0, st, a.plan_merge, (const __half*)a.ws_o, (const float*)a.ws_ml,
                    a.out, (long long)ldo, lse, (int)rows_per_batch, (int)heads);
    if (le != cudaSuccess) return fail(int(le), "attention combine launch: %s", cudaGetErrorString(le));
    rc = check_launch("attn4_combine");
  }
  return rc;
}

static int attention_impl(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int32_t nbatch, int32_t rows_per_batch,
                          int32_t heads, const int32_t* seq_lens, float scale, void* workspace, int64_t workspace_bytes,
                          float* lse, oron_stream_t stream) {
  if (!qkv || !out || nbatch <= 0 || rows_per_batch <= 0 || heads <= 0)
    return fail(ORON_ERR_BAD_ARG, "attention: bad argument");
  if (ldo % 8 != 0) return fail(ORON_ERR_BAD_ARG, "attention: ldo must be a multiple of 8");
  CUtensorMap tq;
  int rc = make_tmap_bf16(&tq, qkv, uint64_t(3 * heads * ATT_D), uint64_t(rows_per_batch), uint64_t(nbatch),
                          uint64_t(ld_qkv), uint64_t(ld_qkv) * uint64_t(rows_per_batch), ATT_TILE, 3);
  if (rc) return rc;
  if (attn_version() == 4)
    return attention4_impl(tq, out, ldo, nbatch, rows_per_batch, heads, seq_lens, scale, workspace, workspace_bytes, lse,
                           reinterpret_cast<cudaStream_t>(stream));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ATT_SMEM_BYTES);
    if (e != cudaSuccess) return fail(int(e), "attention smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const AttnWsLayout w = attn_ws_layout(nbatch, rows_per_batch, heads);
  const bool have_ws = workspace != nullptr && workspace_bytes >= w.bytes && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0;
  static int env_mode = -2;
  if (env_mode == -2) { const char* e = getenv("ORON_ATT_BALANCED"); env_mode = e ? atoi(e) : -1; }
  const int mode = g_attn_schedule >= 0 ? g_attn_schedule : env_mode;
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.rows_per_batch = rows_per_batch;
  a.nbatch = nbatch;
  a.heads = heads;
  a.seq_lens = seq_lens;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.dbg = g_attn_dbg;
  a.lse = lse;
  a.q_tiles = (rows_per_batch + ATT_TILE - 1) / ATT_TILE;
  const long long items = (long long)a.q_tiles * heads * nbatch;
  // the balanced schedule only pays when there are more items than resident CTA slots
  const bool balanced = have_ws && (mode == 1 || (mode != 0 && items > attn_slots()));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (balanced) {
    char* p = reinterpret_cast<char*>(workspace);
    a.plan_hdr = reinterpret_cast<const AttnPlanHeader*>(p);
    a.plan_nseg = reinterpret_cast<const int*>(p + w.off_nseg);
    a.plan_segs = reinterpret_cast<const AttnSeg*>(p + w.off_segs);
    a.plan_merge = reinterpret_cast<const AttnMergeEnt*>(p + w.off_merge);
    a.ws_cnt = reinterpret_cast<int*>(p + w.off_cnt);
    a.ws_ml = reinterpret_cast<float*>(p + w.off_ml);
    a.ws_o = reinterpret_cast<__half*>(p + w.off_o);
  }
  dim3 grid(balanced ? unsigned(w.grid) : unsigned(items));
  cudaError_t le = launch_pdl(attn_fwd_tcgen05_kernel, grid, dim3(ATT_THREADS), ATT_SMEM_BYTES, st, tq, a);
  if (le != cudaSuccess) return fail(int(le), "attention launch: %s", cudaGetErrorString(le));
  return check_launch("attn_fwd_tcgen05");
}

extern "C" int oron_attention_bf16(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int32_t nbatch,
                                   int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale,
                                   void* workspace, int64_t workspace_bytes, oron_stream_t stream) {
  return attention_impl(qkv, ld_qkv, out, ldo, nbatch, rows_per_batch, heads, seq_lens, scale, workspace, workspace_bytes, nullptr,
                        stream);
}
// Training forward (include/oron_b200_train.h): one CTA per item, and the log-sum-exp of every row is kept.
extern "C" int oron_attention_fwd_lse(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int32_t nbatch,
                                      int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale, float* lse,
                                      void* workspace, int64_t workspace_bytes, oron_stream_t stream) {
  if (!lse) return fail(ORON_ERR_BAD_ARG, "attention_fwd_lse: lse is NULL");
  // the round-1 kernel writes lse for whole items only: planned launches need the current one
  const bool ws_ok = attn_version() == 4;
  return attention_impl(qkv, ld_qkv, out, ldo, nbatch, rows_per_batch, heads, seq_lens, scale, ws_ok ? workspace : nullptr,
                        ws_ok ? workspace_bytes : 0, lse, stream);
}

// ---------------------------------------------------------------------------
// row-wise kernels
// ---------------------------------------------------------------------------
extern "C" int oron_ln_modulate(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                                float eps, const float* scale, const float* shift, int64_t mod_ld, int32_t mod_nb,
                                int64_t step_stride, const int32_t* step_ptr, int32_t add_one, void* out_bf16,
                                float* out_f32, int64_t ldo, oron_stream_t stream) {
  if (!x || !scale || (!out_bf16 && !out_f32)) return fail(ORON_ERR_BAD_ARG, "ln_modulate: null pointer");
  if (ldx % 4 != 0 || ldo % 4 != 0) return fail(ORON_ERR_BAD_ARG, "ln_modulate: ld must be a multiple of 4");
  LnArgs a;
  a.x = x; a.ldx = ldx; a.rows_per_batch = rows_per_batch; a.nbatch = nbatch; a.C = C; a.eps = eps;
  a.scale = scale; a.shift = shift; a.mod_ld = mod_ld; a.mod_nb = mod_nb > 0 ? mod_nb : 1;
  a.step_stride = step_stride; a.step_ptr = step_ptr; a.add_one = add_one;
  a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.out_f32 = out_f32; a.ldo = ldo;
  const long long rows = (long long)rows_per_batch * nbatch;
  const int blocks = int((rows + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 1024) launch_pdl(ln_modulate_kernel<1024>, dim3(blocks), dim3(256), 0, st, a);
  else if (C == 512) launch_pdl(ln_modulate_kernel<512>, dim3(blocks), dim3(256), 0, st, a);
  else if (C == 256) launch_pdl(ln_modulate_kernel<256>, dim3(blocks), dim3(256), 0, st, a);
  else if (C == 128) launch_pdl(ln_modulate_kernel<128>, dim3(blocks), dim3(256), 0, st, a);
  else if (C == 768) launch_pdl(ln_modulate_kernel<768>, dim3(blocks), dim3(256), 0, st, a);
  else if (C == 64) launch_pdl(ln_modulate64_kernel, dim3(blocks), dim3(256), 0, st, a);
  else return fail(ORON_ERR_UNSUPPORTED, "ln_modulate: C=%d not supported (64/128/256/512/768/1024)", C);
  return check_launch("ln_modulate");
}

extern "C" int oron_cfg_euler_step(float* x, const float* v, int64_t ldv, int32_t nb, int32_t rows_per_batch,
                                   int32_t n_mels, int32_t has_uncond, float cfg, const float* dt, int32_t* step_ptr,
                                   void* xb_bf16, int64_t ldxb, float* traj, float* v_out, int32_t method,
                                   oron_stream_t stream) {
  if (!x || !v || !dt || !step_ptr || !xb_bf16) return fail(ORON_ERR_BAD_ARG, "cfg_euler_step: null pointer");
  if (method != 0 && method != 1) return fail(ORON_ERR_BAD_ARG, "cfg_euler_step: method must be 0 (Euler) or 1 (midpoint)");
  EulerArgs a;
  a.x = x; a.v = v; a.ldv = ldv; a.nb = nb; a.rows_per_batch = rows_per_batch; a.n_mels = n_mels;
  a.has_uncond = has_uncond; a.cfg = cfg; a.dt = dt; a.step_ptr = step_ptr;
  a.xb = reinterpret_cast<__nv_bfloat16*>(xb_bf16); a.ldxb = ldxb; a.traj = traj; a.v_out = v_out; a.midpoint = method;
  const long long total = (long long)nb * rows_per_batch * n_mels;
  int blocks = int((total + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(cfg_euler_kernel, dim3(blocks), dim3(256), 0, st, a);
  int rc = check_launch("cfg_euler");
  if (rc) return rc;
  launch_pdl(step_advance_kernel, dim3(1), dim3(1), 0, st, step_ptr);
  return check_launch("step_advance");
}

extern "C" int oron_cast_rows_bf16(const float* x, int64_t ldx, int64_t rows, int32_t C, void* out_bf16, int64_t ldo,
                                   int32_t reps, oron_stream_t stream) {
  if (!x || !out_bf16 || rows <= 0 || C <= 0) return fail(ORON_ERR_BAD_ARG, "cast_rows_bf16: bad argument");
  const long long total = rows * C;
  int blocks = int((total + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  cast_rows_bf16_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ldx, rows, C, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, reps > 0 ? reps : 1);
  return check_launch("cast_rows_bf16");
}

extern "C" int oron_time_sinusoid(const float* t, int32_t n, void* out_bf16, int64_t ldo, oron_stream_t stream) {
  if (!t || !out_bf16 || n <= 0) return fail(ORON_ERR_BAD_ARG, "time_sinusoid: bad argument");
  time_sinusoid_kernel<<<n, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      t, n, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo);
  return check_launch("time_sinusoid");
}

extern "C" int oron_text_embed_front(const int32_t* ids, const uint8_t* drop, const float* table,
                                     const float* pos_table, int32_t rows_per_batch, int32_t nb, int32_t C, float* x,
                                     int64_t ldx, uint8_t* row_valid, oron_stream_t stream) {
  if (!ids || !drop || !table || !pos_table || !x || !row_valid)
    return fail(ORON_ERR_BAD_ARG, "text_embed_front: null pointer");
  const long long rows = (long long)rows_per_batch * nb;
  text_embed_front_kernel<<<unsigned(rows), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids, drop, table, pos_table, rows_per_batch, nb, C, x, ldx, row_valid);
  return check_launch("text_embed_front");
}

extern "C" int oron_dwconv7_ln(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                               const int32_t* seq_lens, const float* w, const float* wb, const float* ln_w,
                               const float* ln_b, float eps, void* out_bf16, int64_t ldo, oron_stream_t stream) {
  if (!x || !w || !wb || !ln_w || !ln_b || !out_bf16) return fail(ORON_ERR_BAD_ARG, "dwconv7_ln: null pointer");
  DwLnArgs a;
  a.x = x; a.ldx = ldx; a.rows_per_batch = rows_per_batch; a.nbatch = nbatch; a.seq_lens = seq_lens;
  a.w = w; a.wb = wb; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = eps;
  a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.ldo = ldo;
  const long long rows = (long long)rows_per_batch * nbatch;
  const int blocks = int((rows + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(wb) |
                         reinterpret_cast<uintptr_t>(ln_w) | reinterpret_cast<uintptr_t>(ln_b)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0 && (ldx & 3) == 0 && (ldo & 3) == 0;
  if (aligned && (C == 512 || C == 256)) {
    // runs of consecutive frames per work item: long enough to amortise the 6 halo rows, short enough that the
    // persistent grid (4 resident CTAs per SM) is balanced
    static int occ512 = 0, occ256 = 0;
    if (!occ512) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ512, dwconv7_ln_run_kernel<512>, 128, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ256, dwconv7_ln_run_kernel<256>, 64, 0);
      if (occ512 < 1) occ512 = 1;
      if (occ256 < 1) occ256 = 1;
    }
    const long long slots = (long long)(C == 512 ? occ512 : occ256) * num_sms();
    long long run = (rows + slots - 1) / slots;
    run = run < 16 ? 16 : (run > 64 ? 64 : run);
    run = (run + 1) & ~1LL;
    const long long items = ((rows_per_batch + run - 1) / run) * (long long)nbatch;
    dim3 grid((unsigned)(items < slots ? items : slots));
    if (C == 512) dwconv7_ln_run_kernel<512><<<grid, 128, 0, st>>>(a, int(run));
    else dwconv7_ln_run_kernel<256><<<grid, 64, 0, st>>>(a, int(run));
    return check_launch("dwconv7_ln");
  }
  if (C == 512) dwconv7_ln_kernel<512><<<blocks, 256, 0, st>>>(a);
  else if (C == 256) dwconv7_ln_kernel<256><<<blocks, 256, 0, st>>>(a);
  else if (C == 128) dwconv7_ln_kernel<128><<<blocks, 256, 0, st>>>(a);
  else if (C == 64) dwconv7_ln_kernel<64><<<blocks, 256, 0, st>>>(a);
  else if (C == 32) dwconv7_ln_kernel<32><<<blocks, 256, 0, st>>>(a);
  else return fail(ORON_ERR_UNSUPPORTED, "dwconv7_ln: C=%d not supported", C);
  return check_launch("dwconv7_ln");
}

extern "C" int oron_grn(void* h_bf16, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t C,
                        const int32_t* seq_lens, const float* gamma, const float* beta, float* gx2,
                        oron_stream_t stream) {
  if (!h_bf16 || !gamma || !beta || !gx2) return fail(ORON_ERR_BAD_ARG, "grn: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rpb = 32;
  dim3 grid((rows_per_batch + rpb - 1) / rpb, nb);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(h_bf16);
  grn_sumsq_kernel<<<dim3((C + 31) / 32, nb), 256, 0, st>>>(h, ldh, rows_per_batch, nb, seq_lens, C, gx2);
  int rc = check_launch("grn_sumsq");
  if (rc) return rc;
  grn_apply_kernel<<<grid, 256, 0, st>>>(h, ldh, rows_per_batch, nb, C, rpb, gx2, gamma, beta);
  return check_launch("grn_apply");
}

extern "C" int oron_abi_version(void) { return ORON_ABI_VERSION; }
extern "C" const char* oron_last_error(void) { return g_err; }
extern "C" uint64_t oron_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
