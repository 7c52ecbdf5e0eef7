// C ABI of the smplk library (see include/smplk.h): model packing, forward / backward
// orchestration, error handling.  Host code only glues kernels together; all math on the hot
// path runs in the CUDA kernels of pose_kernels.cuh, blend_gemm.cuh, skinning.cuh, backward.cuh.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/smplk.h"
#include "backward.cuh"
#include "blend_gemm.cuh"
#include "blend_gemm_2cta.cuh"
#include "blend_skin_fused.cuh"
#include "lbs_replay_gemm.cuh"
#include "common.cuh"
#include "pose_kernels.cuh"
#include "skinning.cuh"
#include "skinning8.cuh"
#include "skin_fit.cuh"
#include "picks.cuh"
#include "mesh_ops.cuh"
#include "fit_loss.cuh"

using namespace smplk;

// ------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail((int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                             \
  } while (0)

#define LAUNCH_CHECK(name)                                                               \
  do {                                                                                   \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess)                                                               \
      return fail((int)e_, "launch of %s failed: %s", name, cudaGetErrorString(e_));     \
  } while (0)

// Kernel launch with (pdl = true) or without the programmatic-stream-serialization attribute: see ptx::pdl_wait.
// Only kernels that execute `pdl_wait` before their first access to mutable global memory may be launched with it.
// Errors are picked up by the LAUNCH_CHECK that follows (cudaGetLastError).
template <typename... KArgs, typename... Args>
static void launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
  if (e != cudaSuccess && pdl && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
    // a driver / capture mode without programmatic launch: same kernel with plain stream ordering (its
    // griddepcontrol instructions are no-ops then)
    (void)cudaGetLastError();
    cfg.numAttrs = 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
  }
}

// Every entry point works on its model's (or the named) device and leaves the caller's current device
// as it found it (a single process may drive several GPUs).
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      changed = err == cudaSuccess;
    }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
#define DEVICE_GUARD(dev)                                                                \
  DeviceGuard device_guard_(dev);                                                        \
  if (device_guard_.err != cudaSuccess)                                                  \
    return fail((int)device_guard_.err, "cudaSetDevice(%d) failed: %s", (int)(dev),      \
                cudaGetErrorString(device_guard_.err))

extern "C" const char* smplk_last_error_string(void) { return g_err; }
extern "C" int smplk_version(void) { return SMPLK_VERSION; }
extern "C" uint64_t smplk_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum BlendPath : int { BLEND_SIMT = 0, BLEND_TF32 = 1, BLEND_F16 = 2 };
struct ProfRec {
  cudaEvent_t e0, e1;
  int slot;
};

struct smplk_model {
  ModelDev d;
  int device;
  int num_sms;
  int cc_major;
  std::vector<void*> allocs;
  EncodeTiledFn encode;
  bool has_tma;
  CUtensorMap tmap_pd_hi, tmap_pd_lo;      // forward: B operand rows = vertex coords
  CUtensorMap tmap_pdh_hi, tmap_pdh_lo;    // forward, fp16-split operand
  // CTA-pair kernel: 128-byte rows, 128-row boxes (each CTA loads half of the B tile)
  CUtensorMap tmap2_pd_hi, tmap2_pd_lo, tmap2_pdh_hi, tmap2_pdh_lo, tmap2_pdkn_hi, tmap2_pdkn_lo;
  bool use_2cta;
  CUtensorMap tmapf_pdh_hi, tmapf_pdh_lo;  // fused blend+skinning kernel: 84-vertex column tiles
  // rigged-mesh replay as a GEMM (lbs_replay_gemm.cuh): P = w (x) [v_template; 1], fp16 two-term split
  CUtensorMap tmap_rp_hi, tmap_rp_lo;
  const __half* rp_hi; const __half* rp_lo;   // [V][rp_kp]
  int rp_kp;                                  // round_up(4 J, 32)
  float rp_scale;                             // power of two applied to P
  // skinning pass of the two-kernel forward with the transform blend on the tensor cores (kSkin instance of the
  // replay kernel): W[v][j] as fp16 two-term split
  CUtensorMap tmap_sw_hi, tmap_sw_lo;
  const __half* sw_hi; const __half* sw_lo;   // [V][sw_kp]
  int sw_kp;                                  // round_up(J, 32)
  float sw_scale;
  bool skin_gemm_ok;
  bool use_skin_gemm;                         // option skin_gemm = 1 (default 0: measured 0.164 + 0.015 ms operand pass against
                                              // 0.170 ms of the streaming kernel at 4,096 bodies, profiles/r02_skin_gemm_ab.txt)
  bool replay_gemm_ok;                        // LBS-only handle on sm_100 with the operand built
  bool use_replay_gemm;                       // option replay_gemm = 0 selects the streaming skinning kernel
  bool fused_tma_out;   // option fused_tma_out = 0: the fused kernel stores its result per lane instead of through TMA
  bool use_fused;       // option fused = 0 selects the two-kernel forward (cross-checks, stand-alone kernel timings)
  bool use_pose_block;  // option pose_block = 0 selects the warp-per-body pose kernel + transposition pass
  bool skip_pose;       // option skip_pose = 1 (measurement aid): smplk_forward launches no pose kernel and reuses the workspace rows
                        // (features, transforms) of the previous call -> back-to-back launches of the blend / skinning kernel alone
  bool use_pdl;         // option pdl = 0: every kernel of a call is launched with plain stream ordering (no programmatic dependent launch)
  CUtensorMap tmap_pdkn_hi, tmap_pdkn_lo;  // backward: B operand rows = blend features
  CUtensorMap tmap_pdknh_hi, tmap_pdknh_lo, tmap2_pdknh_hi, tmap2_pdknh_lo;   // same, fp16 two-term split
  CUtensorMap tmap_pdknb_hi, tmap_pdknb_lo, tmap2_pdknb_hi, tmap2_pdknb_lo;   // same, bf16 two-term split
  bool bwd_f16;         // backward GEMM on fp16-split operands (option backward_tf32 keeps 3xTF32)
  // host staging for smplk_forward_host
  void* stage_dev;
  size_t stage_bytes;
  cudaStream_t copy_stream;                 // device->host copies of smplk_forward_host
  std::vector<cudaEvent_t> chunk_events;    // one per chunk of that call
  // optional per-kernel device timing (smplk_profile_*)
  BlendPath default_tc;  // BLEND_F16 unless the option blend_tf32 is set
  int skin_bpb;         // SMPLK_SKIN_BPB: override bodies per block (tuning)
  bool skin_g8;         // 8 vertices per thread (default; SMPLK_SKIN_G8=0 selects the 4-vertex kernel)
  bool skin_tma;        // SMPLK_SKIN_TMA=1: per-warp cp.async.bulk pipeline (measured slower: 0.210 vs 0.182 ms)
  bool force_skin_v1;   // SMPLK_SKIN_V1=1 in the environment: per-vertex gather kernel (A/B testing)
  bool sparse_picks;    // option sparse_picks = 0: keypoint-only gradients take the dense backward (cross-check)
  bool fit_fused;       // fused skinning + loss + skinning-backward kernel of smplk_fit_vertex_l2 (option fit_fused = 0: off)
  bool da_v1;           // SMPLK_DA_V1=1 (A/B builds): the fitting step's dA through dA_kernel instead of dA_seg_kernel
  mutable bool prof_on;
  mutable std::vector<ProfRec> prof_pending;
  mutable double prof_ms[SMPLK_PROF_SLOTS];
  mutable int64_t prof_n[SMPLK_PROF_SLOTS];
};

struct ProfScope {
  const smplk_model* m;
  cudaStream_t st;
  int slot;
  cudaEvent_t e0;
  ProfScope(const smplk_model* m_, cudaStream_t st_, int slot_) : m(m_), st(st_), slot(slot_), e0(nullptr) {
    if (m->prof_on) {
      cudaEventCreate(&e0);
      cudaEventRecord(e0, st);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEvent_t e1;
      cudaEventCreate(&e1);
      cudaEventRecord(e1, st);
      m->prof_pending.push_back({e0, e1, slot});
    }
  }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static float tf32_rn_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0xfffu + ((u >> 13) & 1u);
  u &= ~0x1fffu;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

template <typename T>
static int upload(smplk_model* mdl, const std::vector<T>& h, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
  CUDA_TRY(cudaMalloc(&p, bytes));
  mdl->allocs.push_back(p);
  if (!h.empty()) CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(p);
  return 0;
}

static int make_tmap_2d(const smplk_model* mdl, CUtensorMap* map, const void* ptr, uint64_t inner,
                        uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                        CUtensorMapL2promotion promo, bool f16 = false,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t pitch_elems = 0) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {(pitch_elems ? pitch_elems : inner) * (f16 ? sizeof(__half) : sizeof(float))};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = mdl->encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                           2, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SMPLK_E_DEVICE, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// GEMM operand tile: rows of kRowBytes (the smem swizzle span), `box_rows` rows per TMA box.
static int make_operand_tmap(const smplk_model* mdl, CUtensorMap* map, const void* ptr, uint64_t inner,
                             uint64_t outer, uint32_t box_rows, CUtensorMapL2promotion promo, bool f16) {
  return make_tmap_2d(mdl, map, ptr, inner, outer, kRowBytes / (f16 ? 2 : 4), box_rows, promo, f16,
                      kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// Operand tile of the CTA-pair kernel: always 128-byte rows / 128B swizzle, 128 rows per box.
static int make_operand_tmap_2cta(const smplk_model* mdl, CUtensorMap* map, const void* ptr, uint64_t inner,
                                  uint64_t outer, CUtensorMapL2promotion promo, bool f16) {
  return make_tmap_2d(mdl, map, ptr, inner, outer, 128 / (f16 ? 2 : 4), kBlendBM, promo, f16,
                      CU_TENSOR_MAP_SWIZZLE_128B);
}

// fp16 operand tile of the fused kernel: kFzRowBytes of K per row (= swizzle span), 128 rows per box.
static int make_operand_tmap_fused(const smplk_model* mdl, CUtensorMap* map, const void* ptr, uint64_t inner,
                                   uint64_t outer, CUtensorMapL2promotion promo) {
  return make_tmap_2d(mdl, map, ptr, inner, outer, kFzRowBytes / 2, kBlendBM, promo, true,
                      kFzRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

extern "C" int smplk_model_destroy(smplk_model* model) {
  if (!model) return 0;
  DeviceGuard device_guard_(model->device);
  for (void* p : model->allocs) cudaFree(p);
  if (model->stage_dev) cudaFree(model->stage_dev);
  for (cudaEvent_t e : model->chunk_events) cudaEventDestroy(e);
  if (model->copy_stream) cudaStreamDestroy(model->copy_stream);
  delete model;
  return 0;
}

static int build_model(const smplk_model_desc* desc, smplk_model* mdl) {
  ModelDev& d = mdl->d;
  const int V = desc->num_verts, J = desc->num_joints, NB = desc->num_betas;
  const bool lbs_only = desc->posedirs == nullptr;
  d.V = V; d.J = J; d.NB = lbs_only ? 0 : NB;
  d.P = 9 * (J - 1);
  d.K = lbs_only ? 0 : d.P + d.NB;
  d.Kpad = lbs_only ? 0 : round_up(d.K, kBlendBK);
  d.N = 3 * V;
  d.Npad = round_up(d.N, kBlendBN);
  d.E = desc->extra_vertex_ids ? desc->num_extra_verts : 0;
  d.R = desc->regressor_posed ? desc->num_regressors : 0;
  d.C = (desc->hand_comp_l && desc->hand_comp_r) ? desc->num_pca : 0;
  d.lbs_only = lbs_only ? 1 : 0;

  // ---- tree tables
  std::vector<int> parents(J), depth(J, 0);
  int max_depth = 0;
  for (int j = 0; j < J; ++j) {
    parents[j] = desc->parents[j];
    if (j == 0) {
      if (parents[0] >= 0) return fail(SMPLK_E_ARG, "parents[0] must be -1");
    } else {
      if (parents[j] < 0 || parents[j] >= j)
        return fail(SMPLK_E_ARG, "parents[%d]=%d must satisfy 0 <= parent < joint", j, parents[j]);
      depth[j] = depth[parents[j]] + 1;
      max_depth = std::max(max_depth, depth[j]);
    }
  }
  d.max_depth = max_depth;
  if (int r = upload(mdl, parents, &d.parents)) return r;
  if (int r = upload(mdl, depth, &d.depth)) return r;
  {
    std::vector<int> cptr(J + 1, 0), cidx(std::max(J - 1, 1), 0);
    for (int j = 1; j < J; ++j) cptr[parents[j] + 1]++;
    for (int j = 0; j < J; ++j) cptr[j + 1] += cptr[j];
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int j = 1; j < J; ++j) cidx[fill[parents[j]]++] = j;
    if (int r = upload(mdl, cptr, &d.child_ptr)) return r;
    if (int r = upload(mdl, cidx, &d.child_idx)) return r;
  }
  {
    std::vector<int> order, lstart(max_depth + 2, 0);
    for (int dd = 0; dd <= max_depth; ++dd) {
      lstart[dd] = (int)order.size();
      for (int j = 0; j < J; ++j) if (depth[j] == dd) order.push_back(j);
    }
    lstart[max_depth + 1] = (int)order.size();
    if (int r = upload(mdl, order, &d.order)) return r;
    if (int r = upload(mdl, lstart, &d.level_start)) return r;
  }

  // ---- template / bias
  std::vector<float> bias(d.Npad, 0.f);
  for (int n = 0; n < d.N; ++n) bias[n] = (float)desc->v_template[n];
  if (int r = upload(mdl, bias, &d.bias)) return r;

  // ---- rest joints: J = J_template + J_shapedirs . beta  (== J_regressor (v_template + S beta))
  std::vector<float> Jt(J * 3, 0.f), Js((size_t)J * 3 * std::max(d.NB, 1), 0.f);
  if (lbs_only) {
    if (!desc->joints_fixed) return fail(SMPLK_E_ARG, "rigged-mesh model needs joints_fixed");
    for (int i = 0; i < J * 3; ++i) Jt[i] = (float)desc->joints_fixed[i];
  } else {
    if (!desc->J_regressor || !desc->shapedirs)
      return fail(SMPLK_E_ARG, "blendshape model needs J_regressor and shapedirs");
    std::vector<double> acc(3 + 3 * NB);
    for (int j = 0; j < J; ++j) {
      std::fill(acc.begin(), acc.end(), 0.0);
      const double* row = desc->J_regressor + (size_t)j * V;
      for (int v = 0; v < V; ++v) {
        const double w = row[v];
        if (w == 0.0) continue;
        for (int a = 0; a < 3; ++a) {
          acc[a] += w * desc->v_template[3 * v + a];
          const double* sd = desc->shapedirs + ((size_t)3 * v + a) * NB;
          for (int i = 0; i < NB; ++i) acc[3 + a * NB + i] += w * sd[i];
        }
      }
      for (int a = 0; a < 3; ++a) {
        Jt[3 * j + a] = (float)acc[a];
        for (int i = 0; i < NB; ++i) Js[((size_t)3 * j + a) * NB + i] = (float)acc[3 + a * NB + i];
      }
    }
  }
  if (int r = upload(mdl, Jt, &d.J_template)) return r;
  if (int r = upload(mdl, Js, &d.J_shapedirs)) return r;

  // ---- blend operand: PD[n][k] = posedirs | shapedirs, as tf32 hi/lo (N x K) and exact (K x N)
  if (!lbs_only) {
    const size_t nk = (size_t)d.Npad * d.Kpad;
    std::vector<float> hi(nk, 0.f), lo(nk, 0.f), kn(nk, 0.f), knh(nk, 0.f), knl(nk, 0.f);
    for (int n = 0; n < d.N; ++n) {
      const double* pdrow = desc->posedirs + (size_t)n * d.P;
      const double* sdrow = desc->shapedirs + (size_t)n * NB;
      for (int k = 0; k < d.K; ++k) {
        const float x = (float)(k < d.P ? pdrow[k] : sdrow[k - d.P]);
        const float h = tf32_rn_host(x);
        hi[(size_t)n * d.Kpad + k] = h;
        lo[(size_t)n * d.Kpad + k] = x - h;
        kn[(size_t)k * d.Npad + n] = x;
        knh[(size_t)k * d.Npad + n] = h;
        knl[(size_t)k * d.Npad + n] = x - h;
      }
    }
    {
      // fp16 two-term split of posedirs * 2^e, e chosen so the largest entry lands in [2^14, 2^15)
      float amax = 0.f;
      for (int n = 0; n < d.N; ++n)
        for (int k = 0; k < d.K; ++k) amax = std::max(amax, std::fabs(kn[(size_t)k * d.Npad + n]));
      int e = 0;
      if (amax > 0.f) e = 14 - (int)std::floor(std::log2(amax));
      e = std::max(-20, std::min(e, 40));
      d.pd_scale = std::ldexp(1.0f, e);
      std::vector<__half> hh(nk, __float2half(0.f)), hl(nk, __float2half(0.f));
      for (int n = 0; n < d.N; ++n)
        for (int k = 0; k < d.K; ++k) {
          const float x = kn[(size_t)k * d.Npad + n] * d.pd_scale;
          const __half h = __float2half_rn(x);
          hh[(size_t)n * d.Kpad + k] = h;
          hl[(size_t)n * d.Kpad + k] = __float2half_rn(x - __half2float(h));
        }
      if (int r = upload(mdl, hh, &d.pd_nk_h_hi)) return r;
      if (int r = upload(mdl, hl, &d.pd_nk_h_lo)) return r;
      {   // the same split, feature-major [Kpad][Npad]: B operand of the fp16 backward GEMM
        std::vector<__half> th(nk, __float2half(0.f)), tl(nk, __float2half(0.f));
        for (int n = 0; n < d.N; ++n)
          for (int k = 0; k < d.K; ++k) {
            th[(size_t)k * d.Npad + n] = hh[(size_t)n * d.Kpad + k];
            tl[(size_t)k * d.Npad + n] = hl[(size_t)n * d.Kpad + k];
          }
        if (int r = upload(mdl, th, &d.pd_kn_h_hi)) return r;
        if (int r = upload(mdl, tl, &d.pd_kn_h_lo)) return r;
      }
      {   // bf16 two-term split of the same scaled operand: B of the backward GEMM after
          // smplk_fit_vertex_l2, whose d_v_posed rows are bf16 (tcgen05 kind::f16 faults on mixed
          // fp16 x bf16 operands)
        std::vector<uint16_t> bh(nk, 0), bl(nk, 0);
        for (int n = 0; n < d.N; ++n)
          for (int k = 0; k < d.K; ++k) {
            const float x = kn[(size_t)k * d.Npad + n] * d.pd_scale;
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
            memcpy(&bh[(size_t)k * d.Npad + n], &h, 2);
            memcpy(&bl[(size_t)k * d.Npad + n], &l, 2);
          }
        if (int r = upload(mdl, bh, &d.pd_kn_b_hi)) return r;
        if (int r = upload(mdl, bl, &d.pd_kn_b_lo)) return r;
      }
      // same operand in the fused kernel's column layout: tile t = vertices [84 t, 84 t + 84),
      // row 256 t + c <-> flat coordinate 252 t + c (c < 252), rows 256 t + 252.. = 0
      d.fz_tiles = (V + kFzTileVerts - 1) / kFzTileVerts;
      const size_t nf = (size_t)d.fz_tiles * kBlendBN;
      std::vector<__half> fh(nf * d.Kpad, __float2half(0.f)), fl(nf * d.Kpad, __float2half(0.f));
      std::vector<float> bf(nf + 64, 0.f);
      // within every 12-vertex chunk (36 columns) the columns are ordered as vertex PAIRS,
      // [x_2k x_2k+1 y_2k y_2k+1 z_2k z_2k+1], so that the epilogue's TMEM loads land as packed
      // fp32x2 operands: output column 6k + 3h + d of a chunk sits in GEMM column 6k + 2d + h
      for (int t = 0; t < d.fz_tiles; ++t)
        for (int c = 0; c < kFzTileCols; ++c) {
          const int n = t * kFzTileCols + c;
          if (n >= d.N) break;
          const int cc = c % kFzChunkCols, k6 = cc / 6, h = (cc % 6) / 3, dd = cc % 3;
          const size_t row = (size_t)t * kBlendBN + (c - cc) + 6 * k6 + 2 * dd + h;
          memcpy(&fh[row * d.Kpad], &hh[(size_t)n * d.Kpad], (size_t)d.Kpad * sizeof(__half));
          memcpy(&fl[row * d.Kpad], &hl[(size_t)n * d.Kpad], (size_t)d.Kpad * sizeof(__half));
          bf[row] = (float)desc->v_template[n];
        }
      if (int r = upload(mdl, fh, &d.pdf_h_hi)) return r;
      if (int r = upload(mdl, fl, &d.pdf_h_lo)) return r;
      if (int r = upload(mdl, bf, &d.bias_f)) return r;
    }
    if (int r = upload(mdl, hi, &d.pd_nk_hi)) return r;
    if (int r = upload(mdl, lo, &d.pd_nk_lo)) return r;
    if (int r = upload(mdl, kn, &d.pd_kn)) return r;
    if (int r = upload(mdl, knh, &d.pd_kn_hi)) return r;
    if (int r = upload(mdl, knl, &d.pd_kn_lo)) return r;
  } else {
    d.pd_nk_hi = d.pd_nk_lo = d.pd_kn = d.pd_kn_hi = d.pd_kn_lo = nullptr;
    d.pd_nk_h_hi = d.pd_nk_h_lo = nullptr;
    d.pd_kn_h_hi = d.pd_kn_h_lo = nullptr;
    d.pd_kn_b_hi = d.pd_kn_b_lo = nullptr;
    d.pdf_h_hi = d.pdf_h_lo = nullptr;
    d.bias_f = nullptr;
    d.fz_tiles = 0;
    d.pd_scale = 1.f;
  }

  // ---- LBS weights: ELL (k-major) + packed 4-wide form + CSC for the backward
  {
    int ell_k = 1;
    std::vector<std::vector<std::pair<float, int>>> rows(V);
    for (int v = 0; v < V; ++v) {
      for (int j = 0; j < J; ++j) {
        const float w = (float)desc->weights[(size_t)v * J + j];
        if (w != 0.f) rows[v].push_back({w, j});
      }
      std::sort(rows[v].begin(), rows[v].end(),
                [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
                  return a.first > b.first || (a.first == b.first && a.second < b.second);
                });
      ell_k = std::max<int>(ell_k, (int)rows[v].size());
    }
    d.ell_k = ell_k;
    std::vector<int> eidx((size_t)ell_k * V, 0);
    std::vector<float> ew((size_t)ell_k * V, 0.f);
    std::vector<uint32_t> idx4(V, 0);
    std::vector<float4> w4(V, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<int> cnt(J + 1, 0);
    for (int v = 0; v < V; ++v) {
      float wq[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t packed = 0;
      for (size_t k = 0; k < rows[v].size(); ++k) {
        eidx[k * V + v] = rows[v][k].second;
        ew[k * V + v] = rows[v][k].first;
        cnt[rows[v][k].second + 1]++;
        if (k < 4) {
          wq[k] = rows[v][k].first;
          packed |= (uint32_t)rows[v][k].second << (8 * k);
        }
      }
      idx4[v] = packed;
      w4[v] = make_float4(wq[0], wq[1], wq[2], wq[3]);
    }
    for (int j = 0; j < J; ++j) cnt[j + 1] += cnt[j];
    std::vector<int> cptr(cnt), cvert(cnt[J]);
    std::vector<float> cw(cnt[J]);
    std::vector<int> fillp(cnt.begin(), cnt.end() - 1);
    for (int v = 0; v < V; ++v)
      for (auto& e : rows[v]) {
        const int pos = fillp[e.second]++;
        cvert[pos] = v;
        cw[pos] = e.first;
      }
    // 4-vertex groups: distinct joints of the group (<= kGrpJoints on the fast path), weights per vertex
    {
      const int G = (V + 3) / 4;
      std::vector<uint2> gj(G, make_uint2(0u, 0u));
      std::vector<float4> gw((size_t)G * kGrpJoints, make_float4(0.f, 0.f, 0.f, 0.f));
      bool ok = ell_k <= 4;
      for (int g = 0; g < G && ok; ++g) {
        std::vector<std::pair<float, int>> uniq;   // (total weight, joint)
        for (int i = 0; i < 4; ++i) {
          const int v = 4 * g + i;
          if (v >= V) break;
          for (auto& e : rows[v]) {
            bool found = false;
            for (auto& u : uniq) if (u.second == e.second) { u.first += e.first; found = true; }
            if (!found) uniq.push_back({e.first, e.second});
          }
        }
        if ((int)uniq.size() > kGrpJoints) { ok = false; break; }
        std::sort(uniq.begin(), uniq.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
          return a.first > b.first || (a.first == b.first && a.second < b.second);
        });
        uint32_t packed[2] = {0u, 0u};
        for (size_t u = 0; u < uniq.size(); ++u) {
          packed[u / 4] |= (uint32_t)uniq[u].second << (8 * (u % 4));
          float wv[4] = {0.f, 0.f, 0.f, 0.f};
          for (int i = 0; i < 4; ++i) {
            const int v = 4 * g + i;
            if (v >= V) break;
            for (auto& e : rows[v]) if (e.second == uniq[u].second) wv[i] = e.first;
          }
          gw[(size_t)g * kGrpJoints + u] = make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
        // unused slots repeat the group's first joint with zero weight (no extra smem line touched)
        for (size_t u = uniq.size(); u < (size_t)kGrpJoints; ++u)
          packed[u / 4] |= (uint32_t)(uniq.empty() ? 0 : uniq[0].second) << (8 * (u % 4));
        gj[g] = make_uint2(packed[0], packed[1]);
      }
      d.grp_ok = ok ? 1 : 0;
      if (int r = upload(mdl, gj, &d.grp_joints)) return r;
      if (int r = upload(mdl, gw, &d.grp_w)) return r;
    }
    // 8-vertex groups (skin_grouped8_kernel): same idea, twice the reuse of every fetched transform.  The 8
    // heaviest joints of a group sit in the register slots; a group bound to more (the few groups where the
    // vertex ranges of several joints meet) keeps the rest in an overflow list only its own lane walks.
    {
      const int G = (V + 7) / 8;
      std::vector<uint2> gj(G, make_uint2(0u, 0u));
      std::vector<float4> gw((size_t)G * kGrpJoints * 2, make_float4(0.f, 0.f, 0.f, 0.f));
      std::vector<int> optr(G + 1, 0), ojoint;
      std::vector<float4> ow;
      int ovf_groups = 0;
      for (int g = 0; g < G; ++g) {
        std::vector<std::pair<float, int>> uniq;
        for (int i = 0; i < 8; ++i) {
          const int v = 8 * g + i;
          if (v >= V) break;
          for (auto& e : rows[v]) {
            bool found = false;
            for (auto& u : uniq) if (u.second == e.second) { u.first += e.first; found = true; }
            if (!found) uniq.push_back({e.first, e.second});
          }
        }
        std::sort(uniq.begin(), uniq.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
          return a.first > b.first || (a.first == b.first && a.second < b.second);
        });
        auto weights_of = [&](int joint, float4* lo, float4* hi) {
          float wv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int i = 0; i < 8; ++i) {
            const int v = 8 * g + i;
            if (v >= V) break;
            for (auto& e : rows[v]) if (e.second == joint) wv[i] = e.first;
          }
          *lo = make_float4(wv[0], wv[1], wv[2], wv[3]);
          *hi = make_float4(wv[4], wv[5], wv[6], wv[7]);
        };
        uint32_t packed[2] = {0u, 0u};
        for (size_t u = 0; u < (size_t)kGrpJoints; ++u) {
          const int jj = uniq.empty() ? 0 : (u < uniq.size() ? uniq[u].second : uniq[0].second);
          packed[u / 4] |= (uint32_t)jj << (8 * (u % 4));
          if (u >= uniq.size()) continue;
          weights_of(uniq[u].second, &gw[((size_t)g * kGrpJoints + u) * 2 + 0], &gw[((size_t)g * kGrpJoints + u) * 2 + 1]);
        }
        for (size_t u = kGrpJoints; u < uniq.size(); ++u) {
          float4 lo, hi;
          weights_of(uniq[u].second, &lo, &hi);
          ojoint.push_back(uniq[u].second);
          ow.push_back(lo);
          ow.push_back(hi);
        }
        if (uniq.size() > (size_t)kGrpJoints) ++ovf_groups;
        optr[g + 1] = (int)ojoint.size();
        gj[g] = make_uint2(packed[0], packed[1]);
      }
      // measured on B200 (profiles/r02_skin_variants.txt): with overflowing groups in a warp the 8-vertex kernel
      // is slower than the 4-vertex one (0.208 vs 0.172 ms at 4096 bodies), so it is taken only when no group
      // overflows (rigid / smooth rigs); dense weight matrices stay on the generic ELL kernel
      d.grp8_ok = (ovf_groups == 0 && ell_k <= 4) ? 1 : 0;
      if (int r = upload(mdl, gj, &d.grp8_joints)) return r;
      if (int r = upload(mdl, gw, &d.grp8_w)) return r;
      if (int r = upload(mdl, optr, &d.grp8_ovf_ptr)) return r;
      if (int r = upload(mdl, ojoint, &d.grp8_ovf_joint)) return r;
      if (int r = upload(mdl, ow, &d.grp8_ovf_w)) return r;
    }
    if (int r = upload(mdl, eidx, &d.ell_idx)) return r;
    if (int r = upload(mdl, ew, &d.ell_w)) return r;
    if (int r = upload(mdl, idx4, &d.skin_idx4)) return r;
    if (int r = upload(mdl, w4, &d.skin_w4)) return r;
    if (int r = upload(mdl, cptr, &d.csc_ptr)) return r;
    if (int r = upload(mdl, cvert, &d.csc_vert)) return r;
    if (int r = upload(mdl, cw, &d.csc_w)) return r;
    {   // segments of the joint -> vertex lists (dA_seg_kernel)
      std::vector<int> sbeg, slen, jptr(J + 1, 0);
      for (int j = 0; j < J; ++j) {
        for (int n = cptr[j]; n < cptr[j + 1]; n += kDASeg) {
          sbeg.push_back(n);
          slen.push_back(std::min(kDASeg, cptr[j + 1] - n));
        }
        jptr[j + 1] = (int)sbeg.size();
      }
      d.seg_count = (int)sbeg.size();
      double worst = 0.0;
      for (int v = 0; v < V; ++v) {
        double sum = 0.0;
        for (auto& e : rows[v]) sum += (double)e.first;
        worst = std::max(worst, std::fabs(sum - 1.0));
      }
      d.w_rows_normalised = worst <= 1e-6 ? 1 : 0;
      if (int r = upload(mdl, sbeg, &d.seg_beg)) return r;
      if (int r = upload(mdl, slen, &d.seg_len)) return r;
      if (int r = upload(mdl, jptr, &d.joint_seg_ptr)) return r;
    }
    // fused epilogue tables: for every 12-vertex chunk of every 84-vertex tile, the distinct joints
    // its vertices are bound to and, per joint, the chunk's weights (0 where a vertex does not use it)
    if (!lbs_only) {
      const int tiles = (V + kFzTileVerts - 1) / kFzTileVerts;
      std::vector<int> off(1, 0), cj;
      std::vector<float4> cwv;
      for (int t = 0; t < tiles; ++t)
        for (int c = 0; c < kFzChunks; ++c) {
          std::vector<int> js;
          for (int i = 0; i < kFzChunkVerts; ++i) {
            const int lv = c * kFzChunkVerts + i, v = t * kFzTileVerts + lv;
            if (lv >= kFzTileVerts || v >= V) break;
            for (auto& e : rows[v])
              if (std::find(js.begin(), js.end(), e.second) == js.end()) js.push_back(e.second);
          }
          std::sort(js.begin(), js.end());
          for (int j : js) {
            float wv[16];
            for (int i = 0; i < 16; ++i) {
              wv[i] = 0.f;
              const int lv = c * kFzChunkVerts + i, v = t * kFzTileVerts + lv;
              if (i >= kFzChunkVerts || lv >= kFzTileVerts || v >= V) continue;
              for (auto& e : rows[v]) if (e.second == j) wv[i] = e.first;
            }
            cj.push_back(j);
            for (int i = 0; i < 4; ++i) cwv.push_back(make_float4(wv[4 * i], wv[4 * i + 1], wv[4 * i + 2], wv[4 * i + 3]));
          }
          off.push_back((int)cj.size());
        }
      // the fused epilogue pays per (chunk, joint) entry; canonical (<= 4 weights, locally coherent)
      // rigs have ~4-8 joints per chunk, dense weight matrices have all J -> keep those on the
      // two-kernel path
      d.fz_ok = (cj.size() <= (size_t)12 * tiles * kFzChunks) ? 1 : 0;
      cj.resize(cj.size() + 64, 0);      // the epilogue reads joint ids in 32-entry windows
      if (int r = upload(mdl, off, &d.fz_off)) return r;
      if (int r = upload(mdl, cj, &d.fz_joint)) return r;
      if (int r = upload(mdl, cwv, &d.fz_w)) return r;
    } else {
      d.fz_ok = 0;
    }
  }

  // ---- hand PCA, pose mean
  {
    std::vector<float> cl, cr, pm;
    if (d.C > 0) {
      if (J < 31) return fail(SMPLK_E_SHAPE, "hand PCA needs a skeleton with 30 hand joints");
      cl.resize((size_t)d.C * 45);
      cr.resize((size_t)d.C * 45);
      for (size_t i = 0; i < cl.size(); ++i) {
        cl[i] = (float)desc->hand_comp_l[i];
        cr[i] = (float)desc->hand_comp_r[i];
      }
    }
    if (desc->pose_mean) {
      pm.resize(3 * J);
      for (int i = 0; i < 3 * J; ++i) pm[i] = (float)desc->pose_mean[i];
    }
    if (int r = upload(mdl, cl, &d.comp_l)) return r;
    if (int r = upload(mdl, cr, &d.comp_r)) return r;
    if (int r = upload(mdl, pm, &d.pose_mean)) return r;
    if (!desc->pose_mean) d.pose_mean = nullptr;
  }

  // ---- vertex picks and posed-vertex regressors
  {
    std::vector<int> ev(d.E);
    for (int e = 0; e < d.E; ++e) {
      ev[e] = desc->extra_vertex_ids[e];
      if (ev[e] < 0 || ev[e] >= V) return fail(SMPLK_E_ARG, "extra_vertex_ids[%d] out of range", e);
    }
    if (int r = upload(mdl, ev, &d.extra_vids)) return r;
    d.pick_pd = nullptr;
    if (!lbs_only && d.E > 0) {   // rows of the blend operand at the picked coordinates (sparse keypoint backward)
      std::vector<float> pp((size_t)3 * d.E * d.Kpad, 0.f);
      for (int e = 0; e < d.E; ++e)
        for (int c = 0; c < 3; ++c) {
          const int n = 3 * ev[e] + c;
          const double* pdrow = desc->posedirs + (size_t)n * d.P;
          const double* sdrow = desc->shapedirs + (size_t)n * NB;
          for (int k = 0; k < d.K; ++k)
            pp[(size_t)(3 * e + c) * d.Kpad + k] = (float)(k < d.P ? pdrow[k] : sdrow[k - d.P]);
        }
      if (int r = upload(mdl, pp, &d.pick_pd)) return r;
    }
    std::vector<int> rptr(d.R + 1, 0), rcol;
    std::vector<float> rval;
    for (int r = 0; r < d.R; ++r) {
      for (int v = 0; v < V; ++v) {
        const double w = desc->regressor_posed[(size_t)r * V + v];
        if (w != 0.0) {
          rcol.push_back(v);
          rval.push_back((float)w);
        }
      }
      rptr[r + 1] = (int)rcol.size();
    }
    if (int r = upload(mdl, rptr, &d.reg_ptr)) return r;
    if (int r = upload(mdl, rcol, &d.reg_col)) return r;
    if (int r = upload(mdl, rval, &d.reg_val)) return r;
    // the same picks / regressor entries, grouped by the vertex they read (gather form of the backward)
    {
      std::vector<std::vector<std::pair<int, float>>> by_vert(V);
      for (int e = 0; e < d.E; ++e) by_vert[ev[e]].push_back({e, 1.0f});
      for (int r = 0; r < d.R; ++r)
        for (int n = rptr[r]; n < rptr[r + 1]; ++n) by_vert[rcol[n]].push_back({d.E + r, rval[n]});
      std::vector<int> sv, sp(1, 0), ss;
      std::vector<float> sw;
      for (int v = 0; v < V; ++v) {
        if (by_vert[v].empty()) continue;
        sv.push_back(v);
        for (auto& e : by_vert[v]) { ss.push_back(e.first); sw.push_back(e.second); }
        sp.push_back((int)ss.size());
      }
      d.sc_T = (int)sv.size();
      if (int r = upload(mdl, sv, &d.sc_vert)) return r;
      if (int r = upload(mdl, sp, &d.sc_ptr)) return r;
      if (int r = upload(mdl, ss, &d.sc_src)) return r;
      if (int r = upload(mdl, sw, &d.sc_w)) return r;
    }
  }

  // ---- rigged-mesh replay operand P[v][4 j + k] = w[v][j] [v_template[v]; 1]_k, scaled by a power of two
  mdl->replay_gemm_ok = false;
  mdl->skin_gemm_ok = false;
  if (lbs_only && mdl->cc_major == 10) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr)
      return fail(SMPLK_E_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
    mdl->encode = reinterpret_cast<EncodeTiledFn>(fn);
    const int Kp = round_up(4 * J, kRpKB);
    double maxabs = 0.0;
    for (int v = 0; v < V; ++v)
      for (int j = 0; j < J; ++j) {
        const double w = desc->weights[(size_t)v * J + j];
        if (w == 0.0) continue;
        maxabs = std::max(maxabs, std::fabs(w));
        for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::fabs(w * desc->v_template[3 * v + k]));
      }
    float scale = 1.f;
    if (maxabs > 0.0 && std::isfinite(maxabs)) scale = std::ldexp(1.f, (int)std::floor(std::log2(1024.0 / maxabs)));
    std::vector<__half> ph((size_t)V * Kp, __float2half(0.f)), pl((size_t)V * Kp, __float2half(0.f));
    for (int v = 0; v < V; ++v)
      for (int j = 0; j < J; ++j) {
        const double w = desc->weights[(size_t)v * J + j];
        if (w == 0.0) continue;
        for (int k = 0; k < 4; ++k) {
          const float x = (float)(w * (k < 3 ? desc->v_template[3 * v + k] : 1.0)) * scale;
          const __half h = __float2half_rn(x);
          ph[(size_t)v * Kp + 4 * j + k] = h;
          pl[(size_t)v * Kp + 4 * j + k] = __float2half_rn(x - __half2float(h));
        }
      }
    if (int r = upload(mdl, ph, &mdl->rp_hi)) return r;
    if (int r = upload(mdl, pl, &mdl->rp_lo)) return r;
    mdl->rp_kp = Kp; mdl->rp_scale = scale;
    const CUtensorMapL2promotion p256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (int r = make_tmap_2d(mdl, &mdl->tmap_rp_hi, mdl->rp_hi, Kp, V, kRpKB, kBlendBM, p256, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
    if (int r = make_tmap_2d(mdl, &mdl->tmap_rp_lo, mdl->rp_lo, Kp, V, kRpKB, kBlendBM, p256, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
    CUDA_TRY(cudaFuncSetAttribute(lbs_replay_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRpSmemAlloc));
    mdl->replay_gemm_ok = true;
  }

  // ---- TMA descriptors of the constant GEMM operand
  mdl->has_tma = false;
  if (!lbs_only && mdl->cc_major == 10) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr)
      return fail(SMPLK_E_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
    mdl->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pd_hi, d.pd_nk_hi, d.Kpad, d.Npad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pd_lo, d.pd_nk_lo, d.Kpad, d.Npad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdkn_hi, d.pd_kn_hi, d.Npad, d.Kpad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdkn_lo, d.pd_kn_lo, d.Npad, d.Kpad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdh_hi, d.pd_nk_h_hi, d.Kpad, d.Npad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, true)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdh_lo, d.pd_nk_h_lo, d.Kpad, d.Npad, kBlendBN,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, true)) return r;
    const CUtensorMapL2promotion p256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pd_hi, d.pd_nk_hi, d.Kpad, d.Npad, p256, false)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pd_lo, d.pd_nk_lo, d.Kpad, d.Npad, p256, false)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdh_hi, d.pd_nk_h_hi, d.Kpad, d.Npad, p256, true)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdh_lo, d.pd_nk_h_lo, d.Kpad, d.Npad, p256, true)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdkn_hi, d.pd_kn_hi, d.Npad, d.Kpad, p256, false)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdkn_lo, d.pd_kn_lo, d.Npad, d.Kpad, p256, false)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdknh_hi, d.pd_kn_h_hi, d.Npad, d.Kpad, p256, true)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdknh_lo, d.pd_kn_h_lo, d.Npad, d.Kpad, p256, true)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdknh_hi, d.pd_kn_h_hi, d.Npad, d.Kpad, kBlendBN, p256, true)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdknh_lo, d.pd_kn_h_lo, d.Npad, d.Kpad, kBlendBN, p256, true)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdknb_hi, d.pd_kn_b_hi, d.Npad, d.Kpad, p256, true)) return r;
    if (int r = make_operand_tmap_2cta(mdl, &mdl->tmap2_pdknb_lo, d.pd_kn_b_lo, d.Npad, d.Kpad, p256, true)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdknb_hi, d.pd_kn_b_hi, d.Npad, d.Kpad, kBlendBN, p256, true)) return r;
    if (int r = make_operand_tmap(mdl, &mdl->tmap_pdknb_lo, d.pd_kn_b_lo, d.Npad, d.Kpad, kBlendBN, p256, true)) return r;
    if (int r = make_operand_tmap_fused(mdl, &mdl->tmapf_pdh_hi, d.pdf_h_hi, d.Kpad, (uint64_t)d.fz_tiles * kBlendBN, p256)) return r;
    if (int r = make_operand_tmap_fused(mdl, &mdl->tmapf_pdh_lo, d.pdf_h_lo, d.Kpad, (uint64_t)d.fz_tiles * kBlendBN, p256)) return r;
    CUDA_TRY(cudaFuncSetAttribute(blend_skin_fused_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFzSmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_skin_fused_kernel<3 * 6890, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFzSmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_skin_fused_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFzSmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_tcgen05_2cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  k2SmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_tcgen05_2cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  k2SmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kGemmSmemAlloc));
    CUDA_TRY(cudaFuncSetAttribute(blend_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kGemmSmemAlloc));
    {
      // skin weights as a GEMM operand (lbs_replay_gemm.cuh, kSkin): W[v][j], power-of-two scale, fp16 hi + lo
      const int Kp = round_up(J, kRpKB);
      double maxabs = 0.0;
      for (size_t i = 0; i < (size_t)V * J; ++i) maxabs = std::max(maxabs, std::fabs(desc->weights[i]));
      float scale = 1.f;
      if (maxabs > 0.0 && std::isfinite(maxabs)) scale = std::ldexp(1.f, (int)std::floor(std::log2(1024.0 / maxabs)));
      std::vector<__half> wh((size_t)V * Kp, __float2half(0.f)), wl((size_t)V * Kp, __float2half(0.f));
      for (int v = 0; v < V; ++v)
        for (int j = 0; j < J; ++j) {
          const float x = (float)desc->weights[(size_t)v * J + j] * scale;
          if (x == 0.f) continue;
          const __half h = __float2half_rn(x);
          wh[(size_t)v * Kp + j] = h;
          wl[(size_t)v * Kp + j] = __float2half_rn(x - __half2float(h));
        }
      if (int r = upload(mdl, wh, &mdl->sw_hi)) return r;
      if (int r = upload(mdl, wl, &mdl->sw_lo)) return r;
      mdl->sw_kp = Kp; mdl->sw_scale = scale;
      const CUtensorMapL2promotion p256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
      if (int r = make_tmap_2d(mdl, &mdl->tmap_sw_hi, mdl->sw_hi, Kp, V, kRpKB, kBlendBM, p256, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
      if (int r = make_tmap_2d(mdl, &mdl->tmap_sw_lo, mdl->sw_lo, Kp, V, kRpKB, kBlendBM, p256, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
      CUDA_TRY(cudaFuncSetAttribute(lbs_replay_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRpSmemAlloc));
      mdl->skin_gemm_ok = true;
    }
    mdl->has_tma = true;
  }
  {
    // Dynamic shared-memory limits are attributes of the KERNEL, not of a model handle: they are
    // raised to the device's opt-in maximum once, so that a handle created later for a smaller
    // skeleton cannot lower the limit a larger model's launches rely on.
    int max_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, mdl->device));
    // (a kernel's static shared memory counts against the same per-block limit)
#define SMPLK_MAX_DYN_SMEM(kernel)                                                                        \
    do {                                                                                                  \
      cudaFuncAttributes fa_;                                                                             \
      CUDA_TRY(cudaFuncGetAttributes(&fa_, kernel));                                                      \
      CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                                    max_optin - (int)fa_.sharedSizeBytes));                               \
    } while (0)
    SMPLK_MAX_DYN_SMEM((skin_grouped_kernel<true>));
    SMPLK_MAX_DYN_SMEM((skin_grouped_kernel<false>));
    SMPLK_MAX_DYN_SMEM((skin_grouped8_kernel<true>));
    SMPLK_MAX_DYN_SMEM((skin_grouped8_kernel<false>));
#ifdef SMPLK_AB
    SMPLK_MAX_DYN_SMEM((skin_tma_kernel));
#endif
    SMPLK_MAX_DYN_SMEM((pose_forward_block_kernel<1>));
    SMPLK_MAX_DYN_SMEM((pose_forward_block_kernel<2>));
    SMPLK_MAX_DYN_SMEM((pose_backward_kernel<1>));
    SMPLK_MAX_DYN_SMEM((pose_backward_kernel<2>));
    SMPLK_MAX_DYN_SMEM((skin_backward_grouped_kernel<kSkinBwdStages, false>));
    SMPLK_MAX_DYN_SMEM((skin_backward_grouped_kernel<kSkinBwdStages, true>));
    SMPLK_MAX_DYN_SMEM((divide_faces_kernel));
    SMPLK_MAX_DYN_SMEM((skin_fit_l2_kernel));
    SMPLK_MAX_DYN_SMEM((pick_backward_kernel));
    SMPLK_MAX_DYN_SMEM((pick_forward_kernel));
#undef SMPLK_MAX_DYN_SMEM
  }
  return 0;
}

extern "C" int smplk_model_create(const smplk_model_desc* desc, int device, smplk_model** out) {
  if (!desc || !out) return fail(SMPLK_E_ARG, "null argument");
  *out = nullptr;
  if (desc->num_joints < 1 || desc->num_joints > kMaxJoints)
    return fail(SMPLK_E_SHAPE, "num_joints=%d unsupported (1..%d)", desc->num_joints, kMaxJoints);
  if (desc->num_verts < 1) return fail(SMPLK_E_SHAPE, "num_verts must be positive");
  if (!desc->v_template || !desc->weights || !desc->parents)
    return fail(SMPLK_E_ARG, "v_template, weights and parents are required");
  if (desc->posedirs && (desc->num_betas < 0 || desc->num_betas > 300))
    return fail(SMPLK_E_SHAPE, "num_betas=%d unsupported", desc->num_betas);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SMPLK_E_DEVICE, "no CUDA device available (%s); smplk has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(SMPLK_E_DEVICE, "device %d out of range", device);
  DEVICE_GUARD(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  smplk_model* mdl = new (std::nothrow) smplk_model();
  if (!mdl) return fail(SMPLK_E_ARG, "out of host memory");
  memset(&mdl->d, 0, sizeof(mdl->d));
  mdl->device = device;
  mdl->num_sms = prop.multiProcessorCount;
  mdl->cc_major = prop.major;
  mdl->stage_dev = nullptr;
  mdl->stage_bytes = 0;
  mdl->copy_stream = nullptr;
  mdl->encode = nullptr;
  mdl->prof_on = false;
  // kernel choices: defaults are the product path; smplk_model_set_option changes them per handle (cross-checks of the
  // parity tests, stand-alone kernel timings of the bench).  The product library reads NO environment variable.
  mdl->skin_bpb = 0; mdl->skin_g8 = true; mdl->skin_tma = false; mdl->force_skin_v1 = false; mdl->da_v1 = false;
  mdl->fit_fused = true; mdl->sparse_picks = true; mdl->use_2cta = true; mdl->use_fused = true; mdl->fused_tma_out = true; mdl->use_replay_gemm = true; mdl->use_skin_gemm = false;
  mdl->use_pose_block = true; mdl->use_pdl = true; mdl->skip_pose = false; mdl->bwd_f16 = true; mdl->default_tc = BLEND_F16;
#ifdef SMPLK_AB   // A/B builds (tools/): tuning switches of kernels that are on no default path
  { const char* e = getenv("SMPLK_DA_V1"); mdl->da_v1 = e && e[0] == '1'; }
  { const char* e = getenv("SMPLK_SKIN_BPB"); mdl->skin_bpb = e ? atoi(e) : 0; }
  { const char* e = getenv("SMPLK_SKIN_G8"); mdl->skin_g8 = !(e && e[0] == '0'); }
  { const char* e = getenv("SMPLK_SKIN_TMA"); mdl->skin_tma = (e && e[0] == '1'); }
  { const char* e = getenv("SMPLK_SKIN_V1"); mdl->force_skin_v1 = e && e[0] == '1'; }
#endif
  for (int i = 0; i < SMPLK_PROF_SLOTS; ++i) { mdl->prof_ms[i] = 0.0; mdl->prof_n[i] = 0; }
  if (prop.major != 10) {
    delete mdl;
    return fail(SMPLK_E_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  }
  int r = build_model(desc, mdl);
  if (r != 0) {
    smplk_model_destroy(mdl);
    return r;
  }
  *out = mdl;
  return 0;
}

extern "C" int smplk_model_set_option(smplk_model* model, const char* name, int value) {
  if (!model || !name) return fail(SMPLK_E_ARG, "null argument");
  const bool on = value != 0;
  if (!strcmp(name, "fused_tma_out")) model->fused_tma_out = on;
  else if (!strcmp(name, "replay_gemm")) model->use_replay_gemm = on;   // rigged-mesh replay on the tensor cores
  else if (!strcmp(name, "skin_gemm")) model->use_skin_gemm = on;       // two-kernel forward: transform blend on the tensor cores
  else if (!strcmp(name, "fused")) model->use_fused = on;                    // fused blend + skinning forward kernel
  else if (!strcmp(name, "skip_pose")) model->skip_pose = on;           // measurement aid: reuse the previous call's pose-kernel output
  else if (!strcmp(name, "pdl")) model->use_pdl = on;                    // programmatic dependent launch between a call's kernels
  else if (!strcmp(name, "pose_block")) model->use_pose_block = on;     // block-level pose kernel (else warp per body)
  else if (!strcmp(name, "blend_tf32")) model->default_tc = on ? BLEND_TF32 : BLEND_F16;   // 3xTF32 forward operands
  else if (!strcmp(name, "backward_tf32")) model->bwd_f16 = !on;        // 3xTF32 backward GEMM
  else if (!strcmp(name, "gemm_2cta")) model->use_2cta = on;            // CTA-pair GEMM kernels
  else if (!strcmp(name, "fit_fused")) model->fit_fused = on;           // one-kernel skinning + loss + skinning backward
  else if (!strcmp(name, "sparse_picks")) model->sparse_picks = on;     // sparse backward of joints-only losses
  else return fail(SMPLK_E_ARG, "unknown option '%s'", name);
  return 0;
}

extern "C" int smplk_model_get_info(const smplk_model* model, smplk_model_info* info) {
  if (!model || !info) return fail(SMPLK_E_ARG, "null argument");
  const ModelDev& d = model->d;
  info->num_verts = d.V; info->num_joints = d.J; info->num_betas = d.NB;
  info->num_pose_feats = d.P; info->num_extra_verts = d.E; info->num_regressors = d.R;
  info->num_pca = d.C; info->max_weights_per_vertex = d.ell_k; info->lbs_only = d.lbs_only;
  info->device = model->device; info->has_tcgen05_path = model->has_tma ? 1 : 0;
  return 0;
}

// ------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------
constexpr int kDefaultChunk = 8192;

struct WsLayout {
  int chunk;
  size_t off_fhi, off_flo, off_A, off_At, off_vposed, off_dvp, total;
};

static WsLayout ws_layout(const ModelDev& d, int batch, uint32_t flags) {
  WsLayout w;
  if (flags & SMPLK_FLAG_FIT_VERTEX_L2) flags |= SMPLK_FLAG_SAVE_FOR_BACKWARD;
  w.chunk = (flags & SMPLK_FLAG_SAVE_FOR_BACKWARD) ? batch : std::min(batch, kDefaultChunk);
  if (w.chunk < 1) w.chunk = 1;
  const size_t rows_pad = (size_t)round_up(w.chunk, kBlendBM);
  size_t off = 0;
  w.off_fhi = off; off += align_up(rows_pad * d.Kpad * sizeof(float), 1024);
  w.off_flo = off; off += align_up(rows_pad * d.Kpad * sizeof(float), 1024);
  w.off_A = off;   off += align_up((size_t)w.chunk * d.J * 12 * sizeof(float), 1024);
  // transposed transforms of the fused blend+skinning kernel (256-body blocks)
  // ... or, in the two-kernel forward, the transforms as the operand rows of the skinning GEMM (12 per body, hi + lo)
  w.off_At = off;  if (!d.lbs_only) off += align_up(std::max((size_t)round_up(w.chunk, 2 * kBlendBM) * d.J * 12 * sizeof(float),
                                                             (size_t)2 * 12 * w.chunk * round_up(d.J, kRpKB) * sizeof(__half)), 1024);
  // rigged-mesh replay: the frames' transforms as the GEMM's second operand, fp16 hi rows then lo rows (3 per frame)
  else off += align_up((size_t)2 * 3 * w.chunk * round_up(4 * d.J, kRpKB) * sizeof(__half), 1024);
  w.off_vposed = off;
  if (!d.lbs_only) off += align_up((size_t)w.chunk * d.Npad * sizeof(float), 1024);
  // d_v_posed of the fused fitting step: bf16 hi rows, then bf16 lo rows
  w.off_dvp = off;
  if ((flags & SMPLK_FLAG_FIT_VERTEX_L2) && !d.lbs_only) off += align_up((size_t)2 * w.chunk * d.Npad * sizeof(uint16_t), 1024);
  w.total = off;
  return w;
}

extern "C" size_t smplk_workspace_bytes(const smplk_model* model, int32_t batch, uint32_t flags) {
  if (!model || batch < 1) return 0;
  return ws_layout(model->d, batch, flags).total;
}

extern "C" int smplk_workspace_layout(const smplk_model* model, int32_t batch, uint32_t flags,
                                      size_t offsets[4], int32_t* chunk) {
  if (!model || !offsets || batch < 1) return fail(SMPLK_E_ARG, "bad argument");
  const WsLayout w = ws_layout(model->d, batch, flags);
  offsets[0] = w.off_fhi; offsets[1] = w.off_flo; offsets[2] = w.off_A; offsets[3] = w.off_vposed;
  if (chunk) *chunk = w.chunk;
  return 0;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// The block kernel (tables in smem, 32 bodies per block, writes At itself) when its shared memory fits.
static bool pose_block_applies(const smplk_model* mdl) {
  return mdl->use_pose_block && (size_t)pose_block_layout(mdl->d).total * sizeof(float) <= 200 * 1024;
}

static int launch_pose_forward(const smplk_model* mdl, const PoseFwdArgs& pa, cudaStream_t st) {
  const ModelDev& d = mdl->d;
  // small batches: the warp-per-body kernel skips staging the tables (0.019 vs 0.024 ms at 37 bodies)
  if (pose_block_applies(mdl) && (pa.At != nullptr || pa.B >= 128)) {
    // 32 bodies per block when the block also transposes the transforms (lane = body); otherwise 8, so that
    // mid-size batches (the 1,024-body fitting step: 32 blocks before, 22 us on 32 SMs) use every SM
    // At is written for whole 256-body blocks (the fused kernel reads zero transforms for padding rows)
    const int bodies = pa.At ? pa.At_rows : pa.B;
    int nw = (pa.At != nullptr || pa.B >= 4096) ? kPoseBlockWarps : 8;
    // whole waves: spread the bodies evenly over every SM instead of filling 32-body blocks (4,096 bodies: 147 blocks of
    // 28 instead of 128 of 32; an 8,192-body chunk: 293 blocks of 28 = two full waves instead of 256 of 32 = 1.73 --
    // measured: no change at 4,096 bodies, 11 % of the forward at 1,024 bodies, where 32-body blocks used 32 SMs)
    if (nw == kPoseBlockWarps) {
      const int waves = (bodies + kPoseBlockWarps * mdl->num_sms - 1) / (kPoseBlockWarps * mdl->num_sms);
      const int slots = waves * mdl->num_sms;
      nw = std::max(8, std::min(kPoseBlockWarps, (bodies + slots - 1) / slots));
    }
    size_t smem = (size_t)pose_block_layout(d, nw).total * sizeof(float);
    const int blocks = (bodies + nw - 1) / nw;
    // One block per SM when the grid is a single wave: under a programmatic dependent launch the blocks are placed while
    // the previous kernel drains, and small blocks would otherwise pile up (4 deep) on the SMs that free up first
    // (measured: 1,024-body forward 0.0767 ms without, 0.0810 ms with PDL before this request was padded).
    if (mdl->use_pdl && blocks <= mdl->num_sms) smem = std::max(smem, (size_t)116 * 1024);
    ProfScope prof(mdl, st, SMPLK_PROF_POSE_FWD);
    if (d.J <= 32)
      launch_k(mdl->use_pdl, pose_forward_block_kernel<1>, blocks, nw * 32, smem, st, d, pa);
    else
      launch_k(mdl->use_pdl, pose_forward_block_kernel<2>, blocks, nw * 32, smem, st, d, pa);
    LAUNCH_CHECK("pose_forward_block_kernel");
    return 0;
  }
  const int blocks = (pa.B + kPoseWarps - 1) / kPoseWarps;
  const size_t smem = (size_t)kPoseWarps * (std::max(d.Kpad, 32) + d.J * 12) * sizeof(float);
  ProfScope prof(mdl, st, SMPLK_PROF_POSE_FWD);
  if (d.J <= 32)
    launch_k(mdl->use_pdl, pose_forward_kernel<1>, blocks, kPoseWarps * 32, smem, st, d, pa);
  else
    launch_k(mdl->use_pdl, pose_forward_kernel<2>, blocks, kPoseWarps * 32, smem, st, d, pa);
  LAUNCH_CHECK("pose_forward_kernel");
  return 0;
}

static BlendPath choose_blend(const smplk_model* mdl, int rows, uint32_t flags) {
  if (flags & SMPLK_FLAG_BLEND_SIMT) return BLEND_SIMT;
  if (flags & SMPLK_FLAG_BLEND_TF32) return BLEND_TF32;
  if (flags & SMPLK_FLAG_BLEND_TCGEN05) return mdl->default_tc;
  // Tensor cores at every batch size: even a single body (127 of 128 tile rows are padding) takes
  // 19 us on the tcgen05 kernel against 60-135 us for the SIMT kernel, whose 81 blocks cannot pull the
  // 40 MB operand fast enough -- and batch 1 is what the reference's fitting loop runs
  // (lib/Gen_SMPLH/fit_single_frame.py:97).  SMPLK_FLAG_BLEND_SIMT still selects the exact-fp32 kernel.
  return mdl->has_tma ? mdl->default_tc : BLEND_SIMT;
}

static int launch_blend(const smplk_model* mdl, int rows, BlendPath path, float* F_hi, float* F_lo,
                        float* v_posed, cudaStream_t st) {
  const ModelDev& d = mdl->d;
  if (path != BLEND_SIMT) {
    if (!mdl->has_tma) return fail(SMPLK_E_DEVICE, "tcgen05 blend path unavailable on this device");
    const bool f16 = path == BLEND_F16;
    CUtensorMap tm_fhi, tm_flo, tm_out;
    if (mdl->use_2cta && rows > kBlendBM) {
      // CTA pairs: 256-body x 256-coord tiles, B operand halved per CTA
      if (int r = make_operand_tmap_2cta(mdl, &tm_fhi, F_hi, d.Kpad, rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
      if (int r = make_operand_tmap_2cta(mdl, &tm_flo, F_lo, d.Kpad, rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
      if (int r = make_tmap_2d(mdl, &tm_out, v_posed, d.Npad, rows, kEpiCols, kBlendBM,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE)) return r;
      BlendGemmArgs ga;
      const int epb = 128 / (f16 ? 2 : 4);
      ga.num_m_blocks = (rows + 2 * kBlendBM - 1) / (2 * kBlendBM);
      ga.num_n_blocks = d.Npad / kBlendBN;
      ga.num_k_blocks = (d.Kpad + epb - 1) / epb;
      ga.num_splits = 1;
      ga.k_blocks_per_split = ga.num_k_blocks;
      ga.out_rows_per_split = 0;
      ga.k_elems = d.Kpad;
      ga.out_scale = f16 ? 1.0f / d.pd_scale : 1.0f;
      ga.bias = d.bias;
      ga.row_scale = nullptr; ga.row_scale_rows = 0;
      ga.out = v_posed; ga.out_ld = d.Npad; ga.out_rows = rows; ga.out_cols = d.Npad;
      const int tiles = ga.num_m_blocks * ga.num_n_blocks;
      const int grid = 2 * std::min(tiles, mdl->num_sms / 2);
      ProfScope prof(mdl, st, SMPLK_PROF_BLEND_TCGEN05);
      if (f16)
        launch_k(mdl->use_pdl, blend_tcgen05_2cta_kernel<true>, grid, kGemmThreads, k2SmemAlloc, st,
                 tm_fhi, tm_flo, mdl->tmap2_pdh_hi, mdl->tmap2_pdh_lo, tm_out, ga);
      else
        launch_k(mdl->use_pdl, blend_tcgen05_2cta_kernel<false>, grid, kGemmThreads, k2SmemAlloc, st,
                 tm_fhi, tm_flo, mdl->tmap2_pd_hi, mdl->tmap2_pd_lo, tm_out, ga);
      LAUNCH_CHECK("blend_tcgen05_2cta_kernel");
      return 0;
    }
    if (int r = make_operand_tmap(mdl, &tm_fhi, F_hi, d.Kpad, rows, kBlendBM,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
    if (int r = make_operand_tmap(mdl, &tm_flo, F_lo, d.Kpad, rows, kBlendBM,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
    if (int r = make_tmap_2d(mdl, &tm_out, v_posed, d.Npad, rows, kEpiCols, kBlendBM,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE)) return r;
    BlendGemmArgs ga;
    const int elems_per_block = kRowBytes / (f16 ? 2 : 4);
    ga.num_m_blocks = (rows + kBlendBM - 1) / kBlendBM;
    ga.num_n_blocks = d.Npad / kBlendBN;
    ga.num_k_blocks = (d.Kpad + elems_per_block - 1) / elems_per_block;
    ga.num_splits = 1;
    ga.k_blocks_per_split = ga.num_k_blocks;
    ga.out_rows_per_split = 0;
    ga.k_elems = d.Kpad;
    ga.out_scale = f16 ? 1.0f / d.pd_scale : 1.0f;
    ga.bias = d.bias;
    ga.row_scale = nullptr; ga.row_scale_rows = 0;
    ga.out = v_posed; ga.out_ld = d.Npad; ga.out_rows = rows; ga.out_cols = d.Npad;
    const int tiles = ga.num_m_blocks * ga.num_n_blocks;
    const int grid = std::min(tiles, mdl->num_sms);
    ProfScope prof(mdl, st, SMPLK_PROF_BLEND_TCGEN05);
    if (f16)
      blend_tcgen05_kernel<true><<<grid, kGemmThreads, kGemmSmemAlloc, st>>>(
          tm_fhi, tm_flo, mdl->tmap_pdh_hi, mdl->tmap_pdh_lo, tm_out, ga);
    else
      blend_tcgen05_kernel<false><<<grid, kGemmThreads, kGemmSmemAlloc, st>>>(
          tm_fhi, tm_flo, mdl->tmap_pd_hi, mdl->tmap_pd_lo, tm_out, ga);
    LAUNCH_CHECK("blend_tcgen05_kernel");
  } else {
#ifdef SMPLK_AB
    BlendSimtArgs sa;
    sa.M = rows; sa.F_hi = F_hi; sa.F_lo = F_lo; sa.out = v_posed;
    dim3 grid((d.Npad / 4 + kSimtThreads - 1) / kSimtThreads, (rows + kSimtBodies - 1) / kSimtBodies);
    const size_t smem = (size_t)d.Kpad * kSimtBodies * sizeof(float);
    ProfScope prof(mdl, st, SMPLK_PROF_BLEND_SIMT);
    blend_simt_kernel<<<grid, kSimtThreads, smem, st>>>(d, sa);
    LAUNCH_CHECK("blend_simt_kernel");
#else
    return fail(SMPLK_E_UNSUPPORTED, "the exact-fp32 SIMT blend kernel (SMPLK_FLAG_BLEND_SIMT) is an A/B variant: "
                                      "build the library with -DSMPLK_AB");
#endif
  }
  return 0;
}

// Fused forward (blend GEMM + skinning epilogue): true when this call can take it.
static bool fused_applies(const smplk_model* mdl, int rows, BlendPath path, uint32_t flags, bool want_verts) {
  const ModelDev& d = mdl->d;
  return mdl->use_fused && mdl->has_tma && mdl->use_2cta && !d.lbs_only && d.fz_ok && want_verts &&
         path == BLEND_F16 && rows > kBlendBM && !(flags & SMPLK_FLAG_SAVE_FOR_BACKWARD);
}

static int launch_fused(const smplk_model* mdl, int rows, float* F_hi, float* F_lo, const float* A,
                        float* At, const float* transl, float* out, cudaStream_t st) {
  const ModelDev& d = mdl->d;
  const int rows_pad = round_up(rows, 2 * kBlendBM);
  if (A != nullptr) {                        // null: the pose kernel already wrote At
    ProfScope prof(mdl, st, SMPLK_PROF_TRANSPOSE);
    const int n4 = rows_pad * d.J * 3;
    transpose_transforms_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(rows, rows_pad, d.J, A, transl, At);
    LAUNCH_CHECK("transpose_transforms_kernel");
  }
  CUtensorMap tm_fhi, tm_flo;
  if (int r = make_operand_tmap_fused(mdl, &tm_fhi, F_hi, d.Kpad, rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return r;
  if (int r = make_operand_tmap_fused(mdl, &tm_flo, F_lo, d.Kpad, rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return r;
  FusedArgs fa;
  fa.num_m_blocks = rows_pad / (2 * kBlendBM);
  fa.num_n_blocks = d.fz_tiles;
  fa.num_k_blocks = (d.Kpad + kFzElemsPerBlock - 1) / kFzElemsPerBlock;
  fa.k_elems = d.Kpad;
  fa.out_scale = 1.0f / d.pd_scale;
  fa.bias = d.bias_f;
  fa.ch_off = d.fz_off; fa.ch_joint = d.fz_joint; fa.ch_w = d.fz_w;
  fa.At = At; fa.J = d.J;
  fa.out = out; fa.rows = rows; fa.N = d.N;
  fa.out_odd_shift = (d.V % 4 == 2) ? 2 : 0;
  fa.dbg = nullptr;
  static long long* dbg_dev = nullptr;
  const bool dbg = SMPLK_FZ_TIMELINE && getenv("SMPLK_FZ_DEBUG") != nullptr;   // tools/fz_timeline.py
  if (dbg) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, (2 + kFzEpiWarps) * kFzDbgTiles * 4 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, (2 + kFzEpiWarps) * kFzDbgTiles * 4 * sizeof(long long), st);
    fa.dbg = dbg_dev;
  }
  const int tiles = fa.num_m_blocks * fa.num_n_blocks;
  const int grid = 2 * std::min(tiles, mdl->num_sms / 2);
  // Output as TMA tensor stores when two (B, V, 3) rows make a 16-byte multiple (V even) and the caller's
  // buffer is 16-byte aligned: see the kernel's kTmaOut note.  Otherwise the per-lane store path.
  const bool tma_out = mdl->fused_tma_out && (d.V % 2 == 0) && rows >= 2 &&
                       (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  CUtensorMap tm_oe, tm_oo, tm_oo32;
  if (tma_out) {
    const auto promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (int r = make_tmap_2d(mdl, &tm_oe, out, (uint64_t)d.N, (uint64_t)(rows + 1) / 2, kFzChunkCols, 16, promo,
                             false, CU_TENSOR_MAP_SWIZZLE_NONE, 2 * (uint64_t)d.N)) return r;
    if (int r = make_tmap_2d(mdl, &tm_oo, out, 2 * (uint64_t)d.N, (uint64_t)rows / 2, kFzChunkCols, 16, promo,
                             false, CU_TENSOR_MAP_SWIZZLE_NONE, 2 * (uint64_t)d.N)) return r;
    if (int r = make_tmap_2d(mdl, &tm_oo32, out, 2 * (uint64_t)d.N, (uint64_t)rows / 2, 32, 16, promo,
                             false, CU_TENSOR_MAP_SWIZZLE_128B, 2 * (uint64_t)d.N)) return r;
  } else {
    tm_oe = tm_fhi; tm_oo = tm_fhi; tm_oo32 = tm_fhi;          // never dereferenced
  }
  ProfScope prof(mdl, st, SMPLK_PROF_BLEND_SKIN_FUSED);
  // programmatic dependent launch only straight after the block pose kernel (with `A` given, the transposition pass,
  // which has no trigger, sits in between and the attribute would buy nothing)
  const bool pdl = mdl->use_pdl && A == nullptr && !dbg;
  if (tma_out)
    launch_k(pdl, blend_skin_fused_kernel<0, true>, grid, kFzThreads, kFzSmemAlloc, st, tm_fhi, tm_flo, mdl->tmapf_pdh_hi,
             mdl->tmapf_pdh_lo, tm_oe, tm_oo, tm_oo32, fa);
  else if (d.N == 3 * 6890)   // canonical SMPL-family vertex count: row pitch folded into the store addresses
    launch_k(pdl, blend_skin_fused_kernel<3 * 6890, false>, grid, kFzThreads, kFzSmemAlloc, st, tm_fhi, tm_flo,
             mdl->tmapf_pdh_hi, mdl->tmapf_pdh_lo, tm_oe, tm_oo, tm_oo32, fa);
  else
    launch_k(pdl, blend_skin_fused_kernel<0, false>, grid, kFzThreads, kFzSmemAlloc, st, tm_fhi, tm_flo,
             mdl->tmapf_pdh_hi, mdl->tmapf_pdh_lo, tm_oe, tm_oo, tm_oo32, fa);
  LAUNCH_CHECK("blend_skin_fused_kernel");
  if (dbg) {   // tuning aid: per-tile timeline of CTA 0 (cycles relative to the first stamp)
    std::vector<long long> h((2 + kFzEpiWarps) * kFzDbgTiles * 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long t0 = h[(1 * kFzDbgTiles + 0) * 4 + 0];
    for (int it = 0; it < kFzDbgTiles; ++it) {
      if (!h[(1 * kFzDbgTiles + it) * 4 + 0]) break;
      fprintf(stderr, "tile %2d mma[wait %7lld go %7lld done %7lld]", it, h[(1 * kFzDbgTiles + it) * 4 + 0] - t0,
              h[(1 * kFzDbgTiles + it) * 4 + 1] - t0, h[(1 * kFzDbgTiles + it) * 4 + 2] - t0);
      for (int w = 2; w < 2 + kFzEpiWarps; ++w) {
        const long long* r = &h[(w * kFzDbgTiles + it) * 4];
        fprintf(stderr, " | w%d %lld %lld %lld %lld", w, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0);
      }
      fprintf(stderr, "\n");
    }
    for (int i = 0; i < 28; ++i) {
      const long long* r = &h[i * 4];
      if (r[0]) fprintf(stderr, "tile %d chunk %d: start %lld load+entries %lld stsA %lld store_window %lld\n", i / 7, i % 7,
                        r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2]);
    }
  }
  return 0;
}

// Bodies per block of the streaming skinning kernels: the grid (tiles x body groups) should fill a
// whole number of waves of `resident` co-resident blocks -- measured on the fitting-step kernel at
// 1,024 bodies: 0.095 ms at 25 bodies per block (287 blocks, one wave of 296) against 0.106 ms at
// 8 (3.03 waves), 0.118-0.121 ms at 21 / 32 (1.16 / 0.76 waves).  Fewest waves with <= max_bpb bodies.
static int pick_bpb(int rows, int tiles, int resident, int max_bpb) {
  for (int k = 1; k <= 32; ++k) {
    const int groups = (resident * k) / tiles;
    if (groups < 1) continue;
    const int bpb = (rows + groups - 1) / groups;
    if (bpb <= max_bpb) return std::max(bpb, 1);
  }
  return max_bpb;
}

// Rigged-mesh replay of `rows` frames on the tensor cores: transforms -> fp16 operand rows, then one GEMM whose
// epilogue writes the (rows, V, 3) vertices (lbs_replay_gemm.cuh).  With `vsrc` (per-body v_posed rows): the skinning
// pass of the two-kernel forward, transform blend as the GEMM, applied to v_posed in the epilogue (kSkin).
static int launch_replay_gemm(const smplk_model* mdl, int rows, const float* A, const float* transl, float* out,
                              __half* T, cudaStream_t st, const float* vsrc = nullptr, size_t vstride = 0) {
  const ModelDev& d = mdl->d;
  const bool skin = vsrc != nullptr;
  const int Kp = skin ? mdl->sw_kp : mdl->rp_kp;
  const int rows_per = skin ? 12 : 3;                  // operand rows (= accumulator columns) per body / frame
  __half* T_hi = T;
  __half* T_lo = T + (size_t)rows_per * rows * Kp;
  {
    ProfScope prof(mdl, st, SMPLK_PROF_TRANSPOSE);      // the transforms' re-layout pass, as for the fused forward
    if (skin) {
      launch_k(mdl->use_pdl, skin_operand_kernel, rows, kSkinOpThreads, 0, st, rows, d.J, Kp, A, T_hi, T_lo);
    } else {
      const long n = (long)3 * rows * (Kp / 4);
      launch_k(mdl->use_pdl, replay_operand_kernel, (unsigned)((n + 255) / 256), 256, 0, st, rows, d.J, Kp, A, T_hi, T_lo);
    }
    LAUNCH_CHECK("replay_operand_kernel");
  }
  CUtensorMap tm_hi, tm_lo;
  const CUtensorMapL2promotion p128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (int r = make_tmap_2d(mdl, &tm_hi, T_hi, Kp, (uint64_t)rows_per * rows, kRpKB, kRpBN / 2, p128, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
  if (int r = make_tmap_2d(mdl, &tm_lo, T_lo, Kp, (uint64_t)rows_per * rows, kRpKB, kRpBN / 2, p128, true, CU_TENSOR_MAP_SWIZZLE_64B)) return r;
  ReplayArgs ra;
  ra.V = d.V; ra.F = rows;
  ra.num_m_blocks = (d.V + 2 * kBlendBM - 1) / (2 * kBlendBM);
  const int per_tile = skin ? kRpTileBodies : kRpTileFrames;
  ra.num_n_blocks = (rows + per_tile - 1) / per_tile;
  ra.num_k_blocks = Kp / kRpKB;
  ra.out_scale = 1.0f / (skin ? mdl->sw_scale : mdl->rp_scale);
  ra.transl = transl; ra.out = out;
  ra.vsrc = vsrc; ra.vsrc_stride = vstride;
  const int tiles = ra.num_m_blocks * ra.num_n_blocks;
  const int grid = 2 * std::min(tiles, mdl->num_sms / 2);
  ProfScope prof(mdl, st, SMPLK_PROF_SKIN);
  if (skin)
    launch_k(mdl->use_pdl, lbs_replay_gemm_kernel<true>, grid, kRpThreads, kRpSmemAlloc, st, mdl->tmap_sw_hi, mdl->tmap_sw_lo, tm_hi, tm_lo, ra);
  else
    launch_k(mdl->use_pdl, lbs_replay_gemm_kernel<false>, grid, kRpThreads, kRpSmemAlloc, st, mdl->tmap_rp_hi, mdl->tmap_rp_lo, tm_hi, tm_lo, ra);
  LAUNCH_CHECK("lbs_replay_gemm_kernel");
  return 0;
}

// frames from which the replay GEMM is used: below, its fixed costs (operand pass, 80-frame tiles) outweigh the
// streaming kernel's
constexpr int kReplayGemmMinRows = 64;

static int launch_skin(const smplk_model* mdl, int rows, const float* vsrc, size_t vstride,
                       const float* A, const float* transl, float* out, cudaStream_t st, __half* replay_T = nullptr) {
  const ModelDev& d = mdl->d;
  if (replay_T != nullptr && d.lbs_only && vstride == 0 && mdl->replay_gemm_ok && mdl->use_replay_gemm &&
      rows >= kReplayGemmMinRows)
    return launch_replay_gemm(mdl, rows, A, transl, out, replay_T, st);
  if (replay_T != nullptr && !d.lbs_only && vstride != 0 && (vstride % 2) == 0 && mdl->skin_gemm_ok && mdl->use_skin_gemm &&
      rows >= kReplayGemmMinRows && (reinterpret_cast<uintptr_t>(vsrc) % 8) == 0)
    return launch_replay_gemm(mdl, rows, A, transl, out, replay_T, st, vsrc, vstride);
  SkinArgs sa;
  sa.B = rows;
  const int tiles = (d.V + kSkinTileVerts - 1) / kSkinTileVerts;
  sa.vsrc = vsrc; sa.vsrc_stride = vstride; sa.A = A; sa.transl = transl; sa.out = out;
  sa.debug_copy_only = 0;
#ifdef SMPLK_AB
  { const char* e = getenv("SMPLK_SKIN_COPYONLY"); sa.debug_copy_only = (e && e[0] == '1') ? 1 : 0; }
#endif
  const bool grouped = d.grp_ok && !mdl->force_skin_v1;
  // bodies per block: see pick_bpb (whole waves; 0.177 -> 0.169 ms at B=4096, 0.056 -> 0.049 ms at 1024)
  const int resident = (grouped ? 2 : 4) * mdl->num_sms;
  int bpb = 32;
  if (grouped) {
    // rigged-mesh replay: the block's transforms stay in shared memory for its whole body loop -- as many
    // bodies as fit two blocks per SM
    const int fit = (SMPLK_SKIN_NOBAR && vstride == 0) ? skin_nobar_max_bodies(d.J, (size_t)110 * 1024) : 32;
    bpb = pick_bpb(rows, tiles, resident, std::max(1, std::min(32, fit)));
  } else {
    while (bpb > 8 && (long)tiles * ((rows + bpb - 1) / bpb) < 6L * resident) bpb >>= 1;
    while (bpb > 1 && (long)tiles * ((rows + bpb - 1) / bpb) < resident) bpb >>= 1;
  }
  if (mdl->skin_bpb > 0) bpb = mdl->skin_bpb;
  sa.bodies_per_block = bpb;
  const size_t smem_grouped = (SMPLK_SKIN_NOBAR && vstride == 0) ? skin_nobar_smem_bytes(d.J, bpb)
                                                                   : skin_grouped_smem_bytes(d.J);
  const size_t smem = grouped ? smem_grouped
                              : (size_t)(2 * kSkinTileVerts * 3 + 2 * ((d.J * 12 + 3) & ~3)) * sizeof(float);
  dim3 grid(tiles, (rows + bpb - 1) / bpb);
  ProfScope prof(mdl, st, SMPLK_PROF_SKIN);
  if (grouped && d.grp8_ok && mdl->skin_g8) {
    const int res8 = 3 * mdl->num_sms;
    int bpb8 = pick_bpb(rows, tiles, res8, 32);
    if (mdl->skin_bpb > 0) bpb8 = mdl->skin_bpb;
    sa.bodies_per_block = bpb8;
    dim3 grid8(tiles, (rows + bpb8 - 1) / bpb8);
    if (vstride == 0) skin_grouped8_kernel<true><<<grid8, k8Threads, skin_grouped8_smem_bytes(d.J), st>>>(d, sa);
    else skin_grouped8_kernel<false><<<grid8, k8Threads, skin_grouped8_smem_bytes(d.J), st>>>(d, sa);
#ifdef SMPLK_AB
  } else if (grouped && vstride != 0 && mdl->skin_tma) {
    skin_tma_kernel<<<grid, kGrpThreads, skin_tma_smem_bytes(d.J), st>>>(d, sa);
#endif
  } else if (grouped) {
    if (vstride == 0) skin_grouped_kernel<true><<<grid, kGrpThreads, smem, st>>>(d, sa);
    else skin_grouped_kernel<false><<<grid, kGrpThreads, smem, st>>>(d, sa);
  } else if (d.ell_k <= 4) {
    skin_kernel<true><<<grid, kSkinThreads, smem, st>>>(d, sa);
  } else {
    skin_kernel<false><<<grid, kSkinThreads, smem, st>>>(d, sa);
  }
  LAUNCH_CHECK("skin_kernel");
  return 0;
}

// smplk_fit_vertex_l2: what replaces the plain skinning launch
struct FitL2 {
  const float* target;
  float scale;
  float* loss;
};

// the fused skinning + loss + skinning-backward kernel needs the 4-vertex group tables, 8-byte
// aligned (B,V,3) rows and the fp16-class backward GEMM
static bool fit_fused_applies(const smplk_model* mdl) {
  const ModelDev& d = mdl->d;
  return !d.lbs_only && d.grp_ok && !mdl->force_skin_v1 && mdl->has_tma && mdl->bwd_f16 && mdl->fit_fused &&
         ((d.V * 3) % 2 == 0);
}

static int forward_impl(const smplk_model* model, const smplk_forward_args* a, uint32_t flags, const FitL2* fit);
static int vertex_l2_impl(int32_t batch, int32_t floats_per_body, const float* verts, const float* target,
                          float scale, float* grad, float* loss, int loss_stride, int device, smplk_stream stream);

extern "C" int smplk_forward(const smplk_model* model, const smplk_forward_args* a) {
  if (!model || !a) return fail(SMPLK_E_ARG, "null argument");
  if (a->flags & SMPLK_FLAG_FIT_VERTEX_L2)
    return fail(SMPLK_E_ARG, "SMPLK_FLAG_FIT_VERTEX_L2 belongs to smplk_fit_vertex_l2");
  return forward_impl(model, a, a->flags, nullptr);
}

extern "C" int smplk_fit_vertex_l2(const smplk_model* model, const smplk_forward_args* a, const float* target,
                                   float scale, float* loss) {
  if (!model || !a || !target || !loss) return fail(SMPLK_E_ARG, "null argument");
  if (!a->verts) return fail(SMPLK_E_ARG, "the verts buffer (it receives the vertex gradient) is required");
  if (a->joints_regressed || (a->joints && model->d.E > 0))
    return fail(SMPLK_E_ARG, "vertex-pick / regressed joints are not available from smplk_fit_vertex_l2");
  if (a->flags & SMPLK_FLAG_TRANSFORMS_ONLY) return fail(SMPLK_E_ARG, "SMPLK_FLAG_TRANSFORMS_ONLY excludes the loss");
  if (fit_fused_applies(model) && ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(a->verts)) & 7))
    return fail(SMPLK_E_ARG, "target and verts must be 8-byte aligned");
  FitL2 fit;
  fit.target = target; fit.scale = scale; fit.loss = loss;
  return forward_impl(model, a, a->flags | SMPLK_FLAG_FIT_VERTEX_L2 | SMPLK_FLAG_SAVE_FOR_BACKWARD, &fit);
}

static int forward_impl(const smplk_model* model, const smplk_forward_args* a, const uint32_t flags, const FitL2* fit) {
  const ModelDev& d = model->d;
  if (a->batch < 1) return fail(SMPLK_E_ARG, "batch must be >= 1");
  if (!a->pose) return fail(SMPLK_E_ARG, "pose is required");
  if (a->betas && a->betas_batch != 1 && a->betas_batch != a->batch)
    return fail(SMPLK_E_SHAPE, "betas_batch must be 1 or batch (got %d for batch %d)",
                a->betas_batch, a->batch);
  if ((a->hand_pca_l || a->hand_pca_r) && d.C == 0)
    return fail(SMPLK_E_ARG, "hand PCA coefficients given but the model has no PCA components");
  if ((flags & SMPLK_FLAG_TRANSFORMS_ONLY) && (a->verts || a->joints_regressed))
    return fail(SMPLK_E_ARG, "SMPLK_FLAG_TRANSFORMS_ONLY excludes the verts / joints_regressed outputs");
  // joints without vertices: the E picked vertices alone are blended and skinned (pick_forward_kernel)
  const bool picks_fwd = !(flags & SMPLK_FLAG_TRANSFORMS_ONLY) && !a->verts && a->joints && d.E > 0 &&
                         !a->joints_regressed && (d.lbs_only || d.pick_pd != nullptr) && fit == nullptr;
  if (!(flags & SMPLK_FLAG_TRANSFORMS_ONLY) && !picks_fwd &&
      ((a->joints && d.E > 0 && !a->verts) || (a->joints_regressed && !a->verts)))
    return fail(SMPLK_E_ARG, "regressed joints need the verts output buffer");
  if (a->joints_regressed && d.R == 0)
    return fail(SMPLK_E_ARG, "joints_regressed requested but the model has no regressor_posed");
  const WsLayout w = ws_layout(d, a->batch, flags);
  if (!a->workspace || a->workspace_bytes < w.total)
    return fail(SMPLK_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
                a->workspace_bytes);
  if (reinterpret_cast<uintptr_t>(a->workspace) & 255)
    return fail(SMPLK_E_WORKSPACE, "workspace must be 256-byte aligned");
  if (reinterpret_cast<uintptr_t>(a->verts) & 7)     // the skinning kernels store 8-byte pairs
    return fail(SMPLK_E_ARG, "verts must be 8-byte aligned");
  DEVICE_GUARD(model->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  float* F_hi = reinterpret_cast<float*>(ws + w.off_fhi);
  float* F_lo = reinterpret_cast<float*>(ws + w.off_flo);
  float* A = reinterpret_cast<float*>(ws + w.off_A);
  float* At = reinterpret_cast<float*>(ws + w.off_At);
  float* v_posed = reinterpret_cast<float*>(ws + w.off_vposed);
  const int joints_ld = 3 * (d.J + d.E);

  for (int c0 = 0; c0 < a->batch; c0 += w.chunk) {
    const int rows = std::min(w.chunk, a->batch - c0);
    PoseFwdArgs pa;
    pa.B = rows;
    pa.betas = a->betas ? (a->betas_batch == 1 ? a->betas : a->betas + (size_t)c0 * d.NB) : nullptr;
    pa.betas_B = a->betas_batch == 1 ? 1 : rows;
    pa.pose = a->pose + (size_t)c0 * 3 * d.J;
    pa.pca_l = a->hand_pca_l ? a->hand_pca_l + (size_t)c0 * d.C : nullptr;
    pa.pca_r = a->hand_pca_r ? a->hand_pca_r + (size_t)c0 * d.C : nullptr;
    pa.add_mean = (flags & SMPLK_FLAG_ADD_POSE_MEAN) ? 1 : 0;
    pa.transl = a->transl ? a->transl + (size_t)c0 * 3 : nullptr;
    const BlendPath path = d.lbs_only ? BLEND_SIMT : choose_blend(model, rows, flags);
    const bool f16 = !d.lbs_only && path == BLEND_F16;
    pa.F_hi = (d.lbs_only || f16) ? nullptr : F_hi;
    pa.F_lo = (d.lbs_only || f16) ? nullptr : F_lo;
    pa.H_hi = f16 ? reinterpret_cast<__half*>(F_hi) : nullptr;   // fp16 rows alias the fp32 regions
    pa.H_lo = f16 ? reinterpret_cast<__half*>(F_lo) : nullptr;
    pa.A = A;
    pa.At = nullptr;
    pa.joints = a->joints ? a->joints + (size_t)c0 * joints_ld : nullptr;
    pa.joints_ld = joints_ld;
    pa.full_pose = a->full_pose ? a->full_pose + (size_t)c0 * 3 * d.J : nullptr;
    const bool fused = fused_applies(model, rows, path, flags, a->verts != nullptr);
    const bool at_from_pose = fused && pose_block_applies(model);
    pa.At_rows = round_up(rows, 2 * kBlendBM);
    if (at_from_pose) pa.At = At;            // the block pose kernel writes the transposed transforms itself
    if (fit != nullptr && a->verts && !fused && fit_fused_applies(model)) {
      // the loss accumulator of skin_fit_l2_kernel: zeroed here, ahead of the chunk's first kernel, so that no memset
      // node sits between two kernels of the chain (it would end the programmatic dependent launch there)
      const int loss_stride = (flags & SMPLK_FLAG_LOSS_SUM) ? 0 : 1;
      CUDA_TRY(cudaMemsetAsync(fit->loss + (size_t)c0 * loss_stride, 0, (size_t)(loss_stride ? rows : 1) * sizeof(float), st));
    }
    if (!model->skip_pose) {
      if (int r = launch_pose_forward(model, pa, st)) return r;
    }
    if (flags & SMPLK_FLAG_TRANSFORMS_ONLY) continue;    // pose / FK kernel only: A, joints, full_pose
    if (picks_fwd) {
      PickFwdArgs pf;
      pf.B = rows; pf.H_hi = pa.H_hi; pf.H_lo = pa.H_lo; pf.F_hi = pa.F_hi; pf.F_lo = pa.F_lo;
      pf.A = A; pf.transl = pa.transl;
      pf.vposed = d.lbs_only ? nullptr : v_posed;     // the picked entries are what the sparse backward reads
      pf.joints = pa.joints; pf.joints_ld = joints_ld;
      ProfScope prof(model, st, SMPLK_PROF_SKIN);
      pick_forward_kernel<<<rows, kPickFwdThreads, pick_fwd_smem_bytes(d.J, d.E, std::max(d.Kpad, 1)), st>>>(d, pf);
      LAUNCH_CHECK("pick_forward_kernel");
      continue;
    }
    if (!fused && !d.lbs_only && (a->verts || (flags & SMPLK_FLAG_SAVE_FOR_BACKWARD))) {
      if (int r = launch_blend(model, rows, path, F_hi, F_lo, v_posed, st)) return r;
    }
    if (a->verts) {
      float* vout = a->verts + (size_t)c0 * d.V * 3;
      if (fused) {
        if (int r = launch_fused(model, rows, F_hi, F_lo, at_from_pose ? nullptr : A, At, pa.transl, vout, st)) return r;
      } else {
        if (fit != nullptr && fit_fused_applies(model)) {
          // skinning + loss + gradient + skinning backward in one kernel; vout receives the gradient
          const int loss_stride = (flags & SMPLK_FLAG_LOSS_SUM) ? 0 : 1;      // (zeroed before the pose kernel, above)
          SkinFitArgs fa;
          fa.B = rows;
          fa.vposed = v_posed; fa.vposed_stride = (size_t)d.Npad; fa.A = A; fa.transl = pa.transl;
          fa.target = fit->target + (size_t)c0 * d.V * 3; fa.scale = fit->scale;
          fa.grad = vout; fa.loss = fit->loss + (size_t)c0 * loss_stride; fa.loss_stride = loss_stride;
          fa.dvp_hi = reinterpret_cast<__nv_bfloat16*>(ws + w.off_dvp);
          fa.dvp_lo = fa.dvp_hi + (size_t)w.chunk * d.Npad;
          const int tiles = (d.V + kSkinTileVerts - 1) / kSkinTileVerts;
          int bpb = pick_bpb(rows, tiles, 2 * model->num_sms, 32);
          if (model->skin_bpb > 0) bpb = model->skin_bpb;
          fa.bodies_per_block = bpb;
          ProfScope prof(model, st, SMPLK_PROF_SKIN);
          launch_k(model->use_pdl, skin_fit_l2_kernel, dim3(tiles, (rows + bpb - 1) / bpb), kGrpThreads,
                   skin_grouped_smem_bytes(d.J), st, d, fa);
          LAUNCH_CHECK("skin_fit_l2_kernel");
        } else {
          if (int r = launch_skin(model, rows, d.lbs_only ? d.bias : v_posed,
                                  d.lbs_only ? 0 : (size_t)d.Npad, A, pa.transl, vout, st,
                                  reinterpret_cast<__half*>(At))) return r;
          if (fit != nullptr) {   // generic weights: stand-alone loss kernel, gradient in place
            const int ls = (flags & SMPLK_FLAG_LOSS_SUM) ? 0 : 1;
            if (int r = vertex_l2_impl(rows, d.V * 3, vout, fit->target + (size_t)c0 * d.V * 3, fit->scale, vout,
                                       fit->loss + (size_t)c0 * ls, ls, model->device, a->stream)) return r;
          }
        }
      }
      if (a->joints && d.E > 0) {
        const int n = rows * d.E;
        gather_extra_joints_kernel<<<(n + 127) / 128, 128, 0, st>>>(d, rows, vout, pa.joints, joints_ld);
        LAUNCH_CHECK("gather_extra_joints_kernel");
      }
      if (a->joints_regressed) {
        const long nthreads = (long)rows * d.R * 32;
        regress_joints_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(
            d, rows, vout, a->joints_regressed + (size_t)c0 * d.R * 3);
        LAUNCH_CHECK("regress_joints_kernel");
      }
    }
  }
  return 0;
}

extern "C" int smplk_regress_joints(const smplk_model* model, int32_t batch, const float* verts,
                                    float* out, smplk_stream stream) {
  if (!model || !verts || !out || batch < 1) return fail(SMPLK_E_ARG, "bad argument");
  const ModelDev& d = model->d;
  if (d.R == 0) return fail(SMPLK_E_ARG, "model has no regressor_posed");
  DEVICE_GUARD(model->device);
  const long nthreads = (long)batch * d.R * 32;
  regress_joints_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(d, batch, verts, out);
  LAUNCH_CHECK("regress_joints_kernel");
  return 0;
}

// A[b,j] = [G_R | G_t - G_R J_j]: the rest-pose removal of do_skinning (models/smplh_np.py:73-78)
__global__ void rest_removal_kernel(int n, const float* __restrict__ G, const float* __restrict__ Jr,
                                    float* __restrict__ A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* g = G + (size_t)i * 16;
  const float* j = Jr + (size_t)i * 3;
  float* o = A + (size_t)i * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float r0 = g[4 * r], r1 = g[4 * r + 1], r2 = g[4 * r + 2];
    o[4 * r] = r0; o[4 * r + 1] = r1; o[4 * r + 2] = r2;
    o[4 * r + 3] = g[4 * r + 3] - (r0 * j[0] + r1 * j[1] + r2 * j[2]);
  }
}

extern "C" int smplk_remove_rest(int32_t batch, int32_t num_joints, const float* G, const float* joints_rest,
                                 float* A, int device, smplk_stream stream) {
  if (batch < 1 || num_joints < 1 || !G || !joints_rest || !A) return fail(SMPLK_E_ARG, "bad argument");
  DEVICE_GUARD(device);
  const int n = batch * num_joints;
  rest_removal_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, G, joints_rest, A);
  LAUNCH_CHECK("rest_removal_kernel");
  return 0;
}

extern "C" int smplk_skin_transforms(const smplk_model* model, int32_t batch, const float* G,
                                     const float* joints_rest, const float* v_posed, int32_t v_posed_ld,
                                     const float* transl, float* A, float* verts, smplk_stream stream) {
  if (!model || batch < 1 || !G || !joints_rest || !A || !verts) return fail(SMPLK_E_ARG, "bad argument");
  const ModelDev& d = model->d;
  if (v_posed && (v_posed_ld < 3 * d.V || (v_posed_ld & 3) || (reinterpret_cast<uintptr_t>(v_posed) & 15)))
    return fail(SMPLK_E_ARG, "v_posed rows must be 16-byte aligned: v_posed_ld a multiple of 4 floats >= 3V");
  if (!v_posed && !d.lbs_only) return fail(SMPLK_E_ARG, "v_posed is required for a blendshape model");
  if (reinterpret_cast<uintptr_t>(verts) & 7) return fail(SMPLK_E_ARG, "verts must be 8-byte aligned");
  DEVICE_GUARD(model->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = batch * d.J;
  rest_removal_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, G, joints_rest, A);
  LAUNCH_CHECK("rest_removal_kernel");
  return launch_skin(model, batch, v_posed ? v_posed : d.bias, v_posed ? (size_t)v_posed_ld : 0, A, transl, verts, st);
}

extern "C" int smplk_batch_rodrigues(int32_t n, const float* axis_angle, float* rotmats, int device,
                                     smplk_stream stream) {
  if (n < 1 || !axis_angle || !rotmats) return fail(SMPLK_E_ARG, "bad argument");
  DEVICE_GUARD(device);
  rodrigues_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      n, axis_angle, rotmats);
  LAUNCH_CHECK("rodrigues_kernel");
  return 0;
}

// Host-buffer forward.  The device->host copy of the vertices (82,680 B per body) is ~25x the kernel
// time, so the batch is cut into chunks: chunk c is computed on the caller's stream while chunk c-1
// drains over PCIe on a second (model-owned) stream, ordered by one event per chunk.  The device
// staging buffer holds the whole batch, so a chunk's copy never blocks a later chunk's kernels.
constexpr int kHostChunk = 1024;

extern "C" int smplk_forward_host(smplk_model* model, int32_t batch, uint32_t flags,
                                  const float* betas, int32_t betas_batch, const float* pose,
                                  const float* transl, float* verts, float* joints,
                                  smplk_stream stream) {
  if (!model || !pose || batch < 1) return fail(SMPLK_E_ARG, "bad argument");
  if (betas && betas_batch != 1 && betas_batch != batch)
    return fail(SMPLK_E_SHAPE, "betas_batch must be 1 or batch (got %d for batch %d)", betas_batch, batch);
  const ModelDev& d = model->d;
  DEVICE_GUARD(model->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  flags &= ~SMPLK_FLAG_SAVE_FOR_BACKWARD;
  const int chunk = std::min<int>(batch, kHostChunk);
  const int nchunks = (batch + chunk - 1) / chunk;
  const size_t ws_bytes = smplk_workspace_bytes(model, chunk, flags);
  const size_t nb_betas = betas ? align_up((size_t)betas_batch * d.NB * 4, 256) : 0;
  const size_t nb_pose = align_up((size_t)batch * 3 * d.J * 4, 256);
  const size_t nb_tr = transl ? align_up((size_t)batch * 12, 256) : 0;
  const size_t nb_verts = align_up((size_t)batch * d.V * 12, 256);
  const size_t nb_joints = joints ? align_up((size_t)batch * (d.J + d.E) * 12, 256) : 0;
  const size_t total = align_up(ws_bytes, 256) + nb_betas + nb_pose + nb_tr + nb_verts + nb_joints;
  if (total > model->stage_bytes) {
    if (model->stage_dev) CUDA_TRY(cudaFree(model->stage_dev));
    model->stage_dev = nullptr;
    model->stage_bytes = 0;
    CUDA_TRY(cudaMalloc(&model->stage_dev, total));
    model->stage_bytes = total;
  }
  if (!model->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&model->copy_stream, cudaStreamNonBlocking));
  while ((int)model->chunk_events.size() < nchunks) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    model->chunk_events.push_back(e);
  }
  uint8_t* p = reinterpret_cast<uint8_t*>(model->stage_dev);
  void* ws = p; p += align_up(ws_bytes, 256);
  float* d_betas = betas ? reinterpret_cast<float*>(p) : nullptr; p += nb_betas;
  float* d_pose = reinterpret_cast<float*>(p); p += nb_pose;
  float* d_tr = transl ? reinterpret_cast<float*>(p) : nullptr; p += nb_tr;
  float* d_verts = reinterpret_cast<float*>(p); p += nb_verts;
  float* d_joints = joints ? reinterpret_cast<float*>(p) : nullptr;
  if (betas) CUDA_TRY(cudaMemcpyAsync(d_betas, betas, (size_t)betas_batch * d.NB * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_pose, pose, (size_t)batch * 3 * d.J * 4, cudaMemcpyHostToDevice, st));
  if (transl) CUDA_TRY(cudaMemcpyAsync(d_tr, transl, (size_t)batch * 12, cudaMemcpyHostToDevice, st));
  const size_t jl = (size_t)(d.J + d.E) * 3;
  for (int c = 0; c < nchunks; ++c) {
    const int c0 = c * chunk, rows = std::min(chunk, batch - c0);
    smplk_forward_args fa;
    memset(&fa, 0, sizeof(fa));
    fa.batch = rows; fa.flags = flags;
    fa.betas = d_betas ? (betas_batch == 1 ? d_betas : d_betas + (size_t)c0 * d.NB) : nullptr;
    fa.betas_batch = (betas && betas_batch != 1) ? rows : 1;
    fa.pose = d_pose + (size_t)c0 * 3 * d.J;
    fa.transl = d_tr ? d_tr + (size_t)c0 * 3 : nullptr;
    fa.verts = d_verts + (size_t)c0 * d.V * 3;
    fa.joints = d_joints ? d_joints + (size_t)c0 * jl : nullptr;
    fa.workspace = ws; fa.workspace_bytes = ws_bytes; fa.stream = stream;
    if (int r = smplk_forward(model, &fa)) return r;
    CUDA_TRY(cudaEventRecord(model->chunk_events[c], st));
    CUDA_TRY(cudaStreamWaitEvent(model->copy_stream, model->chunk_events[c], 0));
    if (verts)
      CUDA_TRY(cudaMemcpyAsync(verts + (size_t)c0 * d.V * 3, fa.verts, (size_t)rows * d.V * 12, cudaMemcpyDeviceToHost,
                               model->copy_stream));
    if (joints)
      CUDA_TRY(cudaMemcpyAsync(joints + (size_t)c0 * jl, fa.joints, (size_t)rows * jl * 4, cudaMemcpyDeviceToHost,
                               model->copy_stream));
  }
  CUDA_TRY(cudaStreamSynchronize(model->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int smplk_vertex_l2(int32_t batch, int32_t floats_per_body, const float* verts,
                               const float* target, float scale, float* grad, float* loss,
                               int device, smplk_stream stream) {
  return vertex_l2_impl(batch, floats_per_body, verts, target, scale, grad, loss, 1, device, stream);
}

static int vertex_l2_impl(int32_t batch, int32_t floats_per_body, const float* verts, const float* target,
                          float scale, float* grad, float* loss, int loss_stride, int device, smplk_stream stream) {
  if (batch < 1 || floats_per_body < 1 || !verts || !target || !loss)
    return fail(SMPLK_E_ARG, "bad argument");
  DEVICE_GUARD(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemsetAsync(loss, 0, (size_t)(loss_stride ? batch : 1) * sizeof(float), st));
  int gx = 8;
  while (gx > 1 && (long)gx * batch > 16384) gx >>= 1;
  const int vec2 = (floats_per_body % 2 == 0) && ((reinterpret_cast<uintptr_t>(verts) | reinterpret_cast<uintptr_t>(target) |
                                                   reinterpret_cast<uintptr_t>(grad)) & 7) == 0;
  vertex_l2_kernel<<<dim3(gx, batch), 256, 0, st>>>(floats_per_body, verts, target, scale, grad, loss, vec2, loss_stride);
  LAUNCH_CHECK("vertex_l2_kernel");
  return 0;
}

extern "C" int smplk_profile_enable(smplk_model* model, int enable) {
  if (!model) return fail(SMPLK_E_ARG, "null model");
  model->prof_on = enable != 0;
  return 0;
}

extern "C" int smplk_profile_read(smplk_model* model, double ms[SMPLK_PROF_SLOTS],
                                  int64_t counts[SMPLK_PROF_SLOTS], int reset) {
  if (!model || !ms || !counts) return fail(SMPLK_E_ARG, "null argument");
  DEVICE_GUARD(model->device);
  for (ProfRec& r : model->prof_pending) {
    CUDA_TRY(cudaEventSynchronize(r.e1));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, r.e0, r.e1));
    model->prof_ms[r.slot] += t;
    model->prof_n[r.slot] += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  model->prof_pending.clear();
  for (int i = 0; i < SMPLK_PROF_SLOTS; ++i) {
    ms[i] = model->prof_ms[i];
    counts[i] = model->prof_n[i];
    if (reset) { model->prof_ms[i] = 0.0; model->prof_n[i] = 0; }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// mesh operations either side of the forward
// ------------------------------------------------------------------------------------------
extern "C" int smplk_inverse_lbs(const smplk_model* model, int32_t batch, const float* A,
                                 const float* verts, const float* transl, float* v_rest,
                                 smplk_stream stream) {
  if (!model || batch < 1 || !A || !verts || !v_rest) return fail(SMPLK_E_ARG, "bad argument");
  const ModelDev& d = model->d;
  DEVICE_GUARD(model->device);
  dim3 grid((d.V + 255) / 256, batch);
  inverse_lbs_kernel<<<grid, 256, (size_t)d.J * 12 * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      d, batch, A, verts, transl, v_rest);
  LAUNCH_CHECK("inverse_lbs_kernel");
  return 0;
}

extern "C" int smplk_inverse_joints(int32_t batch, int32_t num_joints, const float* A, const float* joints,
                                    int32_t joints_ld, const float* transl, float* out, int device,
                                    smplk_stream stream) {
  if (batch < 1 || num_joints < 1 || !A || !joints || !out || joints_ld < 3 * num_joints)
    return fail(SMPLK_E_ARG, "bad argument");
  DEVICE_GUARD(device);
  const int n = batch * num_joints;
  inverse_joints_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      batch, num_joints, A, joints, joints_ld, transl, out);
  LAUNCH_CHECK("inverse_joints_kernel");
  return 0;
}

extern "C" int smplk_vertex_normals(int32_t batch, int32_t num_verts, const int32_t* faces,
                                    const int32_t* vf_ptr, const int32_t* vf_face, const float* verts,
                                    float* normals, int device, smplk_stream stream) {
  if (batch < 1 || num_verts < 1 || !faces || !vf_ptr || !vf_face || !verts || !normals)
    return fail(SMPLK_E_ARG, "bad argument");
  DEVICE_GUARD(device);
  dim3 grid((num_verts + 255) / 256, batch);
  vertex_normals_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      batch, num_verts, faces, vf_ptr, vf_face, verts, normals);
  LAUNCH_CHECK("vertex_normals_kernel");
  return 0;
}

extern "C" int smplk_divide_faces(int32_t batch, int32_t num_verts, int32_t num_faces, const int32_t* faces,
                                  const float* verts, int32_t* faces_out, int32_t* vidx_out,
                                  int32_t* counts, int device, smplk_stream stream) {
  if (batch < 1 || num_verts < 1 || num_faces < 1 || !faces || !verts || !faces_out || !vidx_out || !counts)
    return fail(SMPLK_E_ARG, "bad argument");
  const size_t smem = ((size_t)2 * num_verts + 32) * sizeof(int);
  if (smem > 200 * 1024) return fail(SMPLK_E_SHAPE, "divide_faces: %d vertices exceed the shared-memory table", num_verts);
  DEVICE_GUARD(device);
  {   // smplk_divide_faces needs no model handle: raise the kernel's limit here (idempotent)
    int max_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaFuncSetAttribute(divide_faces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  }
  divide_faces_kernel<<<2 * batch, kDivThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      batch, num_verts, num_faces, faces, verts, faces_out, vidx_out, counts);
  LAUNCH_CHECK("divide_faces_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// fitting loss around the body model
// ------------------------------------------------------------------------------------------
extern "C" int smplk_reprojection_loss(const smplk_reprojection_args* a) {
  if (!a || a->batch < 1 || a->num_joints < 1 || !a->joints || !a->rotation || !a->translation ||
      !a->focal || !a->center || !a->gt_joints || !a->loss)
    return fail(SMPLK_E_ARG, "bad argument");
  if ((a->camera_batch != 1 && a->camera_batch != a->batch) ||
      (a->weights && a->weights_batch != 1 && a->weights_batch != a->batch))
    return fail(SMPLK_E_SHAPE, "camera_batch / weights_batch must be 1 or batch");
  DEVICE_GUARD(a->device);
  ReprojArgs r;
  r.B = a->batch; r.Jn = a->num_joints; r.joints = a->joints; r.rotation = a->rotation;
  r.translation = a->translation; r.focal = a->focal; r.center = a->center; r.cam_batch = a->camera_batch;
  r.gt = a->gt_joints; r.weights = a->weights; r.w_batch = a->weights ? a->weights_batch : 1;
  r.rho = a->rho; r.data_weight = a->data_weight; r.loss = a->loss; r.d_joints = a->d_joints;
  r.d_translation = a->d_translation;
  const int warps_per_block = 4;
  reprojection_loss_kernel<<<(a->batch + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0,
                             reinterpret_cast<cudaStream_t>(a->stream)>>>(r);
  LAUNCH_CHECK("reprojection_loss_kernel");
  return 0;
}

extern "C" int smplk_fit_priors(const smplk_prior_args* a) {
  if (!a || a->batch < 1 || !a->loss) return fail(SMPLK_E_ARG, "bad argument");
  if (a->body_pose && a->num_body_pose < 56) return fail(SMPLK_E_SHAPE, "body_pose needs >= 56 columns for the angle prior");
  DEVICE_GUARD(a->device);
  PriorArgs p;
  p.B = a->batch; p.betas = a->betas; p.nb = a->num_betas; p.pose_embedding = a->pose_embedding;
  p.ne = a->num_embedding; p.body_pose = a->body_pose; p.np = a->num_body_pose; p.lhand = a->left_hand_pose;
  p.rhand = a->right_hand_pose; p.nh = a->num_hand; p.shape_weight = a->shape_weight;
  p.body_pose_weight = a->body_pose_weight; p.bending_weight = a->bending_prior_weight;
  p.hand_weight = a->hand_prior_weight; p.loss = a->loss; p.d_betas = a->d_betas;
  p.d_pose_embedding = a->d_pose_embedding; p.d_body_pose = a->d_body_pose; p.d_lhand = a->d_left_hand_pose;
  p.d_rhand = a->d_right_hand_pose;
  fit_priors_kernel<<<a->batch, 64, 0, reinterpret_cast<cudaStream_t>(a->stream)>>>(p);
  LAUNCH_CHECK("fit_priors_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
#include "backward_host.inl"
