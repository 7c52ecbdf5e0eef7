// Blendshape contraction  v_posed[b, n] = v_template[n] + sum_k F[b,k] * PD[n,k]
//   F  = [ (R_j - I) j>=1 | betas ]           (B x K,  K = 9(J-1) + NB, padded to Kpad)
//   PD = [ posedirs | shapedirs ]^T            (3V x K)
// i.e. models/smplh_np.py:50 (shape blend) and :59 (pose blend) in ONE GEMM whose epilogue adds
// the template; upstream smplx: einsum('bl,mkl->bmk') + matmul(pose_feature, posedirs).
//
// Two kernels:
//  * blend_tcgen05_kernel -- the product path.  Persistent, warp-specialised sm_100a GEMM:
//      warp 0  TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//      warp 1  tcgen05.mma issuer (kind::tf32, M=128 N=256 K=8, accumulators in TMEM)
//      warps 2-5 epilogue (tcgen05.ld -> + v_template -> swizzled smem -> TMA store)
//    fp32 accuracy from TF32 tensor cores via 3xTF32 operand splitting: both operands are stored
//    as hi = tf32(x) and lo = x - hi; each k-step issues hi*hi + lo*hi + hi*lo into the same TMEM
//    accumulator (the lo*lo term, ~2^-22 relative, is dropped).  TMEM holds two 128x256 fp32
//    accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//    The same kernel has a second operand format (template kF16): both operands split into two
//    fp16 terms (hi = half(x), lo = half(x - hi); posedirs pre-scaled by a power of two so the lo
//    terms stay in the normal fp16 range) and kind::f16 MMAs (K=16 per instruction).  Same three
//    products, same ~2^-22 relative accuracy, but half the operand bytes and twice the MMA rate.
//  * blend_simt_kernel -- exact fp32 CUDA-core kernel for tiny batches (B < 32: the batch-1
//    fitting loop of lib/Gen_SMPLH/fitting.py, where a 128-row MMA tile would be >75 % padding
//    and the contraction is bound by streaming posedirs once) and for validating the GEMM.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace smplk {

// ------------------------------------------------------------------------------------------
// tcgen05 path
// ------------------------------------------------------------------------------------------
#ifndef SMPLK_GEMM_ROW_BYTES
#define SMPLK_GEMM_ROW_BYTES 128
#endif
constexpr int kRowBytes = SMPLK_GEMM_ROW_BYTES;        // bytes of K per smem row = swizzle span (128 or 64)
static_assert(kRowBytes == 128 || kRowBytes == 64, "row bytes must match a TMA/UMMA swizzle mode");
constexpr int kGemmStages = kRowBytes == 128 ? 2 : 4;  // 192 KB of operand stages either way
constexpr int kKSteps = kRowBytes / 32;                // UMMA k-steps (32 B each) per k-block
constexpr int kGemmThreads = 192;
constexpr int kTileABytes = kBlendBM * kRowBytes;      // 16 KB (8 KB at 64 B rows)
constexpr int kTileBBytes = kBlendBN * kRowBytes;      // 32 KB (16 KB)
constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;  // hi+lo of both operands
constexpr int kEpiCols = 32;
constexpr int kEpiBufBytes = kBlendBM * kEpiCols * 4;  // 16 KB
constexpr int kGemmSmemBytes = kGemmStages * kStageBytes + 2 * kEpiBufBytes + kBlendBN * 4 + 256;
constexpr int kGemmSmemAlloc = kGemmSmemBytes + 1024;  // slack for manual 1024-byte alignment
constexpr int kTmemCols = 512;

struct BlendGemmArgs {
  int num_m_blocks, num_n_blocks;
  int num_k_blocks;        // total k-blocks of the contraction
  int num_splits;          // split-K factor (1 for the forward blend)
  int k_blocks_per_split;  // ceil(num_k_blocks / num_splits)
  int out_rows_per_split;  // output row offset per split (partials stacked along rows)
  int k_elems;             // contraction length in elements (multiple of the MMA K: 8 tf32 / 16 f16)
  float out_scale;         // accumulator scale applied in the epilogue (1/posedirs scale for f16)
  const float* row_scale;  // [row_scale_rows] extra per-row scale (row = m block row, split independent), or null
  int row_scale_rows;
  const float* bias;       // [num_n_blocks * 256] added in the epilogue, or null
  float* out;              // output matrix for the direct (register -> global) epilogue
  int out_ld;              // floats per output row
  int out_rows, out_cols;  // valid extent (rows / columns beyond it are not written)
  int a_bf16 = 0;          // f16 kernels: both operands hold bf16 (instruction descriptor a_format = b_format = 1)
};

// tile -> (m block, n block, k range); m fastest so CTAs running together share the B operand
struct TileCoord {
  int m0, n0, kb0, kb1, out_row0;
};
__device__ __forceinline__ TileCoord tile_coord(const BlendGemmArgs& a, int tile) {
  TileCoord t;
  const int mb = tile % a.num_m_blocks;
  const int rest = tile / a.num_m_blocks;
  const int nb = rest % a.num_n_blocks;
  const int sp = rest / a.num_n_blocks;
  t.m0 = mb * kBlendBM;
  t.n0 = nb * kBlendBN;
  t.kb0 = sp * a.k_blocks_per_split;
  t.kb1 = min(a.num_k_blocks, t.kb0 + a.k_blocks_per_split);
  t.out_row0 = sp * a.out_rows_per_split + t.m0;
  return t;
}

template <bool kF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
blend_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_f_hi,
                     const __grid_constant__ CUtensorMap tmap_f_lo,
                     const __grid_constant__ CUtensorMap tmap_pd_hi,
                     const __grid_constant__ CUtensorMap tmap_pd_lo,
                     const __grid_constant__ CUtensorMap tmap_out, const BlendGemmArgs args) {
  extern __shared__ uint8_t gemm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + kGemmStages * kStageBytes;
  float* bias_s = reinterpret_cast<float*>(epi_base + 2 * kEpiBufBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + 2 * kEpiBufBytes + kBlendBN * 4);
  uint64_t* full_bar = bars;                      // [kGemmStages]
  uint64_t* empty_bar = bars + kGemmStages;       // [kGemmStages]
  uint64_t* tmem_full = bars + 2 * kGemmStages;   // [2]
  uint64_t* tmem_empty = bars + 2 * kGemmStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kGemmStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = args.num_m_blocks * args.num_n_blocks * args.num_splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_f_hi);
    ptx::prefetch_tmap(&tmap_f_lo);
    ptx::prefetch_tmap(&tmap_pd_hi);
    ptx::prefetch_tmap(&tmap_pd_lo);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < kGemmStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<kTmemCols>(tmem_ptr);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(args, tile);
        const int m0 = tc.m0, n0 = tc.n0;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * kStageBytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          const int k0 = kb * (kRowBytes / (kF16 ? 2 : 4));       // elements per smem row
          ptx::tma_load_2d(st, &tmap_f_hi, &full_bar[stage], k0, m0);
          ptx::tma_load_2d(st + kTileABytes, &tmap_f_lo, &full_bar[stage], k0, m0);
          ptx::tma_load_2d(st + 2 * kTileABytes, &tmap_pd_hi, &full_bar[stage], k0, n0);
          ptx::tma_load_2d(st + 2 * kTileABytes + kTileBBytes, &tmap_pd_lo, &full_bar[stage], k0, n0);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (kF16 ? ptx::make_idesc_f16(kBlendBM, kBlendBN)
                                   : ptx::make_idesc_tf32(kBlendBM, kBlendBN)) | ((kF16 && args.a_bf16) ? ((1u << 7) | (1u << 10)) : 0u);
      constexpr int kElemsPerBlock = kRowBytes / (kF16 ? 2 : 4);
      constexpr int kUmmaK = kF16 ? 16 : 8;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlendBN;
        const TileCoord tc = tile_coord(args, tile);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(stage_base + stage * kStageBytes);
          const uint64_t a_hi = ptx::make_kmajor_desc<kRowBytes>(st);
          const uint64_t a_lo = ptx::make_kmajor_desc<kRowBytes>(st + kTileABytes);
          const uint64_t b_hi = ptx::make_kmajor_desc<kRowBytes>(st + 2 * kTileABytes);
          const uint64_t b_lo = ptx::make_kmajor_desc<kRowBytes>(st + 2 * kTileABytes + kTileBBytes);
          // the last k-block may be partial (TMA zero-fills past k_elems; skip those MMAs)
          const int ksteps = min(kKSteps, (args.k_elems - kb * kElemsPerBlock) / kUmmaK);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            if (k < ksteps) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);  // +32 B per UMMA_K
              const uint32_t first = (kb != tc.kb0 || k != 0) ? 1u : 0u;
              ptx::umma<kF16>(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
              ptx::umma<kF16>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              ptx::umma<kF16>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);   // smem slot free once these MMAs retire
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int ewarp = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = ewarp * 32 + lane;             // row of the 128-row tile == TMEM lane
    const int etid = threadIdx.x - 64;             // 0..127
    const bool store_thread = (etid == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int ebuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(args, tile);
      const int m0 = tc.out_row0, n0 = tc.n0;
      float oscale = args.out_scale;
      if (args.row_scale != nullptr) oscale *= (tc.m0 + row < args.row_scale_rows) ? args.row_scale[tc.m0 + row] : 0.f;
      ptx::named_bar_sync(1, 128);                 // previous tile finished reading bias_s
      bias_s[etid] = args.bias ? args.bias[n0 + etid] : 0.f;
      bias_s[etid + 128] = args.bias ? args.bias[n0 + etid + 128] : 0.f;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < kBlendBN / kEpiCols; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) +
                               static_cast<uint32_t>(acc * kBlendBN + c * kEpiCols);
        ptx::tmem_ld_32x32b_x32(taddr, v);
        ptx::tmem_ld_wait();
        if (store_thread) ptx::tma_store_wait_read<1>();   // buffer `ebuf` no longer being read
        ptx::named_bar_sync(1, 128);
        uint8_t* ebase = epi_base + ebuf * kEpiBufBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(bias_s + c * kEpiCols + 4 * q);
          float4 o;
          o.x = fmaf(__uint_as_float(v[4 * q + 0]), oscale, bq.x);
          o.y = fmaf(__uint_as_float(v[4 * q + 1]), oscale, bq.y);
          o.z = fmaf(__uint_as_float(v[4 * q + 2]), oscale, bq.z);
          o.w = fmaf(__uint_as_float(v[4 * q + 3]), oscale, bq.w);
          *reinterpret_cast<float4*>(ebase + ((q ^ (row & 7)) << 4)) = o;   // 128B swizzle
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (store_thread) {
          ptx::tma_store_2d(&tmap_out, epi_base + ebuf * kEpiBufBytes, n0 + c * kEpiCols, m0);
          ptx::tma_store_commit();
        }
        ebuf ^= 1;
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (store_thread) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

#ifdef SMPLK_AB   // validation kernel (SMPLK_FLAG_BLEND_SIMT): built only with -DSMPLK_AB
// ------------------------------------------------------------------------------------------
// exact-fp32 SIMT path (tiny batches, validation)
// ------------------------------------------------------------------------------------------
constexpr int kSimtThreads = 64;
constexpr int kSimtBodies = 8;

struct BlendSimtArgs {
  int M;                 // bodies
  const float* F_hi;     // [M][Kpad]
  const float* F_lo;
  float* out;            // [M][Npad]
};

__global__ void __launch_bounds__(kSimtThreads)
blend_simt_kernel(const ModelDev m, const BlendSimtArgs a) {
  extern __shared__ __align__(16) float simt_smem[];   // [Kpad][kSimtBodies]
  const int n = (blockIdx.x * kSimtThreads + threadIdx.x) * 4;
  const int b0 = blockIdx.y * kSimtBodies;
  for (int i = threadIdx.x; i < m.Kpad * kSimtBodies; i += kSimtThreads) {
    const int k = i / kSimtBodies, bb = i % kSimtBodies;
    const int b = b0 + bb;
    simt_smem[i] = (b < a.M) ? a.F_hi[(size_t)b * m.Kpad + k] + a.F_lo[(size_t)b * m.Kpad + k] : 0.f;
  }
  __syncthreads();
  if (n >= m.Npad) return;
  float4 acc[kSimtBodies];
  const float4 bias = *reinterpret_cast<const float4*>(m.bias + n);
#pragma unroll
  for (int bb = 0; bb < kSimtBodies; ++bb) acc[bb] = bias;
  const float* pd = m.pd_kn + n;
#pragma unroll 4
  for (int k = 0; k < m.K; ++k) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(pd + (size_t)k * m.Npad));
    const float4 f0 = *reinterpret_cast<const float4*>(simt_smem + k * kSimtBodies);
    const float4 f1 = *reinterpret_cast<const float4*>(simt_smem + k * kSimtBodies + 4);
    const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
    for (int bb = 0; bb < kSimtBodies; ++bb) {
      acc[bb].x = fmaf(f[bb], p.x, acc[bb].x);
      acc[bb].y = fmaf(f[bb], p.y, acc[bb].y);
      acc[bb].z = fmaf(f[bb], p.z, acc[bb].z);
      acc[bb].w = fmaf(f[bb], p.w, acc[bb].w);
    }
  }
#pragma unroll
  for (int bb = 0; bb < kSimtBodies; ++bb) {
    const int b = b0 + bb;
    if (b < a.M) *reinterpret_cast<float4*>(a.out + (size_t)b * m.Npad + n) = acc[bb];
  }
}
#endif  // SMPLK_AB

}  // namespace smplk
