// CTA-pair version of the blend GEMM (tcgen05 cta_group::2).
//
// ncu on the single-CTA kernel (profiles/r01_*) shows the main loop limited by shared-memory
// bandwidth, not by the tensor pipe: per 96 KB stage the TMA writes 96 KB and the 12 MMAs read
// 144 KB of operands through the same 128 B/cycle port (measured 2013 cycles/stage in tf32 mode =
// (96+144+17) KB / 128 B, against 1536 cycles of MMA).  A CTA pair computes a 256-body x 256-coord
// tile: each CTA loads its own 128 feature rows and only HALF of the posedirs tile (128 rows), the
// pair's tensor cores share the B halves, so per-CTA smem traffic per stage drops to 64 KB written
// + 96 KB read and the stage budget (64 KB) allows 3 stages.
//
// Protocol (cute / DeepGEMM 2-SM convention):
//   full[s]       lives in the leader CTA, 2 arrivals (one producer per CTA) + bytes of both CTAs
//   empty[s]      one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[a]  one per CTA, same multicast commit
//   tmem_empty[a] in the leader, 256 arrivals (the epilogue threads of both CTAs, remote arrive)
#pragma once
#include "blend_gemm.cuh"

namespace smplk {

#ifndef SMPLK_EPI_DIRECT
#define SMPLK_EPI_DIRECT 0
#endif
// Epilogue store path.  0 (default): swizzled smem staging + TMA store.  1: each thread writes its
// row's 32 fp32 (one 128-byte line) straight from registers with 16-byte stores -- measured SLOWER
// on B200 (f16 blend 0.213 ms vs 0.160 ms at B=4096): 32 partial lines per store instruction cost
// more L2 write transactions than the smem round trip saves.  Kept for A/B runs only.
constexpr bool k2DirectStore = SMPLK_EPI_DIRECT != 0;
constexpr int k2Stages = 3;
constexpr int k2TileABytes = kBlendBM * 128;        // this CTA's 128 rows of the A operand, 128 B of K
constexpr int k2TileBBytes = (kBlendBN / 2) * 128;  // this CTA's half (128 rows) of the B operand
constexpr int k2StageBytes = 2 * k2TileABytes + 2 * k2TileBBytes;   // 64 KB
constexpr int k2SmemBytes = k2Stages * k2StageBytes + 2 * kEpiBufBytes + kBlendBN * 4 + 256;
constexpr int k2SmemAlloc = k2SmemBytes + 1024;

template <bool kF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
blend_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_f_hi,
                          const __grid_constant__ CUtensorMap tmap_f_lo,
                          const __grid_constant__ CUtensorMap tmap_pd_hi,
                          const __grid_constant__ CUtensorMap tmap_pd_lo,
                          const __grid_constant__ CUtensorMap tmap_out, const BlendGemmArgs args) {
  extern __shared__ uint8_t gemm2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(gemm2_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + k2Stages * k2StageBytes;
  float* bias_s = reinterpret_cast<float*>(epi_base + 2 * kEpiBufBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + 2 * kEpiBufBytes + kBlendBN * 4);
  uint64_t* full_bar = bars;                        // [k2Stages]   (used in the leader)
  uint64_t* empty_bar = bars + k2Stages;            // [k2Stages]
  uint64_t* tmem_full = bars + 2 * k2Stages;        // [2]
  uint64_t* tmem_empty = bars + 2 * k2Stages + 2;   // [2]          (used in the leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * k2Stages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  // m blocks of this kernel are 256 rows (one per CTA pair)
  const int num_tiles = args.num_m_blocks * args.num_n_blocks * args.num_splits;
  constexpr int kElemsPerBlock = 128 / (kF16 ? 2 : 4);
  constexpr int kUmmaK = kF16 ? 16 : 8;

  ptx::pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_f_hi);
    ptx::prefetch_tmap(&tmap_f_lo);
    ptx::prefetch_tmap(&tmap_pd_hi);
    ptx::prefetch_tmap(&tmap_pd_lo);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < k2Stages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 256);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta<kTmemCols>(tmem_ptr);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();            // barrier inits + TMEM allocation visible in both CTAs
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::pdl_wait();                // the set-up above overlaps the previous kernel's tail; its output is read from here on

  auto coord = [&](int tile, int& m0, int& n0, int& kb0, int& kb1, int& out_row0) {
    const int mb = tile % args.num_m_blocks;
    const int rest = tile / args.num_m_blocks;
    const int nb = rest % args.num_n_blocks;
    const int sp = rest / args.num_n_blocks;
    m0 = mb * 2 * kBlendBM + (int)rank * kBlendBM;       // this CTA's 128 rows of the 256-row tile
    n0 = nb * kBlendBN;
    kb0 = sp * args.k_blocks_per_split;
    kb1 = min(args.num_k_blocks, kb0 + args.k_blocks_per_split);
    out_row0 = sp * args.out_rows_per_split + m0;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m0, n0, kb0, kb1, orow;
        coord(tile, m0, n0, kb0, kb1, orow);
        const int nb0 = n0 + (int)rank * (kBlendBN / 2);   // this CTA's half of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * k2StageBytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * k2StageBytes);
          else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
          const int k0 = kb * kElemsPerBlock;
          ptx::tma_load_2d_2sm(st, &tmap_f_hi, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + k2TileABytes, &tmap_f_lo, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + 2 * k2TileABytes, &tmap_pd_hi, &full_bar[stage], k0, nb0);
          ptx::tma_load_2d_2sm(st + 2 * k2TileABytes + k2TileBBytes, &tmap_pd_lo, &full_bar[stage], k0, nb0);
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = (kF16 ? ptx::make_idesc_f16(2 * kBlendBM, kBlendBN)
                                   : ptx::make_idesc_tf32(2 * kBlendBM, kBlendBN)) | ((kF16 && args.a_bf16) ? ((1u << 7) | (1u << 10)) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m0, n0, kb0, kb1, orow;
        coord(tile, m0, n0, kb0, kb1, orow);
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlendBN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(stage_base + stage * k2StageBytes);
          const uint64_t a_hi = ptx::make_kmajor_desc<128>(st);
          const uint64_t a_lo = ptx::make_kmajor_desc<128>(st + k2TileABytes);
          const uint64_t b_hi = ptx::make_kmajor_desc<128>(st + 2 * k2TileABytes);
          const uint64_t b_lo = ptx::make_kmajor_desc<128>(st + 2 * k2TileABytes + k2TileBBytes);
          const int ksteps = min(4, (args.k_elems - kb * kElemsPerBlock) / kUmmaK);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < ksteps) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);
              const uint32_t first = (kb != kb0 || k != 0) ? 1u : 0u;
              ptx::umma_2cta<kF16>(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
              ptx::umma_2cta<kF16>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              ptx::umma_2cta<kF16>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          ptx::umma_commit_2cta(&empty_bar[stage]);    // frees the slot in BOTH CTAs
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2cta(&tmem_full[acc]);        // accumulators complete in BOTH CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, both CTAs; own 128 rows) =====================
    const int ewarp = warp & 3;
    const int row = ewarp * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool store_thread = (etid == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int ebuf = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m0, n0, kb0, kb1, orow;
      coord(tile, m0, n0, kb0, kb1, orow);
      float oscale = args.out_scale;
      if (args.row_scale != nullptr) oscale *= (m0 + row < args.row_scale_rows) ? args.row_scale[m0 + row] : 0.f;
      ptx::named_bar_sync(1, 128);
      bias_s[etid] = args.bias ? args.bias[n0 + etid] : 0.f;
      bias_s[etid + 128] = args.bias ? args.bias[n0 + etid + 128] : 0.f;
      if (k2DirectStore) ptx::named_bar_sync(1, 128);   // bias visible to all epilogue warps
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tcgen05_fence_after();
      if (k2DirectStore) {
        const int grow = orow + row;                     // global output row of this thread
        float* orow_ptr = args.out + (size_t)grow * args.out_ld + n0;
        const bool row_ok = grow < args.out_rows;
#pragma unroll 2
        for (int c = 0; c < kBlendBN / kEpiCols; ++c) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) +
                                 static_cast<uint32_t>(acc * kBlendBN + c * kEpiCols);
          ptx::tmem_ld_32x32b_x32(taddr, v);
          ptx::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int col = c * kEpiCols + 4 * q;
              const float4 bq = *reinterpret_cast<const float4*>(bias_s + col);
              float4 o;
              o.x = fmaf(__uint_as_float(v[4 * q + 0]), oscale, bq.x);
              o.y = fmaf(__uint_as_float(v[4 * q + 1]), oscale, bq.y);
              o.z = fmaf(__uint_as_float(v[4 * q + 2]), oscale, bq.z);
              o.w = fmaf(__uint_as_float(v[4 * q + 3]), oscale, bq.w);
              if (n0 + col + 4 <= args.out_cols) *reinterpret_cast<float4*>(orow_ptr + col) = o;
            }
          }
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < kBlendBN / kEpiCols; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) +
                               static_cast<uint32_t>(acc * kBlendBN + c * kEpiCols);
        ptx::tmem_ld_32x32b_x32(taddr, v);
        ptx::tmem_ld_wait();
        if (store_thread) ptx::tma_store_wait_read<1>();
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
          *reinterpret_cast<float4*>(ebase + ((q ^ (row & 7)) << 4)) = o;
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (store_thread) {
          ptx::tma_store_2d(&tmap_out, epi_base + ebuf * kEpiBufBytes, n0 + c * kEpiCols, orow);
          ptx::tma_store_commit();
        }
        ebuf ^= 1;
      }
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);    // leader's barrier collects both CTAs
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (store_thread) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();            // both CTAs done with TMEM and with each other's barriers
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

}  // namespace smplk
