// Fused forward: blendshape GEMM (tcgen05, CTA pairs) whose epilogue applies linear-blend skinning
// straight from the TMEM accumulators, so v_posed never goes to HBM.
//
//   verts[b,v] = sum_k w[v,k] A[b, j[v,k]] [v_template[v] + sum_f F[b,f] PD[3v.., f]; 1] + transl[b]
//
// Reference math: models/smplh_np.py:50,59 (shape + pose blend), :79-82 (T = W.A; v = T [v_posed;1];
// + trans); upstream smplx lbs().  Algorithmic HBM traffic per body drops from 84,580 (GEMM) +
// 167,856 (skinning) bytes to 84,004 (SURVEY 8d "forward, fused end-to-end").
//
// Mapping (10 warps per CTA, one CTA per SM, CTA pairs)
//   * GEMM main loop as in blend_tcgen05_2cta_kernel (TMA producer warp, one MMA-issuing thread in
//     the leader CTA, 256-body x 256-column tiles, two 256-column TMEM accumulators), but with
//     64-byte operand rows x 4 stages (128 KB in flight) so that the epilogue's buffers fit.
//   * The B operand is re-packed so that a 256-column tile holds 84 WHOLE vertices (252 columns,
//     interleaved xyz, + 4 zero columns): column c of tile t is flat output coordinate 252 t + c.
//   * Epilogue = 8 warps; a thread is one body (TMEM lane), a warp the 32 bodies of its TMEM lane
//     quarter.  The two warps of a quarter split EVERY tile's 7 chunks of 12 vertices 3/4
//     (alternating), so an accumulator is held for half an epilogue: with two accumulators the
//     per-tile time is max(T_mma, (T_mma + T_hold) / 2).
//   * Per chunk (36 TMEM columns -> 36 registers + v_template from smem): the chunk's distinct
//     joints come from a table built at model-create time.  For each (chunk, joint) entry the body's
//     3x4 transform is read from the joint-major transform array At[128-body block][joint][128][12]
//     (the 32 bodies of a warp are 1,536 contiguous bytes per joint) and applied to the 12 vertices
//     with the entry's weights (0 where a vertex does not use the joint; branch-free).
//   * Transforms and weights travel through a per-warp cp.async ring in shared memory (kFzRing
//     entries of 13 lines): ptxas tracks ALL global loads of a warp with ONE scoreboard, so a
//     register prefetch cannot run more than one entry ahead -- cp.async groups can.
//   * Results leave through a rolling 32-column staging window per warp (STS.128 rows, LDS columns)
//     and are written as 128-byte row segments.  (The caller's (B, V, 3) fp32 rows are 82,680 bytes
//     = 8 mod 16, so a 2-D TMA store cannot target them directly.)
#pragma once
#include "blend_gemm_2cta.cuh"

namespace smplk {

// Epilogue warps: 8 (two per TMEM lane quarter; the product configuration) or 12 (three per quarter).
// ncu on the 8-warp kernel (profiles/r02_ncu_fused.txt) shows stall_wait as the largest stall (29 % of the
// samples) at 39 % issue-active, which suggested more warps; measured on B200 (profiles/r02_fused_variants.txt)
// the 12-warp build is SLOWER: 0.236 vs 0.213 ms at 4,096 bodies.  14 warps get only 128 registers each
// (the launch fails with 136 or 144: the register file is allocated as for 16 warps) and spill, and the
// extra per-warp buffers cost one GEMM stage (4 -> 3 stages alone: 0.227 ms) and one ring entry; the
// shared-memory pipe (l1tex 67 %), shared with the GEMM's operand traffic, is the second limiter.
#ifndef SMPLK_FZ_EPI_WARPS
#define SMPLK_FZ_EPI_WARPS 8
#endif
#ifndef SMPLK_FZ_RING
#define SMPLK_FZ_RING (SMPLK_FZ_EPI_WARPS == 12 ? 3 : 4)   // transform entries in flight per epilogue warp (cp.async ring in smem)
#endif
#ifndef SMPLK_FZ_BACKOFF_NS
#define SMPLK_FZ_BACKOFF_NS 40
#endif
#ifndef SMPLK_FZ_EPI_BACKOFF_NS
#define SMPLK_FZ_EPI_BACKOFF_NS 20
#endif
#ifndef SMPLK_FZ_PROBE_NO_EPILOGUE
#define SMPLK_FZ_PROBE_NO_EPILOGUE 0
#endif
#ifndef SMPLK_FZ_ROW_BYTES
#define SMPLK_FZ_ROW_BYTES 64
#endif
// GEMM operand stages: rows of kFzRowBytes of K (= swizzle span).  64-byte rows x 4 stages hold
// 128 KB in flight (3 stages of prefetch ~ 2400 MMA cycles, enough to cover the TMA latency) and
// leave room for the epilogue's transform ring; 128-byte rows x 3 stages (the stand-alone GEMM's
// choice) would fill shared memory.
constexpr int kFzRowBytes = SMPLK_FZ_ROW_BYTES;
#ifndef SMPLK_FZ_STAGES
#define SMPLK_FZ_STAGES (SMPLK_FZ_ROW_BYTES == 64 ? (SMPLK_FZ_EPI_WARPS == 12 ? 3 : 4) : 2)
#endif
constexpr int kFzStages = SMPLK_FZ_STAGES;
constexpr int kFzKSteps = kFzRowBytes / 32;                 // UMMA k-steps (16 fp16 = 32 B) per k-block
constexpr int kFzElemsPerBlock = kFzRowBytes / 2;           // fp16 elements per k-block
constexpr int kFzTileABytes = kBlendBM * kFzRowBytes;       // this CTA's 128 feature rows
constexpr int kFzTileBBytes = (kBlendBN / 2) * kFzRowBytes; // this CTA's half of the posedirs tile
constexpr int kFzStageBytes = 2 * kFzTileABytes + 2 * kFzTileBBytes;
constexpr int kFzRing = SMPLK_FZ_RING;
constexpr int kFzRingEntryBytes = 13 * 128;                 // 12 transform lines + 1 weight line per warp
constexpr int kFzEpiWarps = SMPLK_FZ_EPI_WARPS;
constexpr int kFzParts = kFzEpiWarps / 4;           // epilogue warps per TMEM lane quarter
static_assert(kFzEpiWarps == 8 || kFzEpiWarps == 12, "two or three epilogue warps per TMEM lane quarter");
constexpr int kFzThreads = 64 + kFzEpiWarps * 32;   // producer warp, MMA warp, kFzParts x 4 epilogue warps
constexpr int kFzTileVerts = 84;                    // whole vertices per 256-column tile
constexpr int kFzTileCols = 3 * kFzTileVerts;       // 252 output coordinates per tile (+ 4 zero columns)
constexpr int kFzChunkVerts = 12;
constexpr int kFzChunkCols = 36;
constexpr int kFzChunks = 7;                        // 7 x 12 vertices
constexpr int kFzStageStride = 36;                  // words per staged row of the 32-column window: 16-byte
                                                    // aligned rows, conflict-free for STS.128 rows and LDS columns
constexpr int kFzStageWords = 32 * kFzStageStride;
constexpr int kFzBiasBytes = 4 * kFzChunkCols * 4;          // v_template of a warp's (<= 4) chunks
constexpr int kFzSmemBytes = kFzStages * kFzStageBytes + kFzEpiWarps * kFzStageWords * 4 +
                             kFzEpiWarps * kFzRing * kFzRingEntryBytes + kFzEpiWarps * kFzBiasBytes + 256;
constexpr int kFzSmemAlloc = kFzSmemBytes + 1024;
static_assert(kFzSmemAlloc <= 232448, "fused kernel shared memory exceeds the sm_100 limit");

struct FusedArgs {
  int num_m_blocks;        // 256-body blocks
  int num_n_blocks;        // 84-vertex tiles
  int num_k_blocks;
  int k_elems;
  float out_scale;         // 1 / pd_scale
  const float* bias;       // [num_n_blocks * 256 + 64] v_template in the fused column layout
  const int* ch_off;       // [num_n_blocks * 7 + 1] first entry of every 12-vertex chunk
  const int* ch_joint;     // [entries] joint id
  const float4* ch_w;      // [entries][4] weight of that joint for the chunk's 12 vertices (+ 4 zeros)
  const float* At;         // [ceil(rows/256)*2][J][128][12] joint-major transforms per 128-body block (transl folded in)
  int J;
  float* out;              // (rows, N) posed vertices
  int rows;
  int N;                   // 3 V
  int out_odd_shift;       // TMA output: 2 when the odd bodies' rows start 8 bytes off a 16-byte boundary (V = 2 mod 4), else 0
  long long* dbg;          // tuning aid (SMPLK_FZ_TIMELINE builds): per-tile clock64 stamps of CTA 0
};
constexpr int kFzDbgTiles = 32;
#ifndef SMPLK_FZ_TIMELINE
#define SMPLK_FZ_TIMELINE 0
#endif
#if SMPLK_FZ_TIMELINE
#define FZ_STAMP(cond, ptr_expr) do { if (cond) *(ptr_expr) = clock64(); } while (0)
#else
#define FZ_STAMP(cond, ptr_expr) do { } while (0)
#endif

// A [rows][J][12] -> At [ceil(rows/128)][J][128][12] (+ transl on the translation column): joint-major
// within every 128-body block, so that the 32 bodies of a fused-kernel warp are 1,536 contiguous bytes
// per joint.  Only used with the warp-per-body pose kernel; the block pose kernel writes At itself.
__global__ void __launch_bounds__(256)
transpose_transforms_kernel(int rows, int rows_pad, int J, const float* __restrict__ A,
                            const float* __restrict__ transl, float* __restrict__ At) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // one float4 (transform row) each
  if (i >= rows_pad * J * 3) return;
  const int r = i % 3, j = (i / 3) % J, b = i / (3 * J);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b < rows) {
    v = *reinterpret_cast<const float4*>(A + ((size_t)b * J + j) * 12 + 4 * r);
    if (transl != nullptr) v.w += transl[3 * b + r];
  }
  *reinterpret_cast<float4*>(At + (((size_t)(b >> 7) * J + j) * 128 + (b & 127)) * 12 + 4 * r) = v;
}

// kN > 0: floats per output row known at compile time (3 * 6890 for SMPL / SMPL-H), so the 32 row
// addresses of a window store are immediates; kN == 0: row pitch from args.N.
// 14 warps: __launch_bounds__(448) gives 128 registers per thread; a __maxnreg__ of 136 or 144 compiles
// without spills but the launch is refused (too many resources), see above.
#ifndef SMPLK_FZ_MAXNREG
#define SMPLK_FZ_MAXNREG 0
#endif
#if SMPLK_FZ_EPI_WARPS == 12 && SMPLK_FZ_MAXNREG > 0
#define SMPLK_FZ_BOUNDS __maxnreg__(SMPLK_FZ_MAXNREG)
#else
#define SMPLK_FZ_BOUNDS __launch_bounds__(kFzThreads, 1)
#endif
// kTmaOut: results leave as TMA tensor stores, one 16-row x 36-column box per chunk for the even and one
// for the odd bodies of the warp.  A (B, V, 3) fp32 row is 12 V bytes -- 82,680 = 8 mod 16 for V = 6,890 -- so
// no tensor map can have ONE body per row (strides must be multiples of 16 bytes); TWO bodies per row (24 V
// bytes, V even) can: `tmap_out_even` sees the even bodies (inner extent 3 V, so nothing spills into the odd
// body), `tmap_out_odd` / `tmap_out_odd32` the same rows with inner extent 6 V, addressed at column 3 V + c.
// Chunks cut by the end of the row take the per-lane store path.
template <int kN, bool kTmaOut>
__global__ void __cluster_dims__(2, 1, 1) SMPLK_FZ_BOUNDS
blend_skin_fused_kernel(const __grid_constant__ CUtensorMap tmap_f_hi,
                        const __grid_constant__ CUtensorMap tmap_f_lo,
                        const __grid_constant__ CUtensorMap tmap_pd_hi,
                        const __grid_constant__ CUtensorMap tmap_pd_lo,
                        const __grid_constant__ CUtensorMap tmap_out_even,
                        const __grid_constant__ CUtensorMap tmap_out_odd,
                        const __grid_constant__ CUtensorMap tmap_out_odd32, const FusedArgs args) {
  extern __shared__ uint8_t fz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(fz_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  float* epi_base = reinterpret_cast<float*>(smem + kFzStages * kFzStageBytes);
  uint8_t* ring_base = smem + kFzStages * kFzStageBytes + kFzEpiWarps * kFzStageWords * 4;
  uint8_t* bias_base = ring_base + kFzEpiWarps * kFzRing * kFzRingEntryBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_base + kFzEpiWarps * kFzBiasBytes);
  uint64_t* full_bar = bars;                        // [kFzStages]   (used in the leader)
  uint64_t* empty_bar = bars + kFzStages;           // [kFzStages]
  uint64_t* tmem_full = bars + 2 * kFzStages;       // [2]
  uint64_t* tmem_empty = bars + 2 * kFzStages + 2;  // [2]          (used in the leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kFzStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int outN = kN > 0 ? kN : args.N;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int num_tiles = args.num_m_blocks * args.num_n_blocks;
  constexpr int kUmmaK = 16;

  ptx::pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_f_hi);
    ptx::prefetch_tmap(&tmap_f_lo);
    ptx::prefetch_tmap(&tmap_pd_hi);
    ptx::prefetch_tmap(&tmap_pd_lo);
    if (kTmaOut) {
      ptx::prefetch_tmap(&tmap_out_even);
      ptx::prefetch_tmap(&tmap_out_odd);
      ptx::prefetch_tmap(&tmap_out_odd32);
    }
    for (int s = 0; s < kFzStages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 2 * 32 * kFzEpiWarps);   // every epilogue thread, in both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta<kTmemCols>(tmem_ptr);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // barriers, TMEM and the descriptor prefetches above overlap the pose kernel's tail (programmatic dependent launch);
  // its feature rows and transforms are read from here on
  ptx::pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int mb = tile % args.num_m_blocks;
        const int nb = tile / args.num_m_blocks;
        const int m0 = mb * 2 * kBlendBM + (int)rank * kBlendBM;
        const int nb0 = nb * kBlendBN + (int)rank * (kBlendBN / 2);   // this CTA's half of the B tile
        for (int kb = 0; kb < args.num_k_blocks; ++kb) {
          ptx::mbar_wait_backoff(&empty_bar[stage], phase ^ 1, SMPLK_FZ_BACKOFF_NS);
          uint8_t* st = stage_base + stage * kFzStageBytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kFzStageBytes);
          else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
          const int k0 = kb * kFzElemsPerBlock;
          ptx::tma_load_2d_2sm(st, &tmap_f_hi, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + kFzTileABytes, &tmap_f_lo, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + 2 * kFzTileABytes, &tmap_pd_hi, &full_bar[stage], k0, nb0);
          ptx::tma_load_2d_2sm(st + 2 * kFzTileABytes + kFzTileBBytes, &tmap_pd_lo, &full_bar[stage], k0, nb0);
          if (++stage == kFzStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * kBlendBM, kBlendBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        [[maybe_unused]] const int dbg_it = (tile - cluster_id) / num_clusters;
        FZ_STAMP(args.dbg && blockIdx.x == 0 && dbg_it < kFzDbgTiles, &args.dbg[(1 * kFzDbgTiles + dbg_it) * 4 + 0]);
        ptx::mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1, SMPLK_FZ_BACKOFF_NS);
        ptx::tcgen05_fence_after();
        FZ_STAMP(args.dbg && blockIdx.x == 0 && dbg_it < kFzDbgTiles, &args.dbg[(1 * kFzDbgTiles + dbg_it) * 4 + 1]);
        const uint32_t d_tmem = tmem_base + acc * kBlendBN;
        for (int kb = 0; kb < args.num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(stage_base + stage * kFzStageBytes);
          const uint64_t a_hi = ptx::make_kmajor_desc<kFzRowBytes>(st);
          const uint64_t a_lo = ptx::make_kmajor_desc<kFzRowBytes>(st + kFzTileABytes);
          const uint64_t b_hi = ptx::make_kmajor_desc<kFzRowBytes>(st + 2 * kFzTileABytes);
          const uint64_t b_lo = ptx::make_kmajor_desc<kFzRowBytes>(st + 2 * kFzTileABytes + kFzTileBBytes);
          const int ksteps = min(kFzKSteps, (args.k_elems - kb * kFzElemsPerBlock) / kUmmaK);
#pragma unroll
          for (int k = 0; k < kFzKSteps; ++k) {
            if (k < ksteps) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);
              const uint32_t first = (kb != 0 || k != 0) ? 1u : 0u;
              ptx::umma_2cta<true>(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
              ptx::umma_2cta<true>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              ptx::umma_2cta<true>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          ptx::umma_commit_2cta(&empty_bar[stage]);
          if (++stage == kFzStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2cta(&tmem_full[acc]);
        FZ_STAMP(args.dbg && blockIdx.x == 0 && dbg_it < kFzDbgTiles, &args.dbg[(1 * kFzDbgTiles + dbg_it) * 4 + 2]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== skinning epilogue (warps 2..9) =====================
    const int part = (warp - 2) >> 2;         // which part of every tile's chunk range this warp takes
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const uint32_t stage_u32 = ptx::smem_u32(epi_base + (warp - 2) * kFzStageWords);
    // this body's staged row; TMA output: even bodies in rows 0..15, odd bodies in rows 16..31 (one box each)
    const uint32_t stage_row = stage_u32 + (kTmaOut ? ((lane & 1) * 16 + (lane >> 1)) : lane) * (kFzStageStride * 4);
    [[maybe_unused]] const uint32_t stage_odd = stage_u32 + 16 * kFzStageStride * 4;    // TMA output: the odd bodies' rows
    [[maybe_unused]] float cy0 = 0.f, cy1 = 0.f;                                        // ... and their two carried columns
    const uint32_t stage_col = stage_u32 + lane * 4;                      // this lane's staged column
    const uint32_t ring_warp = ptx::smem_u32(ring_base) + (warp - 2) * (kFzRing * kFzRingEntryBytes);
    const uint32_t ring_piece = ring_warp + lane * 16;    // this lane's 16-byte piece of lines l/8 + 4k
    const uint32_t bias_warp = ptx::smem_u32(bias_base) + (warp - 2) * kFzBiasBytes;
    int ps = 0, cs = 0;                       // ring slots: next to fill / next to apply
    const float oscale = args.out_scale;
    const int JC128 = args.J * 12 * 128;
    const float* wflat = reinterpret_cast<const float*>(args.ch_w);
    constexpr unsigned kFull = 0xffffffffu;
    for (int it = 0;; ++it) {
      const int tile = cluster_id + it * num_clusters;
      if (tile >= num_tiles) break;
      const int mb = tile % args.num_m_blocks;
      const int nb = tile / args.num_m_blocks;
      const int m0 = mb * 2 * kBlendBM + (int)rank * kBlendBM;
      const int acc = it & 1;                   // TMEM accumulator buffer of this tile
      const uint32_t acc_phase = (it >> 1) & 1;
      // All warps of a TMEM lane quarter work on EVERY tile, each on a contiguous range of its 7 chunks
      // (3/4 for two warps, 2/2/3 for three; the longer share rotates with the tile), so an accumulator is
      // held for a fraction of an epilogue: with 2 accumulators the per-tile time is
      // max(T_mma, (T_mma + T_hold) / 2).
      int c_lo = 0, c_hi = 0;
      {
        constexpr int base = kFzChunks / kFzParts, rem = kFzChunks % kFzParts;
        int acc_c = 0;
#pragma unroll
        for (int pp = 0; pp < kFzParts; ++pp) {
          const int sz = base + ((((pp + it) % kFzParts) < rem) ? 1 : 0);
          if (pp == part) { c_lo = acc_c; c_hi = acc_c + sz; }
          acc_c += sz;
        }
      }
      // joint-major per 128-body block: a joint's transforms of this warp's 32 bodies are 96 contiguous
      // 16-byte pieces; lane l copies pieces l, l + 32, l + 64
      const float* At_w = args.At + (size_t)(m0 >> 7) * JC128 + q * 32 * 12 + lane * 4;
      const float* bias_t = args.bias + nb * kBlendBN;
      const int row0 = m0 + q * 32;
      const int nrows = min(32, args.rows - row0);
      float* out_t = args.out + (size_t)row0 * outN + (size_t)nb * kFzTileCols;
      const int cols_left = outN - nb * kFzTileCols;     // valid output columns from this tile on
      // chunk offsets (lane i <- off[i], i <= 7), first window of joint ids, first bias lines: all
      // independent of the accumulator, so issue them before waiting for the MMAs
      const int ol = __ldg(args.ch_off + nb * kFzChunks + min(lane, kFzChunks));
      int e = __shfl_sync(kFull, ol, c_lo);
      const int e_end = __shfl_sync(kFull, ol, c_hi);
      int wb = e;                                          // base of the joint-id window
      int jl = __ldg(args.ch_joint + wb + lane);
      // v_template of this warp's chunks -> smem (first cp.async group of the tile)
      {
        const int npieces = (c_hi - c_lo) * (kFzChunkCols / 4);
        for (int i = lane; i < npieces; i += 32)
          ptx::cp_async_16(bias_warp + i * 16, bias_t + c_lo * kFzChunkCols + i * 4);
        ptx::cp_async_commit();
      }
      // packed pairs: p2[3k + d] = coordinate d of vertices (2k, 2k+1) of the chunk (the tile's columns
      // are packed in that order), o2 likewise -> the skinning math runs on FFMA2 (fma.rn.f32x2)
      uint64_t p2[kFzChunkCols / 2], o2[kFzChunkCols / 2];

      // (chunk, joint) entry ee -> ring slot: the joint's 12 transform lines of the warp's 32 bodies
      // (128 bytes each, lane = body) + the entry's 64-byte weight line, as 16-byte cp.async pieces
      // (3 per lane + 4 lanes for the weights).  No register scoreboard is held while the loads fly
      // (ptxas tracks ALL global loads of a warp with one scoreboard, so register prefetching cannot
      // run more than one entry ahead; cp.async groups can).
      auto issue_entry = [&](int ee) {
        if (ee < e_end) {
          if (ee >= wb + 32) {
            wb += 32;
            jl = __ldg(args.ch_joint + wb + lane);
          }
          const int jj = __shfl_sync(kFull, jl, ee - wb);
          const float* ap = At_w + (size_t)jj * (12 * 128);
          const uint32_t dst = ring_piece + ps * kFzRingEntryBytes;
#pragma unroll
          for (int k = 0; k < 3; ++k) ptx::cp_async_16(dst + k * 512, ap + k * 128);
          if (lane < 4) ptx::cp_async_16(dst + 12 * 128, wflat + (size_t)ee * 16 + lane * 4);
        }
        ptx::cp_async_commit();                 // one group per entry, empty past the tile's end
        ps = (ps + 1 == kFzRing) ? 0 : ps + 1;
      };
      // Apply entry e (slot cs) and refill the slot of entry e-1 with entry e + kFzRing - 1.
      // Branch-free: a vertex that does not use the joint has w = 0 and adds exactly 0 (p and the
      // transforms are finite, padding rows / columns included).
      auto apply_entry = [&](int ee) {
        ptx::cp_async_wait<kFzRing - 2>();      // this lane's pieces of entry e have landed ...
        __syncwarp();                           // ... and so have everyone else's; slot e-1 is drained
        issue_entry(ee + kFzRing - 1);
        const uint32_t src = ring_warp + cs * kFzRingEntryBytes;
        cs = (cs + 1 == kFzRing) ? 0 : cs + 1;
        float a[12];                              // this body's 3x4 transform: 48 contiguous bytes
#pragma unroll
        for (int i = 0; i < 3; ++i)
          ptx::ld_shared_v4(src + lane * 48 + i * 16, a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
        float wv[kFzChunkVerts];
#pragma unroll
        for (int i = 0; i < kFzChunkVerts / 4; ++i)
          ptx::ld_shared_v4(src + 12 * 128 + i * 16, wv[4 * i], wv[4 * i + 1], wv[4 * i + 2], wv[4 * i + 3]);
        // two vertices per instruction: the transform components are scalar (broadcast) operands
        uint64_t a2[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) a2[i] = ptx::pack_f32x2(a[i], a[i]);
#pragma unroll
        for (int k = 0; k < kFzChunkVerts / 2; ++k) {
          const uint64_t w2 = ptx::pack_f32x2(wv[2 * k], wv[2 * k + 1]);
          const uint64_t x = p2[3 * k], y = p2[3 * k + 1], z = p2[3 * k + 2];
          const uint64_t qx = ptx::fma_f32x2(a2[0], x, ptx::fma_f32x2(a2[1], y, ptx::fma_f32x2(a2[2], z, a2[3])));
          const uint64_t qy = ptx::fma_f32x2(a2[4], x, ptx::fma_f32x2(a2[5], y, ptx::fma_f32x2(a2[6], z, a2[7])));
          const uint64_t qz = ptx::fma_f32x2(a2[8], x, ptx::fma_f32x2(a2[9], y, ptx::fma_f32x2(a2[10], z, a2[11])));
          o2[3 * k] = ptx::fma_f32x2(w2, qx, o2[3 * k]);
          o2[3 * k + 1] = ptx::fma_f32x2(w2, qy, o2[3 * k + 1]);
          o2[3 * k + 2] = ptx::fma_f32x2(w2, qz, o2[3 * k + 2]);
        }
      };
      // 32 staged columns (one 128-byte segment per body row) -> global
      auto store_window = [&](int wnd, int lane_lo, int lane_hi) {
        const int col = 32 * wnd + lane;
        if (lane >= lane_lo && lane < lane_hi && col < cols_left) {
          float* dst = out_t + col;
          if (nrows == 32) {
            // batches of 8: the shared loads are volatile asm, so without explicit batching every
            // store would wait for its own load (load latency x 32 per window)
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = ptx::ld_shared_f32(stage_col + (r0 + j) * (kFzStageStride * 4));
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[(r0 + j) * outN] = v[j];
            }
          } else {
            for (int r = 0; r < nrows; ++r)
              dst[r * outN] = ptx::ld_shared_f32(stage_col + r * (kFzStageStride * 4));
          }
        }
      };

      ps = 0;                                   // every group of the previous tile has been consumed
      cs = 0;
#pragma unroll
      for (int i = 0; i < kFzRing - 1; ++i) issue_entry(e + i);
      ptx::cp_async_wait<kFzRing - 1>();        // the v_template group (older than the primed entries)
      __syncwarp();
      [[maybe_unused]] const bool dbg_on = args.dbg && blockIdx.x == 0 && lane == 0 && it < kFzDbgTiles;
      FZ_STAMP(dbg_on, &args.dbg[(warp * kFzDbgTiles + it) * 4 + 0]);
      ptx::mbar_wait_backoff(&tmem_full[acc], acc_phase, SMPLK_FZ_EPI_BACKOFF_NS);
      ptx::tcgen05_fence_after();
      FZ_STAMP(dbg_on, &args.dbg[(warp * kFzDbgTiles + it) * 4 + 1]);
#if SMPLK_FZ_PROBE_NO_EPILOGUE
      // measurement probe (never in the product build): hand the accumulator straight back, so the kernel's time is
      // the TMA + MMA pipeline of THIS configuration alone (64-byte operand rows x 4 stages); the output is garbage
      ptx::cp_async_wait<0>();
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
      continue;
#endif

#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        [[maybe_unused]] const bool dbg_c = dbg_on && (warp & 3) == 2 && it < 4;
        [[maybe_unused]] long long* dbg_p = args.dbg ? args.dbg + (0 * kFzDbgTiles + it * 7 + c) * 4 : nullptr;
        FZ_STAMP(dbg_c, &dbg_p[0]);
        // ---- TMEM accumulator columns of chunk c -> p (+ v_template), o = 0
        {
          uint32_t pr[kFzChunkCols];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(acc * kBlendBN + c * kFzChunkCols);
          ptx::tmem_ld_32x32b_x16(taddr, pr);
          ptx::tmem_ld_32x32b_x16(taddr + 16, pr + 16);
          ptx::tmem_ld_32x32b_x4(taddr + 32, pr + 32);
          ptx::tmem_ld_wait();
          if (c == c_hi - 1) {                  // this warp's last read of the accumulator
            ptx::tcgen05_fence_before();
            ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
            FZ_STAMP(dbg_on, &args.dbg[(warp * kFzDbgTiles + it) * 4 + 2]);
          }
          const uint32_t bsrc = bias_warp + (c - c_lo) * (kFzChunkCols * 4);
          const uint64_t os2 = ptx::pack_f32x2(oscale, oscale);
#pragma unroll
          for (int i = 0; i < kFzChunkCols; i += 4) {
            float b0, b1, b2, b3;
            ptx::ld_shared_v4(bsrc + i * 4, b0, b1, b2, b3);      // warp-uniform address: broadcast
            p2[i / 2] = ptx::fma_f32x2(ptx::pack_b32x2(pr[i], pr[i + 1]), os2, ptx::pack_f32x2(b0, b1));
            p2[i / 2 + 1] = ptx::fma_f32x2(ptx::pack_b32x2(pr[i + 2], pr[i + 3]), os2, ptx::pack_f32x2(b2, b3));
            o2[i / 2] = o2[i / 2 + 1] = 0ull;
          }
        }
        // ---- the chunk's (joint, weights) entries: entry e is applied from its ring slot while
        // entries e+1 .. e+kFzRing-1 -- of this chunk or the next -- are in flight
        const int cend = __shfl_sync(kFull, ol, c + 1);

        for (; e < cend; ++e) apply_entry(e);
        FZ_STAMP(dbg_c, &dbg_p[1]);
        // ---- o (32 bodies x 36 columns) -> rolling 32-column staging window -> global.  Chunk c
        // starts at window column w0 = 36 c mod 32; its first 32 - w0 columns complete the window.
        // unpack: output column 6k + 3h + d of the chunk = half h of o2[3k + d]
        float o[kFzChunkCols];
#pragma unroll
        for (int k = 0; k < kFzChunkVerts / 2; ++k)
#pragma unroll
          for (int d = 0; d < 3; ++d) ptx::unpack_f32x2(o2[3 * k + d], o[6 * k + d], o[6 * k + 3 + d]);
        if constexpr (kTmaOut) {
          const int cbeg = c * kFzChunkCols;
          // odd: this lane's body starts 8 bytes off a 16-byte boundary (V = 2 mod 4; for V = 0 mod 4 every row is aligned)
          const bool odd = (lane & 1) != 0 && args.out_odd_shift != 0;
          if (cbeg + kFzChunkCols <= cols_left) {
            // Whole chunk inside the row.  A TMA store must START on a 16-byte boundary (measured:
            // tools/micro/tma_store_probe.cu, illegal instruction otherwise) and an odd body's row starts 8 bytes
            // off one, so the odd bodies' boxes sit two columns to the left of the chunk: [36 c - 2, 36 c + 34),
            // the two leading columns carried over from the previous chunk.  The first chunk of a warp's range has
            // no predecessor: its columns 0, 1 are stored by the lanes, columns 2..33 as a 32-column box
            // (128-byte rows, 128B swizzle); the last carry of the range is stored by the lanes as well.
            if (lane == 0) ptx::tma_store_wait_read<0>();     // the buffer's previous boxes were issued a chunk ago
            __syncwarp();
            const bool first = c == c_lo && args.out_odd_shift != 0;
            if (first) {
              if (!odd) {
#pragma unroll
                for (int g = 0; g < kFzChunkCols / 4; ++g)
                  ptx::st_shared_v4(stage_row + g * 16, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
              } else {
                if (lane < nrows) *reinterpret_cast<float2*>(out_t + (size_t)lane * outN + cbeg) = make_float2(o[0], o[1]);
                const uint32_t row32 = stage_odd + (lane >> 1) * 128;
                const uint32_t sw = (row32 >> 7) & 7;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  ptx::st_shared_v4(row32 + ((g ^ sw) << 4), o[4 * g + 2], o[4 * g + 3], o[4 * g + 4], o[4 * g + 5]);
              }
            } else {
              float s0 = odd ? cy0 : o[0], s1 = odd ? cy1 : o[1], s2 = odd ? o[0] : o[2], s3 = odd ? o[1] : o[3];
              ptx::st_shared_v4(stage_row, s0, s1, s2, s3);
#pragma unroll
              for (int g = 1; g < kFzChunkCols / 4; ++g) {
                s0 = odd ? o[4 * g - 2] : o[4 * g]; s1 = odd ? o[4 * g - 1] : o[4 * g + 1];
                s2 = odd ? o[4 * g] : o[4 * g + 2]; s3 = odd ? o[4 * g + 1] : o[4 * g + 3];
                ptx::st_shared_v4(stage_row + g * 16, s0, s1, s2, s3);
              }
            }
            cy0 = o[kFzChunkCols - 2]; cy1 = o[kFzChunkCols - 1];
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int col = nb * kFzTileCols + cbeg;
              ptx::tma_store_2d(&tmap_out_even, epi_base + (warp - 2) * kFzStageWords, col, row0 >> 1);
              if (first) ptx::tma_store_2d(&tmap_out_odd32, epi_base + (warp - 2) * kFzStageWords + 16 * kFzStageStride,
                                           outN + col + 2, row0 >> 1);
              else ptx::tma_store_2d(&tmap_out_odd, epi_base + (warp - 2) * kFzStageWords + 16 * kFzStageStride,
                                     outN + col - args.out_odd_shift, row0 >> 1);
              ptx::tma_store_commit();
            }
            // the range's (or the row's) last full chunk: its two trailing columns have no box to ride in
            if ((c == c_hi - 1 || cbeg + 2 * kFzChunkCols > cols_left) && odd && lane < nrows)
              *reinterpret_cast<float2*>(out_t + (size_t)lane * outN + cbeg + kFzChunkCols - 2) = make_float2(cy0, cy1);
          } else if (cbeg < cols_left && lane < nrows) {
            // the row's last vertices (one chunk of the last tile): straight from the registers
            float* dst = out_t + (size_t)lane * outN + cbeg;
#pragma unroll
            for (int i = 0; i < kFzChunkCols; ++i)
              if (cbeg + i < cols_left) dst[i] = o[i];
          }
          FZ_STAMP(dbg_c, &dbg_p[2]);
          FZ_STAMP(dbg_c, &dbg_p[3]);
        } else {
          const int w0 = (c * kFzChunkCols) & 31;
          const int g0 = w0 >> 2;                  // window position of the chunk's first 4-column group
          __syncwarp();
#pragma unroll
          for (int g = 0; g < kFzChunkCols / 4; ++g)
            if (g0 + g < 8) ptx::st_shared_v4(stage_row + (g0 + g) * 16, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
          __syncwarp();
          FZ_STAMP(dbg_c, &dbg_p[2]);
          store_window((c * kFzChunkCols) >> 5, c == c_lo ? w0 : 0, 32);   // columns below w0 of the range's first window are the other warp's
          FZ_STAMP(dbg_c, &dbg_p[3]);
          __syncwarp();
#pragma unroll
          for (int g = 1; g < kFzChunkCols / 4; ++g)
            if (g0 + g >= 8) ptx::st_shared_v4(stage_row + (g0 + g - 8) * 16, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
          if (c == c_hi - 1) {                     // the last window of this warp's range is partial
            __syncwarp();
            store_window(((c + 1) * kFzChunkCols) >> 5, 0, ((c + 1) * kFzChunkCols) & 31);
          }
        }
      }
      FZ_STAMP(dbg_on, &args.dbg[(warp * kFzDbgTiles + it) * 4 + 3]);
    }
  }

  if (kTmaOut && warp >= 2 && lane == 0) ptx::tma_store_wait<0>();
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

}  // namespace smplk
