// Rigged-mesh replay (LBS only: fixed template, fixed joints, any vertex count) as ONE tensor-core GEMM.
//
//   verts[f, v, c] = sum_j w[v, j] (R_j[f] v_t[v] + t_j[f])[c] + transl[f, c]
//                  = sum_{j, k} ( w[v, j] [v_t[v]; 1]_k ) * A[f, j, c, k]  + transl[f, c]
//                        P[v, 4 j + k]  (model constant)     T[3 f + c, 4 j + k]  (the frame's transforms)
//
// Reference math: lib/model2video.py:55-85 = lib/mesh2smpl_model.py:268-313 (RecoverModel: T = W . A, v = T [v_t; 1],
// + trans), models/smplh_np.py:79-82.  With the template fixed the blend of transforms and its application to the
// vertex are one bilinear form whose vertex factor P is constant, so the whole replay is verts = P . T^T with
// K = 4 J (96 for the 24-joint rig) -- 60 CUDA-core FMAs per vertex and frame (the streaming kernel: issue-bound at
// 0.36-0.51 of HBM) become 1,728 tensor flops (three fp16 two-term passes, fp32-equivalent accuracy as in the
// blend GEMM: P_hi T_hi + P_lo T_hi + P_hi T_lo), 0.57 of the tensor pipe at the HBM write rate.
//
// Orientation: D[M = vertex, N = (frame, c)].  A TMEM lane is a vertex, so a thread's three consecutive accumulator
// columns are the (x, y, z) of ITS vertex in one frame: 12 contiguous bytes of the (F, V, 3) output, and the 32 lanes
// of a warp write 384 contiguous bytes per frame straight from registers -- no shared-memory transposition, no
// barrier in the epilogue.  (tools/micro/store_pattern_probe.cu: this store pattern alone runs at 4.4 / 5.4 TB/s at
// 6,890 / 50,000 vertices, within 3-6 % of fully coalesced 4-byte stores.)
//
// Mapping: CTA pairs (cta_group::2), 256 vertices x 240 columns (80 frames) per tile, two TMEM accumulators;
// operands in 64-byte K rows (32 fp16, 64B swizzle) through a 6-stage TMA ring that runs across tiles; one
// MMA-issuing thread in the leader; 8 epilogue warps (TMEM lane quarter x half of the tile's frames).
// Tile order: replay_tile_coords below.
//
// kSkin = true: the SAME kernel as the skinning pass of the two-kernel forward (blend GEMM -> v_posed -> skinning;
// every forward that keeps v_posed for a backward).  v_posed differs per body, so only the BLEND of the transforms
// is a GEMM: Tm[v, (b, 4 c + k)] = sum_j w[v, j] A[b, j, c, k] (K = J, 12 accumulator columns per body, 20 bodies
// per tile); the epilogue thread -- still one vertex -- reads its 12 bytes of v_posed[b], applies Tm (12 FMAs
// instead of the streaming kernel's 60 and no shared-memory transform gathers) and stores its 12 bytes.
#pragma once
#include "blend_gemm_2cta.cuh"

namespace smplk {

constexpr int kRpRowBytes = 64;                       // K bytes per operand row and k-block (= swizzle span)
constexpr int kRpKB = kRpRowBytes / 2;                // fp16 elements per k-block
constexpr int kRpTileFrames = 80;
constexpr int kRpBN = 3 * kRpTileFrames;              // 240 accumulator columns per tile (UMMA N)
constexpr int kRpStages = 6;
constexpr int kRpTileABytes = kBlendBM * kRpRowBytes;         // this CTA's 128 vertex rows: 8 KB
constexpr int kRpTileBBytes = (kRpBN / 2) * kRpRowBytes;      // this CTA's half of the frame rows: 7,680 B
constexpr int kRpStageBytes = 2 * kRpTileABytes + 2 * kRpTileBBytes;
constexpr int kRpEpiWarps = 8;
constexpr int kRpThreads = 64 + 32 * kRpEpiWarps;
constexpr int kRpSmemBytes = kRpStages * kRpStageBytes + 256;
constexpr int kRpSmemAlloc = kRpSmemBytes + 1024;
static_assert(kRpStageBytes % 1024 == 0, "stage alignment");
static_assert(kRpSmemAlloc <= 232448, "replay kernel shared memory exceeds the sm_100 limit");
static_assert(kRpBN % 16 == 0 && kRpBN <= 256, "UMMA N of a CTA pair");
static_assert((kRpBN / 2) % 16 == 8, "the epilogue reads its 120 columns as 7 x 16 + 8");
static_assert(kRpBN / 2 == 120 && (kRpBN / 12) % 4 == 0, "kSkin: two rounds of 5 bodies per warp");

constexpr int kRpTileBodies = kRpBN / 12;             // kSkin: 20 bodies per tile

struct ReplayArgs {
  int V, F;                // vertices, frames (kSkin: bodies) of this launch
  int num_m_blocks;        // 256-vertex blocks
  int num_n_blocks;        // 80-frame tiles
  int num_k_blocks;        // ceil(4 J / 32)
  float out_scale;         // 1 / (power-of-two scale of P)
  const float* transl;     // (F, 3) or null
  float* out;              // (F, V, 3)
  const float* vsrc;       // kSkin: v_posed rows (F, vsrc_stride), 8-byte aligned, even stride
  size_t vsrc_stride;
};

// kSkin operand: A (B, J, 12) fp32 -> T_hi / T_lo (12 B, Kp) fp16, row 12 b + r = [A[b, j, r] for j] (zero padded).
// One block per body: its J x 12 floats are read coalesced into shared memory and leave as 8-byte pieces of four j.
constexpr int kSkinOpThreads = 192;
__global__ void __launch_bounds__(kSkinOpThreads)
skin_operand_kernel(int B, int J, int Kp, const float* __restrict__ A, __half* __restrict__ T_hi,
                    __half* __restrict__ T_lo) {
  __shared__ float sa[kMaxJoints * 12 + 48];
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int b = blockIdx.x;
  const float* src = A + (size_t)b * J * 12;
  for (int i = threadIdx.x; i < Kp * 12; i += kSkinOpThreads) sa[i] = i < J * 12 ? src[i] : 0.f;
  __syncthreads();
  const int q4 = Kp / 4;
  for (int i = threadIdx.x; i < 12 * q4; i += kSkinOpThreads) {
    const int r = i / q4, j = 4 * (i % q4);
    float x[4];
    __half h[4], l[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      x[n] = sa[(j + n) * 12 + r];
      h[n] = __float2half_rn(x[n]);
      l[n] = __float2half_rn(x[n] - __half2float(h[n]));
    }
    const size_t o = ((size_t)b * 12 + r) * Kp + j;
    __half2* ph = reinterpret_cast<__half2*>(T_hi + o);
    __half2* pl = reinterpret_cast<__half2*>(T_lo + o);
    ph[0] = __halves2half2(h[0], h[1]); ph[1] = __halves2half2(h[2], h[3]);
    pl[0] = __halves2half2(l[0], l[1]); pl[1] = __halves2half2(l[2], l[3]);
  }
}

// Tile order.  Vertex blocks are taken in groups of `group` (= the number of clusters): within a group the vertex
// block runs fastest, then the frame tile.  The clusters running at one time therefore write a few whole frames'
// rows of the group's vertices (DRAM locality of the write stream), and a cluster meets the same vertex block again
// one round later, so a group's P tiles (98 KB each) stay in L2 while its frames stream by -- with the vertex block
// running over ALL blocks (one group), a 200,000-vertex mesh re-reads its 77 MB of P from L2 / HBM for every frame
// tile (measured 3.96 TB/s against 4.79 at 50,000 vertices).
__device__ __forceinline__ void replay_tile_coords(int tile, int num_mb, int num_nb, int group, int& mb, int& nb) {
  const int per_group = group * num_nb;
  const int g = tile / per_group;
  const int rem = tile - g * per_group;
  const int gsize = min(group, num_mb - g * group);
  nb = rem / gsize;
  mb = g * group + (rem - nb * gsize);
}

// A (F, J, 3, 4) fp32 -> T_hi / T_lo (3 F, Kp) fp16, row 3 f + c = [A[f, j, c, 0..3] for j] (zero padded to Kp).
__global__ void __launch_bounds__(256)
replay_operand_kernel(int F, int J, int Kp, const float* __restrict__ A, __half* __restrict__ T_hi,
                      __half* __restrict__ T_lo) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int q4 = Kp / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;       // one 4-element group of one row
  if (i >= (long)3 * F * q4) return;
  const int j = (int)(i % q4);
  const long row = i / q4;
  const int c = (int)(row % 3);
  const long f = row / 3;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < J) x = *reinterpret_cast<const float4*>(A + ((size_t)f * J + j) * 12 + 4 * c);
  const __half h0 = __float2half_rn(x.x), h1 = __float2half_rn(x.y), h2 = __float2half_rn(x.z), h3 = __float2half_rn(x.w);
  const __half l0 = __float2half_rn(x.x - __half2float(h0)), l1 = __float2half_rn(x.y - __half2float(h1));
  const __half l2 = __float2half_rn(x.z - __half2float(h2)), l3 = __float2half_rn(x.w - __half2float(h3));
  const size_t o = (size_t)row * Kp + 4 * j;
  __half2* ph = reinterpret_cast<__half2*>(T_hi + o);
  __half2* pl = reinterpret_cast<__half2*>(T_lo + o);
  ph[0] = __halves2half2(h0, h1); ph[1] = __halves2half2(h2, h3);
  pl[0] = __halves2half2(l0, l1); pl[1] = __halves2half2(l2, l3);
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

template <bool kSkin>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRpThreads, 1)
lbs_replay_gemm_kernel(const __grid_constant__ CUtensorMap tmap_p_hi, const __grid_constant__ CUtensorMap tmap_p_lo,
                       const __grid_constant__ CUtensorMap tmap_t_hi, const __grid_constant__ CUtensorMap tmap_t_lo,
                       const ReplayArgs args) {
  extern __shared__ uint8_t rp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(rp_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRpStages * kRpStageBytes);
  uint64_t* full_bar = bars;                        // [kRpStages]   (used in the leader)
  uint64_t* empty_bar = bars + kRpStages;           // [kRpStages]
  uint64_t* tmem_full = bars + 2 * kRpStages;       // [2]
  uint64_t* tmem_empty = bars + 2 * kRpStages + 2;  // [2]          (used in the leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kRpStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int num_tiles = args.num_m_blocks * args.num_n_blocks;
  constexpr int kUmmaK = 16;
  constexpr int kKSteps = kRpKB / kUmmaK;

  ptx::pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_p_hi);
    ptx::prefetch_tmap(&tmap_p_lo);
    ptx::prefetch_tmap(&tmap_t_hi);
    ptx::prefetch_tmap(&tmap_t_lo);
    for (int s = 0; s < kRpStages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 2 * 32 * kRpEpiWarps);   // every epilogue thread, in both CTAs
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
  ptx::pdl_wait();         // set-up above under the operand kernel's tail; its rows (and kSkin: v_posed) are read from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int mb, nb;
        replay_tile_coords(tile, args.num_m_blocks, args.num_n_blocks, num_clusters, mb, nb);
        const int m0 = mb * 2 * kBlendBM + (int)rank * kBlendBM;       // this CTA's vertex rows
        const int n0 = nb * kRpBN + (int)rank * (kRpBN / 2);           // this CTA's half of the (frame, c) rows
        for (int kb = 0; kb < args.num_k_blocks; ++kb) {
          ptx::mbar_wait_backoff(&empty_bar[stage], phase ^ 1, 40);
          uint8_t* st = stage_base + stage * kRpStageBytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kRpStageBytes);
          else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
          const int k0 = kb * kRpKB;
          ptx::tma_load_2d_2sm(st, &tmap_p_hi, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + kRpTileABytes, &tmap_p_lo, &full_bar[stage], k0, m0);
          ptx::tma_load_2d_2sm(st + 2 * kRpTileABytes, &tmap_t_hi, &full_bar[stage], k0, n0);
          ptx::tma_load_2d_2sm(st + 2 * kRpTileABytes + kRpTileBBytes, &tmap_t_lo, &full_bar[stage], k0, n0);
          if (++stage == kRpStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * kBlendBM, kRpBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        ptx::mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1, 40);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlendBN;
        for (int kb = 0; kb < args.num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(stage_base + stage * kRpStageBytes);
          const uint64_t a_hi = ptx::make_kmajor_desc<kRpRowBytes>(st);
          const uint64_t a_lo = ptx::make_kmajor_desc<kRpRowBytes>(st + kRpTileABytes);
          const uint64_t b_hi = ptx::make_kmajor_desc<kRpRowBytes>(st + 2 * kRpTileABytes);
          const uint64_t b_lo = ptx::make_kmajor_desc<kRpRowBytes>(st + 2 * kRpTileABytes + kRpTileBBytes);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);
            const uint32_t first = (kb != 0 || k != 0) ? 1u : 0u;
            ptx::umma_2cta<true>(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
            ptx::umma_2cta<true>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            ptx::umma_2cta<true>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
          }
          ptx::umma_commit_2cta(&empty_bar[stage]);
          if (++stage == kRpStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2cta(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): TMEM -> registers -> global =====================
    const int half = (warp - 2) >> 2;         // which 40 frames of the tile
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const float oscale = args.out_scale;
    const int V = args.V;
    [[maybe_unused]] float2 na[kRpTileBodies / 2];      // kSkin: the next tile's v_posed, in flight
    [[maybe_unused]] float nbv[kRpTileBodies / 2];
    for (int it = 0;; ++it) {
      const int tile = cluster_id + it * num_clusters;
      if (tile >= num_tiles) break;
      int mb, nb;
      replay_tile_coords(tile, args.num_m_blocks, args.num_n_blocks, num_clusters, mb, nb);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int v = mb * 2 * kBlendBM + (int)rank * kBlendBM + q * 32 + lane;
      const bool v_ok = v < V;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(acc * kBlendBN + half * (kRpBN / 2));
      // the warp's whole share of the accumulator (120 columns) in one go: the loads fly together, the accumulator
      // goes back to the MMA thread at once, and the stores below depend on nothing but registers
      if constexpr (!kSkin) {
        uint32_t r[kRpBN / 2];
        const int f_base = nb * kRpTileFrames + half * (kRpTileFrames / 2);
        ptx::mbar_wait_backoff(&tmem_full[acc], acc_phase, 20);
        ptx::tcgen05_fence_after();
#pragma unroll
        for (int i = 0; i < (kRpBN / 2) / 16; ++i) ptx::tmem_ld_32x32b_x16(taddr0 + 16 * i, r + 16 * i);
        tmem_ld_32x32b_x8(taddr0 + 16 * ((kRpBN / 2) / 16), r + 16 * ((kRpBN / 2) / 16));
        // the frames' translations: lane l of load i holds transl[3 f_base + 32 i + l] (broadcast by shuffle below)
        float tv[(kRpBN / 2 + 31) / 32];
#pragma unroll
        for (int i = 0; i < (kRpBN / 2 + 31) / 32; ++i) {
          const long idx = 3 * (long)f_base + 32 * i + lane;
          tv[i] = (args.transl != nullptr && 32 * i + lane < kRpBN / 2 && idx < 3 * (long)args.F) ? __ldg(args.transl + idx) : 0.f;
        }
        ptx::tmem_ld_wait();
        ptx::tcgen05_fence_before();
        ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
        // 12 bytes per lane and frame as one 8-byte and one 4-byte store, whichever order is aligned: element index
        // e = 3 (f V + v); e even -> (x, y) | z, e odd -> x | (y, z).  Branch-free: the lanes of a warp alternate.
        const int nf = min(kRpTileFrames / 2, args.F - f_base);
        size_t e = ((size_t)f_base * V + (v_ok ? v : 0)) * 3;
        const size_t e_step = (size_t)V * 3;
#pragma unroll
        for (int k = 0; k < kRpTileFrames / 2; ++k) {
          const float tx = __shfl_sync(0xffffffffu, tv[(3 * k) >> 5], (3 * k) & 31);
          const float ty = __shfl_sync(0xffffffffu, tv[(3 * k + 1) >> 5], (3 * k + 1) & 31);
          const float tz = __shfl_sync(0xffffffffu, tv[(3 * k + 2) >> 5], (3 * k + 2) & 31);
          if (k < nf && v_ok) {
            const float x = fmaf(__uint_as_float(r[3 * k]), oscale, tx);
            const float y = fmaf(__uint_as_float(r[3 * k + 1]), oscale, ty);
            const float z = fmaf(__uint_as_float(r[3 * k + 2]), oscale, tz);
            const bool al = (e & 1) == 0;
            float* p = args.out + e;
            *reinterpret_cast<float2*>(p + (al ? 0 : 1)) = make_float2(al ? x : y, al ? y : z);
            p[al ? 2 : 0] = al ? z : x;
          }
          e += e_step;
        }
      } else {
        constexpr int kHalfBodies = kRpTileBodies / 2;
        const int b_base = nb * kRpTileBodies + half * kHalfBodies;
        const int nbod = min(kHalfBodies, args.F - b_base);
        // this vertex's v_posed of the warp's bodies: 12 bytes per lane as an 8-byte and a 4-byte load, whichever order
        // is aligned (row strides are even, so the parity is the vertex's).  The loads of the NEXT tile are issued
        // before this tile's accumulator is waited for, so their HBM latency hides behind a whole tile.  Lanes past
        // the last vertex and bodies past the last body read a clamped (valid) address and store nothing.
        const bool al = (v & 1) == 0;
        float2 pa[kHalfBodies];
        float pb[kHalfBodies];
        auto load_vp = [&](int vv, int bb0) {
          const int vc = min(vv, V - 1);
          const bool a2 = (vc & 1) == 0;
          const float* p = args.vsrc + (size_t)min(bb0, args.F - 1) * args.vsrc_stride + 3 * (size_t)vc;
          const int o2 = a2 ? 0 : 1, o1 = a2 ? 2 : 0;
          const size_t last = (size_t)max(0, min(kHalfBodies, args.F - bb0) - 1) * args.vsrc_stride;
          size_t off = 0;
#pragma unroll
          for (int i = 0; i < kHalfBodies; ++i) {
            na[i] = *reinterpret_cast<const float2*>(p + off + o2);
            nbv[i] = p[off + o1];
            off = min(off + args.vsrc_stride, last);
          }
        };
        if (it == 0) load_vp(v, b_base);
#pragma unroll
        for (int i = 0; i < kHalfBodies; ++i) { pa[i] = na[i]; pb[i] = nbv[i]; }
        {
          const int tile2 = tile + num_clusters;
          if (tile2 < num_tiles) {
            int mb2, nb2;
            replay_tile_coords(tile2, args.num_m_blocks, args.num_n_blocks, num_clusters, mb2, nb2);
            load_vp(mb2 * 2 * kBlendBM + (int)rank * kBlendBM + q * 32 + lane, nb2 * kRpTileBodies + half * kHalfBodies);
          }
        }
        const long tidx = 3 * (long)b_base + lane;
        const float tv = (args.transl != nullptr && lane < 3 * kHalfBodies && tidx < 3 * (long)args.F) ? __ldg(args.transl + tidx) : 0.f;
        // output: element index e = 3 (b V + v); the 8-byte store goes where e is even
        const int vodd = V & 1;
        int par = (int)(((size_t)b_base * vodd + (size_t)v) & 1);
        float* po = args.out + ((size_t)b_base * V + (v_ok ? v : 0)) * 3;
        const size_t po_step = (size_t)V * 3;
        ptx::mbar_wait_backoff(&tmem_full[acc], acc_phase, 20);
        ptx::tcgen05_fence_after();
        // the warp's 120 columns in two rounds of 5 bodies (60 registers; the register file gives a 10-warp block 168)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[60];
          const uint32_t ta = taddr0 + 60 * h;
          ptx::tmem_ld_32x32b_x16(ta, r);
          ptx::tmem_ld_32x32b_x16(ta + 16, r + 16);
          ptx::tmem_ld_32x32b_x16(ta + 32, r + 32);
          tmem_ld_32x32b_x8(ta + 48, r + 48);
          ptx::tmem_ld_32x32b_x4(ta + 56, r + 56);
          ptx::tmem_ld_wait();
          if (h == 1) {
            ptx::tcgen05_fence_before();
            ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
          }
#pragma unroll
          for (int ii = 0; ii < kHalfBodies / 2; ++ii) {
            const int i = h * (kHalfBodies / 2) + ii;
            const float tx = __shfl_sync(0xffffffffu, tv, 3 * i);
            const float ty = __shfl_sync(0xffffffffu, tv, 3 * i + 1);
            const float tz = __shfl_sync(0xffffffffu, tv, 3 * i + 2);
            const float px = al ? pa[i].x : pb[i], py = al ? pa[i].y : pa[i].x, pz = al ? pb[i] : pa[i].y;
            float t[12];
#pragma unroll
            for (int n = 0; n < 12; ++n) t[n] = __uint_as_float(r[12 * ii + n]);
            const float x = fmaf(fmaf(t[0], px, fmaf(t[1], py, fmaf(t[2], pz, t[3]))), oscale, tx);
            const float y = fmaf(fmaf(t[4], px, fmaf(t[5], py, fmaf(t[6], pz, t[7]))), oscale, ty);
            const float z = fmaf(fmaf(t[8], px, fmaf(t[9], py, fmaf(t[10], pz, t[11]))), oscale, tz);
            if (i < nbod && v_ok) {
              const bool sal = par == 0;
              *reinterpret_cast<float2*>(po + (sal ? 0 : 1)) = make_float2(sal ? x : y, sal ? y : z);
              po[sal ? 2 : 0] = sal ? z : x;
            }
            po += po_step;
            par ^= vodd;
          }
        }
      }
    }
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

}  // namespace smplk
