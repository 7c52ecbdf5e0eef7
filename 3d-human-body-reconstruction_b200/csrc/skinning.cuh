// Fused linear-blend skinning: verts[b,v] = (sum_k w[v,k] A[b, j[v,k]]) [v_posed[b,v]; 1] + transl[b]
//
// Reference math: models/smplh_np.py:79-82 (T = tensordot(W, A); v = T [v_posed;1]; + trans),
// lib/model2video.py:77-81 for the rigged-mesh variant (v_template instead of v_posed).
//
// HBM-bound design: a block owns one tile of 1024 consecutive vertices and walks over a group of
// bodies.  The vertex tile (12 KB contiguous) and the body's 3x4 transforms are staged into shared
// memory with 16-byte cp.async, double buffered so the next body's loads overlap this body's math;
// each thread keeps the (<=4) joint ids and weights of its 4 vertices in registers across bodies;
// results are written back in place and streamed out as fully coalesced 8-byte stores.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

// packed fp32x2 FMAs in the skinning loop: measured SLOWER here (0.208 vs 0.172 ms at 4096 bodies; the kernel is
// bound by shared-memory wavefronts and latency, not FMA issue) -- kept for A/B builds only
#ifndef SMPLK_SKIN_FFMA2
#define SMPLK_SKIN_FFMA2 0
#endif

namespace smplk {

constexpr int kSkinThreads = 256;
constexpr int kSkinVPT = kSkinTileVerts / kSkinThreads;  // 4 vertices per thread

struct SkinArgs {
  int B;                  // bodies
  int bodies_per_block;
  const float* vsrc;      // v_posed rows, or the shared v_template row
  size_t vsrc_stride;     // floats between bodies (0 = shared template)
  const float* A;         // (B,J,12)
  const float* transl;    // (B,3) or null
  float* out;             // (B,V,3)
  int debug_copy_only;    // tuning aid: skip the transform blend (measures the streaming skeleton)
};

template <bool kReg4>
__global__ void __launch_bounds__(kSkinThreads)
skin_kernel(const ModelDev m, const SkinArgs a) {
  extern __shared__ __align__(16) float skin_smem[];
  const int tid = threadIdx.x;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int nv = min(kSkinTileVerts, m.V - v0);
  const int nfloat = nv * 3;
  const int nchunk = (nfloat + 3) >> 2;     // 16-byte chunks of the vertex tile
  const int a_floats = m.J * 12;
  const int a_chunks = m.J * 3;
  const int a_pad = (a_floats + 3) & ~3;
  float* vt[2] = {skin_smem, skin_smem + kSkinTileVerts * 3};
  float* As[2] = {skin_smem + 2 * kSkinTileVerts * 3, skin_smem + 2 * kSkinTileVerts * 3 + a_pad};

  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  // per-thread skin weights of vertices v0 + tid + 256 i
  uint32_t idx4[kSkinVPT];
  float4 w4[kSkinVPT];
  if (kReg4) {
#pragma unroll
    for (int i = 0; i < kSkinVPT; ++i) {
      const int v = v0 + tid + kSkinThreads * i;
      if (v < m.V) {
        idx4[i] = m.skin_idx4[v];
        w4[i] = m.skin_w4[v];
      } else {
        idx4[i] = 0;
        w4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  auto prefetch = [&](int b, int buf) {
    const float* src = a.vsrc + (size_t)b * a.vsrc_stride + (size_t)v0 * 3;
    for (int c = tid; c < nchunk; c += kSkinThreads) ptx::cp_async_16(vt[buf] + 4 * c, src + 4 * c);
    const float* asrc = a.A + (size_t)b * a_floats;
    for (int c = tid; c < a_chunks; c += kSkinThreads) ptx::cp_async_16(As[buf] + 4 * c, asrc + 4 * c);
    ptx::cp_async_commit();
  };

  prefetch(b0, 0);
  const bool even_rows = ((m.V * 3) & 1) == 0;
  for (int b = b0; b < b1; ++b) {
    const int buf = (b - b0) & 1;
    if (b + 1 < b1) {
      prefetch(b + 1, buf ^ 1);
      ptx::cp_async_wait<1>();
    } else {
      ptx::cp_async_wait<0>();
    }
    __syncthreads();
    float* vtile = vt[buf];
    const float* Ab = As[buf];
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (a.transl) {
      tx = a.transl[3 * b + 0];
      ty = a.transl[3 * b + 1];
      tz = a.transl[3 * b + 2];
    }
#pragma unroll
    for (int i = 0; i < kSkinVPT; ++i) {
      const int lv = tid + kSkinThreads * i;
      if (lv < nv) {
        const float x = vtile[3 * lv + 0], y = vtile[3 * lv + 1], z = vtile[3 * lv + 2];
        float T[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) T[q] = 0.f;
        if (kReg4) {
          const float wk[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = (idx4[i] >> (8 * k)) & 0xff;
            const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
            const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
            const float w = wk[k];
            T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]);
            T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
            T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]);
            T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
            T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]);
            T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
          }
        } else {
          const int v = v0 + lv;
          for (int k = 0; k < m.ell_k; ++k) {
            const float w = m.ell_w[(size_t)k * m.V + v];
            if (w != 0.f) {
              const int j = m.ell_idx[(size_t)k * m.V + v];
              const float* Aj = Ab + j * 12;
#pragma unroll
              for (int q = 0; q < 12; ++q) T[q] = fmaf(w, Aj[q], T[q]);
            }
          }
        }
        vtile[3 * lv + 0] = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3]))) + tx;
        vtile[3 * lv + 1] = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7]))) + ty;
        vtile[3 * lv + 2] = fmaf(T[8], x, fmaf(T[9], y, fmaf(T[10], z, T[11]))) + tz;
      }
    }
    __syncthreads();
    float* orow = a.out + (size_t)b * m.V * 3 + (size_t)v0 * 3;
    if (even_rows) {  // row base and tile offset are both even -> 8-byte aligned
      const int n2 = nfloat >> 1;
      float2* o2 = reinterpret_cast<float2*>(orow);
      const float2* s2 = reinterpret_cast<const float2*>(vtile);
      for (int c = tid; c < n2; c += kSkinThreads) __stcs(o2 + c, s2[c]);
      if ((nfloat & 1) && tid == 0) orow[nfloat - 1] = vtile[nfloat - 1];
    } else {
      for (int c = tid; c < nfloat; c += kSkinThreads) __stcs(orow + c, vtile[c]);
    }
    __syncthreads();  // tile buffer `buf` is refilled by the prefetch of body b+2
  }
}

// ------------------------------------------------------------------------------------------
// skin_grouped_kernel -- the product path for sparse weights.
//
// ncu on skin_kernel (profiles/r01_ncu_full_first_path.csv) showed it bound by shared-memory
// wavefronts and block barriers, not HBM: every vertex gathered 4 x 48 B of transforms per body
// (61 M wavefronts per launch at B=4096) and every body cost three __syncthreads.  Here:
//  * a thread owns a group of 4 CONSECUTIVE vertices; neighbouring vertices share most of their
//    joints, so the packer stores the <= 8 distinct joints of the group with a 4x8 weight block and
//    each transform is fetched from shared memory once per group (typ. 4-6 x 48 B per 4 vertices
//    instead of 16 x 48 B);
//  * each WARP streams its own 128 vertices (1536 contiguous bytes per body) through a private
//    ring of cp.async stages: coalesced 16-byte copies two bodies ahead, in-place compute, coalesced
//    8-byte stores; only __syncwarp inside the body loop;
//  * the bodies' 3x4 transforms are staged for 8 bodies at a time (double buffered), so the block
//    meets at a barrier once per 8 bodies.
// ------------------------------------------------------------------------------------------
constexpr int kGrpThreads = 256;     // 8 warps x 128 vertices = one 1024-vertex tile
#ifndef SMPLK_GRP_STAGES
#define SMPLK_GRP_STAGES 5
#endif
constexpr int kGrpStages = SMPLK_GRP_STAGES;   // cp.async ring depth per warp (prefetch S-1 bodies ahead)
constexpr int kGrpABodies = 8;       // transforms staged per A-group
constexpr int kWarpFloats = 384;     // 128 vertices x 3

__host__ __device__ inline int grp_a_pad(int J) { return (J * 12 + 3) & ~3; }
__host__ __device__ inline size_t skin_grouped_smem_bytes(int J) {
  return (size_t)(kGrpStages * kSkinTileVerts * 3 + 2 * kGrpABodies * grp_a_pad(J) + 4 * kGrpABodies * 2) *
         sizeof(float);
}
// SMPLK_SKIN_NOBAR (rigged-mesh replay, kSharedTemplate): the transforms of ALL the block's bodies are staged up
// front (one barrier per block instead of one per 8 bodies: ncu showed the barrier as the largest stall of
// skin_grouped_kernel, 1.1-1.2 cycles per issued instruction) and the thread's template vertices stay in
// registers; costs bodies_per_block x J x 48 bytes of shared memory.  Measured (profiles/r02_skin_variants.txt):
// 0.283 -> 0.269 ms at 8,192 frames x 6,890 vertices, 1.687 -> 1.605 ms at 50,000 vertices; with per-body
// v_posed rows (J = 52: only 20 bodies fit) it is no gain (0.173 -> 0.176 ms), so that path keeps the 8-body groups.
#ifndef SMPLK_SKIN_NOBAR
#define SMPLK_SKIN_NOBAR 1
#endif
#ifndef SMPLK_SKIN_PREFETCH
#define SMPLK_SKIN_PREFETCH 0
#endif
__host__ __device__ inline size_t skin_nobar_smem_bytes(int J, int bodies_per_block) {
  return (size_t)(kGrpStages * kSkinTileVerts * 3 + bodies_per_block * (grp_a_pad(J) + 4)) * sizeof(float);
}
// most bodies per block whose transforms fit beside the ring with two blocks per SM
__host__ inline int skin_nobar_max_bodies(int J, size_t smem_per_block) {
  const size_t ring = (size_t)kGrpStages * kSkinTileVerts * 3 * sizeof(float);
  if (smem_per_block <= ring) return 0;
  return (int)((smem_per_block - ring) / ((size_t)(grp_a_pad(J) + 4) * sizeof(float)));
}

template <bool kSharedTemplate>
__global__ void __launch_bounds__(kGrpThreads, 2)
skin_grouped_kernel(const ModelDev m, const SkinArgs a) {
  extern __shared__ __align__(16) float sg_smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int a_floats = m.J * 12;
  const int a_chunks = m.J * 3;
  const int a_pad = grp_a_pad(m.J);
  float* ring = sg_smem;                                         // [stages][1024*3]
  // kNoBar (rigged-mesh replay): [bodies_per_block][a_pad] + [bodies_per_block][4]; else [2][8][a_pad] + [2][8][4]
  constexpr bool kNoBar = kSharedTemplate && (SMPLK_SKIN_NOBAR != 0);
  float* Abuf = sg_smem + kGrpStages * kSkinTileVerts * 3;
  float* Tbuf = Abuf + (kNoBar ? a.bodies_per_block : 2 * kGrpABodies) * a_pad;

  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  // this warp's slice of the row: floats [wf0, wf0 + 384)
  const int wf0 = v0 * 3 + warp * kWarpFloats;
  const int w_nfloat = max(0, min(kWarpFloats, m.V * 3 - wf0));   // valid output floats of the warp
  const int g = (v0 >> 2) + tid;                                  // global 4-vertex group of this thread
  const bool g_valid = 4 * g < m.V;
  uint2 jid = make_uint2(0u, 0u);
  float4 w[kGrpJoints];
  uint32_t used = 0;
#pragma unroll
  for (int u = 0; u < kGrpJoints; ++u) {
    w[u] = g_valid ? m.grp_w[(size_t)g * kGrpJoints + u] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (w[u].x != 0.f || w[u].y != 0.f || w[u].z != 0.f || w[u].w != 0.f) used |= 1u << u;
  }
  if (g_valid) jid = m.grp_joints[g];
  used = __reduce_or_sync(0xffffffffu, used);   // warp-uniform: no divergence in the joint loop
  if (a.debug_copy_only) used = 0;

  // ---- async copy helpers (every thread commits the same number of groups)
  auto issue_A = [&](int grp) {                 // transforms + translations of bodies b0+8grp ..
    if (kNoBar) {                               // ... or of ALL the block's bodies
      const int nb = b1 - b0;
      for (int c = tid; c < nb * a_chunks; c += kGrpThreads) {
        const int bi = c / a_chunks, cc = c - bi * a_chunks;
        ptx::cp_async_16(Abuf + bi * a_pad + 4 * cc, a.A + (size_t)(b0 + bi) * a_floats + 4 * cc);
      }
      for (int c = tid; c < nb * 3; c += kGrpThreads) {
        const int bi = c / 3, k = c - bi * 3;
        Tbuf[bi * 4 + k] = a.transl ? a.transl[(size_t)(b0 + bi) * 3 + k] : 0.f;
      }
      return;
    }
    const int bb0 = b0 + grp * kGrpABodies;
    const int nb = min(kGrpABodies, b1 - bb0);
    float* dstA = Abuf + (grp & 1) * kGrpABodies * a_pad;
    for (int c = tid; c < nb * a_chunks; c += kGrpThreads) {
      const int bi = c / a_chunks, cc = c - bi * a_chunks;
      ptx::cp_async_16(dstA + bi * a_pad + 4 * cc, a.A + (size_t)(bb0 + bi) * a_floats + 4 * cc);
    }
    if (tid < nb * 3) {
      float* dstT = Tbuf + (grp & 1) * kGrpABodies * 4;
      const int bi = tid / 3, k = tid - bi * 3;
      dstT[bi * 4 + k] = a.transl ? a.transl[(size_t)(bb0 + bi) * 3 + k] : 0.f;
    }
  };
  auto issue_v = [&](int b) {                   // this warp's 1536 B of body b -> ring slot
    if (b < b1 && (!kSharedTemplate || b == b0)) {
      const float* src = a.vsrc + (size_t)b * a.vsrc_stride + wf0;
      float* dst = ring + ((b - b0) % kGrpStages) * (kSkinTileVerts * 3) + warp * kWarpFloats;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = lane + 32 * i;                          // 16-byte chunk of the warp slice
        if (wf0 + 4 * c + 4 <= m.Npad) ptx::cp_async_16(dst + 4 * c, src + 4 * c);
      }
    }
    ptx::cp_async_commit();
  };

  issue_A(0);                                   // joins the first commit group
#pragma unroll
  for (int i = 0; i < kGrpStages - 1; ++i) issue_v(b0 + i);
  const bool even_rows = ((m.V * 3) & 1) == 0;
  float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, c2 = c0;

  for (int b = b0; b < b1; ++b) {
    const int rel = b - b0;
    const int agrp = rel / kGrpABodies;
    issue_v(b + kGrpStages - 1);
    ptx::cp_async_wait<kGrpStages - 1>();                     // body b (and everything older) landed
    if (kNoBar) {
      if (rel == 0) __syncthreads();  // every thread's share of the transforms (first commit group) is published
      else __syncwarp();
    } else if ((rel % kGrpABodies) == 0) {
      // this thread's share of the group's transforms was issued >= 8 bodies ago (or in the
      // prologue) and is covered by the wait above; the barrier publishes it block-wide
      __syncthreads();
      if (b + kGrpABodies < b1) issue_A(agrp + 1);            // overwrites group agrp-1; joins next commit
    } else {
      __syncwarp();
    }
    const int arow = kNoBar ? rel : (agrp & 1) * kGrpABodies + (rel % kGrpABodies);
    const float* Ab = Abuf + arow * a_pad;
    const float* Tb = Tbuf + arow * 4;
    const float tx = Tb[0], ty = Tb[1], tz = Tb[2];
    float* slot = ring + (kSharedTemplate ? 0 : (rel % kGrpStages)) * (kSkinTileVerts * 3) + warp * kWarpFloats;
    float4* mine = reinterpret_cast<float4*>(slot) + 3 * lane;
    if (!kSharedTemplate || rel == 0) { c0 = mine[0]; c1 = mine[1]; c2 = mine[2]; }   // a shared template stays in registers
    // vertices of the group: (c0.x c0.y c0.z) (c0.w c1.x c1.y) (c1.z c1.w c2.x) (c2.y c2.z c2.w)
    const float vx[4] = {c0.x, c0.w, c1.z, c2.y};
    const float vy[4] = {c0.y, c1.x, c1.w, c2.z};
    const float vz[4] = {c0.z, c1.y, c2.x, c2.w};
    float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f}, oz[4] = {0.f, 0.f, 0.f, 0.f};
#if SMPLK_SKIN_FFMA2
    // vertex pairs (0,1) and (2,3) as packed fp32x2 operands (FFMA2): half the FMA issue slots
    const uint64_t x2[2] = {ptx::pack_f32x2(vx[0], vx[1]), ptx::pack_f32x2(vx[2], vx[3])};
    const uint64_t y2[2] = {ptx::pack_f32x2(vy[0], vy[1]), ptx::pack_f32x2(vy[2], vy[3])};
    const uint64_t z2[2] = {ptx::pack_f32x2(vz[0], vz[1]), ptx::pack_f32x2(vz[2], vz[3])};
    uint64_t ox2[2], oy2[2], oz2[2];
    ox2[0] = ox2[1] = oy2[0] = oy2[1] = oz2[0] = oz2[1] = ptx::pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const uint64_t w2[2] = {ptx::pack_f32x2(w[u].x, w[u].y), ptx::pack_f32x2(w[u].z, w[u].w)};
        const uint64_t a00 = ptx::pack_f32x2(r0.x, r0.x), a01 = ptx::pack_f32x2(r0.y, r0.y), a02 = ptx::pack_f32x2(r0.z, r0.z),
                       a03 = ptx::pack_f32x2(r0.w, r0.w), a10 = ptx::pack_f32x2(r1.x, r1.x), a11 = ptx::pack_f32x2(r1.y, r1.y),
                       a12 = ptx::pack_f32x2(r1.z, r1.z), a13 = ptx::pack_f32x2(r1.w, r1.w), a20 = ptx::pack_f32x2(r2.x, r2.x),
                       a21 = ptx::pack_f32x2(r2.y, r2.y), a22 = ptx::pack_f32x2(r2.z, r2.z), a23 = ptx::pack_f32x2(r2.w, r2.w);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t px = ptx::fma_f32x2(a00, x2[k], ptx::fma_f32x2(a01, y2[k], ptx::fma_f32x2(a02, z2[k], a03)));
          const uint64_t py = ptx::fma_f32x2(a10, x2[k], ptx::fma_f32x2(a11, y2[k], ptx::fma_f32x2(a12, z2[k], a13)));
          const uint64_t pz = ptx::fma_f32x2(a20, x2[k], ptx::fma_f32x2(a21, y2[k], ptx::fma_f32x2(a22, z2[k], a23)));
          ox2[k] = ptx::fma_f32x2(w2[k], px, ox2[k]);
          oy2[k] = ptx::fma_f32x2(w2[k], py, oy2[k]);
          oz2[k] = ptx::fma_f32x2(w2[k], pz, oz2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      ptx::unpack_f32x2(ox2[k], ox[2 * k], ox[2 * k + 1]);
      ptx::unpack_f32x2(oy2[k], oy[2 * k], oy[2 * k + 1]);
      ptx::unpack_f32x2(oz2[k], oz[2 * k], oz[2 * k + 1]);
    }
#elif SMPLK_SKIN_PREFETCH
    // software pipelining across the joint slots: the transform of slot u+1 is fetched from shared memory
    // before slot u's 48 FMAs (every slot is its own basic block behind the warp-uniform `used` test, so the
    // compiler cannot hoist the loads itself).  A lane's non-zero slots are a prefix (the packer sorts a group's
    // joints by weight), so `used` is a prefix mask too.
    {
      float4 r0, r1, r2;
      {
        const float4* Aj = reinterpret_cast<const float4*>(Ab + (jid.x & 0xff) * 12);
        r0 = Aj[0]; r1 = Aj[1]; r2 = Aj[2];
      }
#pragma unroll
      for (int u = 0; u < kGrpJoints; ++u) {
        if (used & (1u << u)) {
          float4 n0 = r0, n1 = r1, n2 = r2;
          if (u + 1 < kGrpJoints && (used & (2u << u))) {
            const int jn = (((u + 1) < 4 ? jid.x : jid.y) >> (8 * ((u + 1) & 3))) & 0xff;
            const float4* An = reinterpret_cast<const float4*>(Ab + jn * 12);
            n0 = An[0]; n1 = An[1]; n2 = An[2];
          }
          const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float px = fmaf(r0.x, vx[i], fmaf(r0.y, vy[i], fmaf(r0.z, vz[i], r0.w)));
            const float py = fmaf(r1.x, vx[i], fmaf(r1.y, vy[i], fmaf(r1.z, vz[i], r1.w)));
            const float pz = fmaf(r2.x, vx[i], fmaf(r2.y, vy[i], fmaf(r2.z, vz[i], r2.w)));
            ox[i] = fmaf(wu[i], px, ox[i]);
            oy[i] = fmaf(wu[i], py, oy[i]);
            oz[i] = fmaf(wu[i], pz, oz[i]);
          }
          r0 = n0; r1 = n1; r2 = n2;
        }
      }
    }
#else
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float px = fmaf(r0.x, vx[i], fmaf(r0.y, vy[i], fmaf(r0.z, vz[i], r0.w)));
          const float py = fmaf(r1.x, vx[i], fmaf(r1.y, vy[i], fmaf(r1.z, vz[i], r1.w)));
          const float pz = fmaf(r2.x, vx[i], fmaf(r2.y, vy[i], fmaf(r2.z, vz[i], r2.w)));
          ox[i] = fmaf(wu[i], px, ox[i]);
          oy[i] = fmaf(wu[i], py, oy[i]);
          oz[i] = fmaf(wu[i], pz, oz[i]);
        }
      }
    }
#endif
    float* orow = a.out + (size_t)b * m.V * 3 + wf0;
    float* stage = kSharedTemplate ? ring + kSkinTileVerts * 3 + warp * kWarpFloats : slot;
    float4* ot = reinterpret_cast<float4*>(stage) + 3 * lane;
    ot[0] = make_float4(ox[0] + tx, oy[0] + ty, oz[0] + tz, ox[1] + tx);
    ot[1] = make_float4(oy[1] + ty, oz[1] + tz, ox[2] + tx, oy[2] + ty);
    ot[2] = make_float4(oz[2] + tz, ox[3] + tx, oy[3] + ty, oz[3] + tz);
    __syncwarp();
    if (even_rows) {                 // row base, tile and warp offsets are all even -> 8-byte aligned
      const float2* s2 = reinterpret_cast<const float2*>(stage);
      float2* o2 = reinterpret_cast<float2*>(orow);
      if (w_nfloat == kWarpFloats) {
#pragma unroll
        for (int i = 0; i < 6; ++i) __stcs(o2 + lane + 32 * i, s2[lane + 32 * i]);
      } else {
        for (int c = lane; c < (w_nfloat >> 1); c += 32) __stcs(o2 + c, s2[c]);
        if ((w_nfloat & 1) && lane == 0) orow[w_nfloat - 1] = stage[w_nfloat - 1];
      }
    } else {
      for (int c = lane; c < w_nfloat; c += 32) __stcs(orow + c, stage[c]);
    }
    __syncwarp();                    // slot is refilled by a later issue_v of this warp
  }
}

#ifdef SMPLK_AB   // A/B variant, not on any default path: built only with -DSMPLK_AB
// ------------------------------------------------------------------------------------------
// skin_tma_kernel -- skin_grouped_kernel with the streaming data moved off the LSU pipe.
//
// ncu on skin_grouped_kernel (profiles/r01_ncu_skin_grouped_v3.csv): L1TEX data pipe 63 % busy
// (the cp.async writes, the LDS.64/STG.64 copy-out and the transform gathers all share it) and
// 17 % of the stall samples on the block barrier.  Here every warp runs a private TMA pipeline:
//   lane 0 issues one 1536-byte cp.async.bulk (global -> this warp's ring slot, mbarrier
//   complete_tx) three bodies ahead; the warp waits on the slot's mbarrier, computes in place,
//   and lane 0 issues one cp.async.bulk smem -> global for the result.  Output rows of (B,V,3)
//   are only 8-byte aligned for odd bodies: those results are written 8 bytes shifted inside the
//   (16-byte padded) slot so the bulk store's source and destination are both 16-byte aligned,
//   with the first and last 8 bytes stored by two lanes.
// The 8-body transform groups are bulk-loaded by warp 0 behind full/empty mbarriers, so there is
// no __syncthreads in the body loop.
// ------------------------------------------------------------------------------------------
constexpr int kTmaStages = 5;                  // ring slots per warp; loads run kTmaStages-2 ahead
constexpr int kTmaDist = kTmaStages - 2;
constexpr int kSlotFloats = kWarpFloats + 4;   // 1536 B + 16 B pad (shifted odd-row results)

__host__ __device__ inline size_t skin_tma_smem_bytes(int J) {
  return (size_t)(8 * kTmaStages * kSlotFloats + 2 * kGrpABodies * grp_a_pad(J)) * sizeof(float) +
         (8 * kTmaStages + 4) * sizeof(uint64_t) + 16;
}

__global__ void __launch_bounds__(kGrpThreads, 2)
skin_tma_kernel(const ModelDev m, const SkinArgs a) {
  extern __shared__ __align__(16) float st_smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int a_floats = m.J * 12;
  const int a_pad = grp_a_pad(m.J);
  float* ring = st_smem + warp * (kTmaStages * kSlotFloats);      // this warp's slots
  float* Abuf = st_smem + 8 * kTmaStages * kSlotFloats;           // [2][8][a_pad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Abuf + 2 * kGrpABodies * a_pad);
  uint64_t* vfull = bars + warp * kTmaStages;                     // [kTmaStages] per warp
  uint64_t* afull = bars + 8 * kTmaStages;                        // [2]
  uint64_t* aempty = afull + 2;                                   // [2]

  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  const int wf0 = v0 * 3 + warp * kWarpFloats;
  const int w_nfloat = max(0, min(kWarpFloats, m.V * 3 - wf0));
  const uint32_t in_bytes = (uint32_t)max(0, min(kWarpFloats, m.Npad - wf0)) * 4u;   // multiple of 16
  const bool full_slice = w_nfloat == kWarpFloats;
  const int g = (v0 >> 2) + tid;
  const bool g_valid = 4 * g < m.V;
  uint2 jid = make_uint2(0u, 0u);
  float4 w[kGrpJoints];
  uint32_t used = 0;
#pragma unroll
  for (int u = 0; u < kGrpJoints; ++u) {
    w[u] = g_valid ? m.grp_w[(size_t)g * kGrpJoints + u] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (w[u].x != 0.f || w[u].y != 0.f || w[u].z != 0.f || w[u].w != 0.f) used |= 1u << u;
  }
  if (g_valid) jid = m.grp_joints[g];
  used = __reduce_or_sync(0xffffffffu, used);

  if (tid == 0) {
    for (int i = 0; i < 8 * kTmaStages; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::mbar_init(&afull[0], 1); ptx::mbar_init(&afull[1], 1);
    ptx::mbar_init(&aempty[0], 8); ptx::mbar_init(&aempty[1], 8);
    ptx::fence_barrier_init();
  }
  __syncthreads();

  auto issue_v = [&](int b) {            // lane 0 only
    if (b < b1 && in_bytes > 0) {
      const int s = (b - b0) % kTmaStages;
      ptx::mbar_arrive_expect_tx(&vfull[s], in_bytes);
      ptx::bulk_load(ring + s * kSlotFloats, a.vsrc + (size_t)b * a.vsrc_stride + wf0, in_bytes, &vfull[s]);
    }
  };
  auto issue_A = [&](int grp) {          // warp 0 lane 0 only
    const int bb0 = b0 + grp * kGrpABodies;
    const int nb = min(kGrpABodies, b1 - bb0);
    const uint32_t bytes = (uint32_t)(nb * a_floats) * 4u;      // J*48 bytes per body: multiple of 16
    float* dst = Abuf + (grp & 1) * kGrpABodies * a_pad;
    ptx::mbar_arrive_expect_tx(&afull[grp & 1], bytes);
    if (a_pad == a_floats) {
      ptx::bulk_load(dst, a.A + (size_t)bb0 * a_floats, bytes, &afull[grp & 1]);
    } else {
      for (int i = 0; i < nb; ++i)
        ptx::bulk_load(dst + i * a_pad, a.A + (size_t)(bb0 + i) * a_floats, (uint32_t)a_floats * 4u, &afull[grp & 1]);
    }
  };

  if (lane == 0) {
    if (warp == 0) issue_A(0);
    for (int i = 0; i < kTmaDist; ++i) issue_v(b0 + i);
  }
  const bool even_rows = ((m.V * 3) & 1) == 0;
  float tn[3] = {0.f, 0.f, 0.f};         // translation of the next body, prefetched
  if (a.transl && lane < 3) tn[0] = a.transl[(size_t)b0 * 3 + lane];

  for (int b = b0; b < b1; ++b) {
    const int rel = b - b0;
    const int agrp = rel / kGrpABodies;
    const int s = rel % kTmaStages;
    if (lane == 0) {
      ptx::tma_store_wait_read<1>();                 // stores of bodies <= b-2 have left their slots
      issue_v(b + kTmaDist);
    }
    // translation: lanes 0..2 hold x,y,z of this body; prefetch the next one
    const float tcur = tn[0];
    if (a.transl && lane < 3 && b + 1 < b1) tn[0] = a.transl[(size_t)(b + 1) * 3 + lane];
    const float tx = __shfl_sync(0xffffffffu, tcur, 0), ty = __shfl_sync(0xffffffffu, tcur, 1),
                tz = __shfl_sync(0xffffffffu, tcur, 2);
    if ((rel % kGrpABodies) == 0) {
      ptx::mbar_wait(&afull[agrp & 1], (agrp >> 1) & 1);
      if (warp == 0 && lane == 0 && b + kGrpABodies < b1) {
        if (agrp >= 1) ptx::mbar_wait(&aempty[(agrp + 1) & 1], ((agrp - 1) >> 1) & 1);
        issue_A(agrp + 1);
      }
    }
    if (in_bytes > 0) ptx::mbar_wait(&vfull[s], (rel / kTmaStages) & 1);
    const float* Ab = Abuf + ((agrp & 1) * kGrpABodies + (rel % kGrpABodies)) * a_pad;
    float* slot = ring + s * kSlotFloats;
    const float4* mine = reinterpret_cast<const float4*>(slot) + 3 * lane;
    const float4 c0 = mine[0], c1 = mine[1], c2 = mine[2];
    const float vx[4] = {c0.x, c0.w, c1.z, c2.y};
    const float vy[4] = {c0.y, c1.x, c1.w, c2.z};
    const float vz[4] = {c0.z, c1.y, c2.x, c2.w};
    float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f}, oz[4] = {0.f, 0.f, 0.f, 0.f};
#if SMPLK_SKIN_FFMA2
    // vertex pairs (0,1) and (2,3) as packed fp32x2 operands (FFMA2): half the FMA issue slots
    const uint64_t x2[2] = {ptx::pack_f32x2(vx[0], vx[1]), ptx::pack_f32x2(vx[2], vx[3])};
    const uint64_t y2[2] = {ptx::pack_f32x2(vy[0], vy[1]), ptx::pack_f32x2(vy[2], vy[3])};
    const uint64_t z2[2] = {ptx::pack_f32x2(vz[0], vz[1]), ptx::pack_f32x2(vz[2], vz[3])};
    uint64_t ox2[2], oy2[2], oz2[2];
    ox2[0] = ox2[1] = oy2[0] = oy2[1] = oz2[0] = oz2[1] = ptx::pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const uint64_t w2[2] = {ptx::pack_f32x2(w[u].x, w[u].y), ptx::pack_f32x2(w[u].z, w[u].w)};
        const uint64_t a00 = ptx::pack_f32x2(r0.x, r0.x), a01 = ptx::pack_f32x2(r0.y, r0.y), a02 = ptx::pack_f32x2(r0.z, r0.z),
                       a03 = ptx::pack_f32x2(r0.w, r0.w), a10 = ptx::pack_f32x2(r1.x, r1.x), a11 = ptx::pack_f32x2(r1.y, r1.y),
                       a12 = ptx::pack_f32x2(r1.z, r1.z), a13 = ptx::pack_f32x2(r1.w, r1.w), a20 = ptx::pack_f32x2(r2.x, r2.x),
                       a21 = ptx::pack_f32x2(r2.y, r2.y), a22 = ptx::pack_f32x2(r2.z, r2.z), a23 = ptx::pack_f32x2(r2.w, r2.w);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t px = ptx::fma_f32x2(a00, x2[k], ptx::fma_f32x2(a01, y2[k], ptx::fma_f32x2(a02, z2[k], a03)));
          const uint64_t py = ptx::fma_f32x2(a10, x2[k], ptx::fma_f32x2(a11, y2[k], ptx::fma_f32x2(a12, z2[k], a13)));
          const uint64_t pz = ptx::fma_f32x2(a20, x2[k], ptx::fma_f32x2(a21, y2[k], ptx::fma_f32x2(a22, z2[k], a23)));
          ox2[k] = ptx::fma_f32x2(w2[k], px, ox2[k]);
          oy2[k] = ptx::fma_f32x2(w2[k], py, oy2[k]);
          oz2[k] = ptx::fma_f32x2(w2[k], pz, oz2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      ptx::unpack_f32x2(ox2[k], ox[2 * k], ox[2 * k + 1]);
      ptx::unpack_f32x2(oy2[k], oy[2 * k], oy[2 * k + 1]);
      ptx::unpack_f32x2(oz2[k], oz[2 * k], oz[2 * k + 1]);
    }
#elif SMPLK_SKIN_PREFETCH
    // software pipelining across the joint slots: the transform of slot u+1 is fetched from shared memory
    // before slot u's 48 FMAs (every slot is its own basic block behind the warp-uniform `used` test, so the
    // compiler cannot hoist the loads itself).  A lane's non-zero slots are a prefix (the packer sorts a group's
    // joints by weight), so `used` is a prefix mask too.
    {
      float4 r0, r1, r2;
      {
        const float4* Aj = reinterpret_cast<const float4*>(Ab + (jid.x & 0xff) * 12);
        r0 = Aj[0]; r1 = Aj[1]; r2 = Aj[2];
      }
#pragma unroll
      for (int u = 0; u < kGrpJoints; ++u) {
        if (used & (1u << u)) {
          float4 n0 = r0, n1 = r1, n2 = r2;
          if (u + 1 < kGrpJoints && (used & (2u << u))) {
            const int jn = (((u + 1) < 4 ? jid.x : jid.y) >> (8 * ((u + 1) & 3))) & 0xff;
            const float4* An = reinterpret_cast<const float4*>(Ab + jn * 12);
            n0 = An[0]; n1 = An[1]; n2 = An[2];
          }
          const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float px = fmaf(r0.x, vx[i], fmaf(r0.y, vy[i], fmaf(r0.z, vz[i], r0.w)));
            const float py = fmaf(r1.x, vx[i], fmaf(r1.y, vy[i], fmaf(r1.z, vz[i], r1.w)));
            const float pz = fmaf(r2.x, vx[i], fmaf(r2.y, vy[i], fmaf(r2.z, vz[i], r2.w)));
            ox[i] = fmaf(wu[i], px, ox[i]);
            oy[i] = fmaf(wu[i], py, oy[i]);
            oz[i] = fmaf(wu[i], pz, oz[i]);
          }
          r0 = n0; r1 = n1; r2 = n2;
        }
      }
    }
#else
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float px = fmaf(r0.x, vx[i], fmaf(r0.y, vy[i], fmaf(r0.z, vz[i], r0.w)));
          const float py = fmaf(r1.x, vx[i], fmaf(r1.y, vy[i], fmaf(r1.z, vz[i], r1.w)));
          const float pz = fmaf(r2.x, vx[i], fmaf(r2.y, vy[i], fmaf(r2.z, vz[i], r2.w)));
          ox[i] = fmaf(wu[i], px, ox[i]);
          oy[i] = fmaf(wu[i], py, oy[i]);
          oz[i] = fmaf(wu[i], pz, oz[i]);
        }
      }
    }
#endif
    const float r[12] = {ox[0] + tx, oy[0] + ty, oz[0] + tz, ox[1] + tx, oy[1] + ty, oz[1] + tz,
                         ox[2] + tx, oy[2] + ty, oz[2] + tz, ox[3] + tx, oy[3] + ty, oz[3] + tz};
    float* orow = a.out + (size_t)b * m.V * 3 + wf0;
    const uint32_t misalign = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(orow) & 15);
    if (full_slice && even_rows && misalign == 0) {
      float4* ot = reinterpret_cast<float4*>(slot) + 3 * lane;
      ot[0] = make_float4(r[0], r[1], r[2], r[3]);
      ot[1] = make_float4(r[4], r[5], r[6], r[7]);
      ot[2] = make_float4(r[8], r[9], r[10], r[11]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::bulk_store(orow, slot, kWarpFloats * 4);
        ptx::tma_store_commit();
      }
    } else if (full_slice && even_rows && misalign == 8) {
      // row base = 8 (mod 16): results shifted by 8 bytes inside the padded slot
      __syncwarp();                                   // all lanes hold their inputs in registers
      float2* o2 = reinterpret_cast<float2*>(slot + 2) + 6 * lane;
#pragma unroll
      for (int i = 0; i < 6; ++i) o2[i] = make_float2(r[2 * i], r[2 * i + 1]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::bulk_store(orow + 2, slot + 4, kWarpFloats * 4 - 16);     // 16-byte aligned middle
        ptx::tma_store_commit();
        *reinterpret_cast<float2*>(orow) = *reinterpret_cast<const float2*>(slot + 2);
      } else if (lane == 31) {
        *reinterpret_cast<float2*>(orow + kWarpFloats - 2) =
            *reinterpret_cast<const float2*>(slot + kWarpFloats);
      }
    } else {
      // partial slice (mesh tail) or odd 3V: coalesced LSU copy-out
      float4* ot = reinterpret_cast<float4*>(slot) + 3 * lane;
      ot[0] = make_float4(r[0], r[1], r[2], r[3]);
      ot[1] = make_float4(r[4], r[5], r[6], r[7]);
      ot[2] = make_float4(r[8], r[9], r[10], r[11]);
      __syncwarp();
      for (int c = lane; c < w_nfloat; c += 32) __stcs(orow + c, slot[c]);
      __syncwarp();
      if (lane == 0) ptx::tma_store_commit();         // keep bulk-group accounting uniform
    }
    if (lane == 0 && ((rel % kGrpABodies) == kGrpABodies - 1 || b + 1 == b1))
      ptx::mbar_arrive(&aempty[agrp & 1]);
  }
  if (lane == 0) ptx::tma_store_wait<0>();
}
#endif  // SMPLK_AB

// joints[b, J + e] = verts[b, extra_vids[e]]  (upstream VertexJointSelector; verts already + transl)
__global__ void gather_extra_joints_kernel(const ModelDev m, int B, const float* __restrict__ verts,
                                           float* __restrict__ joints, int joints_ld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * m.E) return;
  int b = i / m.E, e = i % m.E;
  const float* v = verts + ((size_t)b * m.V + m.extra_vids[e]) * 3;
  float* o = joints + (size_t)b * joints_ld + 3 * (m.J + e);
  o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
}

// out[b, r] = sum_n val[n] verts[b, col[n]]   (CSR rows of regressor_posed; warp per (b, r))
__global__ void regress_joints_kernel(const ModelDev m, int B, const float* __restrict__ verts,
                                      float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * m.R) return;
  const int b = warp / m.R, r = warp % m.R;
  float x = 0.f, y = 0.f, z = 0.f;
  const float* vb = verts + (size_t)b * m.V * 3;
  for (int n = m.reg_ptr[r] + lane; n < m.reg_ptr[r + 1]; n += 32) {
    const float w = m.reg_val[n];
    const float* v = vb + (size_t)m.reg_col[n] * 3;
    x = fmaf(w, v[0], x); y = fmaf(w, v[1], y); z = fmaf(w, v[2], z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
    z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  if (lane == 0) {
    float* o = out + ((size_t)b * m.R + r) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
}

}  // namespace smplk
