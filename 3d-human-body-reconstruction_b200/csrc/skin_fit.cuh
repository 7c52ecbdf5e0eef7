// skin_fit_l2_kernel -- the vertex-L2 fitting step's middle (BASELINE config 3) in one pass:
//
//   verts  = LBS(v_posed, A) + transl                (skin_grouped_kernel's loop)
//   loss_b += scale * ||verts - target||^2           (lib/Gen_SMPLH/fitting.py:491-495 on vertices)
//   g      = 2 scale (verts - target)                -> `grad` (B,V,3), read by dA_kernel
//   d_v_posed = sum_u w_u R_u^T g                    (skin_backward_grouped_kernel's loop)
//             -> bf16 two-term split rows, the A operand of the backward blend GEMM
//
// instead of skinning (writes verts), the loss kernel (reads verts + target, writes g) and the
// skinning backward (reads g, writes d_v_posed): the vertices and the transforms a thread already
// holds never go back to HBM.  bf16 instead of the fp16 split of the stand-alone backward: the
// row scale of the fp16 path needs max|g| of the whole body before the first row is written, bf16
// has fp32's exponent range (two terms keep 16 mantissa bits; gradient error ~1e-6 relative).
#pragma once
#include <cuda_bf16.h>
#include "skinning.cuh"

namespace smplk {

struct SkinFitArgs {
  int B;
  int bodies_per_block;
  const float* vposed;        // (B, vposed_stride)
  size_t vposed_stride;       // floats
  const float* A;             // (B,J,12)
  const float* transl;        // (B,3) or null
  const float* target;        // (B,V,3)
  float scale;
  float* grad;                // (B,V,3)
  float* loss;                // (B) -- or one float with loss_stride 0 --, zeroed by the host, accumulated with atomics
  int loss_stride;
  __nv_bfloat16* dvp_hi;      // (B,Npad)
  __nv_bfloat16* dvp_lo;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, float& ra, float& rb) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
  ra = a - __bfloat162float(ha);
  rb = b - __bfloat162float(hb);
  return (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
}

// Same streaming skeleton as skin_grouped_kernel (a thread owns 4 consecutive vertices with the
// <= 8 distinct joints of the group; each warp streams its 128 vertices through a cp.async ring).
__global__ void __launch_bounds__(kGrpThreads, 2)
skin_fit_l2_kernel(const ModelDev m, const SkinFitArgs a) {
  extern __shared__ __align__(16) float sf_smem[];
  ptx::pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int a_floats = m.J * 12;
  const int a_chunks = m.J * 3;
  const int a_pad = grp_a_pad(m.J);
  float* ring = sf_smem;                                         // [stages][1024*3]
  float* Abuf = sf_smem + kGrpStages * kSkinTileVerts * 3;       // [2][8][a_pad]
  float* Tbuf = Abuf + 2 * kGrpABodies * a_pad;                  // [2][8][4] translations

  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  const int wf0 = v0 * 3 + warp * kWarpFloats;
  const int w_nfloat = max(0, min(kWarpFloats, m.V * 3 - wf0));   // even: the host requires 3V even
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

  auto issue_A = [&](int grp) {
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
  auto issue_v = [&](int b) {
    if (b < b1) {
      const float* src = a.vposed + (size_t)b * a.vposed_stride + wf0;
      float* dst = ring + ((b - b0) % kGrpStages) * (kSkinTileVerts * 3) + warp * kWarpFloats;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = lane + 32 * i;
        if (wf0 + 4 * c + 4 <= m.Npad) ptx::cp_async_16(dst + 4 * c, src + 4 * c);
      }
    }
    ptx::cp_async_commit();
  };

  // never-copied tail floats of the ring must be finite (they meet zero weights)
  for (int i = tid; i < kGrpStages * kSkinTileVerts * 3; i += kGrpThreads) ring[i] = 0.f;
  __syncthreads();
  ptx::pdl_wait();          // group tables and the ring set-up above overlap the blend GEMM's tail; v_posed / A from here on
  issue_A(0);
#pragma unroll
  for (int i = 0; i < kGrpStages - 1; ++i) issue_v(b0 + i);
  const float s2 = 2.f * a.scale;
  const int col0 = wf0 + 12 * lane;                       // first d_v_posed column of this thread
  float loss_run = 0.f;

  for (int b = b0; b < b1; ++b) {
    const int rel = b - b0;
    const int agrp = rel / kGrpABodies;
    issue_v(b + kGrpStages - 1);
    // this body's target values at the warp's store positions: in flight under the skinning loop
    const float2* t2 = reinterpret_cast<const float2*>(a.target + (size_t)b * m.V * 3 + wf0);
    float2 tg[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int t = lane + 32 * i;
      tg[i] = (2 * t + 2 <= w_nfloat) ? __ldcs(t2 + t) : make_float2(0.f, 0.f);
    }
    ptx::cp_async_wait<kGrpStages - 1>();
    if ((rel % kGrpABodies) == 0) {
      __syncthreads();
      if (b + kGrpABodies < b1) issue_A(agrp + 1);
    } else {
      __syncwarp();
    }
    const float* Ab = Abuf + ((agrp & 1) * kGrpABodies + (rel % kGrpABodies)) * a_pad;
    const float* Tb = Tbuf + ((agrp & 1) * kGrpABodies + (rel % kGrpABodies)) * 4;
    const float tx = Tb[0], ty = Tb[1], tz = Tb[2];
    float* slot = ring + (rel % kGrpStages) * (kSkinTileVerts * 3) + warp * kWarpFloats;
    float4* mine = reinterpret_cast<float4*>(slot) + 3 * lane;
    float vx[4], vy[4], vz[4];
    {
      const float4 c0 = mine[0], c1 = mine[1], c2 = mine[2];
      vx[0] = c0.x; vx[1] = c0.w; vx[2] = c1.z; vx[3] = c2.y;
      vy[0] = c0.y; vy[1] = c1.x; vy[2] = c1.w; vy[3] = c2.z;
      vz[0] = c0.z; vz[1] = c1.y; vz[2] = c2.x; vz[3] = c2.w;
    }
    // blended transforms of the 4 vertices, T_i = sum_u w_ui A_u: applied to v_posed now and,
    // transposed, to the gradient below (one pass over the joints instead of two)
    float T[4][12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int q = 0; q < 12; ++q) T[i][q] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const float Aq[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
        const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int q = 0; q < 12; ++q) T[i][q] = fmaf(wu[i], Aq[q], T[i][q]);
        }
      }
    }
    float ox[4], oy[4], oz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ox[i] = fmaf(T[i][0], vx[i], fmaf(T[i][1], vy[i], fmaf(T[i][2], vz[i], T[i][3])));
      oy[i] = fmaf(T[i][4], vx[i], fmaf(T[i][5], vy[i], fmaf(T[i][6], vz[i], T[i][7])));
      oz[i] = fmaf(T[i][8], vx[i], fmaf(T[i][9], vy[i], fmaf(T[i][10], vz[i], T[i][11])));
    }
    mine[0] = make_float4(ox[0] + tx, oy[0] + ty, oz[0] + tz, ox[1] + tx);
    mine[1] = make_float4(oy[1] + ty, oz[1] + tz, ox[2] + tx, oy[2] + ty);
    mine[2] = make_float4(oz[2] + tz, ox[3] + tx, oy[3] + ty, oz[3] + tz);
    __syncwarp();
    // residual, loss and gradient on the coalesced (float2 per lane) view of the warp's slice; the
    // gradient goes to HBM for dA_kernel and back into the slot for this thread's own 4 vertices
    float2* g2 = reinterpret_cast<float2*>(a.grad + (size_t)b * m.V * 3 + wf0);
    float2* sl2 = reinterpret_cast<float2*>(slot);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int t = lane + 32 * i;
      const bool ok = 2 * t + 2 <= w_nfloat;
      const float2 ov = sl2[t];
      const float dx = ok ? ov.x - tg[i].x : 0.f, dy = ok ? ov.y - tg[i].y : 0.f;
      acc = fmaf(dx, dx, fmaf(dy, dy, acc));
      const float2 gv = make_float2(s2 * dx, s2 * dy);
      sl2[t] = gv;
      if (ok) g2[t] = gv;
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    if (a.loss_stride == 0) loss_run += acc;            // one float for all bodies: one atomic per warp, after the loop
    else if (lane == 0 && w_nfloat > 0) atomicAdd(a.loss + b, a.scale * acc);
    __syncwarp();
    {
      const float4 c0 = mine[0], c1 = mine[1], c2 = mine[2];
      vx[0] = c0.x; vx[1] = c0.w; vx[2] = c1.z; vx[3] = c2.y;
      vy[0] = c0.y; vy[1] = c1.x; vy[2] = c1.w; vy[3] = c2.z;
      vz[0] = c0.z; vz[1] = c1.y; vz[2] = c2.x; vz[3] = c2.w;
    }
    __syncwarp();                                        // the slot is free for the copy of body b + stages
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // d_v_posed = T_R^T g
      ox[i] = fmaf(T[i][0], vx[i], fmaf(T[i][4], vy[i], T[i][8] * vz[i]));
      oy[i] = fmaf(T[i][1], vx[i], fmaf(T[i][5], vy[i], T[i][9] * vz[i]));
      oz[i] = fmaf(T[i][2], vx[i], fmaf(T[i][6], vy[i], T[i][10] * vz[i]));
    }
    // bf16 two-term split; a thread's 12 columns are 24 contiguous bytes per term
    const float o[12] = {ox[0], oy[0], oz[0], ox[1], oy[1], oz[1], ox[2], oy[2], oz[2], ox[3], oy[3], oz[3]};
    uint2* oh = reinterpret_cast<uint2*>(a.dvp_hi + (size_t)b * m.Npad + col0);
    uint2* ol = reinterpret_cast<uint2*>(a.dvp_lo + (size_t)b * m.Npad + col0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float ra, rb, rc, rd, d0, d1;
      const uint32_t h0 = pack_bf16x2(o[4 * i], o[4 * i + 1], ra, rb);
      const uint32_t h1 = pack_bf16x2(o[4 * i + 2], o[4 * i + 3], rc, rd);
      const uint32_t l0 = pack_bf16x2(ra, rb, d0, d1);
      const uint32_t l1 = pack_bf16x2(rc, rd, d0, d1);
      if (col0 + 4 * i + 4 <= m.Npad) {
        oh[i] = make_uint2(h0, h1);
        ol[i] = make_uint2(l0, l1);
      }
    }
  }
  if (a.loss_stride == 0 && lane == 0 && w_nfloat > 0) atomicAdd(a.loss, a.scale * loss_run);
}

}  // namespace smplk
