// skin_grouped8_kernel -- skin_grouped_kernel with 8 consecutive vertices per thread.
//
// The transform gathers dominate the L1/shared-memory traffic of skin_grouped_kernel (each fetched
// 3x4 transform costs 12 shared-memory wavefronts per warp no matter how many lanes share the
// address: 512 B must be written back to registers).  Owning 8 vertices per thread halves the
// fetches per vertex again: the packer stores the <= 8 distinct joints of every 8-vertex group.
// Block = 4 warps x 256 vertices (same 1024-vertex tile), transforms staged 4 bodies at a time.
// A lane's 96 bytes are kept at a 112-byte pitch in shared memory so its 16-byte accesses are
// bank-conflict free ((7 l + i) mod 8 is a permutation over a quarter warp).
#pragma once
#include "skinning.cuh"

// packed fp32x2 FMAs in the skinning loop: measured SLOWER here (0.208 vs 0.172 ms at 4096 bodies; the kernel is
// bound by shared-memory wavefronts and latency, not FMA issue) -- kept for A/B builds only
#ifndef SMPLK_SKIN_FFMA2
#define SMPLK_SKIN_FFMA2 0
#endif

namespace smplk {

constexpr int k8Threads = 128;
constexpr int k8Warps = 4;
constexpr int k8Stages = 3;
constexpr int k8ABodies = 4;
constexpr int k8WarpFloats = 768;            // 256 vertices x 3
constexpr int k8WarpPitch = 32 * 28;         // floats per warp slice in smem (7 x 16 B per lane)
constexpr int k8StageFloats = k8Warps * k8WarpPitch;

__host__ __device__ inline size_t skin_grouped8_smem_bytes(int J) {
  return (size_t)(k8Stages * k8StageFloats + 2 * k8ABodies * grp_a_pad(J) + 2 * k8ABodies * 4) * sizeof(float);
}

// 16-byte chunk k (0..191) of a warp slice -> float offset in the padded smem layout
__device__ __forceinline__ int pad_pos(int k) { return (7 * (k / 6) + (k % 6)) * 4; }

template <bool kSharedTemplate>
__global__ void __launch_bounds__(k8Threads, 3)
skin_grouped8_kernel(const ModelDev m, const SkinArgs a) {
  extern __shared__ __align__(16) float s8_smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int a_floats = m.J * 12;
  const int a_chunks = m.J * 3;
  const int a_pad = grp_a_pad(m.J);
  float* ring = s8_smem;
  float* Abuf = s8_smem + k8Stages * k8StageFloats;      // [2][4][a_pad]
  float* Tbuf = Abuf + 2 * k8ABodies * a_pad;            // [2][4][4]

  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  const int wf0 = v0 * 3 + warp * k8WarpFloats;
  const int w_nfloat = max(0, min(k8WarpFloats, m.V * 3 - wf0));
  const int g = (v0 >> 3) + tid;                          // global 8-vertex group
  const bool g_valid = 8 * g < m.V;
  uint2 jid = make_uint2(0u, 0u);
  float4 wa[kGrpJoints], wb[kGrpJoints];
  uint32_t used = 0;
#pragma unroll
  for (int u = 0; u < kGrpJoints; ++u) {
    wa[u] = g_valid ? m.grp8_w[((size_t)g * kGrpJoints + u) * 2 + 0] : make_float4(0.f, 0.f, 0.f, 0.f);
    wb[u] = g_valid ? m.grp8_w[((size_t)g * kGrpJoints + u) * 2 + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (wa[u].x != 0.f || wa[u].y != 0.f || wa[u].z != 0.f || wa[u].w != 0.f || wb[u].x != 0.f ||
        wb[u].y != 0.f || wb[u].z != 0.f || wb[u].w != 0.f)
      used |= 1u << u;
  }
  if (g_valid) jid = m.grp8_joints[g];
  used = __reduce_or_sync(0xffffffffu, used);
  const int ovf0 = g_valid ? m.grp8_ovf_ptr[g] : 0, ovf1 = g_valid ? m.grp8_ovf_ptr[g + 1] : 0;

  auto issue_A = [&](int grp) {
    const int bb0 = b0 + grp * k8ABodies;
    const int nb = min(k8ABodies, b1 - bb0);
    float* dstA = Abuf + (grp & 1) * k8ABodies * a_pad;
    for (int c = tid; c < nb * a_chunks; c += k8Threads) {
      const int bi = c / a_chunks, cc = c - bi * a_chunks;
      ptx::cp_async_16(dstA + bi * a_pad + 4 * cc, a.A + (size_t)(bb0 + bi) * a_floats + 4 * cc);
    }
    if (tid < nb * 3) {
      float* dstT = Tbuf + (grp & 1) * k8ABodies * 4;
      const int bi = tid / 3, k = tid - bi * 3;
      dstT[bi * 4 + k] = a.transl ? a.transl[(size_t)(bb0 + bi) * 3 + k] : 0.f;
    }
  };
  auto issue_v = [&](int b) {
    if (b < b1 && (!kSharedTemplate || b == b0)) {
      const float* src = a.vsrc + (size_t)b * a.vsrc_stride + wf0;
      float* dst = ring + ((b - b0) % k8Stages) * k8StageFloats + warp * k8WarpPitch;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int c = lane + 32 * i;
        if (wf0 + 4 * c + 4 <= m.Npad) ptx::cp_async_16(dst + pad_pos(c), src + 4 * c);
      }
    }
    ptx::cp_async_commit();
  };

  issue_A(0);
#pragma unroll
  for (int i = 0; i < k8Stages - 1; ++i) issue_v(b0 + i);
  const bool even_rows = ((m.V * 3) & 1) == 0;

  for (int b = b0; b < b1; ++b) {
    const int rel = b - b0;
    const int agrp = rel / k8ABodies;
    issue_v(b + k8Stages - 1);
    ptx::cp_async_wait<k8Stages - 1>();
    if ((rel % k8ABodies) == 0) {
      __syncthreads();
      if (b + k8ABodies < b1) issue_A(agrp + 1);
    } else {
      __syncwarp();
    }
    const float* Ab = Abuf + ((agrp & 1) * k8ABodies + (rel % k8ABodies)) * a_pad;
    const float* Tb = Tbuf + ((agrp & 1) * k8ABodies + (rel % k8ABodies)) * 4;
    const float tx = Tb[0], ty = Tb[1], tz = Tb[2];
    float* slot = ring + (kSharedTemplate ? 0 : (rel % k8Stages)) * k8StageFloats + warp * k8WarpPitch;
    const float4* mine = reinterpret_cast<const float4*>(slot) + 7 * lane;
    float c[24];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 q = mine[i];
      c[4 * i + 0] = q.x; c[4 * i + 1] = q.y; c[4 * i + 2] = q.z; c[4 * i + 3] = q.w;
    }
    float o[24];
#if SMPLK_SKIN_FFMA2
    // two vertices per instruction: pairs (2k, 2k+1) as packed fp32x2 operands; the transform component is a
    // scalar-broadcast source, the weight pair sits in adjacent registers of the float4 it was loaded as
    uint64_t x2[4], y2[4], z2[4], ox2[4], oy2[4], oz2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x2[k] = ptx::pack_f32x2(c[6 * k], c[6 * k + 3]);
      y2[k] = ptx::pack_f32x2(c[6 * k + 1], c[6 * k + 4]);
      z2[k] = ptx::pack_f32x2(c[6 * k + 2], c[6 * k + 5]);
      ox2[k] = oy2[k] = oz2[k] = ptx::pack_f32x2(0.f, 0.f);
    }
    auto apply2 = [&](const float4* Aj, const float4& wa_, const float4& wb_) {
      const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
      const uint64_t w2[4] = {ptx::pack_f32x2(wa_.x, wa_.y), ptx::pack_f32x2(wa_.z, wa_.w),
                              ptx::pack_f32x2(wb_.x, wb_.y), ptx::pack_f32x2(wb_.z, wb_.w)};
      const uint64_t a00 = ptx::pack_f32x2(r0.x, r0.x), a01 = ptx::pack_f32x2(r0.y, r0.y), a02 = ptx::pack_f32x2(r0.z, r0.z),
                     a03 = ptx::pack_f32x2(r0.w, r0.w), a10 = ptx::pack_f32x2(r1.x, r1.x), a11 = ptx::pack_f32x2(r1.y, r1.y),
                     a12 = ptx::pack_f32x2(r1.z, r1.z), a13 = ptx::pack_f32x2(r1.w, r1.w), a20 = ptx::pack_f32x2(r2.x, r2.x),
                     a21 = ptx::pack_f32x2(r2.y, r2.y), a22 = ptx::pack_f32x2(r2.z, r2.z), a23 = ptx::pack_f32x2(r2.w, r2.w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t px = ptx::fma_f32x2(a00, x2[k], ptx::fma_f32x2(a01, y2[k], ptx::fma_f32x2(a02, z2[k], a03)));
        const uint64_t py = ptx::fma_f32x2(a10, x2[k], ptx::fma_f32x2(a11, y2[k], ptx::fma_f32x2(a12, z2[k], a13)));
        const uint64_t pz = ptx::fma_f32x2(a20, x2[k], ptx::fma_f32x2(a21, y2[k], ptx::fma_f32x2(a22, z2[k], a23)));
        ox2[k] = ptx::fma_f32x2(w2[k], px, ox2[k]);
        oy2[k] = ptx::fma_f32x2(w2[k], py, oy2[k]);
        oz2[k] = ptx::fma_f32x2(w2[k], pz, oz2[k]);
      }
    };
#else
#pragma unroll
    for (int i = 0; i < 24; ++i) o[i] = 0.f;
    auto apply2 = [&](const float4* Aj, const float4& wa_, const float4& wb_) {
      const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
      const float wu[8] = {wa_.x, wa_.y, wa_.z, wa_.w, wb_.x, wb_.y, wb_.z, wb_.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = c[3 * i], y = c[3 * i + 1], z = c[3 * i + 2];
        const float px = fmaf(r0.x, x, fmaf(r0.y, y, fmaf(r0.z, z, r0.w)));
        const float py = fmaf(r1.x, x, fmaf(r1.y, y, fmaf(r1.z, z, r1.w)));
        const float pz = fmaf(r2.x, x, fmaf(r2.y, y, fmaf(r2.z, z, r2.w)));
        o[3 * i + 0] = fmaf(wu[i], px, o[3 * i + 0]);
        o[3 * i + 1] = fmaf(wu[i], py, o[3 * i + 1]);
        o[3 * i + 2] = fmaf(wu[i], pz, o[3 * i + 2]);
      }
    };
#endif
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        apply2(reinterpret_cast<const float4*>(Ab + j * 12), wa[u], wb[u]);
      }
    }
    // the rare group bound to more than 8 joints (where vertex ranges of several joints meet): its extra
    // (joint, weights) entries come from a per-group list; only that lane runs the loop
    for (int e = ovf0; e < ovf1; ++e)
      apply2(reinterpret_cast<const float4*>(Ab + m.grp8_ovf_joint[e] * 12), m.grp8_ovf_w[2 * e], m.grp8_ovf_w[2 * e + 1]);
#if SMPLK_SKIN_FFMA2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ptx::unpack_f32x2(ox2[k], o[6 * k], o[6 * k + 3]);
      ptx::unpack_f32x2(oy2[k], o[6 * k + 1], o[6 * k + 4]);
      ptx::unpack_f32x2(oz2[k], o[6 * k + 2], o[6 * k + 5]);
    }
#endif
    float* orow = a.out + (size_t)b * m.V * 3 + wf0;
    float* stage = kSharedTemplate ? ring + k8StageFloats + warp * k8WarpPitch : slot;
    float4* ot = reinterpret_cast<float4*>(stage) + 7 * lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float t0 = (4 * i + 0) % 3 == 0 ? tx : ((4 * i + 0) % 3 == 1 ? ty : tz);
      const float t1 = (4 * i + 1) % 3 == 0 ? tx : ((4 * i + 1) % 3 == 1 ? ty : tz);
      const float t2 = (4 * i + 2) % 3 == 0 ? tx : ((4 * i + 2) % 3 == 1 ? ty : tz);
      const float t3 = (4 * i + 3) % 3 == 0 ? tx : ((4 * i + 3) % 3 == 1 ? ty : tz);
      ot[i] = make_float4(o[4 * i] + t0, o[4 * i + 1] + t1, o[4 * i + 2] + t2, o[4 * i + 3] + t3);
    }
    __syncwarp();
    if (even_rows) {
      float2* o2 = reinterpret_cast<float2*>(orow);
      if (w_nfloat == k8WarpFloats) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int t = lane + 32 * i;
          __stcs(o2 + t, *reinterpret_cast<const float2*>(stage + pad_pos(t >> 1) + 2 * (t & 1)));
        }
      } else {
        for (int t = lane; t < (w_nfloat >> 1); t += 32)
          __stcs(o2 + t, *reinterpret_cast<const float2*>(stage + pad_pos(t >> 1) + 2 * (t & 1)));
        if ((w_nfloat & 1) && lane == 0) {
          const int t = w_nfloat - 1;
          orow[t] = stage[pad_pos(t >> 2) + (t & 3)];
        }
      }
    } else {
      for (int t = lane; t < w_nfloat; t += 32) __stcs(orow + t, stage[pad_pos(t >> 2) + (t & 3)]);
    }
    __syncwarp();
  }
}

}  // namespace smplk
