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
