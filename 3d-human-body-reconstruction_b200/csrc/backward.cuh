// Backward (vector-Jacobian product) of the body-model forward -- replaces the autograd pass of
// total_loss.backward() at lib/Gen_SMPLH/fitting.py:256 through upstream smplx lbs().
//
//   d_verts ----> skin_backward_kernel : d_v_posed = (sum_k w_k R_A[j_k])^T d_verts   (tf32 hi/lo)
//           \---> dA_kernel            : dA[b,j] = sum_v w[v,j] d_verts[b,v] (x) [v_posed[b,v];1]
//   d_v_posed --> blend GEMM (split-K) : d_feat = d_v_posed . [posedirs | shapedirs]
//   dA, d_feat, d_joints --> pose_backward_kernel : chain / rest-joint / Rodrigues / PCA backward
#pragma once
#include "common.cuh"
#include "pose_kernels.cuh"
#include "ptx_sm100.cuh"
#include "skinning.cuh"

namespace smplk {

// d_verts_eff[b, v] += sum over the picks / regressor rows that read vertex v (ModelDev::sc_*): one
// thread per (body, touched vertex), terms added in table order -> no atomics, bit-reproducible.
__global__ void scatter_joint_grads_kernel(const ModelDev m, int B, const float* __restrict__ d_joints,
                                           int joints_ld, const float* __restrict__ d_jreg,
                                           float* __restrict__ dverts) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || t >= m.sc_T) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int n = m.sc_ptr[t]; n < m.sc_ptr[t + 1]; ++n) {
    const int src = m.sc_src[n];
    const float* g;
    if (src < m.E) {
      if (d_joints == nullptr) continue;
      g = d_joints + (size_t)b * joints_ld + 3 * (m.J + src);
    } else {
      if (d_jreg == nullptr) continue;
      g = d_jreg + ((size_t)b * m.R + (src - m.E)) * 3;
    }
    const float w = m.sc_w[n];
    s0 = fmaf(w, g[0], s0); s1 = fmaf(w, g[1], s1); s2 = fmaf(w, g[2], s2);
  }
  float* o = dverts + ((size_t)b * m.V + m.sc_vert[t]) * 3;
  o[0] += s0; o[1] += s1; o[2] += s2;
}

// Keypoint fitting (the reference's actual closure: the data term of lib/Gen_SMPLH/fitting.py:369-381
// reads only joints): the vertex gradient is non-zero at the E vertex-pick joints alone (nose, eyes,
// ears, toes, heels, finger tips), so the dense backward -- a (B,V,3) gradient buffer, dA over every
// vertex, the skinning backward and a 20,736-deep GEMM -- collapses to one small exact-fp32 kernel:
//   per pick e (vertex v_e, gradient g_e = d_joints[J + e]):
//     dA[b, j]   += w g_e (x) [v_posed[b, v_e]; 1]              over the <= ell_k joints of v_e
//     d_e         = sum_k w_k R_{j_k}^T g_e                       (d_v_posed at the pick)
//   d_feat[b, k]  = sum_{e,c} d_e[c] * pd[k, 3 v_e + c]           (ModelDev::pick_pd, gathered at create)
//   dtr[b]        = sum_e g_e
// One block per body.
struct PickBwdArgs {
  int B;
  const float* d_joints;    // (B, joints_ld); picks start at column 3J
  int joints_ld;
  const float* A;           // (B,J,12)
  const float* vsrc;        // v_posed rows or the shared template
  size_t vsrc_stride;
  float* dA;                // (B,J,12)
  float* dtr;               // (B,3)
  float* d_feat;            // (B,Kpad) or null
};

constexpr int kPickThreads = 256;
__host__ __device__ inline size_t pick_bwd_smem_bytes(int J, int E) {
  return (size_t)(2 * J * 12 + 3 * E + 4) * sizeof(float);
}

__global__ void __launch_bounds__(kPickThreads) pick_backward_kernel(const ModelDev m, const PickBwdArgs a) {
  extern __shared__ __align__(16) float pk_smem[];
  float* sA = pk_smem;                    // [J][12] transforms of this body
  float* sdA = sA + m.J * 12;             // [J][12]
  float* sd = sdA + m.J * 12;             // [3E] d_v_posed at the picks
  float* sdtr = sd + 3 * m.E;             // [3]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < m.J * 12; i += kPickThreads) {
    sA[i] = a.A[(size_t)b * m.J * 12 + i];
    sdA[i] = 0.f;
  }
  if (tid < 3) sdtr[tid] = 0.f;
  __syncthreads();
  for (int e = tid; e < m.E; e += kPickThreads) {
    const int v = m.extra_vids[e];
    const float* gp = a.d_joints + (size_t)b * a.joints_ld + 3 * (m.J + e);
    const float* xp = a.vsrc + (size_t)b * a.vsrc_stride + 3 * (size_t)v;
    const float g[3] = {gp[0], gp[1], gp[2]};
    const float x[4] = {xp[0], xp[1], xp[2], 1.f};
    float d[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < m.ell_k; ++k) {
      const float w = m.ell_w[(size_t)k * m.V + v];
      if (w == 0.f) continue;
      const int j = m.ell_idx[(size_t)k * m.V + v];
      const float* Aj = sA + j * 12;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float wg = w * g[r];
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(&sdA[j * 12 + r * 4 + c], wg * x[c]);
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = fmaf(wg, Aj[r * 4 + c], d[c]);
      }
    }
    sd[3 * e + 0] = d[0]; sd[3 * e + 1] = d[1]; sd[3 * e + 2] = d[2];
    atomicAdd(&sdtr[0], g[0]); atomicAdd(&sdtr[1], g[1]); atomicAdd(&sdtr[2], g[2]);
  }
  __syncthreads();
  for (int i = tid; i < m.J * 12; i += kPickThreads) a.dA[(size_t)b * m.J * 12 + i] = sdA[i];
  if (tid < 3) a.dtr[3 * b + tid] = sdtr[tid];
  if (a.d_feat != nullptr) {
    const int n = 3 * m.E;
    for (int k = tid; k < m.Kpad; k += kPickThreads) {
      float a0 = 0.f, a1 = 0.f;
      int i = 0;
      for (; i + 2 <= n; i += 2) {
        a0 = fmaf(sd[i], m.pick_pd[(size_t)i * m.Kpad + k], a0);
        a1 = fmaf(sd[i + 1], m.pick_pd[(size_t)(i + 1) * m.Kpad + k], a1);
      }
      if (i < n) a0 = fmaf(sd[i], m.pick_pd[(size_t)i * m.Kpad + k], a0);
      a.d_feat[(size_t)b * m.Kpad + k] = a0 + a1;
    }
  }
}

// Fused vertex L2 data term of the fitting step (config 3; squared-L2 form of
// lib/Gen_SMPLH/fitting.py:491-495 applied to vertices):  loss[b] = scale * sum ||V - V*||^2,
// grad = 2 * scale * (V - V*).  One pass over V and V* instead of ~6 elementwise torch kernels.
__global__ void __launch_bounds__(256)
vertex_l2_kernel(int n_per_body, const float* verts, const float* __restrict__ target,
                 float scale, float* grad, float* __restrict__ loss, int vec2, int loss_stride) {
  // `grad` may alias `verts` (the fused fitting node overwrites the vertices with their gradient):
  // every element is read and written by the same thread, and neither pointer is __restrict__
  __shared__ float red[8];
  const int b = blockIdx.y;
  const size_t base = (size_t)b * n_per_body;
  float acc = 0.f;
  const float s2 = 2.f * scale;
  if (vec2) {             // rows are 8-byte aligned: two floats per access
    const int n2 = n_per_body >> 1;
    const int chunk = (n2 + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * chunk, i1 = min(n2, i0 + chunk);
    const float2* v2 = reinterpret_cast<const float2*>(verts + base);
    const float2* t2 = reinterpret_cast<const float2*>(target + base);
    float2* g2 = reinterpret_cast<float2*>(grad ? grad + base : nullptr);
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
      const float2 v = v2[i], t = __ldcs(t2 + i);
      const float dx = v.x - t.x, dy = v.y - t.y;
      acc = fmaf(dx, dx, fmaf(dy, dy, acc));
      if (grad) g2[i] = make_float2(s2 * dx, s2 * dy);
    }
  } else {
    const int chunk = (n_per_body + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * chunk, i1 = min(n_per_body, i0 + chunk);
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
      const float d = verts[base + i] - __ldcs(target + base + i);
      acc = fmaf(d, d, acc);
      if (grad) grad[base + i] = s2 * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss + (size_t)b * loss_stride, scale * t);
  }
}

struct SkinBwdArgs {
  int B;
  int bodies_per_block;
  const float* dverts;   // (B,V,3)
  const float* A;        // (B,J,12)
  float* dvp_hi;         // (B,Npad) tf32-rounded d_v_posed
  float* dvp_lo;         // (B,Npad) residual
  // fp16 two-term split output (grouped kernel): rows scaled per body by the power of two row_scale[b]
  __half* h_hi;          // (B,Npad) or null
  __half* h_lo;
  const float* row_scale;   // (B) written by dA_kernel
};

template <bool kReg4>
__global__ void __launch_bounds__(kSkinThreads)
skin_backward_kernel(const ModelDev m, const SkinBwdArgs a) {
  extern __shared__ __align__(16) float sb_smem[];
  const int tid = threadIdx.x;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int nv = min(kSkinTileVerts, m.V - v0);
  const int nfloat = nv * 3;
  const int nstore = min(kSkinTileVerts * 3, m.Npad - v0 * 3);   // incl. zeroed K padding
  float* t_in = sb_smem;                              // [3072] d_verts tile
  float* t_hi = sb_smem + kSkinTileVerts * 3;         // [3072]
  float* t_lo = sb_smem + 2 * kSkinTileVerts * 3;     // [3072]
  float* As = sb_smem + 3 * kSkinTileVerts * 3;       // [J*12]
  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);

  uint32_t idx4[kSkinVPT];
  float4 w4[kSkinVPT];
  if (kReg4) {
#pragma unroll
    for (int i = 0; i < kSkinVPT; ++i) {
      const int v = v0 + tid + kSkinThreads * i;
      idx4[i] = v < m.V ? m.skin_idx4[v] : 0u;
      w4[i] = v < m.V ? m.skin_w4[v] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int b = b0; b < b1; ++b) {
    const float* src = a.dverts + (size_t)b * m.V * 3 + (size_t)v0 * 3;
    for (int c = tid; c < kSkinTileVerts * 3; c += kSkinThreads) {
      t_in[c] = c < nfloat ? __ldcs(src + c) : 0.f;
      t_hi[c] = 0.f;
      t_lo[c] = 0.f;
    }
    const float* asrc = a.A + (size_t)b * m.J * 12;
    for (int c = tid; c < m.J * 12; c += kSkinThreads) As[c] = asrc[c];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSkinVPT; ++i) {
      const int lv = tid + kSkinThreads * i;
      if (lv < nv) {
        const float gx = t_in[3 * lv + 0], gy = t_in[3 * lv + 1], gz = t_in[3 * lv + 2];
        float T[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) T[q] = 0.f;
        if (kReg4) {
          const float wk[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = (idx4[i] >> (8 * k)) & 0xff;
            const float* Aj = As + j * 12;
            const float w = wk[k];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              T[3 * r + 0] = fmaf(w, Aj[4 * r + 0], T[3 * r + 0]);
              T[3 * r + 1] = fmaf(w, Aj[4 * r + 1], T[3 * r + 1]);
              T[3 * r + 2] = fmaf(w, Aj[4 * r + 2], T[3 * r + 2]);
            }
          }
        } else {
          const int v = v0 + lv;
          for (int k = 0; k < m.ell_k; ++k) {
            const float w = m.ell_w[(size_t)k * m.V + v];
            if (w != 0.f) {
              const float* Aj = As + m.ell_idx[(size_t)k * m.V + v] * 12;
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                T[3 * r + 0] = fmaf(w, Aj[4 * r + 0], T[3 * r + 0]);
                T[3 * r + 1] = fmaf(w, Aj[4 * r + 1], T[3 * r + 1]);
                T[3 * r + 2] = fmaf(w, Aj[4 * r + 2], T[3 * r + 2]);
              }
            }
          }
        }
        // d_v_posed = T_R^T g
        const float ox = T[0] * gx + T[3] * gy + T[6] * gz;
        const float oy = T[1] * gx + T[4] * gy + T[7] * gz;
        const float oz = T[2] * gx + T[5] * gy + T[8] * gz;
        const float hx = ptx::tf32_round(ox), hy = ptx::tf32_round(oy), hz = ptx::tf32_round(oz);
        t_hi[3 * lv + 0] = hx; t_hi[3 * lv + 1] = hy; t_hi[3 * lv + 2] = hz;
        t_lo[3 * lv + 0] = ox - hx; t_lo[3 * lv + 1] = oy - hy; t_lo[3 * lv + 2] = oz - hz;
      }
    }
    __syncthreads();
    float4* oh = reinterpret_cast<float4*>(a.dvp_hi + (size_t)b * m.Npad + (size_t)v0 * 3);
    float4* ol = reinterpret_cast<float4*>(a.dvp_lo + (size_t)b * m.Npad + (size_t)v0 * 3);
    for (int c = tid; c < (nstore >> 2); c += kSkinThreads) {
      oh[c] = reinterpret_cast<const float4*>(t_hi)[c];
      ol[c] = reinterpret_cast<const float4*>(t_lo)[c];
    }
    __syncthreads();
  }
}

// Grouped / warp-streamed version of skin_backward_kernel (same structure as skin_grouped_kernel):
// each warp streams its 128 vertices of d_verts through a private cp.async ring (8-byte copies:
// rows of (B,V,3) are only 8-byte aligned), a thread owns 4 consecutive vertices and fetches each
// of the group's <= 8 distinct transforms once, d_v_posed = sum_u w_u R_u^T g is written as TF32
// hi/lo rows with 16-byte coalesced stores (pad columns up to Npad are zeroed: they are the K
// padding of the backward GEMM).
constexpr int kSkinBwdStages = 3;

template <int kStages, bool kHalfOut>
__global__ void __launch_bounds__(kGrpThreads, 2)
skin_backward_grouped_kernel(const ModelDev m, const SkinBwdArgs a) {
  extern __shared__ __align__(16) float sbg_smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * kSkinTileVerts;
  const int a_floats = m.J * 12;
  const int a_chunks = m.J * 3;
  const int a_pad = grp_a_pad(m.J);
  float* ring = sbg_smem;                                        // [stages][3072] g in, hi out
  float* lobuf = sbg_smem + kStages * kSkinTileVerts * 3;        // [3072] lo out
  float* Abuf = lobuf + kSkinTileVerts * 3;                      // [2][8][a_pad]
  const int b0 = blockIdx.y * a.bodies_per_block;
  const int b1 = min(a.B, b0 + a.bodies_per_block);
  if (b0 >= b1) return;

  const int wf0 = v0 * 3 + warp * kWarpFloats;
  const int n_in = max(0, min(kWarpFloats, m.V * 3 - wf0));      // valid input floats
  const int n_out = max(0, min(kWarpFloats, m.Npad - wf0));      // output floats incl. zero padding
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
  // never-copied tail floats of the ring must be finite (they are multiplied by zero weights)
  for (int i = tid; i < kStages * kSkinTileVerts * 3; i += kGrpThreads) ring[i] = 0.f;
  __syncthreads();

  auto issue_A = [&](int grp) {
    const int bb0 = b0 + grp * kGrpABodies;
    const int nb = min(kGrpABodies, b1 - bb0);
    float* dstA = Abuf + (grp & 1) * kGrpABodies * a_pad;
    for (int c = tid; c < nb * a_chunks; c += kGrpThreads) {
      const int bi = c / a_chunks, cc = c - bi * a_chunks;
      ptx::cp_async_16(dstA + bi * a_pad + 4 * cc, a.A + (size_t)(bb0 + bi) * a_floats + 4 * cc);
    }
  };
  auto issue_g = [&](int b) {
    if (b < b1) {
      const float* src = a.dverts + (size_t)b * m.V * 3 + wf0;
      float* dst = ring + ((b - b0) % kStages) * (kSkinTileVerts * 3) + warp * kWarpFloats;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int c = lane + 32 * i;                       // 8-byte chunk of the warp slice
        if (2 * c + 2 <= n_in) ptx::cp_async_8(dst + 2 * c, src + 2 * c);
      }
      if ((n_in & 1) && lane == 0) dst[n_in - 1] = src[n_in - 1];
    }
    ptx::cp_async_commit();
  };

  issue_A(0);
#pragma unroll
  for (int i = 0; i < kStages - 1; ++i) issue_g(b0 + i);

  for (int b = b0; b < b1; ++b) {
    const int rel = b - b0;
    const int agrp = rel / kGrpABodies;
    issue_g(b + kStages - 1);
    ptx::cp_async_wait<kStages - 1>();
    if ((rel % kGrpABodies) == 0) {
      __syncthreads();
      if (b + kGrpABodies < b1) issue_A(agrp + 1);
    } else {
      __syncwarp();
    }
    const float* Ab = Abuf + ((agrp & 1) * kGrpABodies + (rel % kGrpABodies)) * a_pad;
    float* slot = ring + (rel % kStages) * (kSkinTileVerts * 3) + warp * kWarpFloats;
    float* lo = lobuf + warp * kWarpFloats;
    float4* mine = reinterpret_cast<float4*>(slot) + 3 * lane;
    const float4 c0 = mine[0], c1 = mine[1], c2 = mine[2];
    const float gx[4] = {c0.x, c0.w, c1.z, c2.y};
    const float gy[4] = {c0.y, c1.x, c1.w, c2.z};
    const float gz[4] = {c0.z, c1.y, c2.x, c2.w};
    float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f}, oz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < kGrpJoints; ++u) {
      if (used & (1u << u)) {
        const int j = ((u < 4 ? jid.x : jid.y) >> (8 * (u & 3))) & 0xff;
        const float4* Aj = reinterpret_cast<const float4*>(Ab + j * 12);
        const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
        const float wu[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // R_u^T g
          const float px = fmaf(r0.x, gx[i], fmaf(r1.x, gy[i], r2.x * gz[i]));
          const float py = fmaf(r0.y, gx[i], fmaf(r1.y, gy[i], r2.y * gz[i]));
          const float pz = fmaf(r0.z, gx[i], fmaf(r1.z, gy[i], r2.z * gz[i]));
          ox[i] = fmaf(wu[i], px, ox[i]);
          oy[i] = fmaf(wu[i], py, oy[i]);
          oz[i] = fmaf(wu[i], pz, oz[i]);
        }
      }
    }
    float o[12] = {ox[0], oy[0], oz[0], ox[1], oy[1], oz[1], ox[2], oy[2], oz[2], ox[3], oy[3], oz[3]};
    if (kHalfOut) {
      // fp16 two-term split of the row scaled into the fp16 range (|d_v_posed| <= sqrt(3) max|g|);
      // a thread's 12 outputs are 24 contiguous bytes per term and the warp's 768 bytes are
      // contiguous, so they leave straight from registers as 8-byte stores
      const float sc = a.row_scale[b];
      uint2 hq[3], lq[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        __half hh[4], ll[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = 4 * i + k;
          const bool ok = 4 * g + idx / 3 < m.V;
          const float x = ok ? o[idx] * sc : 0.f;
          hh[k] = __float2half_rn(x);
          ll[k] = __float2half_rn(x - __half2float(hh[k]));
        }
        const __half2 h01 = __halves2half2(hh[0], hh[1]), h23 = __halves2half2(hh[2], hh[3]);
        const __half2 l01 = __halves2half2(ll[0], ll[1]), l23 = __halves2half2(ll[2], ll[3]);
        hq[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        lq[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      __syncwarp();                                  // everyone has read its inputs from the slot
      uint2* gh = reinterpret_cast<uint2*>(a.h_hi + (size_t)b * m.Npad + wf0) + 3 * lane;
      uint2* gl = reinterpret_cast<uint2*>(a.h_lo + (size_t)b * m.Npad + wf0) + 3 * lane;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (12 * lane + 4 * i + 4 <= n_out) {
          gh[i] = hq[i];
          gl[i] = lq[i];
        }
      }
    } else {
    float h[12], l[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const bool ok = 4 * g + i / 3 < m.V;           // zero (not garbage*0) beyond the last vertex
      const float x = ok ? o[i] : 0.f;
      h[i] = ptx::tf32_round(x);
      l[i] = x - h[i];
    }
    __syncwarp();                                    // everyone has read its inputs from the slot
    float4* hs = reinterpret_cast<float4*>(slot) + 3 * lane;
    float4* ls = reinterpret_cast<float4*>(lo) + 3 * lane;
    hs[0] = make_float4(h[0], h[1], h[2], h[3]); hs[1] = make_float4(h[4], h[5], h[6], h[7]);
    hs[2] = make_float4(h[8], h[9], h[10], h[11]);
    ls[0] = make_float4(l[0], l[1], l[2], l[3]); ls[1] = make_float4(l[4], l[5], l[6], l[7]);
    ls[2] = make_float4(l[8], l[9], l[10], l[11]);
    __syncwarp();
    float4* oh = reinterpret_cast<float4*>(a.dvp_hi + (size_t)b * m.Npad + wf0);
    float4* ol = reinterpret_cast<float4*>(a.dvp_lo + (size_t)b * m.Npad + wf0);
    const float4* sh4 = reinterpret_cast<const float4*>(slot);
    const float4* sl4 = reinterpret_cast<const float4*>(lo);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = lane + 32 * i;
      if (4 * c + 4 <= n_out) {
        oh[c] = sh4[c];
        ol[c] = sl4[c];
      }
    }
    __syncwarp();
    }   // !kHalfOut
  }
}

// dA[b,j] (3x4) = sum over the vertices bound to joint j of  w * g (x) [v_posed; 1];
// also dtr[b] = sum_v g[b,v].  One block per body; warps walk the joint -> vertex (CSC) lists.
struct DAArgs {
  int B;
  const float* dverts;    // (B,V,3)
  const float* vsrc;      // v_posed rows or shared template
  size_t vsrc_stride;
  float* dA;              // (B,J,12)
  float* dtr;             // (B,3)
  float* row_scale;       // (B) or null: power of two s with sqrt(3) max|g[b]| * s in [2^12, 2^13] (fp16 backward GEMM)
  float* row_scale_inv;   // (B) or null: 1 / s
};

#ifndef SMPLK_DA_THREADS
#define SMPLK_DA_THREADS 256
#endif
constexpr int kDAThreads = SMPLK_DA_THREADS;   // one body per block; fewer, fatter blocks keep a body's rows in L1

__global__ void __launch_bounds__(kDAThreads) dA_kernel(const ModelDev m, const DAArgs a) {
  __shared__ float red[4][kDAThreads / 32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* g = a.dverts + (size_t)b * m.V * 3;
  const float* vp = a.vsrc + (size_t)b * a.vsrc_stride;
  // gridDim.y blocks share a body's joints (small batches: one block per body leaves the GPU empty)
  for (int j = warp + (kDAThreads / 32) * blockIdx.y; j < m.J; j += (kDAThreads / 32) * gridDim.y) {
    float acc[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = 0.f;
    // 4 list entries per lane per trip: the index loads, then all 8 gathers, are independent
    // (the plain one-entry loop was bound by two dependent L2 round trips per entry)
    const int beg = m.csc_ptr[j], end = m.csc_ptr[j + 1];
    for (int n0 = beg + lane; n0 < end; n0 += 128) {
      int vv[4];
      float ww[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + 32 * u;
        const bool ok = n < end;
        vv[u] = ok ? m.csc_vert[n] : 0;
        ww[u] = ok ? m.csc_w[n] : 0.f;
      }
      float gq[4][3], xq[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          gq[u][c] = g[3 * vv[u] + c];
          xq[u][c] = vp[3 * vv[u] + c];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float gx = ww[u] * gq[u][0], gy = ww[u] * gq[u][1], gz = ww[u] * gq[u][2];
        const float x = xq[u][0], y = xq[u][1], z = xq[u][2];
        acc[0] = fmaf(gx, x, acc[0]); acc[1] = fmaf(gx, y, acc[1]); acc[2] = fmaf(gx, z, acc[2]); acc[3] += gx;
        acc[4] = fmaf(gy, x, acc[4]); acc[5] = fmaf(gy, y, acc[5]); acc[6] = fmaf(gy, z, acc[6]); acc[7] += gy;
        acc[8] = fmaf(gz, x, acc[8]); acc[9] = fmaf(gz, y, acc[9]); acc[10] = fmaf(gz, z, acc[10]); acc[11] += gz;
      }
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    if (lane == 0) {
      float* o = a.dA + ((size_t)b * m.J + j) * 12;
#pragma unroll
      for (int q = 0; q < 12; ++q) o[q] = acc[q];
    }
  }
  if (blockIdx.y != 0) return;
  // translation gradient: plain sum of the vertex gradients
  float s[3] = {0.f, 0.f, 0.f};
  float gmax = 0.f;
  for (int v = threadIdx.x; v < m.V; v += kDAThreads) {
    const float g0 = g[3 * v + 0], g1 = g[3 * v + 1], g2 = g[3 * v + 2];
    s[0] += g0; s[1] += g1; s[2] += g2;
    gmax = fmaxf(gmax, fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if (lane == 0) red[3][warp] = gmax;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    if (lane == 0) red[q][warp] = s[q];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < kDAThreads / 32; ++w) t += red[threadIdx.x][w];
    a.dtr[3 * b + threadIdx.x] = t;
  }
  if (threadIdx.x == 3 && a.row_scale != nullptr) {
    float mx = 0.f;
    for (int w = 0; w < kDAThreads / 32; ++w) mx = fmaxf(mx, red[3][w]);
    float sc = 1.f;
    if (mx > 0.f && isfinite(mx)) {
      int e;
      frexpf(mx * 1.7320508f, &e);          // mx * sqrt(3) in [2^(e-1), 2^e)
      sc = ldexpf(1.f, max(-100, min(100, 13 - e)));
    }
    a.row_scale[b] = sc;
    a.row_scale_inv[b] = 1.f / sc;
  }
}

// ------------------------------------------------------------------------------------------
// dA_seg_kernel -- the transform gradient of the fitting step, joint-major like dA_kernel but with the
// list traversal turned inside out: a warp owns ONE segment (<= 128 entries) of one joint's vertex
// list for a RUN of bodies.  Its 4 entries per lane (vertex offset, weight) are loaded once and stay in
// registers, so a body costs no dependent index hop (dA_kernel: index list, then gathers -- two global
// round trips per 128 entries, long-scoreboard stall 8.5 per issued instruction), and the gathers of
// body b+1 are in flight while body b is reduced.  The 12 sums of a body are reduced with a halving
// butterfly over 16 values (16 shuffles instead of 60) and written as per-segment partials
// dAp[b][s][12]; the pose backward kernel adds a joint's (adjacent) segments in fixed order (bit-reproducible)
// and forms d_transl = sum_j dA[j][:,3] (valid because every vertex's weights sum to 1: checked at model
// create).  A separate reduction launch cost 9 us for this tiny sum.
// ------------------------------------------------------------------------------------------
constexpr int kDASeg = 128;          // list entries per segment: 4 per lane
constexpr int kDASegWarps = 4;

struct DASegArgs {
  int B;
  int bodies_per_warp;
  const float* dverts;    // (B,V,3)
  const float* vsrc;      // v_posed rows
  size_t vsrc_stride;
  float* dAp;             // (B, S, 12) per-segment partial sums
};

__device__ __forceinline__ float halve_exchange(bool upper, float lo, float hi, int mask) {
  // lanes of the upper half keep `hi` and hand out `lo`; the lower half the other way round
  const float send = upper ? lo : hi;
  const float keep = upper ? hi : lo;
  return keep + __shfl_xor_sync(0xffffffffu, send, mask);
}

template <bool kPair>
__global__ void __launch_bounds__(kDASegWarps * 32) dA_seg_kernel(const ModelDev m, const DASegArgs a) {
  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kDASegWarps + warp;
  if (s >= m.seg_count) return;
  const int beg = m.seg_beg[s], len = m.seg_len[s];
  int off[4];
  float w[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = 32 * u + lane;
    const bool ok = i < len;
    off[u] = ok ? 3 * m.csc_vert[beg + i] : 0;
    w[u] = ok ? m.csc_w[beg + i] : 0.f;
  }
  const int b0 = blockIdx.y * a.bodies_per_warp;
  const int b1 = min(a.B, b0 + a.bodies_per_warp);
  if (b0 >= b1) return;
  ptx::pdl_wait();             // the segment tables above are constants; gradients and v_posed come from earlier kernels
  float gq[4][3], xq[4][3];
  // A vertex's three floats sit at float offset 3v: an 8-byte load of the aligned pair (x, y for an even, y, z for an
  // odd vertex) + a 4-byte load of the third, chosen branch-free -- 4 gathers per entry instead of 6 (the kernel is
  // bound by L1 wavefronts of these 12-byte-stride gathers).  Needs 8-byte aligned rows (kPair; else three 4-byte loads).
  auto load3 = [&](const float* row, int o, float (&v)[3]) {
    if (kPair) {
      const bool odd = o & 1;
      const float2 pr = *reinterpret_cast<const float2*>(row + o + (odd ? 1 : 0));
      const float s1 = row[o + (odd ? 0 : 2)];
      v[0] = odd ? s1 : pr.x; v[1] = odd ? pr.x : pr.y; v[2] = odd ? pr.y : s1;
    } else {
      v[0] = row[o]; v[1] = row[o + 1]; v[2] = row[o + 2];
    }
  };
  auto load = [&](int b, float (&g)[4][3], float (&x)[4][3]) {
    const float* gp = a.dverts + (size_t)b * m.V * 3;
    const float* vp = a.vsrc + (size_t)b * a.vsrc_stride;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load3(gp, off[u], g[u]);
      load3(vp, off[u], x[u]);
    }
  };
  load(b0, gq, xq);
  for (int b = b0; b < b1; ++b) {
    float gn[4][3], xn[4][3];
    load(min(b + 1, b1 - 1), gn, xn);              // next body's gathers fly during this body's reduction
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gx = w[u] * gq[u][0], gy = w[u] * gq[u][1], gz = w[u] * gq[u][2];
      const float x = xq[u][0], y = xq[u][1], z = xq[u][2];
      acc[0] = fmaf(gx, x, acc[0]); acc[1] = fmaf(gx, y, acc[1]); acc[2] = fmaf(gx, z, acc[2]); acc[3] += gx;
      acc[4] = fmaf(gy, x, acc[4]); acc[5] = fmaf(gy, y, acc[5]); acc[6] = fmaf(gy, z, acc[6]); acc[7] += gy;
      acc[8] = fmaf(gz, x, acc[8]); acc[9] = fmaf(gz, y, acc[9]); acc[10] = fmaf(gz, z, acc[10]); acc[11] += gz;
    }
    // halving butterfly: 16 -> 8 -> 4 -> 2 -> 1 values per lane, then the pair (lane, lane ^ 1)
    float r8[8], r4[4], r2[2];
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r8[i] = halve_exchange(u16, acc[i], acc[i + 8], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) r4[i] = halve_exchange(u8, r8[i], r8[i + 4], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) r2[i] = halve_exchange(u4, r4[i], r4[i + 2], 4);
    float r1 = halve_exchange(u2, r2[0], r2[1], 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
    // this lane now holds component  8 [lane&16] + 4 [lane&8] + 2 [lane&4] + [lane&2]
    const int comp = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    if (!(lane & 1) && comp < 12) a.dAp[((size_t)b * m.seg_count + s) * 12 + comp] = r1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { gq[u][c] = gn[u][c]; xq[u][c] = xn[u][c]; }
    }
  }
}

// d r  from  dL/dR for R = rodrigues(r)  (same parametrisation as the forward).
__device__ __forceinline__ void rodrigues_backward(const float* r, const float* dR, float* dr) {
  const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  const float a = sqrtf(ex * ex + ey * ey + ez * ez);
  const float inv = 1.0f / a;
  const float nx = r[0] * inv, ny = r[1] * inv, nz = r[2] * inv;
  float s, c;
  sincosf(a, &s, &c);
  const float sh = sinf(0.5f * a);
  const float omc = 2.0f * sh * sh;
  const float tr = dR[0] + dR[4] + dR[8];
  const float kx = dR[7] - dR[5], ky = dR[2] - dR[6], kz = dR[3] - dR[1];
  const float gK = nx * kx + ny * ky + nz * kz;
  const float sx = 2.f * dR[0] * nx + (dR[1] + dR[3]) * ny + (dR[2] + dR[6]) * nz;
  const float sy = (dR[1] + dR[3]) * nx + 2.f * dR[4] * ny + (dR[5] + dR[7]) * nz;
  const float sz = (dR[2] + dR[6]) * nx + (dR[5] + dR[7]) * ny + 2.f * dR[8] * nz;
  const float nn = nx * nx + ny * ny + nz * nz;
  const float gK2 = 0.5f * (sx * nx + sy * ny + sz * nz) - nn * tr;
  const float dnx = s * kx + omc * (sx - 2.f * tr * nx);
  const float dny = s * ky + omc * (sy - 2.f * tr * ny);
  const float dnz = s * kz + omc * (sz - 2.f * tr * nz);
  const float da = c * gK + s * gK2 - (dnx * r[0] + dny * r[1] + dnz * r[2]) * inv * inv;
  dr[0] = dnx * inv + da * ex * inv;
  dr[1] = dny * inv + da * ey * inv;
  dr[2] = dnz * inv + da * ez * inv;
}

// 0 (product): the backward chain walk of pose_backward_kernel runs one lane per joint; 1: one lane per gradient element
// (A/B build, bitwise-equal gradients).  Measured on B200 at 1,024 bodies (profiles/r02_pose_bwd_walk_ab.txt): the
// element walk is SLOWER, 0.0376 against 0.0334 ms -- a third of the instructions, but two short LDS -> FMA -> STS phases
// and two warp syncs per level instead of one long phase with 12 independent FMA chains per lane; the kernel is bound
// by the latency of its dependent steps (issue-active 26 %, 7 warps per SM), not by instruction issue.
#ifndef SMPLK_POSE_BWD_ELEMWALK
#define SMPLK_POSE_BWD_ELEMWALK 0
#endif

struct PoseBwdArgs {
  int B;
  const float* betas;
  int betas_B;
  const float* pose;
  const float* pca_l;
  const float* pca_r;
  int add_mean;
  const float* dA;           // (B,J,12)
  const float* d_joints;     // (B,joints_ld) or null; first 3J entries = FK joints
  int joints_ld;
  const float* d_feat;       // [splits][split_stride] partial rows of Kpad floats, or null
  int feat_splits;
  size_t feat_split_stride;  // floats
  const float* dtr_verts;    // (B,3) or null
  const float* dAp;          // (B,S,12) per-segment partials of dA_seg_kernel, or null (then dA is read);
                             // with dAp the vertex part of d_transl is sum_j dA[j][:,3] and dtr_verts is null
  float* d_betas;            // (betas_B,NB) or null (atomic accumulate when betas_B == 1)
  float* d_pose;             // (B,3J) or null
  float* d_pca_l;            // (B,C) or null
  float* d_pca_r;
  float* d_transl;           // (B,3) or null
  const float* d_loss;       // (B) or null: scale of body b's parameter gradients
  int d_loss_stride;         // 1, or 0 when d_loss is one float for all bodies
  const float* d_full_pose;  // (B,3J) or null: gradient w.r.t. the assembled pose output
  int staged_segs;           // seg_count when dAp is staged in shared memory (the launch sized it), else 0
};

// smem floats per warp of pose_backward_kernel: world G, local L, dG (12 each), dR (9), dJ, dfull (3 each)
// + the summed blend-GEMM gradient row (Kpad)
// + (segs > 0) the body's per-segment partials of dA_seg_kernel, staged by cp.async while the forward is recomputed
__host__ __device__ inline int pose_bwd_smem_floats(int J, int Kpad, int segs) {
  return ((J * 51 + 3) & ~3) + Kpad + segs * 12;
}

// Split-K partials of the backward blend GEMM, d_feat[sp][row][k], summed into split 0 (in place).
// One thread per (row, k): the loads of a thread are independent and coalesced across the warp, so
// the up-to-74 partials of a small batch cost a few L2 round trips instead of 74 dependent ones in
// the pose kernel's single warp per body.
__global__ void __launch_bounds__(256)
reduce_splits_kernel(int n, int splits, size_t stride, float* __restrict__ d_feat) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int sp = 0;
  for (; sp + 4 <= splits; sp += 4) {
    a0 += d_feat[(size_t)sp * stride + i];
    a1 += d_feat[(size_t)(sp + 1) * stride + i];
    a2 += d_feat[(size_t)(sp + 2) * stride + i];
    a3 += d_feat[(size_t)(sp + 3) * stride + i];
  }
  for (; sp < splits; ++sp) a0 += d_feat[(size_t)sp * stride + i];
  d_feat[i] = (a0 + a1) + (a2 + a3);
}

template <int SLOTS>
__global__ void __launch_bounds__(kPoseWarps * 32)
pose_backward_kernel(const ModelDev m, const PoseBwdArgs a) {
  extern __shared__ __align__(16) float pb_smem[];
  // the skeleton's index tables, staged once per block: the level walks otherwise chain three or
  // four dependent global loads per level (L2 latency each when a single body is fitted)
  __shared__ int s_par[kMaxJoints], s_ord[kMaxJoints], s_lvl[kMaxJoints + 2], s_cptr[kMaxJoints + 1], s_cidx[kMaxJoints];
  ptx::pdl_launch_dependents();
  for (int i = threadIdx.x; i < m.J; i += blockDim.x) {
    s_par[i] = m.parents[i]; s_ord[i] = m.order[i]; s_cptr[i] = m.child_ptr[i];
    if (i < m.J - 1) s_cidx[i] = m.child_idx[i];
  }
  if (threadIdx.x == 0) s_cptr[m.J] = m.child_ptr[m.J];
  for (int i = threadIdx.x; i < m.max_depth + 2; i += blockDim.x) s_lvl[i] = m.level_start[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kPoseWarps + warp;
  if (b >= a.B) return;
  float* Gw = pb_smem + warp * pose_bwd_smem_floats(m.J, m.Kpad, a.staged_segs);   // [J][12] world transforms
  float* Lc = Gw + m.J * 12;                                // [J][12] local [R | Jrel]
  float* dG = Lc + m.J * 12;                                // [J][12]
  float* dRs = dG + m.J * 12;                               // [J][9]
  float* dJ = dRs + m.J * 9;                                // [J][3]
  float* dfull = dJ + m.J * 3;                              // [3J]
  const float* betas_row = a.betas ? a.betas + (size_t)(a.betas_B == 1 ? 0 : b) * m.NB : nullptr;
  // this body's gradient inputs -> smem by cp.async, in flight while the forward is recomputed: the blend GEMM's row
  // and the per-segment partials of dA (read straight from L2 the latter were up to ~10 dependent round trips per lane)
  float* dfeat_sum = Gw + ((m.J * 51 + 3) & ~3);
  float* sdA = dfeat_sum + m.Kpad;
  const bool feat_staged = a.d_feat != nullptr && a.feat_splits == 1;

  // ---- forward recompute (same walk as pose_forward_kernel; cheaper than saving it).  Its first half -- pose
  // assembly, Rodrigues, rest joints: the caller's read-only inputs and constant tables only -- runs BEFORE the
  // programmatic-launch wait, i.e. beside the split reduction and the tail of the backward GEMM.
  float rv[SLOTS][3], Jr[SLOTS][3];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    rv[s][0] = rv[s][1] = rv[s][2] = 0.f;
    Jr[s][0] = Jr[s][1] = Jr[s][2] = 0.f;
    if (j < m.J) {
      float R[9];
      load_joint_pose(m, a.pose, a.pca_l, a.pca_r, a.add_mean, b, j, rv[s]);
      rodrigues(rv[s][0], rv[s][1], rv[s][2], R);
      rest_joint(m, betas_row, j, Jr[s]);
      float4* l4 = reinterpret_cast<float4*>(Lc + j * 12);
      l4[0] = make_float4(R[0], R[1], R[2], Jr[s][0]);
      l4[1] = make_float4(R[3], R[4], R[5], Jr[s][1]);
      l4[2] = make_float4(R[6], R[7], R[8], Jr[s][2]);
    }
  }
  ptx::pdl_wait();
  // this body's gradient inputs -> smem by cp.async, in flight while the chain is walked
  if (feat_staged) {
    const float* f = a.d_feat + (size_t)b * m.Kpad;
    for (int k = 4 * lane; k < m.Kpad; k += 128) ptx::cp_async_16(dfeat_sum + k, f + k);
  }
  if (a.staged_segs > 0) {
    const float* src = a.dAp + (size_t)b * m.seg_count * 12;
    for (int i = 4 * lane; i < a.staged_segs * 12; i += 128) ptx::cp_async_16(sdA + i, src + i);
  }
  ptx::cp_async_commit();
  __syncwarp();
  float pj[SLOTS][3];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    pj[s][0] = pj[s][1] = pj[s][2] = 0.f;
    if (j >= 1 && j < m.J) {
      const float* lp = Lc + s_par[j] * 12;
      pj[s][0] = lp[3]; pj[s][1] = lp[7]; pj[s][2] = lp[11];
    }
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      if (j >= 1) {
        Lc[j * 12 + 3] = Jr[s][0] - pj[s][0];
        Lc[j * 12 + 7] = Jr[s][1] - pj[s][1];
        Lc[j * 12 + 11] = Jr[s][2] - pj[s][2];
      }
      const float4* l4 = reinterpret_cast<const float4*>(Lc + j * 12);
      float4* g4 = reinterpret_cast<float4*>(Gw + j * 12);
      g4[0] = l4[0]; g4[1] = l4[1]; g4[2] = l4[2];
    }
  }
  __syncwarp();
  for (int d = 1; d <= m.max_depth; ++d) {
    const int l0 = s_lvl[d], l1 = s_lvl[d + 1];
    for (int i = l0 + lane; i < l1; i += 32) {
      const int j = s_ord[i];
      const float* Pm = Gw + s_par[j] * 12;
      float out[12];
      affine_mul(Pm, Lc + j * 12, out);
#pragma unroll
      for (int q = 0; q < 12; ++q) Gw[j * 12 + q] = out[q];
    }
    __syncwarp();
  }

  // ---- seed: A_j = [G_R | G_t - G_R J_j], joints_fk = G_t
  float dtr_acc[3] = {0.f, 0.f, 0.f};     // with dAp: sum_j dA[j][:,3] = the vertex part of d_transl (weights sum to 1 per vertex)
  ptx::cp_async_wait<0>();
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      float gl[12];
      if (a.dAp != nullptr) {     // per-segment partials of dA_seg_kernel: a joint's segments are adjacent, summed in order
        const float4* p4 = a.staged_segs > 0
            ? reinterpret_cast<const float4*>(sdA + m.joint_seg_ptr[j] * 12)
            : reinterpret_cast<const float4*>(a.dAp + ((size_t)b * m.seg_count + m.joint_seg_ptr[j]) * 12);
        const int ns = m.joint_seg_ptr[j + 1] - m.joint_seg_ptr[j];
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
        for (int q = 0; q < ns; ++q) {
          const float4 t0 = p4[3 * q], t1 = p4[3 * q + 1], t2 = p4[3 * q + 2];
          s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
          s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
          s2.x += t2.x; s2.y += t2.y; s2.z += t2.z; s2.w += t2.w;
        }
        gl[0] = s0.x; gl[1] = s0.y; gl[2] = s0.z; gl[3] = s0.w; gl[4] = s1.x; gl[5] = s1.y; gl[6] = s1.z; gl[7] = s1.w;
        gl[8] = s2.x; gl[9] = s2.y; gl[10] = s2.z; gl[11] = s2.w;
        dtr_acc[0] += gl[3]; dtr_acc[1] += gl[7]; dtr_acc[2] += gl[11];
      } else {
        const float* gsrc = a.dA + ((size_t)b * m.J + j) * 12;
#pragma unroll
        for (int q = 0; q < 12; ++q) gl[q] = gsrc[q];
      }
      const float* g = gl;
      const float* G = Gw + j * 12;
      float gt[3] = {g[3], g[7], g[11]};
      float djr[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          dG[j * 12 + r * 4 + c] = g[r * 4 + c] - gt[r] * Jr[s][c];
          djr[c] -= G[r * 4 + c] * gt[r];
        }
      }
      if (a.d_joints) {
        const float* gj = a.d_joints + (size_t)b * a.joints_ld + 3 * j;
        gt[0] += gj[0]; gt[1] += gj[1]; gt[2] += gj[2];
      }
      dG[j * 12 + 3] = gt[0]; dG[j * 12 + 7] = gt[1]; dG[j * 12 + 11] = gt[2];
      dJ[j * 3 + 0] = djr[0]; dJ[j * 3 + 1] = djr[1]; dJ[j * 3 + 2] = djr[2];
    }
  }
  __syncwarp();

  // ---- chain backward, deepest level first:  G_j = G_p [R_j | Jrel_j].  No atomics: a joint first
  // gathers what its children (one level down, already final) send up, then derives its own dR and
  // dJrel from the now-final dG_j (the shared-memory float atomics this replaces were CAS loops and
  // 45 % of the kernel's stall samples).
#if SMPLK_POSE_BWD_ELEMWALK
  // A/B variant (see the macro's note): one lane per ELEMENT (joint of the level, row r, column k) of the 3x4 gradient
  // instead of one lane per joint; a level is ceil(12 n / 32) rounds.  Every element keeps the summation order of the
  // lane-per-joint version (bitwise-equal gradients).
  for (int d = m.max_depth; d >= 0; --d) {
    const int l0 = s_lvl[d], n12 = (s_lvl[d + 1] - l0) * 12;
    // phase A: what the children (one level down, final) send up -> dG_j, and their part of dJ_j
    for (int it = lane; it < n12; it += 32) {
      const int i = it / 12, e = it - 12 * i, r = e >> 2, k = e & 3;
      const int j = s_ord[l0 + i];
      const int c0 = s_cptr[j], c1 = s_cptr[j + 1];
      if (c1 > c0) {
        const float* G = Gw + j * 12;
        float acc = dG[j * 12 + e];
        float djr = (k == 3) ? dJ[j * 3 + r] : 0.f;
        for (int n = c0; n < c1; ++n) {
          const int c = s_cidx[n];
          const float* cg = dG + c * 12;
          if (k < 3) {
            const float* L = Lc + c * 12 + k * 4;
            acc += cg[r * 4 + 0] * L[0] + cg[r * 4 + 1] * L[1] + cg[r * 4 + 2] * L[2] + cg[r * 4 + 3] * L[3];
          } else {
            acc += cg[r * 4 + 3];
            djr -= G[0 * 4 + r] * cg[3] + G[1 * 4 + r] * cg[7] + G[2 * 4 + r] * cg[11];   // Jrel_c = J_c - J_j
          }
        }
        dG[j * 12 + e] = acc;
        if (k == 3) dJ[j * 3 + r] = djr;
      }
    }
    __syncwarp();
    // phase B: dR_j and the rest of dJ_j from the now-final dG_j
    for (int it = lane; it < n12; it += 32) {
      const int i = it / 12, e = it - 12 * i, r = e >> 2, k = e & 3;
      const int j = s_ord[l0 + i];
      const float* dg = dG + j * 12;
      if (d >= 1) {
        const float* Pm = Gw + s_par[j] * 12;
        const float v = Pm[0 * 4 + r] * dg[0 * 4 + k] + Pm[1 * 4 + r] * dg[1 * 4 + k] + Pm[2 * 4 + r] * dg[2 * 4 + k];
        if (k < 3) dRs[j * 9 + r * 3 + k] = v;
        else dJ[j * 3 + r] += v;
      } else {          // root: L_0 = [R_0 | J_0]
        if (k < 3) dRs[j * 9 + r * 3 + k] = dg[r * 4 + k];
        else dJ[j * 3 + r] += dg[r * 4 + 3];
      }
    }
    __syncwarp();
  }
#else
  for (int d = m.max_depth; d >= 0; --d) {
    const int l0 = s_lvl[d], l1 = s_lvl[d + 1];
    for (int i = l0 + lane; i < l1; i += 32) {
      const int j = s_ord[i];
      const float* G = Gw + j * 12;
      float dg[12], dj[3];
#pragma unroll
      for (int q = 0; q < 12; ++q) dg[q] = dG[j * 12 + q];
#pragma unroll
      for (int r = 0; r < 3; ++r) dj[r] = dJ[j * 3 + r];
      const int c0 = s_cptr[j], c1 = s_cptr[j + 1];
      for (int n = c0; n < c1; ++n) {
        const int c = s_cidx[n];
        const float* L = Lc + c * 12;
        float cg[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) cg[q] = dG[c * 12 + q];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            dg[r * 4 + k] += cg[r * 4 + 0] * L[k * 4 + 0] + cg[r * 4 + 1] * L[k * 4 + 1] +
                             cg[r * 4 + 2] * L[k * 4 + 2] + cg[r * 4 + 3] * L[k * 4 + 3];
          dg[r * 4 + 3] += cg[r * 4 + 3];
          dj[r] -= G[0 * 4 + r] * cg[3] + G[1 * 4 + r] * cg[7] + G[2 * 4 + r] * cg[11];   // Jrel_c = J_c - J_j
        }
      }
      if (c1 > c0) {
#pragma unroll
        for (int q = 0; q < 12; ++q) dG[j * 12 + q] = dg[q];
      }
      if (d >= 1) {
        const float* Pm = Gw + s_par[j] * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dRs[j * 9 + r * 3 + c] = Pm[0 * 4 + r] * dg[0 * 4 + c] + Pm[1 * 4 + r] * dg[1 * 4 + c] +
                                     Pm[2 * 4 + r] * dg[2 * 4 + c];
          dj[r] += Pm[0 * 4 + r] * dg[3] + Pm[1 * 4 + r] * dg[7] + Pm[2 * 4 + r] * dg[11];
        }
      } else {          // root: L_0 = [R_0 | J_0]
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) dRs[j * 9 + r * 3 + c] = dg[r * 4 + c];
          dj[r] += dg[r * 4 + 3];
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) dJ[j * 3 + r] = dj[r];
    }
    __syncwarp();
  }
#endif
  float dRl[SLOTS][9];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
#pragma unroll
    for (int q = 0; q < 9; ++q) dRl[s][q] = (j < m.J) ? dRs[j * 9 + q] : 0.f;
  }

  // ---- pose-feature gradient from the blend GEMM (split-K partials already summed into the first
  // split's rows by reduce_splits_kernel): stage the body's row in smem with coalesced loads
  if (a.d_feat != nullptr && !feat_staged) {
    __syncwarp();
    const float* f = a.d_feat + (size_t)b * m.Kpad;
    for (int k = lane; k < m.P + m.NB; k += 32) {
      float acc = f[k];
      for (int sp = 1; sp < a.feat_splits; ++sp) acc += f[(size_t)sp * a.feat_split_stride + k];
      dfeat_sum[k] = acc;
    }
    __syncwarp();
  }
  // ---- Rodrigues backward
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      if (a.d_feat != nullptr && j >= 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dRl[s][i] += dfeat_sum[9 * (j - 1) + i];
      }
      float dr[3];
      rodrigues_backward(rv[s], dRl[s], dr);
      if (a.d_full_pose != nullptr) {
        const float* dfp = a.d_full_pose + (size_t)b * 3 * m.J + 3 * j;
        dr[0] += dfp[0]; dr[1] += dfp[1]; dr[2] += dfp[2];
      }
      dfull[3 * j + 0] = dr[0]; dfull[3 * j + 1] = dr[1]; dfull[3 * j + 2] = dr[2];
    }
  }
  __syncwarp();

  const float gs = a.d_loss ? a.d_loss[(size_t)b * a.d_loss_stride] : 1.f;        // upstream gradient of this body's loss
  const int hand0 = m.J - 30;
  if (a.d_pose) {
    for (int i = lane; i < 3 * m.J; i += 32) {
      const int j = i / 3;
      const bool from_pca = (a.pca_l && j >= hand0 && j < hand0 + 15) || (a.pca_r && j >= hand0 + 15);
      a.d_pose[(size_t)b * 3 * m.J + i] = from_pca ? 0.f : gs * dfull[i];
    }
  }
  if (a.d_pca_l && a.pca_l) {
    for (int c = lane; c < m.C; c += 32) {
      float acc = 0.f;
#pragma unroll 9
      for (int i = 0; i < 45; ++i) acc = fmaf(m.comp_l[c * 45 + i], dfull[3 * hand0 + i], acc);
      a.d_pca_l[(size_t)b * m.C + c] = gs * acc;
    }
  }
  if (a.d_pca_r && a.pca_r) {
    for (int c = lane; c < m.C; c += 32) {
      float acc = 0.f;
#pragma unroll 9
      for (int i = 0; i < 45; ++i) acc = fmaf(m.comp_r[c * 45 + i], dfull[3 * (hand0 + 15) + i], acc);
      a.d_pca_r[(size_t)b * m.C + c] = gs * acc;
    }
  }
  // ---- d_betas = J_shapedirs^T dJ + (blend GEMM columns P..P+NB)
  if (a.d_betas) {
    // 16 betas a round, the 3J joint coordinates split over the two half-warps, two sums per lane
    const int nq = 3 * m.J;
    for (int i0 = 0; i0 < m.NB; i0 += 16) {
      const int i = i0 + (lane & 15), h = lane >> 4;
      float acc0 = 0.f, acc1 = 0.f;
      if (i < m.NB) {
#pragma unroll 4
        for (int q = h; q < nq; q += 4) {
          acc0 = fmaf(dJ[q], m.J_shapedirs[(size_t)q * m.NB + i], acc0);
          if (q + 2 < nq) acc1 = fmaf(dJ[q + 2], m.J_shapedirs[(size_t)(q + 2) * m.NB + i], acc1);
        }
      }
      float acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      if (h == 0 && i < m.NB) {
        if (a.d_feat != nullptr) acc += dfeat_sum[m.P + i];
        acc *= gs;
        if (a.betas_B == 1) atomicAdd(&a.d_betas[i], acc);
        else a.d_betas[(size_t)b * m.NB + i] = acc;
      }
    }
  }
  if (a.dAp != nullptr) {                 // lanes hold their joints' translation columns: sum over the warp
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dtr_acc[q] += __shfl_xor_sync(0xffffffffu, dtr_acc[q], o);
    }
  }
  if (a.d_transl && lane < 3) {
    float t = a.dtr_verts ? a.dtr_verts[3 * b + lane] : 0.f;
    if (a.dAp != nullptr) t += lane == 0 ? dtr_acc[0] : (lane == 1 ? dtr_acc[1] : dtr_acc[2]);
    if (a.d_joints)
      for (int j = 0; j < m.J; ++j) t += a.d_joints[(size_t)b * a.joints_ld + 3 * j + lane];
    a.d_transl[3 * b + lane] = gs * t;
  }
}

}  // namespace smplk
