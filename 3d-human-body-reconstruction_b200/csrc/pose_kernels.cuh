// Warp-per-body pose kernels: hand PCA -> Rodrigues -> pose features -> rest joints ->
// kinematic chain -> skinning transforms, and the matching backward.
//
// Reference math: models/smplh_np.py:49-70 (compute_R_G), :73-78 (rest-pose removal),
// :88-109 (rodrigues); upstream smplx lbs.batch_rodrigues / batch_rigid_transform
// (SURVEY.md 8a rows a3-a6); utils/geometry.py:9-45.
//
// Lane l of the warp owns joints l and l+32 (J <= 64).  The chain is walked level by level
// (depth tables from the packer); a child fetches its parent's world transform with warp
// shuffles, so the 51 dependent 4x4 products of the reference become max_depth (<= 10)
// shuffle rounds.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace smplk {

constexpr int kPoseWarps = 4;

struct PoseFwdArgs {
  int B;
  float* At;             // [ceil(B/256)*2][J][128][12] joint-major transforms per 128-body block (+ transl), or null (block kernel only)
  int At_rows;           // rows of At (B rounded up to whole 256-body blocks): bodies in [B, At_rows) get zero transforms
  const float* betas;
  int betas_B;
  const float* pose;     // (B,3J)
  const float* pca_l;    // (B,C) or null
  const float* pca_r;    // (B,C) or null
  int add_mean;
  const float* transl;   // (B,3) or null
  float* F_hi;           // [rows][Kpad] tf32 hi/lo rows, or null
  float* F_lo;
  __half* H_hi;          // [rows][Kpad] fp16 hi/lo rows (f16 blend GEMM), or null
  __half* H_lo;
  float* A;              // [B][J][12]
  float* joints;         // (B, joints_ld) or null; FK joints written to the first 3J entries
  int joints_ld;
  float* full_pose;      // (B,3J) or null
};

// R = I + sin(a) K + (1 - cos a) K^2 with a = ||r + 1e-8|| and K = [r / a]_x  (upstream
// batch_rodrigues; eps added to the vector, utils/geometry.py:16).  1 - cos a is evaluated as
// 2 sin^2(a/2) to avoid the cancellation of the literal form at small angles.
__device__ __forceinline__ void rodrigues(float rx, float ry, float rz, float* R) {
  float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
  float a = sqrtf(ex * ex + ey * ey + ez * ez);
  float inv = 1.0f / a;
  float nx = rx * inv, ny = ry * inv, nz = rz * inv;
  float s, c;
  sincosf(a, &s, &c);
  float sh = sinf(0.5f * a);
  float omc = 2.0f * sh * sh;
  (void)c;
  R[0] = 1.0f - omc * (ny * ny + nz * nz);
  R[1] = -s * nz + omc * nx * ny;
  R[2] = s * ny + omc * nx * nz;
  R[3] = s * nz + omc * nx * ny;
  R[4] = 1.0f - omc * (nx * nx + nz * nz);
  R[5] = -s * nx + omc * ny * nz;
  R[6] = -s * ny + omc * nx * nz;
  R[7] = s * nx + omc * ny * nz;
  R[8] = 1.0f - omc * (nx * nx + ny * ny);
}

// Assembled axis-angle of joint j for body b: `pose` columns, hand PCA override, pose mean.
// Rest joint j = J_template[j] + J_shapedirs[j] . betas, four betas a round with every load issued
// before the first use: for a single cold body the plain loop serialises one L2 round trip per beta.
__device__ __forceinline__ void rest_joint(const ModelDev& m, const float* betas_row, int j, float* out) {
  float v0 = m.J_template[3 * j], v1 = m.J_template[3 * j + 1], v2 = m.J_template[3 * j + 2];
  if (betas_row != nullptr) {
    const float* s0 = m.J_shapedirs + (size_t)(3 * j) * m.NB;
    const float* s1 = s0 + m.NB;
    const float* s2 = s1 + m.NB;
    int i = 0;
    for (; i + 4 <= m.NB; i += 4) {
      const float b0 = betas_row[i], b1 = betas_row[i + 1], b2 = betas_row[i + 2], b3 = betas_row[i + 3];
      const float x0 = s0[i], x1 = s0[i + 1], x2 = s0[i + 2], x3 = s0[i + 3];
      const float y0 = s1[i], y1 = s1[i + 1], y2 = s1[i + 2], y3 = s1[i + 3];
      const float z0 = s2[i], z1 = s2[i + 1], z2 = s2[i + 2], z3 = s2[i + 3];
      v0 = fmaf(x3, b3, fmaf(x2, b2, fmaf(x1, b1, fmaf(x0, b0, v0))));
      v1 = fmaf(y3, b3, fmaf(y2, b2, fmaf(y1, b1, fmaf(y0, b0, v1))));
      v2 = fmaf(z3, b3, fmaf(z2, b2, fmaf(z1, b1, fmaf(z0, b0, v2))));
    }
    for (; i < m.NB; ++i) {
      const float bi = betas_row[i];
      v0 = fmaf(s0[i], bi, v0); v1 = fmaf(s1[i], bi, v1); v2 = fmaf(s2[i], bi, v2);
    }
  }
  out[0] = v0; out[1] = v1; out[2] = v2;
}

__device__ __forceinline__ void load_joint_pose(const ModelDev& m, const float* pose,
                                                const float* pca_l, const float* pca_r,
                                                int add_mean, int b, int j, float* r) {
  const int hand0 = m.J - 30;  // first left-hand joint (22 for SMPL-H)
  if (pca_l != nullptr && j >= hand0 && j < hand0 + 15) {
    const float* comp = m.comp_l + 3 * (j - hand0);
    const float* c = pca_l + (size_t)b * m.C;
    float x = 0.f, y = 0.f, z = 0.f;
#pragma unroll 4
    for (int i = 0; i < m.C; ++i) {
      float ci = c[i];
      x = fmaf(ci, comp[i * 45 + 0], x);
      y = fmaf(ci, comp[i * 45 + 1], y);
      z = fmaf(ci, comp[i * 45 + 2], z);
    }
    r[0] = x; r[1] = y; r[2] = z;
  } else if (pca_r != nullptr && j >= hand0 + 15) {
    const float* comp = m.comp_r + 3 * (j - hand0 - 15);
    const float* c = pca_r + (size_t)b * m.C;
    float x = 0.f, y = 0.f, z = 0.f;
#pragma unroll 4
    for (int i = 0; i < m.C; ++i) {
      float ci = c[i];
      x = fmaf(ci, comp[i * 45 + 0], x);
      y = fmaf(ci, comp[i * 45 + 1], y);
      z = fmaf(ci, comp[i * 45 + 2], z);
    }
    r[0] = x; r[1] = y; r[2] = z;
  } else {
    const float* p = pose + (size_t)b * 3 * m.J + 3 * j;
    r[0] = p[0]; r[1] = p[1]; r[2] = p[2];
  }
  if (add_mean && m.pose_mean != nullptr) {
    r[0] += m.pose_mean[3 * j + 0];
    r[1] += m.pose_mean[3 * j + 1];
    r[2] += m.pose_mean[3 * j + 2];
  }
}

// G (3x4, row-major [R|t]) of the parent joint `p` fetched from the lane/slot that owns it.
template <int SLOTS>
__device__ __forceinline__ void fetch_parent(const float (&G)[SLOTS][12], int p, float* out) {
  const int src = p & 31;
  const int slot = p >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float v0 = __shfl_sync(0xffffffffu, G[0][i], src);
    if (SLOTS > 1) {
      float v1 = __shfl_sync(0xffffffffu, G[SLOTS - 1][i], src);
      out[i] = slot ? v1 : v0;
    } else {
      out[i] = v0;
    }
  }
}

// child = parent * local, all 3x4 affine [R|t].
__device__ __forceinline__ void affine_mul(const float* Pm, const float* L, float* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = Pm[r * 4 + 0] * L[0 * 4 + c];
      v = fmaf(Pm[r * 4 + 1], L[1 * 4 + c], v);
      v = fmaf(Pm[r * 4 + 2], L[2 * 4 + c], v);
      if (c == 3) v += Pm[r * 4 + 3];
      out[r * 4 + c] = v;
    }
  }
}

// Shared forward core: fills R, rest joints Jr and world transforms G of the lane's joints.
template <int SLOTS>
__device__ __forceinline__ void pose_forward_core(const ModelDev& m, const float* betas_row,
                                                  const float (&rv)[SLOTS][3],
                                                  float (&R)[SLOTS][9], float (&Jr)[SLOTS][3],
                                                  float (&Jrel)[SLOTS][3], float (&G)[SLOTS][12],
                                                  int lane) {
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    const bool valid = j < m.J;
    if (valid) {
      rodrigues(rv[s][0], rv[s][1], rv[s][2], R[s]);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float v = m.J_template[3 * j + a];
        if (betas_row != nullptr) {
          const float* sd = m.J_shapedirs + (size_t)(3 * j + a) * m.NB;
          for (int i = 0; i < m.NB; ++i) v = fmaf(sd[i], betas_row[i], v);
        }
        Jr[s][a] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[s][i] = (i % 4 == 0) ? 1.f : 0.f;
      Jr[s][0] = Jr[s][1] = Jr[s][2] = 0.f;
    }
  }
  // parent rest joints -> relative offsets
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    const int p = (j < m.J) ? m.parents[j] : -1;
    const int pp = p < 0 ? 0 : p;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v0 = __shfl_sync(0xffffffffu, Jr[0][a], pp & 31);
      float v1 = __shfl_sync(0xffffffffu, Jr[SLOTS - 1][a], pp & 31);
      float pj = (pp >> 5) ? v1 : v0;
      Jrel[s][a] = (p < 0) ? Jr[s][a] : Jr[s][a] - pj;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      G[s][r * 4 + 0] = R[s][r * 3 + 0];
      G[s][r * 4 + 1] = R[s][r * 3 + 1];
      G[s][r * 4 + 2] = R[s][r * 3 + 2];
      G[s][r * 4 + 3] = Jrel[s][r];
    }
  }
  // walk the tree: after round d every joint of depth <= d holds its world transform
  for (int d = 1; d <= m.max_depth; ++d) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int j = lane + 32 * s;
      const bool act = (j < m.J) && (m.depth[j] == d);
      const int p = act ? m.parents[j] : 0;
      float Pm[12];
      fetch_parent<SLOTS>(G, p, Pm);
      if (act) {
        float out[12];
        affine_mul(Pm, G[s], out);
#pragma unroll
        for (int i = 0; i < 12; ++i) G[s][i] = out[i];
      }
    }
  }
}

// Forward kernel.  The chain runs through a per-warp shared-memory table of 3x4 transforms:
// joints are visited level by level (`order` / `level_start` from the packer), one lane per joint
// of the level, each reading its parent's world transform (3 x LDS.128) and writing its own --
// max_depth (<= 10) short rounds instead of 51 dependent products.  (The backward kernel keeps
// the register/shuffle variant of the walk, `pose_forward_core`, because it needs every lane's
// forward values in registers.)
template <int SLOTS>
__global__ void __launch_bounds__(kPoseWarps * 32)
pose_forward_kernel(const ModelDev m, const PoseFwdArgs a) {
  extern __shared__ __align__(16) float pose_smem[];
  // the skeleton's index tables, staged once per block: the level walk otherwise chains three
  // dependent global loads per level (L2 latency each for a single body)
  __shared__ int s_par[kMaxJoints], s_ord[kMaxJoints], s_lvl[kMaxJoints + 2];
  ptx::pdl_launch_dependents();
  for (int i = threadIdx.x; i < m.J; i += blockDim.x) { s_par[i] = m.parents[i]; s_ord[i] = m.order[i]; }
  for (int i = threadIdx.x; i < m.max_depth + 2; i += blockDim.x) s_lvl[i] = m.level_start[i];
  __syncthreads();
  ptx::pdl_wait();               // the workspace rows written below may still be read by the previous call's kernels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kPoseWarps + warp;
  if (b >= a.B) return;
  const int per_warp = max(m.Kpad, 32) + m.J * 12;
  float* feat = pose_smem + warp * per_warp;                 // [Kpad] blend features
  float* Gs = feat + max(m.Kpad, 32);                        // [J][12] transforms
  const float* betas_row = a.betas ? a.betas + (size_t)(a.betas_B == 1 ? 0 : b) * m.NB : nullptr;

  // ---- per joint: pose -> R, rest joint; local transform [R | J] into the table
  float Jr[SLOTS][3];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    Jr[s][0] = Jr[s][1] = Jr[s][2] = 0.f;
    if (j < m.J) {
      float rv[3], R[9];
      load_joint_pose(m, a.pose, a.pca_l, a.pca_r, a.add_mean, b, j, rv);
      if (a.full_pose) {
        float* fp = a.full_pose + (size_t)b * 3 * m.J + 3 * j;
        fp[0] = rv[0]; fp[1] = rv[1]; fp[2] = rv[2];
      }
      rodrigues(rv[0], rv[1], rv[2], R);
      rest_joint(m, betas_row, j, Jr[s]);
      float4* g = reinterpret_cast<float4*>(Gs + j * 12);
      g[0] = make_float4(R[0], R[1], R[2], Jr[s][0]);
      g[1] = make_float4(R[3], R[4], R[5], Jr[s][1]);
      g[2] = make_float4(R[6], R[7], R[8], Jr[s][2]);
      if ((a.F_hi != nullptr || a.H_hi != nullptr) && j >= 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) feat[9 * (j - 1) + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
      }
    }
  }
  if (a.F_hi != nullptr || a.H_hi != nullptr) {
    for (int i = lane; i < m.Kpad - m.P; i += 32)
      feat[m.P + i] = (i < m.NB && betas_row) ? betas_row[i] : 0.f;
  }
  __syncwarp();
  // ---- translation column -> offset from the parent's rest joint
  float pj[SLOTS][3];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    pj[s][0] = pj[s][1] = pj[s][2] = 0.f;
    if (j >= 1 && j < m.J) {
      const float* gp = Gs + s_par[j] * 12;
      pj[s][0] = gp[3]; pj[s][1] = gp[7]; pj[s][2] = gp[11];
    }
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j >= 1 && j < m.J) {
      Gs[j * 12 + 3] = Jr[s][0] - pj[s][0];
      Gs[j * 12 + 7] = Jr[s][1] - pj[s][1];
      Gs[j * 12 + 11] = Jr[s][2] - pj[s][2];
    }
  }
  __syncwarp();

  // ---- blend-feature rows (independent of the chain; issue the stores before walking it)
  if (a.F_hi != nullptr) {
    float* fh = a.F_hi + (size_t)b * m.Kpad;
    float* fl = a.F_lo + (size_t)b * m.Kpad;
    for (int k = lane; k < m.Kpad; k += 32) {
      float x = feat[k];
      float h = ptx::tf32_round(x);
      fh[k] = h;
      fl[k] = x - h;
    }
  }
  if (a.H_hi != nullptr) {   // two-term fp16 split: x = hi + lo + O(2^-22 |x|)
    __half2* hh = reinterpret_cast<__half2*>(a.H_hi + (size_t)b * m.Kpad);
    __half2* hl = reinterpret_cast<__half2*>(a.H_lo + (size_t)b * m.Kpad);
    for (int k = lane; k < m.Kpad / 2; k += 32) {
      const float2 x = *reinterpret_cast<const float2*>(feat + 2 * k);
      const __half h0 = __float2half_rn(x.x), h1 = __float2half_rn(x.y);
      hh[k] = __halves2half2(h0, h1);
      hl[k] = __halves2half2(__float2half_rn(x.x - __half2float(h0)),
                             __float2half_rn(x.y - __half2float(h1)));
    }
  }

  // ---- walk the tree level by level: G_j = G_parent(j) * L_j
  for (int d = 1; d <= m.max_depth; ++d) {
    const int l0 = s_lvl[d], l1 = s_lvl[d + 1];
    for (int i = l0 + lane; i < l1; i += 32) {
      const int j = s_ord[i];
      const float4* P4 = reinterpret_cast<const float4*>(Gs + s_par[j] * 12);
      float4* L4 = reinterpret_cast<float4*>(Gs + j * 12);
      const float4 p0 = P4[0], p1 = P4[1], p2 = P4[2];
      const float4 q0 = L4[0], q1 = L4[1], q2 = L4[2];
      float4 o0, o1, o2;
      o0.x = fmaf(p0.x, q0.x, fmaf(p0.y, q1.x, p0.z * q2.x));
      o0.y = fmaf(p0.x, q0.y, fmaf(p0.y, q1.y, p0.z * q2.y));
      o0.z = fmaf(p0.x, q0.z, fmaf(p0.y, q1.z, p0.z * q2.z));
      o0.w = fmaf(p0.x, q0.w, fmaf(p0.y, q1.w, fmaf(p0.z, q2.w, p0.w)));
      o1.x = fmaf(p1.x, q0.x, fmaf(p1.y, q1.x, p1.z * q2.x));
      o1.y = fmaf(p1.x, q0.y, fmaf(p1.y, q1.y, p1.z * q2.y));
      o1.z = fmaf(p1.x, q0.z, fmaf(p1.y, q1.z, p1.z * q2.z));
      o1.w = fmaf(p1.x, q0.w, fmaf(p1.y, q1.w, fmaf(p1.z, q2.w, p1.w)));
      o2.x = fmaf(p2.x, q0.x, fmaf(p2.y, q1.x, p2.z * q2.x));
      o2.y = fmaf(p2.x, q0.y, fmaf(p2.y, q1.y, p2.z * q2.y));
      o2.z = fmaf(p2.x, q0.z, fmaf(p2.y, q1.z, p2.z * q2.z));
      o2.w = fmaf(p2.x, q0.w, fmaf(p2.y, q1.w, fmaf(p2.z, q2.w, p2.w)));
      L4[0] = o0; L4[1] = o1; L4[2] = o2;
    }
    __syncwarp();
  }

  // ---- skinning transforms A_j = [G_R | G_t - G_R J_j]  and FK joints
  const float tx = a.transl ? a.transl[3 * b + 0] : 0.f;
  const float ty = a.transl ? a.transl[3 * b + 1] : 0.f;
  const float tz = a.transl ? a.transl[3 * b + 2] : 0.f;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      const float4* G4 = reinterpret_cast<const float4*>(Gs + j * 12);
      float4 g0 = G4[0], g1 = G4[1], g2 = G4[2];
      if (a.joints) {
        float* jo = a.joints + (size_t)b * a.joints_ld + 3 * j;
        jo[0] = g0.w + tx; jo[1] = g1.w + ty; jo[2] = g2.w + tz;
      }
      g0.w -= g0.x * Jr[s][0] + g0.y * Jr[s][1] + g0.z * Jr[s][2];
      g1.w -= g1.x * Jr[s][0] + g1.y * Jr[s][1] + g1.z * Jr[s][2];
      g2.w -= g2.x * Jr[s][0] + g2.y * Jr[s][1] + g2.z * Jr[s][2];
      float4* dst = reinterpret_cast<float4*>(a.A + ((size_t)b * m.J + j) * 12);
      dst[0] = g0; dst[1] = g1; dst[2] = g2;
    }
  }
}


// ------------------------------------------------------------------------------------------
// Block version of the forward kernel: 32 warps = 32 consecutive bodies per block.
//
// ncu on the warp-per-body kernel above (4 bodies per block) shows 53 % of the stall samples on the
// long scoreboard at an IPC of 0.03 per warp: every table access (parents / order / level_start /
// J_template / J_shapedirs / pose_mean / PCA components) is a dependent global load (L1 hit, ~35
// cycles each, a few hundred of them on the critical path).  Here the block stages those tables in
// shared memory once and the warps read them with LDS.  It also writes the joint-major transform
// copy At[128-body block][J][128][12] that the fused blend+skinning kernel reads (a warp's 32 bodies
// are 1,536 contiguous bytes per joint), so there is no separate re-blocking pass.
// ------------------------------------------------------------------------------------------
// 1 (product): the block kernel walks the chain with lane = body; 0: lane = joint, one warp per body (A/B builds,
// bitwise-equal results; tools/gpu_posewalk_ab.sh)
#ifndef SMPLK_POSE_FWD_LANE_BODY
#define SMPLK_POSE_FWD_LANE_BODY 1
#endif
constexpr int kPoseBlockWarps = 32;

struct PoseBlockSmem {
  int per_warp;      // floats per warp: feat[Kpad] + Gs[J*12], padded to 4 mod 32 (rows stay 16-byte
                     // aligned for the float4 table accesses; the transposition reads are 4-way conflicted)
  int off_Jt, off_Js, off_pm, off_cl, off_cr, off_par, off_ord, off_lvl, total;   // float offsets
  int nbp;           // padded betas stride of J_shapedirs rows (odd)
};

// `warps` = bodies per block: 32 when the block also writes the transposed transforms At (lane = body there),
// fewer for mid-size batches so that a few hundred bodies still spread over all SMs.
__host__ __device__ inline PoseBlockSmem pose_block_layout(const ModelDev& m, int warps = kPoseBlockWarps) {
  PoseBlockSmem L;
  const int base = max(m.Kpad, 32) + m.J * 12;
  L.per_warp = base + ((36 - (base & 31)) & 31);             // == 4 (mod 32)
  L.nbp = (max(m.NB, 1) | 1);
  int o = warps * L.per_warp;
  L.off_Jt = o;  o += 3 * m.J;
  L.off_Js = o;  o += 3 * m.J * L.nbp;
  L.off_pm = o;  o += 3 * m.J;
  L.off_cl = o;  o += m.C * 45;
  L.off_cr = o;  o += m.C * 45;
  L.off_par = o; o += m.J;
  L.off_ord = o; o += m.J;
  L.off_lvl = o; o += m.max_depth + 2;
  L.total = o;
  return L;
}

template <int SLOTS>
__global__ void __launch_bounds__(kPoseBlockWarps * 32, 1)
pose_forward_block_kernel(const ModelDev m, const PoseFwdArgs a) {
  extern __shared__ __align__(16) float pose_smem[];
  const int nw = blockDim.x >> 5;                 // bodies of this block (one warp each)
  const PoseBlockSmem L = pose_block_layout(m, nw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * nw + warp;
  const bool live = b < a.B;
  float* feat = pose_smem + warp * L.per_warp;
  float* Gs = feat + max(m.Kpad, 32);
  float* sJt = pose_smem + L.off_Jt;
  float* sJs = pose_smem + L.off_Js;
  float* spm = pose_smem + L.off_pm;
  float* scl = pose_smem + L.off_cl;
  float* scr = pose_smem + L.off_cr;
  int* spar = reinterpret_cast<int*>(pose_smem + L.off_par);
  int* sord = reinterpret_cast<int*>(pose_smem + L.off_ord);
  int* slvl = reinterpret_cast<int*>(pose_smem + L.off_lvl);

  // ---- stage the model tables: 4-byte cp.async, everything in flight at once (as register loads the copies were
  // ~10 dependent L2 round trips per thread: a quarter of the kernel at 1,024 bodies).  Under a programmatic
  // dependent launch this part runs beside the tail of the previous kernel in the stream (constant tables only).
  ptx::pdl_launch_dependents();
  {
    const uint32_t s0 = ptx::smem_u32(pose_smem);
    auto stage = [&](const void* dst, const void* src) {
      ptx::cp_async_4(s0 + static_cast<uint32_t>(reinterpret_cast<const char*>(dst) - reinterpret_cast<const char*>(pose_smem)), src);
    };
    const bool mean = a.add_mean && m.pose_mean;
    for (int i = threadIdx.x; i < 3 * m.J; i += blockDim.x) {
      stage(sJt + i, m.J_template + i);
      if (mean) stage(spm + i, m.pose_mean + i);
      else spm[i] = 0.f;
    }
    for (int i = threadIdx.x; i < 3 * m.J * m.NB; i += blockDim.x) stage(sJs + (i / m.NB) * L.nbp + (i % m.NB), m.J_shapedirs + i);
    if (a.pca_l != nullptr || a.pca_r != nullptr)
      for (int i = threadIdx.x; i < m.C * 45; i += blockDim.x) { stage(scl + i, m.comp_l + i); stage(scr + i, m.comp_r + i); }
    for (int i = threadIdx.x; i < m.J; i += blockDim.x) { stage(spar + i, m.parents + i); stage(sord + i, m.order + i); }
    for (int i = threadIdx.x; i < m.max_depth + 2; i += blockDim.x) stage(slvl + i, m.level_start + i);
    ptx::cp_async_commit();
  }
  ptx::pdl_wait();
  // ---- this body's small inputs into the warp's feature row: betas at feat[P..], PCA coefficients
  // parked at feat[0..2C) until the features overwrite them
  const float* betas_row = (a.betas && live) ? a.betas + (size_t)(a.betas_B == 1 ? 0 : b) * m.NB : nullptr;
  for (int i = lane; i < m.Kpad - m.P; i += 32) feat[m.P + i] = (i < m.NB && betas_row) ? betas_row[i] : 0.f;
  float pca_c = 0.f;                      // lane i < C: left coefficient i; lane 16 + i: right coefficient i
  if (live && a.pca_l && lane < m.C) pca_c = a.pca_l[(size_t)b * m.C + lane];
  if (live && a.pca_r && lane >= 16 && lane - 16 < m.C) pca_c = a.pca_r[(size_t)b * m.C + lane - 16];
  ptx::cp_async_wait<0>();
  __syncthreads();

  const int hand0 = m.J - 30;
  float Jr[SLOTS][3];
  float rvs[SLOTS][3];
  // hand PCA needs every lane for the coefficient broadcasts: evaluate it for all slots first
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    float x = 0.f, y = 0.f, z = 0.f;
    const bool lh = a.pca_l != nullptr && j >= hand0 && j < hand0 + 15;
    const bool rh = a.pca_r != nullptr && j >= hand0 + 15 && j < m.J;
    if (a.pca_l != nullptr || a.pca_r != nullptr) {
      const float* comp = lh ? scl + 3 * (j - hand0) : scr + 3 * (j - hand0 - 15);
      for (int i = 0; i < m.C; ++i) {
        const float cl = __shfl_sync(0xffffffffu, pca_c, i);
        const float cr = __shfl_sync(0xffffffffu, pca_c, 16 + i);
        if (lh || rh) {
          const float ci = lh ? cl : cr;
          x = fmaf(ci, comp[i * 45 + 0], x);
          y = fmaf(ci, comp[i * 45 + 1], y);
          z = fmaf(ci, comp[i * 45 + 2], z);
        }
      }
    }
    if (j < m.J && live) {
      if (!(lh || rh)) {
        const float* pp = a.pose + (size_t)b * 3 * m.J + 3 * j;
        x = pp[0]; y = pp[1]; z = pp[2];
      }
      x += spm[3 * j]; y += spm[3 * j + 1]; z += spm[3 * j + 2];
    }
    rvs[s][0] = x; rvs[s][1] = y; rvs[s][2] = z;
  }
  // ---- per joint: R, rest joint; local transform [R | J] into the table
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    Jr[s][0] = Jr[s][1] = Jr[s][2] = 0.f;
    if (j < m.J) {
      float R[9];
      if (a.full_pose && live) {
        float* fp = a.full_pose + (size_t)b * 3 * m.J + 3 * j;
        fp[0] = rvs[s][0]; fp[1] = rvs[s][1]; fp[2] = rvs[s][2];
      }
      rodrigues(rvs[s][0], rvs[s][1], rvs[s][2], R);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = sJt[3 * j + c];
        const float* sd = sJs + (3 * j + c) * L.nbp;
        for (int i = 0; i < m.NB; ++i) v = fmaf(sd[i], feat[m.P + i], v);
        Jr[s][c] = v;
      }
      float4* g = reinterpret_cast<float4*>(Gs + j * 12);
      g[0] = make_float4(R[0], R[1], R[2], Jr[s][0]);
      g[1] = make_float4(R[3], R[4], R[5], Jr[s][1]);
      g[2] = make_float4(R[6], R[7], R[8], Jr[s][2]);
      if ((a.F_hi != nullptr || a.H_hi != nullptr) && j >= 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) feat[9 * (j - 1) + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
      }
    }
  }
  __syncwarp();
  // ---- translation column -> offset from the parent's rest joint
  float pj[SLOTS][3];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    pj[s][0] = pj[s][1] = pj[s][2] = 0.f;
    if (j >= 1 && j < m.J) {
      const float* gp = Gs + spar[j] * 12;
      pj[s][0] = gp[3]; pj[s][1] = gp[7]; pj[s][2] = gp[11];
    }
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j >= 1 && j < m.J) {
      Gs[j * 12 + 3] = Jr[s][0] - pj[s][0];
      Gs[j * 12 + 7] = Jr[s][1] - pj[s][1];
      Gs[j * 12 + 11] = Jr[s][2] - pj[s][2];
    }
  }
  __syncwarp();

  // ---- blend-feature rows (independent of the chain; issue the stores before walking it)
  if (a.F_hi != nullptr && live) {
    float* fh = a.F_hi + (size_t)b * m.Kpad;
    float* fl = a.F_lo + (size_t)b * m.Kpad;
    for (int k = lane; k < m.Kpad; k += 32) {
      float x = feat[k];
      float h = ptx::tf32_round(x);
      fh[k] = h;
      fl[k] = x - h;
    }
  }
  if (a.H_hi != nullptr && live) {
    __half2* hh = reinterpret_cast<__half2*>(a.H_hi + (size_t)b * m.Kpad);
    __half2* hl = reinterpret_cast<__half2*>(a.H_lo + (size_t)b * m.Kpad);
    for (int k = lane; k < m.Kpad / 2; k += 32) {
      const float2 x = *reinterpret_cast<const float2*>(feat + 2 * k);
      const __half h0 = __float2half_rn(x.x), h1 = __float2half_rn(x.y);
      hh[k] = __halves2half2(h0, h1);
      hl[k] = __halves2half2(__float2half_rn(x.x - __half2float(h0)),
                             __float2half_rn(x.y - __half2float(h1)));
    }
  }

#if SMPLK_POSE_FWD_LANE_BODY
  // ---- walk the tree level by level: G_j = G_parent(j) * L_j, with LANE = BODY of the block and the level's joints
  // dealt to the warps.  (Lane = joint, one warp per body, left most lanes idle: a level of SMPL-H holds 1-10 joints,
  // and the walk was 31 % of the kernel's executed instructions -- ncu source view, 71 instructions x 10 levels per body.
  // Measured on one box against the lane = joint build: 19 % fewer instructions, the SAME kernel time -- DESIGN 6c.)
  // A body's table is per_warp = 4 (mod 32) words from the next, so the 8 lanes of an LDS.128 phase hit 32 distinct banks.
  __syncthreads();
  {
    float* Gb = pose_smem + min(lane, nw - 1) * L.per_warp + max(m.Kpad, 32);
    for (int d = 1; d <= m.max_depth; ++d) {
      const int l0 = slvl[d], l1 = slvl[d + 1];
      if (lane < nw) {
        for (int i = l0 + warp; i < l1; i += nw) {
          const int j = sord[i];
          const float4* P4 = reinterpret_cast<const float4*>(Gb + spar[j] * 12);
          float4* L4 = reinterpret_cast<float4*>(Gb + j * 12);
          const float4 p0 = P4[0], p1 = P4[1], p2 = P4[2];
          const float4 q0 = L4[0], q1 = L4[1], q2 = L4[2];
          float4 o0, o1, o2;
          o0.x = fmaf(p0.x, q0.x, fmaf(p0.y, q1.x, p0.z * q2.x));
          o0.y = fmaf(p0.x, q0.y, fmaf(p0.y, q1.y, p0.z * q2.y));
          o0.z = fmaf(p0.x, q0.z, fmaf(p0.y, q1.z, p0.z * q2.z));
          o0.w = fmaf(p0.x, q0.w, fmaf(p0.y, q1.w, fmaf(p0.z, q2.w, p0.w)));
          o1.x = fmaf(p1.x, q0.x, fmaf(p1.y, q1.x, p1.z * q2.x));
          o1.y = fmaf(p1.x, q0.y, fmaf(p1.y, q1.y, p1.z * q2.y));
          o1.z = fmaf(p1.x, q0.z, fmaf(p1.y, q1.z, p1.z * q2.z));
          o1.w = fmaf(p1.x, q0.w, fmaf(p1.y, q1.w, fmaf(p1.z, q2.w, p1.w)));
          o2.x = fmaf(p2.x, q0.x, fmaf(p2.y, q1.x, p2.z * q2.x));
          o2.y = fmaf(p2.x, q0.y, fmaf(p2.y, q1.y, p2.z * q2.y));
          o2.z = fmaf(p2.x, q0.z, fmaf(p2.y, q1.z, p2.z * q2.z));
          o2.w = fmaf(p2.x, q0.w, fmaf(p2.y, q1.w, fmaf(p2.z, q2.w, p2.w)));
          L4[0] = o0; L4[1] = o1; L4[2] = o2;
        }
      }
      __syncthreads();
    }
  }
#else   // A/B build: the lane = joint walk (one warp per body)
  // ---- walk the tree level by level: G_j = G_parent(j) * L_j
  for (int d = 1; d <= m.max_depth; ++d) {
    const int l0 = slvl[d], l1 = slvl[d + 1];
    for (int i = l0 + lane; i < l1; i += 32) {
      const int j = sord[i];
      const float4* P4 = reinterpret_cast<const float4*>(Gs + spar[j] * 12);
      float4* L4 = reinterpret_cast<float4*>(Gs + j * 12);
      const float4 p0 = P4[0], p1 = P4[1], p2 = P4[2];
      const float4 q0 = L4[0], q1 = L4[1], q2 = L4[2];
      float4 o0, o1, o2;
      o0.x = fmaf(p0.x, q0.x, fmaf(p0.y, q1.x, p0.z * q2.x));
      o0.y = fmaf(p0.x, q0.y, fmaf(p0.y, q1.y, p0.z * q2.y));
      o0.z = fmaf(p0.x, q0.z, fmaf(p0.y, q1.z, p0.z * q2.z));
      o0.w = fmaf(p0.x, q0.w, fmaf(p0.y, q1.w, fmaf(p0.z, q2.w, p0.w)));
      o1.x = fmaf(p1.x, q0.x, fmaf(p1.y, q1.x, p1.z * q2.x));
      o1.y = fmaf(p1.x, q0.y, fmaf(p1.y, q1.y, p1.z * q2.y));
      o1.z = fmaf(p1.x, q0.z, fmaf(p1.y, q1.z, p1.z * q2.z));
      o1.w = fmaf(p1.x, q0.w, fmaf(p1.y, q1.w, fmaf(p1.z, q2.w, p1.w)));
      o2.x = fmaf(p2.x, q0.x, fmaf(p2.y, q1.x, p2.z * q2.x));
      o2.y = fmaf(p2.x, q0.y, fmaf(p2.y, q1.y, p2.z * q2.y));
      o2.z = fmaf(p2.x, q0.z, fmaf(p2.y, q1.z, p2.z * q2.z));
      o2.w = fmaf(p2.x, q0.w, fmaf(p2.y, q1.w, fmaf(p2.z, q2.w, p2.w)));
      L4[0] = o0; L4[1] = o1; L4[2] = o2;
    }
    __syncwarp();
  }
#endif

  // ---- skinning transforms A_j = [G_R | G_t - G_R J_j] (kept in the table), FK joints
  const float tx = (a.transl && live) ? a.transl[3 * b + 0] : 0.f;
  const float ty = (a.transl && live) ? a.transl[3 * b + 1] : 0.f;
  const float tz = (a.transl && live) ? a.transl[3 * b + 2] : 0.f;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      float4* G4 = reinterpret_cast<float4*>(Gs + j * 12);
      float4 g0 = G4[0], g1 = G4[1], g2 = G4[2];
      if (a.joints && live) {
        float* jo = a.joints + (size_t)b * a.joints_ld + 3 * j;
        jo[0] = g0.w + tx; jo[1] = g1.w + ty; jo[2] = g2.w + tz;
      }
      g0.w -= g0.x * Jr[s][0] + g0.y * Jr[s][1] + g0.z * Jr[s][2];
      g1.w -= g1.x * Jr[s][0] + g1.y * Jr[s][1] + g1.z * Jr[s][2];
      g2.w -= g2.x * Jr[s][0] + g2.y * Jr[s][1] + g2.z * Jr[s][2];
      if (!live) g0 = g1 = g2 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.A && live) {
        float4* dst = reinterpret_cast<float4*>(a.A + ((size_t)b * m.J + j) * 12);
        dst[0] = g0; dst[1] = g1; dst[2] = g2;
      }
      G4[0] = g0; G4[1] = g1; G4[2] = g2;
    }
  }
  // ---- joint-major copy for the fused kernel: warp w writes joints w, w + 32; lane = body, so a
  // joint's 32 x 48 bytes leave as one contiguous 1,536-byte run (transl folded into the translation
  // column; bodies past the batch, up to the 256-body block, are zero transforms)
  if (a.At != nullptr) {                          // lane = body of the block (nw <= 32 bodies)
    __syncthreads();
    const int bb = blockIdx.x * nw + lane;
    const float* Gl = pose_smem + lane * L.per_warp + max(m.Kpad, 32);
    float ttx = 0.f, tty = 0.f, ttz = 0.f;
    if (a.transl && bb < a.B) { ttx = a.transl[3 * bb]; tty = a.transl[3 * bb + 1]; ttz = a.transl[3 * bb + 2]; }
    if (lane < nw && bb < a.At_rows)
    for (int j = warp; j < m.J; j += nw) {
      const float4* G4 = reinterpret_cast<const float4*>(Gl + j * 12);
      float4 g0 = G4[0], g1 = G4[1], g2 = G4[2];
      g0.w += ttx; g1.w += tty; g2.w += ttz;
      float4* dst = reinterpret_cast<float4*>(a.At + (((size_t)(bb >> 7) * m.J + j) * 128 + (bb & 127)) * 12);
      dst[0] = g0; dst[1] = g1; dst[2] = g2;
    }
  }
}

// Stand-alone Rodrigues (utils/geometry.py:9-23 batch_rodrigues).
__global__ void rodrigues_kernel(int n, const float* __restrict__ aa, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9];
  rodrigues(aa[3 * i], aa[3 * i + 1], aa[3 * i + 2], R);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[(size_t)9 * i + k] = R[k];
}

}  // namespace smplk
