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
__device__ __forceinline__ void load_joint_pose(const ModelDev& m, const float* pose,
                                                const float* pca_l, const float* pca_r,
                                                int add_mean, int b, int j, float* r) {
  const int hand0 = m.J - 30;  // first left-hand joint (22 for SMPL-H)
  if (pca_l != nullptr && j >= hand0 && j < hand0 + 15) {
    const float* comp = m.comp_l + 3 * (j - hand0);
    const float* c = pca_l + (size_t)b * m.C;
    float x = 0.f, y = 0.f, z = 0.f;
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

template <int SLOTS>
__global__ void __launch_bounds__(kPoseWarps * 32)
pose_forward_kernel(const ModelDev m, const PoseFwdArgs a) {
  extern __shared__ float pose_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kPoseWarps + warp;
  if (b >= a.B) return;
  float* feat = pose_smem + warp * m.Kpad;
  const float* betas_row = a.betas ? a.betas + (size_t)(a.betas_B == 1 ? 0 : b) * m.NB : nullptr;

  float rv[SLOTS][3], R[SLOTS][9], Jr[SLOTS][3], Jrel[SLOTS][3], G[SLOTS][12];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    rv[s][0] = rv[s][1] = rv[s][2] = 0.f;
    if (j < m.J) {
      load_joint_pose(m, a.pose, a.pca_l, a.pca_r, a.add_mean, b, j, rv[s]);
      if (a.full_pose) {
        float* fp = a.full_pose + (size_t)b * 3 * m.J + 3 * j;
        fp[0] = rv[s][0]; fp[1] = rv[s][1]; fp[2] = rv[s][2];
      }
    }
  }
  pose_forward_core<SLOTS>(m, betas_row, rv, R, Jr, Jrel, G, lane);

  // ---- GEMM A-operand row: [ (R_j - I) j=1..J-1 | betas | 0 pad ] split into TF32 hi / lo
  if (a.F_hi != nullptr || a.H_hi != nullptr) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int j = lane + 32 * s;
      if (j >= 1 && j < m.J) {
#pragma unroll
        for (int i = 0; i < 9; ++i) feat[9 * (j - 1) + i] = R[s][i] - ((i % 4 == 0) ? 1.f : 0.f);
      }
    }
    for (int i = lane; i < m.Kpad - m.P; i += 32)
      feat[m.P + i] = (i < m.NB && betas_row) ? betas_row[i] : 0.f;
    __syncwarp();
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
        const float x0 = feat[2 * k], x1 = feat[2 * k + 1];
        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        hh[k] = __halves2half2(h0, h1);
        hl[k] = __halves2half2(__float2half_rn(x0 - __half2float(h0)),
                               __float2half_rn(x1 - __half2float(h1)));
      }
    }
  }

  // ---- skinning transforms A_j = [G_R | G_t - G_R J_j]  and FK joints
  const float tx = a.transl ? a.transl[3 * b + 0] : 0.f;
  const float ty = a.transl ? a.transl[3 * b + 1] : 0.f;
  const float tz = a.transl ? a.transl[3 * b + 2] : 0.f;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int j = lane + 32 * s;
    if (j < m.J) {
      float4* dst = reinterpret_cast<float4*>(a.A + ((size_t)b * m.J + j) * 12);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float t = G[s][r * 4 + 3] - (G[s][r * 4 + 0] * Jr[s][0] + G[s][r * 4 + 1] * Jr[s][1] +
                                     G[s][r * 4 + 2] * Jr[s][2]);
        dst[r] = make_float4(G[s][r * 4 + 0], G[s][r * 4 + 1], G[s][r * 4 + 2], t);
      }
      if (a.joints) {
        float* jo = a.joints + (size_t)b * a.joints_ld + 3 * j;
        jo[0] = G[s][3] + tx;
        jo[1] = G[s][7] + ty;
        jo[2] = G[s][11] + tz;
      }
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
