// pick_forward_kernel -- the joints-only forward (return_verts=False: the camera-initialisation
// closure of lib/Gen_SMPLH/fit_single_frame.py:308-313 and guess_init, fitting.py:82).  Upstream's
// VertexJointSelector needs E (= 21) posed vertices, not 6,890: per body
//   v_posed[3 v_e + c] = v_template + sum_k feature[k] * pd[k, 3 v_e + c]     (63 dot products over the
//                        pre-gathered exact-fp32 operand rows ModelDev::pick_pd)
//   joint[J + e]       = (sum_k w_k A_{j_k}) [v_posed(v_e); 1] + transl
// instead of the 20,736-column blend GEMM and the skinning of every vertex.  The picked v_posed
// entries are also stored at their places in the workspace rows, which is all the sparse
// keypoint backward (pick_backward_kernel) reads.  One block per body.
#pragma once
#include "common.cuh"

namespace smplk {

struct PickFwdArgs {
  int B;
  const __half* H_hi;       // [rows][Kpad] fp16 split feature rows, or null
  const __half* H_lo;
  const float* F_hi;        // [rows][Kpad] tf32 split feature rows, or null (both null: no blendshapes)
  const float* F_lo;
  const float* A;           // (B,J,12)
  const float* transl;      // (B,3) or null
  float* vposed;            // (B,Npad) workspace rows (only the picked entries are written), or null
  float* joints;            // (B, joints_ld); picks go to columns 3J ..
  int joints_ld;
};

constexpr int kPickFwdThreads = 512;
__host__ __device__ inline size_t pick_fwd_smem_bytes(int J, int E, int Kpad) {
  return (size_t)(Kpad + J * 12 + 3 * E + 4) * sizeof(float);
}

__global__ void __launch_bounds__(kPickFwdThreads) pick_forward_kernel(const ModelDev m, const PickFwdArgs a) {
  extern __shared__ __align__(16) float pf_smem[];
  float* feat = pf_smem;                  // [Kpad]
  float* sA = feat + m.Kpad;              // [J][12]
  float* svp = sA + m.J * 12;             // [3E]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool blend = a.H_hi != nullptr || a.F_hi != nullptr;
  for (int k = tid; k < m.Kpad; k += kPickFwdThreads) {
    float x = 0.f;
    if (a.H_hi != nullptr) x = __half2float(a.H_hi[(size_t)b * m.Kpad + k]) + __half2float(a.H_lo[(size_t)b * m.Kpad + k]);
    else if (a.F_hi != nullptr) x = a.F_hi[(size_t)b * m.Kpad + k] + a.F_lo[(size_t)b * m.Kpad + k];
    feat[k] = x;
  }
  for (int i = tid; i < m.J * 12; i += kPickFwdThreads) sA[i] = a.A[(size_t)b * m.J * 12 + i];
  __syncthreads();
  for (int i = warp; i < 3 * m.E; i += kPickFwdThreads / 32) {
    const int col = 3 * m.extra_vids[i / 3] + i % 3;
    float acc = 0.f;
    if (blend) {
      const float* row = m.pick_pd + (size_t)i * m.Kpad;
      for (int k = lane; k < m.Kpad; k += 32) acc = fmaf(feat[k], row[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) {
      const float x = acc + m.bias[col];
      svp[i] = x;
      if (a.vposed != nullptr) a.vposed[(size_t)b * m.Npad + col] = x;
    }
  }
  __syncthreads();
  for (int e = tid; e < m.E; e += kPickFwdThreads) {
    const int v = m.extra_vids[e];
    const float x = svp[3 * e], y = svp[3 * e + 1], z = svp[3 * e + 2];
    float o[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < m.ell_k; ++k) {
      const float w = m.ell_w[(size_t)k * m.V + v];
      if (w == 0.f) continue;
      const float* Aj = sA + m.ell_idx[(size_t)k * m.V + v] * 12;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        o[r] = fmaf(w, fmaf(Aj[4 * r], x, fmaf(Aj[4 * r + 1], y, fmaf(Aj[4 * r + 2], z, Aj[4 * r + 3]))), o[r]);
    }
    float* out = a.joints + (size_t)b * a.joints_ld + 3 * (m.J + e);
#pragma unroll
    for (int r = 0; r < 3; ++r) out[r] = o[r] + (a.transl ? a.transl[3 * b + r] : 0.f);
  }
}

}  // namespace smplk
