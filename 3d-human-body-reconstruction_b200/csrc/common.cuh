// Shared definitions: packed device-side model constants and small helpers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smplk {

constexpr int kMaxJoints = 64;     // two joints per lane in the warp-per-body kernels
constexpr int kBlendBM = 128;      // bodies per blend-GEMM tile (UMMA M)
constexpr int kBlendBN = 256;      // vertex coordinates per blend-GEMM tile (UMMA N)
constexpr int kBlendBK = 32;       // tf32 elements per k-block = one 128-byte swizzle row
constexpr int kSkinTileVerts = 1024;
constexpr int kGrpJoints = 8;      // distinct joints a 4-vertex group may reference on the fast path

// Device pointers + sizes of one packed body model. Passed by value to kernels.
struct ModelDev {
  int V, J, NB, P;       // vertices, joints, betas, pose features 9(J-1)
  int K, Kpad;           // K = P + NB blend-GEMM depth; Kpad = round_up(K, 32)
  int N, Npad;           // N = 3V; Npad = round_up(N, 256)
  int E, R, C;           // extra vertex picks, posed-vertex regressors, hand PCA comps
  int max_depth;
  int ell_k;             // LBS weights per vertex kept by the packer
  int lbs_only;
  const float* bias;          // [Npad] v_template flattened (zero padded)
  const float* pd_nk_hi;      // [Npad][Kpad] tf32-rounded (posedirs|shapedirs)^T, K contiguous
  const float* pd_nk_lo;      // [Npad][Kpad] residual  x - hi
  const __half* pd_nk_h_hi;   // [Npad][Kpad] fp16 hi term of posedirs * pd_scale
  const __half* pd_nk_h_lo;   // [Npad][Kpad] fp16 lo term
  const __half* pd_kn_h_hi;   // [Kpad][Npad] fp16 hi term, feature-major (backward GEMM operand)
  const __half* pd_kn_h_lo;   // [Kpad][Npad] fp16 lo term
  const uint16_t* pd_kn_b_hi; // [Kpad][Npad] bf16 hi term of the same scaled operand (fitting-step backward)
  const uint16_t* pd_kn_b_lo; // [Kpad][Npad] bf16 lo term
  float pd_scale;             // power of two; the f16 GEMM epilogue multiplies by 1/pd_scale
  const float* pd_kn;         // [Kpad][Npad] exact fp32, N contiguous
  const float* pd_kn_hi;      // [Kpad][Npad] tf32-rounded (backward GEMM operand)
  const float* pd_kn_lo;      // [Kpad][Npad] residual
  const float* J_template;    // [J][3]
  const float* J_shapedirs;   // [J][3][NB]
  const int* parents;         // [J]
  const int* depth;           // [J]
  const int* child_ptr;       // [J+1] children of joint j: child_idx[child_ptr[j] .. child_ptr[j+1])
  const int* child_idx;       // [J-1]
  const int* order;           // [J] joints sorted by depth (level-major)
  const int* level_start;     // [max_depth+2] offsets into `order`
  const uint32_t* skin_idx4;  // [V] four u8 joint ids (ell_k <= 4)
  const float4* skin_w4;      // [V] four weights      (ell_k <= 4)
  const int* ell_idx;         // [ell_k][V]
  const float* ell_w;         // [ell_k][V]
  int grp_ok;                 // 1 when every 4-vertex group touches <= 8 distinct joints
  const uint2* grp_joints;    // [ceil(V/4)] eight u8 joint ids of the group (most weight first)
  const float4* grp_w;        // [ceil(V/4)][8] weight of joint u for the group's 4 vertices
  int grp8_ok;                // same for 8-vertex groups
  const uint2* grp8_joints;   // [ceil(V/8)]
  const float4* grp8_w;       // [ceil(V/8)][8 joints][2] weights of the group's 8 vertices
  const int* grp8_ovf_ptr;    // [ceil(V/8)+1] groups bound to more than 8 joints: their remaining (joint, weights) entries
  const int* grp8_ovf_joint;  // [n_ovf]
  const float4* grp8_ovf_w;   // [n_ovf][2]
  const int* csc_ptr;         // [J+1] joint -> (vertex, weight) lists for the backward
  const int* csc_vert;        // [nnz]
  const float* csc_w;         // [nnz]
  // joint -> vertex lists cut into segments of <= 128 entries (dA_seg_kernel): one warp owns one segment for a
  // run of bodies, its 4 entries per lane stay in registers
  int seg_count;              // S
  int w_rows_normalised;      // 1 when every vertex's weights sum to 1 (d_transl = sum_j dA[j][:,3] then)
  const int* seg_beg;         // [S] first entry in csc_vert / csc_w
  const int* seg_len;         // [S] 1..128
  const int* joint_seg_ptr;   // [J+1] segments of joint j: joint_seg_ptr[j] .. joint_seg_ptr[j+1]
  const float* comp_l;        // [C][45]
  const float* comp_r;        // [C][45]
  const float* pose_mean;     // [3J]
  const int* extra_vids;      // [E]
  const float* pick_pd;       // [3E][Kpad] exact fp32 rows of [posedirs | shapedirs] at the picked vertices' coordinates (or null)
  const int* reg_ptr;         // [R+1] CSR of regressor_posed
  const int* reg_col;         // [nnzR]
  const float* reg_val;       // [nnzR]
  // backward of the vertex picks + posed-vertex regressors as a GATHER per touched vertex (deterministic):
  // vertex sc_vert[t] receives sum_n sc_w[n] * g(sc_src[n]),  n in [sc_ptr[t], sc_ptr[t+1]);
  // sc_src < E: d_joints row J + src (a pick);  sc_src >= E: d_joints_regressed row src - E
  int sc_T;
  const int* sc_vert;         // [T]
  const int* sc_ptr;          // [T+1]
  const int* sc_src;          // [nnz]
  const float* sc_w;          // [nnz]
  // fused blend+skinning kernel (blend_skin_fused.cuh): 84 whole vertices per 256-column tile
  int fz_ok;                  // 1 when the per-chunk joint lists are short enough for the fused epilogue
  int fz_tiles;               // ceil(V / 84)
  const __half* pdf_h_hi;     // [fz_tiles*256][Kpad] fp16 hi term, row 256 t + c <-> coordinate 252 t + c
  const __half* pdf_h_lo;     // [fz_tiles*256][Kpad] fp16 lo term
  const float* bias_f;        // [fz_tiles*256 + 64] v_template in the same column layout
  const int* fz_off;          // [fz_tiles*7 + 1] first (joint, weights) entry of every 12-vertex chunk
  const int* fz_joint;        // [entries]
  const float4* fz_w;         // [entries][4] weights of the chunk's 12 vertices for that joint (+ 4 zeros)
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace smplk
