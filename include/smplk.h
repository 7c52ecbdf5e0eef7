/* smplk.h -- C ABI of the B200-native SMPL / SMPL-H body-model path.
 *
 * This is the drop-in boundary for the body-model forward/backward of
 * bokchoy-mian/3D-human-body-reconstruction.  Every entry point cites the reference interface it
 * replaces (paths relative to the reference repo).  Plain pointers and sizes only: no torch types,
 * no C++ types.  All `float*` data arguments of forward/backward/skin/regress are DEVICE pointers
 * on the model's device unless the function name ends in `_host`.
 *
 * Error convention: every function returns 0 on success, a negative SMPLK_E_* code for argument /
 * shape / device errors, or a positive cudaError_t value for CUDA failures.  Nothing throws across
 * the ABI; `smplk_last_error_string()` returns a per-thread description of the last failure.
 *
 * Threading / streams: work is launched on the caller-supplied stream (a cudaStream_t passed as
 * void*), with no internal synchronisation and no hidden allocation in forward/backward; a model
 * handle is bound to one device and immutable after creation (one handle per GPU for multi-GPU).
 */
#ifndef SMPLK_H_
#define SMPLK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPLK_VERSION 2   /* 2: + smplk_skin_transforms, smplk_remove_rest (round 2) */

#define SMPLK_OK 0
#define SMPLK_E_ARG (-1)       /* null / inconsistent argument */
#define SMPLK_E_SHAPE (-2)     /* unsupported shape (e.g. J > 64) */
#define SMPLK_E_WORKSPACE (-3) /* workspace too small or misaligned */
#define SMPLK_E_DEVICE (-4)    /* no CUDA device / wrong device / not sm_100 */
#define SMPLK_E_UNSUPPORTED (-5)

typedef struct smplk_model smplk_model;
typedef void* smplk_stream; /* cudaStream_t */

/* Host-side description of a body model; all arrays are HOST pointers, float64 as in the official
 * pickles (reference loads them at models/smplh_np.py:8-17, models/smpl_np.py:124-133; upstream
 * smplx registers the same tensors as buffers).  A rigged-mesh ("LBS only") model, the
 * RecoverModel of lib/model2video.py:12-40 / lib/mesh2smpl_model.py:131-181, has no blendshapes
 * and FIXED rest joints: set shapedirs = posedirs = J_regressor = NULL and give `joints_fixed`. */
typedef struct {
  int32_t num_verts;            /* V (6890 for SMPL/SMPL-H; arbitrary for a rigged mesh)          */
  int32_t num_joints;           /* J (24 SMPL, 52 SMPL-H; <= 64)                                  */
  int32_t num_betas;            /* NB (10 or 16; 0 for a rigged mesh)                             */
  const double* v_template;     /* (V,3)                                                          */
  const double* shapedirs;      /* (V,3,NB)  or NULL                                              */
  const double* posedirs;       /* (V,3,9*(J-1)) pickle layout, or NULL                           */
  const double* J_regressor;    /* (J,V) dense row-major, or NULL                                 */
  const double* joints_fixed;   /* (J,3) rest joints for a rigged mesh, else NULL                 */
  const double* weights;        /* (V,J) LBS weights                                              */
  const int32_t* parents;       /* (J) parents[0] = -1, parents[i] < i                            */
  int32_t num_pca;              /* hand PCA components C (0 = none); upstream smplx use_pca       */
  const double* hand_comp_l;    /* (C,45) or NULL                                                 */
  const double* hand_comp_r;    /* (C,45) or NULL                                                 */
  const double* pose_mean;      /* (3J) added to the full pose (upstream `pose_mean`), or NULL    */
  int32_t num_extra_verts;      /* E: VertexJointSelector vertex picks appended to the FK joints  */
  const int32_t* extra_vertex_ids; /* (E) or NULL                                                 */
  int32_t num_regressors;       /* number of rows R of `regressor_posed` (0 = none)               */
  const double* regressor_posed;/* (R,V): joints regressed from POSED vertices, i.e. the
                                   J_regressor_extra of models/smplh.py:22-29 and gen_J_3d of
                                   models/smplh_np.py:116-117                                     */
} smplk_model_desc;

/* Replaces: SMPLHModel.__init__ (models/smplh_np.py:7-37), SMPLModel.__init__
 * (models/smpl_np.py:123-156), RecoverModel.__init__ (lib/model2video.py:14-40) and the buffer
 * registration of upstream smplx.SMPL/SMPLH.__init__ behind models/smplh.py:16-24.
 * Packs the constants on `device`: posedirs|shapedirs as two-term-split GEMM operands (fp16 hi/lo,
 * TF32 hi/lo; vertex-major and feature-major; also in the 84-vertex column tiling of the fused
 * blend+skinning kernel) with their TMA descriptors, sparse LBS weights and the per-chunk joint
 * tables of the fused epilogue, J_template / J_shapedirs, tree depth tables. */
int smplk_model_create(const smplk_model_desc* desc, int device, smplk_model** out);
int smplk_model_destroy(smplk_model* model);

/* Model facts a binding needs to size its buffers. */
typedef struct {
  int32_t num_verts, num_joints, num_betas, num_pose_feats; /* V, J, NB, P */
  int32_t num_extra_verts, num_regressors, num_pca;
  int32_t max_weights_per_vertex; /* ELL width chosen by the packer (<=4 takes the register path) */
  int32_t lbs_only;               /* 1 for a rigged mesh                                          */
  int32_t device;
  int32_t has_tcgen05_path;       /* 1 when the TMA descriptors for the blend GEMM were built     */
} smplk_model_info;
int smplk_model_get_info(const smplk_model* model, smplk_model_info* info);

/* Kernel choices of one handle (the defaults are the product path; the library reads no environment variable).
 * Used by the parity tests to cross-check alternative kernels and by the bench to time stand-alone kernels:
 *   "fused" (1)          fused blend GEMM + skinning forward kernel; 0 = blend GEMM, then skinning kernel
 *   "pose_block" (1)     block-level pose kernel; 0 = warp-per-body kernel + transposition pass
 *   "blend_tf32" (0)     1 = 3xTF32 operands in the forward GEMM instead of the fp16 two-term split
 *   "backward_tf32" (0)  1 = 3xTF32 backward GEMM
 *   "gemm_2cta" (1)      CTA-pair GEMM kernels; 0 = 1-CTA kernels
 *   "fit_fused" (1)      smplk_fit_vertex_l2 as one skinning + loss + skinning-backward kernel
 *   "sparse_picks" (1)   joints-only gradients through the sparse pick kernels; 0 = dense vertex backward
 *   "fused_tma_out" (1)  the fused forward kernel writes its result as TMA tensor stores (V even, 16-byte aligned
 *                        verts); 0 = per-lane stores
 *   "skin_gemm" (0)      1 = the skinning pass of the two-kernel forward blends the transforms on the tensor cores
 *                        (same kernel as the replay GEMM, 12 accumulator columns per body); measured on a par with the
 *                        streaming kernel, kept for cross-checks
 *   "replay_gemm" (1)    rigged-mesh (LBS-only) handles replay >= 64 frames as one tensor-core GEMM
 *                        (lib/model2video.py:55-85); 0 = streaming skinning kernel
 *   "pdl" (1)            the kernels of a call are launched with programmatic stream serialization (each kernel's set-up
 *                        overlaps its predecessor's tail; results are bit-identical); 0 = plain stream ordering
 *   "skip_pose" (0)      measurement aid: 1 = smplk_forward launches no pose kernel and reuses the workspace rows (blend
 *                        features, transforms, FK joints) the previous call left there, i.e. it repeats that call's result
 *                        with the blend / skinning kernel alone (bench.py times back-to-back launches of it this way)
 * Set options before the first forward that they affect; unknown names return SMPLK_E_ARG. */
int smplk_model_set_option(smplk_model* model, const char* name, int value);

#define SMPLK_FLAG_SAVE_FOR_BACKWARD 1u /* keep v_posed & transforms of ALL bodies in the workspace */
#define SMPLK_FLAG_ADD_POSE_MEAN 2u     /* full_pose += pose_mean (flat_hand_mean=False upstream)   */
#define SMPLK_FLAG_BLEND_SIMT 4u        /* force the exact-fp32 SIMT blend kernel (small batch / bring-up) */
#define SMPLK_FLAG_BLEND_TCGEN05 8u     /* force a tcgen05 blend GEMM even for tiny batches           */
#define SMPLK_FLAG_BLEND_TF32 16u       /* force the 3xTF32 operand format (default: fp16 two-term split) */
#define SMPLK_FLAG_TRANSFORMS_ONLY 32u  /* run the pose / FK kernel only: skinning transforms A (workspace), FK
                                           joints, full_pose; with E > 0 the vertex-pick joints are left unwritten */
#define SMPLK_FLAG_FIT_VERTEX_L2 64u    /* workspace / backward of smplk_fit_vertex_l2 (implies SAVE_FOR_BACKWARD) */
#define SMPLK_FLAG_LOSS_SUM 128u        /* smplk_fit_vertex_l2: `loss` is ONE float, the sum over the bodies;
                                           smplk_backward: `d_loss` is one float (the gradient of that sum) */

/* Bytes of device workspace `smplk_forward` needs for `batch` bodies (256-byte aligned base). */
size_t smplk_workspace_bytes(const smplk_model* model, int32_t batch, uint32_t flags);

/* Diagnostic: byte offsets of the workspace segments {F_hi, F_lo, A, v_posed} and the chunk size
 * (bodies processed per pass) for (batch, flags).  Used by the parity tests to read intermediates:
 * A = per-joint 3x4 skinning transforms (B,J,12), v_posed rows of stride round_up(3V,256). */
int smplk_workspace_layout(const smplk_model* model, int32_t batch, uint32_t flags,
                           size_t offsets[4], int32_t* chunk);

typedef struct {
  int32_t batch;              /* B                                                                 */
  uint32_t flags;             /* SMPLK_FLAG_*                                                      */
  const float* betas;         /* (betas_batch, NB); NULL = zeros                                   */
  int32_t betas_batch;        /* 1 (broadcast, upstream lbs batch_size = max(...)) or B            */
  const float* pose;          /* (B, 3J) axis-angle: global_orient | body_pose | hands             */
  const float* hand_pca_l;    /* (B, C) or NULL: if given, joints 22..36 = hand_pca_l @ comp_l     */
  const float* hand_pca_r;    /* (B, C) or NULL: joints 37..51                                     */
  const float* transl;        /* (B,3) or NULL                                                     */
  float* verts;               /* out (B,V,3), 8-byte aligned, or NULL                              */
  float* joints;              /* out (B, J+E, 3): FK joints then vertex picks, + transl; or NULL   */
  float* joints_regressed;    /* out (B, R, 3): regressor_posed @ verts; or NULL                   */
  float* full_pose;           /* out (B, 3J) assembled pose (PCA + mean applied); or NULL          */
  void* workspace;
  size_t workspace_bytes;
  smplk_stream stream;
} smplk_forward_args;

/* Replaces, per batch of bodies: upstream smplx.SMPLH.forward / SMPL.forward + lbs() as called at
 * models/smplh.py:26-31 and lib/Gen_SMPLH/fitting.py:243-245, and the numpy twins
 * SMPLHModel.set_params/update/compute_R_G/do_skinning (models/smplh_np.py:39-86),
 * SMPLModel (models/smpl_np.py:158-206), RecoverModel.set_params (lib/model2video.py:42-81).
 * verts == NULL with joints != NULL (return_verts=False, fit_single_frame.py:313): only the E
 * vertex-pick joints' vertices are blended and skinned.  With SMPLK_FLAG_SAVE_FOR_BACKWARD the
 * workspace then holds v_posed at those vertices only, i.e. a later smplk_backward may be given
 * d_joints (FK joints and picks) but not d_verts / d_joints_regressed. */
int smplk_forward(const smplk_model* model, const smplk_forward_args* args);

typedef struct {
  int32_t batch;
  uint32_t flags;             /* same flags as the forward that filled `workspace`                 */
  const float* betas;         /* inputs of the forward (transforms are recomputed from them)       */
  int32_t betas_batch;
  const float* pose;
  const float* hand_pca_l;
  const float* hand_pca_r;
  const float* d_verts;       /* in (B,V,3) dL/dverts or NULL                                      */
  const float* d_joints;      /* in (B,J+E,3) dL/djoints or NULL                                   */
  const float* d_joints_regressed; /* in (B,R,3) or NULL                                           */
  float* d_betas;             /* out (betas_batch, NB) or NULL                                     */
  float* d_pose;              /* out (B,3J) or NULL (w.r.t. the `pose` argument)                   */
  float* d_hand_pca_l;        /* out (B,C) or NULL                                                 */
  float* d_hand_pca_r;        /* out (B,C) or NULL                                                 */
  float* d_transl;            /* out (B,3) or NULL                                                 */
  void* workspace;            /* the forward's workspace (SMPLK_FLAG_SAVE_FOR_BACKWARD)            */
  size_t workspace_bytes;
  void* scratch;              /* smplk_backward_scratch_bytes() bytes                              */
  size_t scratch_bytes;
  smplk_stream stream;
  const float* d_loss;        /* (B) or NULL: every parameter gradient of body b is multiplied by d_loss[b]
                                 (the upstream gradient of a per-body loss whose d_verts was formed for
                                 d_loss = 1, e.g. after smplk_fit_vertex_l2); a shared betas row receives
                                 sum_b d_loss[b] d_betas_b */
  const float* d_full_pose;   /* (B,3J) or NULL: gradient w.r.t. the forward's full_pose output (pose priors
                                 read it, lib/Gen_SMPLH/fitting.py:383-413); flows to d_pose / d_hand_pca_* */
} smplk_backward_args;

size_t smplk_backward_scratch_bytes(const smplk_model* model, int32_t batch);

/* Replaces the autograd backward of the forward above (total_loss.backward() at
 * lib/Gen_SMPLH/fitting.py:256): vector-Jacobian product w.r.t. betas, pose, hand PCA, transl. */
int smplk_backward(const smplk_model* model, const smplk_backward_args* args);

/* Replaces gen_J_3d (models/smplh_np.py:116-117, models/smpl_np.py:230-231,
 * lib/mesh2smpl_model.py:112-113) and vertices2joints(J_regressor_extra, vertices) at
 * models/smplh.py:29: out (B,R,3) = regressor_posed (R,V) @ verts (B,V,3). */
int smplk_regress_joints(const smplk_model* model, int32_t batch, const float* verts, float* out,
                         smplk_stream stream);

/* Replaces do_skinning(G) of the numpy twins (models/smplh_np.py:72-82, models/smpl_np.py:191-202,
 * lib/model2video.py:66-81): the caller hands in the GLOBAL joint transforms G (B,J,4,4) row-major it got
 * from compute_R_G (possibly edited), the rest joints (B,J,3) and the blended vertices;
 *   A[b,j] = [G_R | G_t - G_R J_j]   (written to `A`, (B,J,12), also an output)
 *   verts[b,v] = (sum_k w[v,k] A[b,j_k]) [v_posed[b,v]; 1] + transl[b]
 * v_posed: (B, v_posed_ld) rows, 16-byte aligned (v_posed_ld % 4 == 0, >= 3V); NULL for a rigged mesh
 * (its own v_template is skinned).  All pointers are device pointers. */
int smplk_skin_transforms(const smplk_model* model, int32_t batch, const float* G,
                          const float* joints_rest, const float* v_posed, int32_t v_posed_ld,
                          const float* transl /* (B,3) or NULL */, float* A, float* verts,
                          smplk_stream stream);

/* The rest-pose removal alone: A[b,j] (3x4) = [G_R | G_t - G_R J_j] from global transforms G (B,J,4,4) and
 * rest joints (B,J,3) -- the `G - pack(G . [J;0])` step of do_skinning (models/smplh_np.py:73-78) and of
 * RecoverModel.to_T_pose (lib/mesh2smpl_model.py:194-199), whose result feeds smplk_inverse_lbs /
 * smplk_inverse_joints.  Device pointers. */
int smplk_remove_rest(int32_t batch, int32_t num_joints, const float* G, const float* joints_rest,
                      float* A /* out (B,J,12) */, int device, smplk_stream stream);

/* Replaces utils/geometry.py:9-23 batch_rodrigues (axis-angle (n,3) -> rotation matrices (n,3,3)). */
int smplk_batch_rodrigues(int32_t n, const float* axis_angle, float* rotmats, int device,
                          smplk_stream stream);

/* Fused vertex data term of a fitting step (the squared-L2 loss of
 * lib/Gen_SMPLH/fitting.py:491-495 on vertices; BASELINE config 3): loss[b] = scale * sum over
 * the body's floats of (verts - target)^2 and, if `grad` is given, grad = 2 * scale * (verts -
 * target), in one pass.  Feed `grad` to smplk_backward as d_verts.  `grad` may be the same buffer
 * as `verts` (the gradient then replaces the vertices). */
int smplk_vertex_l2(int32_t batch, int32_t floats_per_body, const float* verts, const float* target,
                    float scale, float* grad, float* loss, int device, smplk_stream stream);

/* End-to-end call with HOST buffers (the reference's numpy API works on host arrays:
 * set_params(pose, beta, trans) -> verts, models/smplh_np.py:39-47).  Copies inputs H2D,
 * runs the forward on `stream`, copies verts/joints D2H and synchronises the stream.  Device
 * staging buffers are owned by the model and grown on demand. */
int smplk_forward_host(smplk_model* model, int32_t batch, uint32_t flags, const float* betas,
                       int32_t betas_batch, const float* pose, const float* transl, float* verts,
                       float* joints, smplk_stream stream);

/* The forward half of a vertex-L2 fitting step (BASELINE config 3; the squared-L2 data term of
 * lib/Gen_SMPLH/fitting.py:491-495 on vertices) in one call: body-model forward as smplk_forward
 * with SMPLK_FLAG_SAVE_FOR_BACKWARD | SMPLK_FLAG_FIT_VERTEX_L2 in a->flags, then
 *   loss[b] = scale * sum ||verts[b] - target[b]||^2,   a->verts <- 2 scale (verts - target)
 * i.e. a->verts (B,V,3) receives the vertex GRADIENT, not the vertices.  Skinning, loss, gradient
 * and the skinning backward run as one kernel where the model allows (sparse weights, 3V even);
 * d_v_posed is then kept in the workspace and smplk_backward, called with the same flags and
 * d_verts = a->verts, skips its own skinning backward.  a->joints may be given only when the model
 * has no vertex-pick joints; a->joints_regressed must be null.  `target` (B,V,3) and a->verts must be
 * 8-byte aligned; `loss` (B), or one float with SMPLK_FLAG_LOSS_SUM. */
int smplk_fit_vertex_l2(const smplk_model* model, const smplk_forward_args* a, const float* target,
                        float scale, float* loss);

/* ---- mesh operations either side of the forward (SURVEY.md 8f "next" rows 2 and 4) ----------- */

/* Inverse LBS (un-posing).  Replaces RecoverModel.to_T_pose / to_rest_pose
 * (lib/mesh2smpl_model.py:183-207, :340-372) and models/smpl_np.py:239-246:
 *   v_rest[b,v] = (sum_k w[v,k] A[b,j_k])^-1 [verts[b,v] - transl[b]; 1]
 * `model` supplies the skin weights (any handle, rigged mesh included); `A` (B,J,12) are the 3x4
 * skinning transforms of the pose being removed (the forward's workspace segment 2, see
 * smplk_workspace_layout, with SMPLK_FLAG_SAVE_FOR_BACKWARD).  All pointers are device pointers. */
int smplk_inverse_lbs(const smplk_model* model, int32_t batch, const float* A, const float* verts,
                      const float* transl /* (B,3) or NULL */, float* v_rest /* out (B,V,3) */,
                      smplk_stream stream);
/* J_rest[b,j] = A[b,j]^-1 [J_posed[b,j] - transl[b]; 1]  (lib/mesh2smpl_model.py:205-207). */
int smplk_inverse_joints(int32_t batch, int32_t num_joints, const float* A, const float* joints,
                         int32_t joints_ld /* floats per body row of `joints` */, const float* transl,
                         float* out /* (B,J,3) */, int device, smplk_stream stream);

/* Per-vertex unit normals of posed meshes: area-weighted sum of the incident triangles' cross
 * products (v1-v0)x(v2-v0), normalised.  Replaces VertNormals(verts, faces, True) at
 * utils/render_model.py:36,63-81 (upstream opendr).  `vf_ptr` (V+1) / `vf_face` are the CSR lists
 * vertex -> incident faces (host-built once per topology). */
int smplk_vertex_normals(int32_t batch, int32_t num_verts, const int32_t* faces /* (F,3) */,
                         const int32_t* vf_ptr, const int32_t* vf_face, const float* verts,
                         float* normals /* out (B,V,3) */, int device, smplk_stream stream);

/* Front / back split of posed meshes.  Replaces SMPLHModel.divide_face (models/smplh_np.py:126-182):
 * a triangle is FRONT when the z of (v1-v0)x(v2-v1) is <= 0, BACK otherwise; each side's vertices
 * are listed in order of first appearance (face order, then corner) and its faces re-indexed.
 *   faces_out (B,2,F,3)  side 0 = front, 1 = back; first counts[b][side][0] rows valid
 *   vidx_out  (B,2,V)    original vertex ids; first counts[b][side][1] entries valid
 *   counts    (B,2,2)    {faces, vertices} per side */
int smplk_divide_faces(int32_t batch, int32_t num_verts, int32_t num_faces, const int32_t* faces,
                       const float* verts, int32_t* faces_out, int32_t* vidx_out, int32_t* counts,
                       int device, smplk_stream stream);

/* ---- fitting loss around the body model (SURVEY.md 8f "next" row 1), batched over bodies ------ */

/* Reprojection data term and its gradient.  Replaces, per closure call, PerspectiveCamera.forward
 * (lib/Gen_SMPLH/camera.py:93-117), GMoF (lib/Gen_SMPLH/util.py:60-71) and the joint term of
 * SMPLifyLoss.forward (lib/Gen_SMPLH/fitting.py:369-381); rho <= 0 gives the squared residual of
 * SMPLifyCameraInitLoss (fitting.py:486-495).  All pointers are device pointers.
 *   loss[b] = data_weight^2 sum_j w[b,j]^2 (rob(gt_x - img_x) + rob(gt_y - img_y)),
 *   img = focal * (R p + t).xy / (R p + t).z + center,  rob(r) = rho^2 r^2 / (r^2 + rho^2) */
typedef struct {
  int32_t batch, num_joints;
  const float* joints;        /* (B,Jn,3) model joints after the joint mapper                    */
  const float* rotation;      /* (camera_batch,3,3) row-major                                    */
  const float* translation;   /* (camera_batch,3)                                                */
  const float* focal;         /* (camera_batch,2) fx, fy                                         */
  const float* center;        /* (camera_batch,2)                                                */
  int32_t camera_batch;       /* 1 (shared camera) or B                                          */
  const float* gt_joints;     /* (B,Jn,2) 2-D detections                                         */
  const float* weights;       /* (weights_batch,Jn) joint_weights * joints_conf, or NULL (= 1)   */
  int32_t weights_batch;      /* 1 or B                                                          */
  float rho, data_weight;
  float* loss;                /* out (B)                                                         */
  float* d_joints;            /* out (B,Jn,3) d loss[b] / d joints, or NULL                      */
  float* d_translation;       /* out (B,3) d loss[b] / d camera translation, or NULL             */
  int device;
  smplk_stream stream;
} smplk_reprojection_args;
int smplk_reprojection_loss(const smplk_reprojection_args* args);

/* Priors and their gradients (lib/Gen_SMPLH/fitting.py:383-413, lib/Gen_SMPLH/prior.py:53-97):
 *   shape_weight^2 |betas|^2 + body_pose_weight^2 |pose_embedding|^2 (or |body_pose|^2 when no
 *   embedding is given: L2Prior) + bending_prior_weight sum_k exp(s_k body_pose[i_k])^2 with
 *   i = {55,58,12,15} - 3, s = {1,-1,-1,-1} + hand_prior_weight^2 (|lh|^2 + |rh|^2).
 * body_pose is full_pose[:, 3:66].  NULL inputs drop their term; NULL gradients are skipped. */
typedef struct {
  int32_t batch;
  const float* betas; int32_t num_betas;
  const float* pose_embedding; int32_t num_embedding;
  const float* body_pose; int32_t num_body_pose;
  const float* left_hand_pose; const float* right_hand_pose; int32_t num_hand;
  float shape_weight, body_pose_weight, bending_prior_weight, hand_prior_weight;
  float* loss;                /* out (B) */
  float* d_betas; float* d_pose_embedding; float* d_body_pose; float* d_left_hand_pose; float* d_right_hand_pose;
  int device;
  smplk_stream stream;
} smplk_prior_args;
int smplk_fit_priors(const smplk_prior_args* args);

/* Per-kernel device timing for benchmarks: while enabled, forward/backward bracket every kernel
 * they launch with a CUDA event pair recorded on the caller's stream (the stream the kernel runs
 * on).  `smplk_profile_read` waits for the pending events and returns accumulated milliseconds and
 * launch counts per slot. */
#define SMPLK_PROF_POSE_FWD 0
#define SMPLK_PROF_BLEND_TCGEN05 1
#define SMPLK_PROF_BLEND_SIMT 2
#define SMPLK_PROF_SKIN 3
#define SMPLK_PROF_DA 4
#define SMPLK_PROF_SKIN_BWD 5
#define SMPLK_PROF_BLEND_BWD 6
#define SMPLK_PROF_POSE_BWD 7
#define SMPLK_PROF_BLEND_SKIN_FUSED 8 /* fused blend GEMM + skinning epilogue (forward without SAVE_FOR_BACKWARD) */
#define SMPLK_PROF_TRANSPOSE 9        /* transform re-layout pass feeding a tensor-core kernel (fused forward without the block
                                        pose kernel; operand rows of the rigged-mesh replay GEMM) */
#define SMPLK_PROF_SLOTS 10
int smplk_profile_enable(smplk_model* model, int enable);
int smplk_profile_read(smplk_model* model, double ms[SMPLK_PROF_SLOTS],
                       int64_t counts[SMPLK_PROF_SLOTS], int reset);

const char* smplk_last_error_string(void);
int smplk_version(void);

/* Number of kernels this library has launched since load (benchmark bookkeeping). */
uint64_t smplk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SMPLK_H_ */
