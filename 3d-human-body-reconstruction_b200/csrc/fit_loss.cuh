// Fitting loss around the body model, batched over bodies: the reference evaluates
// these with ~30 tiny torch ops per closure call at batch_size == 1 (lib/Gen_SMPLH/fit_single_frame.py:97).
//
//  * reprojection data term with its gradient, one warp per body:
//      camera:  pc = R p + t;  img = f * pc.xy / pc.z + c        (lib/Gen_SMPLH/camera.py:93-117)
//      robust:  gmof(r) = rho^2 r^2 / (r^2 + rho^2) per coordinate (lib/Gen_SMPLH/util.py:60-71)
//      loss_b = data_weight^2 * sum_j w_j^2 (gmof(gt_x - img_x) + gmof(gt_y - img_y))
//                                                                 (lib/Gen_SMPLH/fitting.py:369-381)
//      rho <= 0 selects the plain squared residual of SMPLifyCameraInitLoss (fitting.py:486-495).
//      Gradients: d loss / d joints (B,Jn,3) and d loss / d camera translation (B,3).
//  * priors with their gradients, one thread block per body (fitting.py:383-413, prior.py:53-97):
//      shape  : shape_weight^2 * sum betas^2
//      pose   : body_pose_weight^2 * sum x^2  over the pose embedding (VPoser) or the body pose (L2Prior)
//      bending: bending_weight * sum exp(sign_k * body_pose[idx_k])^2, idx = {55,58,12,15} - 3
//      hands  : hand_weight^2 * (sum lh^2 + sum rh^2)
#pragma once
#include "common.cuh"

namespace smplk {

struct ReprojArgs {
  int B, Jn;
  const float* joints;      // (B,Jn,3)
  const float* rotation;    // (Bc,3,3) row-major, Bc = 1 or B
  const float* translation; // (Bc,3)
  const float* focal;       // (Bc,2)
  const float* center;      // (Bc,2)
  int cam_batch;
  const float* gt;          // (B,Jn,2)
  const float* weights;     // (Bw,Jn) joint_weights * conf, Bw = 1 or B; or null (= 1)
  int w_batch;
  float rho, data_weight;
  float* loss;              // (B)
  float* d_joints;          // (B,Jn,3) or null
  float* d_translation;     // (B,3) or null (per body, also when the camera is shared)
};

__global__ void __launch_bounds__(128)
reprojection_loss_kernel(const ReprojArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B) return;
  const int b = warp, cb = a.cam_batch == 1 ? 0 : b;
  const float* R = a.rotation + 9 * cb;
  const float* t = a.translation + 3 * cb;
  const float fx = a.focal[2 * cb], fy = a.focal[2 * cb + 1];
  const float cx = a.center[2 * cb], cy = a.center[2 * cb + 1];
  const float rho2 = a.rho * a.rho, dw2 = a.data_weight * a.data_weight;
  float loss = 0.f, gtx = 0.f, gty = 0.f, gtz = 0.f;
  for (int j = lane; j < a.Jn; j += 32) {
    const float* p = a.joints + ((size_t)b * a.Jn + j) * 3;
    const float x = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
    const float y = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
    const float z = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
    const float iz = 1.0f / z;
    const float u = fx * x * iz + cx, v = fy * y * iz + cy;
    const float rx = a.gt[((size_t)b * a.Jn + j) * 2] - u, ry = a.gt[((size_t)b * a.Jn + j) * 2 + 1] - v;
    float w = a.weights ? a.weights[(size_t)(a.w_batch == 1 ? 0 : b) * a.Jn + j] : 1.f;
    w = w * w * dw2;
    float lx, ly, gx, gy;                   // per-coordinate loss and d loss / d residual
    if (a.rho > 0.f) {
      const float dx = rx * rx + rho2, dy = ry * ry + rho2;
      lx = rho2 * rx * rx / dx; ly = rho2 * ry * ry / dy;
      gx = 2.f * rx * rho2 * rho2 / (dx * dx); gy = 2.f * ry * rho2 * rho2 / (dy * dy);
    } else {
      lx = rx * rx; ly = ry * ry; gx = 2.f * rx; gy = 2.f * ry;
    }
    loss += w * (lx + ly);
    // d loss / d pc:  residual = gt - img  ->  d/d img = -g
    const float du = -w * gx, dv = -w * gy;
    const float dpx = du * fx * iz, dpy = dv * fy * iz;
    const float dpz = -(du * fx * x + dv * fy * y) * iz * iz;
    if (a.d_joints) {
      float* dj = a.d_joints + ((size_t)b * a.Jn + j) * 3;
      dj[0] = R[0] * dpx + R[3] * dpy + R[6] * dpz;     // R^T d pc
      dj[1] = R[1] * dpx + R[4] * dpy + R[7] * dpz;
      dj[2] = R[2] * dpx + R[5] * dpy + R[8] * dpz;
    }
    gtx += dpx; gty += dpy; gtz += dpz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    gtx += __shfl_xor_sync(0xffffffffu, gtx, o);
    gty += __shfl_xor_sync(0xffffffffu, gty, o);
    gtz += __shfl_xor_sync(0xffffffffu, gtz, o);
  }
  if (lane == 0) {
    a.loss[b] = loss;
    if (a.d_translation) {
      a.d_translation[3 * b] = gtx; a.d_translation[3 * b + 1] = gty; a.d_translation[3 * b + 2] = gtz;
    }
  }
}

struct PriorArgs {
  int B;
  const float* betas; int nb;            // (B,nb) or null
  const float* pose_embedding; int ne;   // (B,ne) or null (VPoser latent)
  const float* body_pose; int np;        // (B,np) or null: full_pose[:, 3:66]; L2 term only without an embedding
  const float* lhand; const float* rhand; int nh;   // (B,nh) or null
  float shape_weight, body_pose_weight, bending_weight, hand_weight;
  float* loss;                           // (B)
  float* d_betas; float* d_pose_embedding; float* d_body_pose; float* d_lhand; float* d_rhand;  // or null
};

__global__ void __launch_bounds__(64)
fit_priors_kernel(const PriorArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  float loss = 0.f;
  auto l2 = [&](const float* x, int n, float w, float* g) {
    if (!x) return;
    for (int i = tid; i < n; i += blockDim.x) {
      const float v = x[(size_t)b * n + i];
      loss += w * w * v * v;
      if (g) g[(size_t)b * n + i] = 2.f * w * w * v;
    }
  };
  l2(a.betas, a.nb, a.shape_weight, a.d_betas);
  l2(a.pose_embedding, a.ne, a.body_pose_weight, a.d_pose_embedding);
  l2(a.lhand, a.nh, a.hand_weight, a.d_lhand);
  l2(a.rhand, a.nh, a.hand_weight, a.d_rhand);
  if (a.body_pose) {
    const bool l2_pose = a.pose_embedding == nullptr;
    for (int i = tid; i < a.np; i += blockDim.x) {
      const float v = a.body_pose[(size_t)b * a.np + i];
      float g = 0.f;
      if (l2_pose) {
        loss += a.body_pose_weight * a.body_pose_weight * v * v;
        g = 2.f * a.body_pose_weight * a.body_pose_weight * v;
      }
      // SMPLifyAnglePrior (prior.py:53-97): indices {55,58,12,15} - 3, signs {1,-1,-1,-1}
      float sign = 0.f;
      if (i == 52) sign = 1.f;
      else if (i == 55 || i == 9 || i == 12) sign = -1.f;
      if (sign != 0.f) {
        const float e = expf(v * sign);
        loss += a.bending_weight * e * e;
        g += a.bending_weight * 2.f * e * e * sign;
      }
      if (a.d_body_pose) a.d_body_pose[(size_t)b * a.np + i] = g;
    }
  }
  __shared__ float red[64];
  red[tid] = loss;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) a.loss[b] = red[0];
}

}  // namespace smplk
