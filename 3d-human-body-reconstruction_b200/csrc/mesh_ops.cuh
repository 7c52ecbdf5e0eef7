// Callers either side of the body-model forward (SURVEY 8f "next" rows 2 and 4):
//
//  * inverse LBS / un-posing: v_rest[b,v] = (sum_k w[v,k] A[b,j_k])^-1 [verts[b,v]; 1]
//      lib/mesh2smpl_model.py:183-207 (to_T_pose), :340-372 (to_rest_pose),
//      models/smpl_np.py:239-246; joints: J_rest[b,j] = A[b,j]^-1 [J_posed[b,j]; 1]  (:205-207)
//  * per-vertex normals of the posed mesh (area-weighted sum of incident triangle cross products,
//      normalised): utils/render_model.py:36 VertNormals(verts, faces, True)  [upstream opendr]
//  * front / back face split by the sign of the triangle normal's z:
//      models/smplh_np.py:126-182 divide_face (z <= 0 -> front), with the reference's
//      first-appearance vertex re-indexing.
//
// All three are HBM/L2-bound gather kernels over (body, vertex) or (body, face).
#pragma once
#include "common.cuh"

namespace smplk {

// 3x4 affine inverse applied to a point: out = R^-1 (x - t), R^-1 by the adjugate.
__device__ __forceinline__ void affine_inverse_apply(const float* T, float x, float y, float z, float* out) {
  const float a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], i = T[10];
  const float c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
  const float c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
  const float c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
  const float det = a * c00 + b * c10 + c * c20;
  const float inv = 1.0f / det;
  const float px = x - T[3], py = y - T[7], pz = z - T[11];
  out[0] = (c00 * px + c01 * py + c02 * pz) * inv;
  out[1] = (c10 * px + c11 * py + c12 * pz) * inv;
  out[2] = (c20 * px + c21 * py + c22 * pz) * inv;
}

// grid (ceil(V/256), B); the body's transforms are staged in smem once per block.
__global__ void __launch_bounds__(256)
inverse_lbs_kernel(const ModelDev m, int B, const float* __restrict__ A, const float* __restrict__ verts,
                   const float* __restrict__ transl, float* __restrict__ out) {
  extern __shared__ __align__(16) float inv_smem[];   // [J][12]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < m.J * 12; i += blockDim.x) inv_smem[i] = A[(size_t)b * m.J * 12 + i];
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m.V) return;
  float T[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) T[q] = 0.f;
  for (int k = 0; k < m.ell_k; ++k) {
    const float w = m.ell_w[(size_t)k * m.V + v];
    if (w == 0.f) continue;
    const float* Aj = inv_smem + m.ell_idx[(size_t)k * m.V + v] * 12;
#pragma unroll
    for (int q = 0; q < 12; ++q) T[q] = fmaf(w, Aj[q], T[q]);
  }
  const float* p = verts + ((size_t)b * m.V + v) * 3;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl) { tx = transl[3 * b]; ty = transl[3 * b + 1]; tz = transl[3 * b + 2]; }
  float r[3];
  affine_inverse_apply(T, p[0] - tx, p[1] - ty, p[2] - tz, r);
  float* o = out + ((size_t)b * m.V + v) * 3;
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}

// J_rest[b,j] = A[b,j]^-1 [J_posed[b,j] - transl[b]; 1]
__global__ void inverse_joints_kernel(int B, int J, const float* __restrict__ A,
                                      const float* __restrict__ joints, int joints_ld,
                                      const float* __restrict__ transl, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * J) return;
  const int b = i / J, j = i % J;
  const float* p = joints + (size_t)b * joints_ld + 3 * j;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (transl) { tx = transl[3 * b]; ty = transl[3 * b + 1]; tz = transl[3 * b + 2]; }
  float T[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) T[q] = A[(size_t)i * 12 + q];
  float r[3];
  affine_inverse_apply(T, p[0] - tx, p[1] - ty, p[2] - tz, r);
  out[(size_t)i * 3] = r[0]; out[(size_t)i * 3 + 1] = r[1]; out[(size_t)i * 3 + 2] = r[2];
}

// Vertex normals by gathering over the vertex -> (face) CSR built on the host (deterministic, no
// atomics).  One thread per (body, vertex).
__global__ void __launch_bounds__(256)
vertex_normals_kernel(int B, int V, const int* __restrict__ faces, const int* __restrict__ vf_ptr,
                      const int* __restrict__ vf_face, const float* __restrict__ verts,
                      float* __restrict__ normals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= V) return;
  const float* vb = verts + (size_t)b * V * 3;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  for (int n = vf_ptr[v]; n < vf_ptr[v + 1]; ++n) {
    const int f = vf_face[n];
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float ax = vb[3 * i1] - vb[3 * i0], ay = vb[3 * i1 + 1] - vb[3 * i0 + 1], az = vb[3 * i1 + 2] - vb[3 * i0 + 2];
    const float bx = vb[3 * i2] - vb[3 * i0], by = vb[3 * i2 + 1] - vb[3 * i0 + 1], bz = vb[3 * i2 + 2] - vb[3 * i0 + 2];
    nx += ay * bz - az * by;
    ny += az * bx - ax * bz;
    nz += ax * by - ay * bx;
  }
  const float len = sqrtf(nx * nx + ny * ny + nz * nz);
  const float inv = len > 0.f ? 1.0f / len : 0.f;
  float* o = normals + ((size_t)b * V + v) * 3;
  o[0] = nx * inv; o[1] = ny * inv; o[2] = nz * inv;
}

// divide_face.  One block per (body, side): side 0 = front (z <= 0), 1 = back.
//   faces_out[b][side][n][3]  faces of that side in original order, re-indexed into the side's vertex list
//   vidx_out[b][side][n]      the side's vertices in order of first appearance (face order, then corner)
//   counts[b][side] = {number of faces, number of vertices}
// The reference builds the lists sequentially; first-appearance order is reproduced in parallel by
// an atomicMin of (3 face + corner) per vertex followed by a prefix sum over those keys.
constexpr int kDivThreads = 1024;

__device__ __forceinline__ int block_exclusive_scan(int val, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = val;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;   // inclusive over warps
  }
  __syncthreads();
  const int base = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return base + x - val;
}

__global__ void __launch_bounds__(kDivThreads)
divide_faces_kernel(int B, int V, int F, const int* __restrict__ faces, const float* __restrict__ verts,
                    int* __restrict__ faces_out, int* __restrict__ vidx_out, int* __restrict__ counts) {
  extern __shared__ int div_smem[];        // first_key[V] | new_index[V] | warp sums[32]
  int* first_key = div_smem;
  int* new_index = div_smem + V;
  int* warp_sums = div_smem + 2 * V;
  const int b = blockIdx.x >> 1, side = blockIdx.x & 1;
  const float* vb = verts + (size_t)b * V * 3;
  int* fo = faces_out + ((size_t)b * 2 + side) * F * 3;
  int* vo = vidx_out + ((size_t)b * 2 + side) * V;
  for (int v = threadIdx.x; v < V; v += blockDim.x) first_key[v] = 0x7fffffff;
  __syncthreads();
  auto face_side = [&](int f) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    // evaluated in double: differences and products of fp32 values are exact there, so the sign --
    // an index decision -- is the one the reference's float64 arithmetic gives for these vertices
    const double m0 = (double)vb[3 * i1] - vb[3 * i0], m1 = (double)vb[3 * i1 + 1] - vb[3 * i0 + 1];
    const double n0 = (double)vb[3 * i2] - vb[3 * i1], n1 = (double)vb[3 * i2 + 1] - vb[3 * i1 + 1];
    const double z = m0 * n1 - n0 * m1;    // models/smplh_np.py:152
    return z <= 0.0 ? 0 : 1;
  };
  // pass 1: first appearance of every vertex among this side's faces
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    if (face_side(f) == side) {
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicMin(&first_key[faces[3 * f + c]], 3 * f + c);
    }
  }
  __syncthreads();
  // pass 2: rank the first-appearance keys (prefix sum over the 3F key space, blockDim keys a round)
  int running = 0;
  for (int k0 = 0; k0 < 3 * F; k0 += blockDim.x) {
    const int key = k0 + threadIdx.x;
    int flag = 0, vert = -1;
    if (key < 3 * F) {
      vert = faces[key];
      flag = (first_key[vert] == key) ? 1 : 0;
    }
    int total;
    const int pos = block_exclusive_scan(flag, warp_sums, &total);
    if (flag) {
      new_index[vert] = running + pos;
      vo[running + pos] = vert;
    }
    running += total;
  }
  const int nverts = running;
  __syncthreads();
  // pass 3: compact the side's faces in order, re-indexed
  running = 0;
  for (int f0 = 0; f0 < F; f0 += blockDim.x) {
    const int f = f0 + threadIdx.x;
    const int flag = (f < F && face_side(f) == side) ? 1 : 0;
    int total;
    const int pos = block_exclusive_scan(flag, warp_sums, &total);
    if (flag) {
#pragma unroll
      for (int c = 0; c < 3; ++c) fo[(size_t)(running + pos) * 3 + c] = new_index[faces[3 * f + c]];
    }
    running += total;
  }
  if (threadIdx.x == 0) {
    counts[((size_t)b * 2 + side) * 2] = running;
    counts[((size_t)b * 2 + side) * 2 + 1] = nverts;
  }
}

}  // namespace smplk
