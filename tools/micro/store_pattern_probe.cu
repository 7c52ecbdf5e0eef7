// Probe: HBM write bandwidth of the store patterns a "lane = vertex" tensor-core replay epilogue could use for the
// (F, V, 3) fp32 output.  A warp owns 32 consecutive vertices and walks frames; per frame it writes 384 contiguous bytes.
//   pattern 0: three STG.32, lane l writes float 3 l + c   (12-byte lane stride: every instruction touches 12 sectors, a third each)
//   pattern 1: three STG.32, lane l writes float 32 t + l  (fully coalesced; needs a transposition in front)
//   pattern 2: one 12-byte store per lane as STG.64 + STG.32 when 8-byte aligned (even lanes) / STG.32 + STG.64 (odd lanes)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_pattern_probe tools/micro/store_pattern_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int kPattern>
__global__ void __launch_bounds__(256) store_kernel(float* __restrict__ out, int V, int F, int frames_per_warp) {
  const int lane = threadIdx.x & 31;
  const int vblocks = (V + 31) / 32;
  const long warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long items = (long)vblocks * ((F + frames_per_warp - 1) / frames_per_warp);
  for (long it = warp_global; it < items; it += nwarps) {
    const int vb = (int)(it % vblocks);
    const int f0 = (int)(it / vblocks) * frames_per_warp;
    const int v0 = vb * 32;
    const int nv = min(32, V - v0);
    for (int f = f0; f < min(F, f0 + frames_per_warp); ++f) {
      float* row = out + ((size_t)f * V + v0) * 3;
      const float x = (float)(f + lane), y = x + 1.f, z = x + 2.f;
      if (kPattern == 0) {
        if (lane < nv) { row[3 * lane] = x; row[3 * lane + 1] = y; row[3 * lane + 2] = z; }
      } else if (kPattern == 1) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (32 * t + lane < 3 * nv) row[32 * t + lane] = x + t;
      } else {
        if (lane < nv) {
          float* p = row + 3 * lane;
          if (((reinterpret_cast<size_t>(p) >> 2) & 1) == 0) {
            *reinterpret_cast<float2*>(p) = make_float2(x, y);
            p[2] = z;
          } else {
            p[0] = x;
            *reinterpret_cast<float2*>(p + 1) = make_float2(y, z);
          }
        }
      }
    }
  }
}

template <int kPattern>
static void run(float* d, int V, int F, int fpw, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) store_kernel<kPattern><<<blocks, 256>>>(d, V, F, fpw);
  cudaEventRecord(e0);
  const int n = 10;
  for (int i = 0; i < n; ++i) store_kernel<kPattern><<<blocks, 256>>>(d, V, F, fpw);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  ms /= n;
  printf("pattern %d V=%d F=%d frames/warp=%d blocks=%d: %.4f ms  %.2f TB/s  (%s)\n", kPattern, V, F, fpw, blocks, ms,
         (double)F * V * 12 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int V = argc > 1 ? atoi(argv[1]) : 6890;
  const int F = argc > 2 ? atoi(argv[2]) : 16384;
  float* d; cudaMalloc(&d, (size_t)F * V * 12);
  for (int fpw : {16, 64}) for (int blocks : {148 * 4, 148 * 8}) {
    run<0>(d, V, F, fpw, blocks);
    run<1>(d, V, F, fpw, blocks);
    run<2>(d, V, F, fpw, blocks);
  }
  return 0;
}
