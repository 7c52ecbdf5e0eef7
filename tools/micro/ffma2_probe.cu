// Microbenchmark: issue rate / latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, in the operand forms the
// skinning epilogue uses (scalar-broadcast multiplicand).  Prints cycles per instruction per warp for 1, 2, 4, 8
// warps per scheduler and ILP 1..12 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/micro/ffma2_probe.cu && /tmp/ffma2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int ILP, bool PACKED, bool BCAST>
__global__ void probe(float* out, long long* cyc, int iters, float s) {
  float a = s + threadIdx.x * 1e-9f;
  uint64_t a2 = BCAST ? pack(a, a) : pack(a, a + 1e-7f);
  uint64_t v2[ILP];
  float v1[2 * ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { v2[i] = pack(s * i, s * i + 1.f); v1[2 * i] = s * i; v1[2 * i + 1] = s * i + 1.f; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v2[i] = fma2(a2, v2[i], v2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) v1[i] = fma1(a, v1[i], v1[i]);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v2[i])); acc += lo + hi + v1[2 * i] + v1[2 * i + 1]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, bool PACKED, bool BCAST>
static void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps_per_sched = 1; warps_per_sched <= 8; warps_per_sched *= 2) {
    const int threads = 128 * warps_per_sched;
    if (threads > 1024) break;
    probe<ILP, PACKED, BCAST><<<148, threads>>>(out, cyc, iters, 0.5f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 8 * (PACKED ? ILP : 2 * ILP);
    printf("%-28s ILP %2d  warps/scheduler %d : %.2f cycles per instruction per warp, %.2f FMA lanes x2 /cycle/scheduler\n", name,
           PACKED ? ILP : 2 * ILP, warps_per_sched, c / instr, warps_per_sched * instr * (PACKED ? 64 : 32) / c / 32.0);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long));
  run<1, true, true>("FFMA2 bcast", out, cyc);   run<3, true, true>("FFMA2 bcast", out, cyc);
  run<6, true, true>("FFMA2 bcast", out, cyc);   run<12, true, true>("FFMA2 bcast", out, cyc);
  run<6, true, false>("FFMA2 packed-a", out, cyc);
  run<1, false, true>("FFMA", out, cyc);         run<3, false, true>("FFMA", out, cyc);
  run<6, false, true>("FFMA", out, cyc);         run<12, false, true>("FFMA", out, cyc);
  return 0;
}
