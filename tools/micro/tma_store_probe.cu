// Probe: which start addresses does a TMA tensor store (cp.async.bulk.tensor.2d.global.shared::cta) accept?
// The (B, V, 3) fp32 output rows of the fused kernel are 8 mod 16 bytes apart, so the boxes of the odd bodies start
// 8 bytes off a 16-byte boundary.  Each case runs in its own process (a faulting store poisons the context):
//   tma_store_probe <elem_bytes 4|8> <start_coord> <box_inner_elems> [pitch_elems]
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_store_probe tools/micro/tma_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void store_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int words) {
  extern __shared__ __align__(128) uint32_t buf[];
  for (int i = threadIdx.x; i < words; i += blockDim.x) buf[i] = 1000u + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(&tm)),
                 "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  const int eb = argc > 1 ? atoi(argv[1]) : 4;
  const int c0 = argc > 2 ? atoi(argv[2]) : 0;
  const int box = argc > 3 ? atoi(argv[3]) : 36;
  const long pitch = argc > 4 ? atol(argv[4]) : 41340 * 4 / eb;      // elements per row (two bodies)
  const int rows = 64, box_rows = 16;
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  uint32_t* d = nullptr;
  const size_t total_words = (size_t)rows * pitch * eb / 4;
  cudaMalloc(&d, total_words * 4);
  cudaMemset(d, 0, total_words * 4);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * eb};
  cuuint32_t bx[2] = {(cuuint32_t)box, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode(&tm, eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, dims, strides, bx,
                      es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("elem %d start %d box %d: encode failed %d\n", eb, c0, box, (int)r); return 1; }
  const int words = box * box_rows * eb / 4;
  store_kernel<<<1, 128, words * 4>>>(tm, c0, 8, words);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("elem %d start %d box %d: %s\n", eb, c0, box, cudaGetErrorString(e)); return 1; }
  std::vector<uint32_t> h(total_words);
  cudaMemcpy(h.data(), d, total_words * 4, cudaMemcpyDeviceToHost);
  long bad = 0, set = 0;
  const int wpe = eb / 4;
  for (int rr = 0; rr < rows; ++rr)
    for (long w = 0; w < pitch * wpe; ++w) {
      const uint32_t v = h[(size_t)rr * pitch * wpe + w];
      uint32_t want = 0;
      if (rr >= 8 && rr < 8 + box_rows && w >= (long)c0 * wpe && w < (long)(c0 + box) * wpe)
        want = 1000u + (rr - 8) * box * wpe + (uint32_t)(w - (long)c0 * wpe);
      if (v != want) ++bad;
      if (v) ++set;
    }
  printf("elem %d start %d (byte offset %ld mod 16 = %ld) box %d: ok, %ld words written, %ld wrong\n", eb, c0, (long)c0 * eb,
         ((long)c0 * eb) % 16, box, set, bad);
  return bad ? 1 : 0;
}
