// GPU probe: where does a 5-D TMA box {16, 4, T, 1, 1} with SWIZZLE_128B land in shared memory?  (fused im2col
// of the patch embedding, csrc/gemm_tcgen05.cu load_a_im2col).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_box_probe tools/tma_box_probe.cu -lcuda && /tmp/tma_box_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, uint16_t* out, int x, int y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), s = (uint32_t)__cvta_generic_to_shared(smem);
  s = (s + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s),
                 "l"(reinterpret_cast<uint64_t>(&map)), "r"(b), "r"(x), "r"(y), "r"(0), "r"(0), "r"(0)
                 : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(b) : "memory");
  }
  __syncthreads();
  const uint16_t* p = reinterpret_cast<const uint16_t*>(smem + (s - (uint32_t)__cvta_generic_to_shared(smem)));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = p[i];
}

int main() {
  const int W = 48, H = 32, T = 8;
  std::vector<uint16_t> h(W * H * T);
  // value encodes (t, y, x): t*4096 + y*64 + x  (x < 64, y < 64, t < 16)
  for (int t = 0; t < T; ++t) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(t * H + y) * W + x] = (uint16_t)(t * 4096 + y * 64 + x);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 1024);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)T * H * W * 2, (cuuint64_t)T * H * W * 2};
  cuuint32_t box[5] = {16, 4, (cuuint32_t)T, 1, 1}, es[5] = {1, 1, 1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  const int x0 = 16, y0 = 8;
  probe<<<1, 128, 4096>>>(map, o, x0, y0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint16_t> got(512);
  cudaMemcpy(got.data(), o, 1024, cudaMemcpyDeviceToHost);
  int bad_dense = 0, bad_padded = 0;
  for (int t = 0; t < T; ++t) for (int py = 0; py < 4; ++py) for (int px = 0; px < 16; ++px) {
    const uint16_t want = (uint16_t)(t * 4096 + (y0 + py) * 64 + x0 + px);
    // hypothesis A (dense, address swizzle): byte = t*128 + py*32 + px*2, 16-byte chunk index ^= (t & 7)
    const int byteA = t * 128 + py * 32 + px * 2;
    const int swA = (byteA & ~0x70) | ((((byteA >> 4) & 7) ^ ((byteA >> 7) & 7)) << 4);
    if (got[swA / 2] != want) ++bad_dense;
    // hypothesis B: no swizzle at all (dense)
    if (got[byteA / 2] != want) ++bad_padded;
  }
  printf("mismatches: dense+address-swizzle %d, dense no swizzle %d (of 512)\n", bad_dense, bad_padded);
  for (int i = 0; i < 128; ++i) {
    const uint16_t v = got[i];
    printf("%s[t%d y%d x%d]", (i % 8 == 0) ? "\n" : " ", v >> 12, (v >> 6) & 63, v & 63);
  }
  printf("\n");
  return 0;
}
