// GPU probe: where do 5-D TMA boxes of a [T, H, W] uint16 volume land in shared memory?  (fused im2col of the
// patch embedding, csrc/gemm_tcgen05.cu load_a_im2col).  Findings on B200 (profiles/r02f_tma_box_probe.log):
//   SWIZZLE_128B, box {16, 4, T}: every 32-byte inner line occupies its OWN 128-byte shared-memory row (the row pitch
//     is the swizzle span, not the box width) - four pixel rows can NOT be packed into one 128-byte K-major row;
//   SWIZZLE_32B, box {16, 1, T}: T rows x 32 bytes, dense - exactly one SWIZZLE_32B UMMA atom per 8 frames.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_box_probe tools/tma_box_probe.cu -lcuda && /tmp/tma_box_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, uint16_t* out, int x, int y, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t s = (s0 + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint16_t*>(smem + (s - s0))[i] = 0xFFFF;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s),
                 "l"(reinterpret_cast<uint64_t>(&map)), "r"(b), "r"(x), "r"(y), "r"(0), "r"(0), "r"(0)
                 : "memory");
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < 2000000; ++spin)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    out[2048] = (uint16_t)ok;
  }
  __syncthreads();
  const uint16_t* p = reinterpret_cast<const uint16_t*>(smem + (s - s0));
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = p[i];
}

int main() {
  const int W = 48, H = 32, T = 8;
  std::vector<uint16_t> h(W * H * T);
  for (int t = 0; t < T; ++t) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(t * H + y) * W + x] = (uint16_t)(t * 4096 + y * 64 + x);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 4100 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)T * H * W * 2, (cuuint64_t)T * H * W * 2};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  cuInit(0);
  const int x0 = 16, y0 = 8;
  std::vector<uint16_t> got(2049);
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap map;
    cuuint32_t box[5] = {16, mode == 0 ? 4u : 1u, (cuuint32_t)T, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = 16 * (int)box[1] * T * 2;
    probe<<<1, 128, 8192>>>(map, o, x0, y0, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got.data(), o, 2049 * 2, cudaMemcpyDeviceToHost);
    printf("== %s box {16,%d,%d}: encode rc=%d kernel %s barrier completed=%d (expect_tx %d bytes)\n", mode == 0 ? "SWIZZLE_128B" : "SWIZZLE_32B",
           (int)box[1], T, (int)r, cudaGetErrorString(e), (int)got[2048], bytes);
    int written = 0;
    for (int i = 0; i < 2048; ++i) written += got[i] != 0xFFFF;
    printf("   elements written: %d\n", written);
    if (mode == 1) {
      int bad = 0, bad_ns = 0;
      for (int t = 0; t < T; ++t) for (int px = 0; px < 16; ++px) {
        const uint16_t want = (uint16_t)(t * 4096 + y0 * 64 + x0 + px);
        const int off = t * 32 + px * 2;
        const int sw = (off & ~0x10) | ((((off >> 4) & 1) ^ ((off >> 7) & 1)) << 4);   // Swizzle<1,4,3>
        bad += got[sw / 2] != want;
        bad_ns += got[off / 2] != want;
      }
      printf("   dense rows of 32 B, 16-byte chunk ^= address bit 7: %d mismatches; no swizzle: %d mismatches (of 128)\n", bad, bad_ns);
    }
    for (int i = 0; i < 160; ++i) {
      const uint16_t v = got[i];
      if (i % 8 == 0) printf("\n   %4d: ", i * 2);
      if (v == 0xFFFF) printf("[   --    ]"); else printf("[t%d y%d x%2d]", v >> 12, (v >> 6) & 63, v & 63);
    }
    printf("\n");
  }
  return 0;
}
