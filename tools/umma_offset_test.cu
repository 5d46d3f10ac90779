// Does a K-major SWIZZLE_128B UMMA operand tolerate a start address that is 128-byte (one row) past a 1024-byte
// swizzle-atom boundary?  (Needed to alias "chunk 1 of group w" onto "chunk 0 of group w+1" in the fused conv.)
// A is written with the address-based swizzle (byte ^= ((byte >> 7) & 7) << 4), the MMA reads 128 rows starting at
// row `shift`; tries descriptor base_offset = 0 and base_offset = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_offset_test tools/umma_offset_test.cu && ./umma_offset_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* A /*[160][64]*/, const __nv_bfloat16* B /*[64][64]*/,
                                          float* D /*[128][64]*/, int shift, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 160 rows x 128 B
  uint8_t* sB = smem + 160 * 128;     // 64 rows x 128 B (1024-aligned: 160*128 = 20480 = 20 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 8; i += 128) {          // 16-byte granules, address-based swizzle
    const int r = i >> 3, g = i & 7;
    const uint32_t byte = r * 128 + ((g ^ (r & 7)) << 4);
    *(uint4*)(sA + byte) = *(const uint4*)(A + r * 64 + g * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, g = i & 7;
    *(uint4*)(sB + r * 128 + ((g ^ (r & 7)) << 4)) = *(const uint4*)(B + r * 64 + g * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint64_t HI = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t a_addr = smem_u32(sA) + shift * 128, b_addr = smem_u32(sB);
    const uint64_t bo = use_base_offset ? ((uint64_t)((a_addr >> 7) & 7) << 49) : 0ull;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t ad = HI | bo | (uint64_t)(((a_addr & 0x3FFFF) >> 4) + 2 * kk);
      const uint64_t bd = HI | (uint64_t)(((b_addr & 0x3FFFF) >> 4) + 2 * kk);
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(kk ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 4; ++c) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * 64 + c * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  std::vector<float> hA(160 * 64), hB(64 * 64);
  std::vector<__nv_bfloat16> bA(160 * 64), bB(64 * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { bA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); hA[i] = __bfloat162float(bA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { bB[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); hB[i] = __bfloat162float(bB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, bA.size() * 2); cudaMalloc(&dB, bB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hD(128 * 64);
  for (int shift : {0, 1, 3, 8, 9}) {
    for (int ubo : {0, 1}) {
      cudaMemset(dD, 0, 128 * 64 * 4);
      k<<<1, 128, 40 * 1024>>>(dA, dB, dD, shift, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, ubo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int kk = 0; kk < 64; ++kk) ref += (double)hA[(m + shift) * 64 + kk] * hB[n * 64 + kk];
          const double err = fabs(ref - hD[m * 64 + n]);
          if (err > maxerr) maxerr = err;
          if (err > 1e-3) ++bad;
        }
      printf("shift %d rows, descriptor base_offset %s: max |err| = %.4g, wrong entries = %d / 8192\n", shift,
             ubo ? "= (addr>>7)&7" : "= 0", maxerr, bad);
    }
  }
  return 0;
}
