// Shared-memory fill-rate microbenchmark (B200): how fast can ONE persistent CTA per SM stream global data into a
// ring of shared-memory stages with (a) 2-D tiled TMA, (b) 1-D bulk TMA, (c) cp.async 16 B (LDGSTS),
// (d) LDG.128 -> STS.128 through registers.   Decides the A-operand load path of the conv kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench tools/membench.cu -lcuda && ./membench
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

// ---------------- (a) 2-D tiled TMA: box {64 bf16, R rows}, S slots ----------------
__global__ void __launch_bounds__(64) k_tma2d(const __grid_constant__ CUtensorMap tm, int R, int S, int boxes_per_cta, int total_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * R * 128);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {          // producer
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(empty0 + 8 * slot, ((i / S) & 1) ^ 1);
      mbar_expect(full0 + 8 * slot, R * 128);
      const int row = (int)(((long long)(blockIdx.x + (long long)i * gridDim.x) * R) % (total_rows - R));
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + (size_t)slot * R * 128)), "l"(&tm), "r"(0), "r"(row), "r"(full0 + 8 * slot) : "memory");
    }
  } else if (threadIdx.x == 32) {  // consumer
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(full0 + 8 * slot, (i / S) & 1);
      mbar_arrive(empty0 + 8 * slot);
    }
  }
}

// ---------------- (a') 3-D tiled TMA with overlapping x windows: box {64, BW, R}, like the conv A planes ----------------
__global__ void __launch_bounds__(64) k_tma3d(const __grid_constant__ CUtensorMap tm, int BW, int R, int S, int boxes_per_cta, int nx, int ny) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int slot_bytes = BW * R * 128;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * slot_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(empty0 + 8 * slot, ((i / S) & 1) ^ 1);
      mbar_expect(full0 + 8 * slot, slot_bytes);
      const long long b = blockIdx.x + (long long)i * gridDim.x;
      const int x = (int)(b % nx) * BW, y = (int)((b / nx) % ny) * 16;
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(smem + (size_t)slot * slot_bytes)), "l"(&tm), "r"(0), "r"(x), "r"(y), "r"(full0 + 8 * slot) : "memory");
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(full0 + 8 * slot, (i / S) & 1);
      mbar_arrive(empty0 + 8 * slot);
    }
  }
}

// ---------------- (b) 1-D bulk copies of `bytes` each ----------------
__global__ void __launch_bounds__(64) k_bulk(const uint8_t* src, size_t src_bytes, int bytes, int S, int per_cta) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(empty0 + 8 * slot, ((i / S) & 1) ^ 1);
      mbar_expect(full0 + 8 * slot, bytes);
      const size_t off = ((size_t)(blockIdx.x + (size_t)i * gridDim.x) * bytes) % (src_bytes - bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + (size_t)slot * bytes)), "l"(src + off), "r"(bytes), "r"(full0 + 8 * slot) : "memory");
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < per_cta; ++i) {
      const int slot = i % S;
      mbar_wait(full0 + 8 * slot, (i / S) & 1);
      mbar_arrive(empty0 + 8 * slot);
    }
  }
}

// ---------------- (c) cp.async 16 B, THREADS threads, stage = 16 KB, LAG groups in flight ----------------
template <int THREADS, int LAG>
__global__ void __launch_bounds__(THREADS) k_ldgsts(const uint8_t* src, size_t src_bytes, int per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int STAGE = 16384, S = LAG + 1, PER = STAGE / 16 / THREADS;
  for (int i = 0; i < per_cta; ++i) {
    const size_t off = ((size_t)(blockIdx.x + (size_t)i * gridDim.x) * STAGE) % (src_bytes - STAGE);
    const uint32_t dst = smem_u32(smem + (i % S) * STAGE);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int g = threadIdx.x + j * THREADS;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + g * 16), "l"(src + off + (size_t)g * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------- (c') cp.async 16 B in the conv-gather pattern: 8 lanes per 128-byte row, rows ROW_STRIDE apart,
// zero-fill form, swizzled destination, fence.proxy.async + mbarrier arrive per stage ----------------
template <int LAG, bool ZFILL, bool FENCE>
__global__ void __launch_bounds__(128) k_gather(const uint8_t* src, size_t src_bytes, int row_stride, int per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  constexpr int STAGE = 16384, S = LAG + 1;
  const int tid = threadIdx.x, j = tid & 7, rb = tid >> 3;
  if (tid == 0) mbar_init(smem_u32(&bar), 128);
  __syncthreads();
  const uint32_t dst_off = (uint32_t)rb * 128 + (uint32_t)((j ^ (rb & 7)) << 4);
  for (int i = 0; i < per_cta; ++i) {
    const size_t tile = ((size_t)(blockIdx.x + (size_t)i * gridDim.x) * 128 * row_stride) % (src_bytes - (size_t)129 * row_stride);
    const uint32_t dst = smem_u32(smem + (i % S) * STAGE) + dst_off;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint8_t* p = src + tile + (size_t)(rb + 16 * r) * row_stride + j * 16;
      if (ZFILL) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + r * 16 * 128), "l"(p), "r"(16u) : "memory");
      else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * 16 * 128), "l"(p) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
    if (FENCE) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&bar));
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------- (d) LDG.128 -> STS.128, UNROLL loads in flight per thread ----------------
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_ldg_sts(const uint4* src, size_t src_vecs, int per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint4* s4 = (uint4*)smem;
  constexpr int CHUNK = THREADS * UNROLL;    // uint4 per iteration
  for (int i = 0; i < per_cta; ++i) {
    const size_t off = ((size_t)(blockIdx.x + (size_t)i * gridDim.x) * CHUNK) % (src_vecs - CHUNK);
    uint4 v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) v[j] = __ldg(src + off + threadIdx.x + j * THREADS);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) s4[(threadIdx.x + j * THREADS) % 4096] = v[j];
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, clock attr %d MHz\n", prop.name, sms, clk_khz / 1000);
  const size_t big = (size_t)2 << 30, small = (size_t)32 << 20;   // DRAM-streaming vs L2-resident footprints
  uint8_t* d = nullptr;
  CK(cudaMalloc(&d, big + 4096));
  CK(cudaMemset(d, 1, big));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, double bytes, float ms) {
    printf("  %-46s %8.1f GB/s  %6.1f B/clk/SM (at 1.9 GHz)\n", name, bytes / ms / 1e6, bytes / (ms * 1e-3) / sms / 1.9e9);
  };
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  for (int pass = 0; pass < 2; ++pass) {
    const size_t foot = pass ? small : big;
    printf("== footprint %zu MB (%s) ==\n", foot >> 20, pass ? "L2-resident" : "DRAM streaming");
    // (a) TMA 2-D
    for (int R : {64, 128, 176}) {
      for (int S : {2, 4, 8}) {
        CUtensorMap tm;
        const cuuint64_t gdim[2] = {64, foot / 128};
        const cuuint64_t gstr[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)R}, es[2] = {1, 1};
        if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        const size_t smem = 1024 + (size_t)S * R * 128 + 512;
        if (smem > 227 * 1024) continue;
        CK(cudaFuncSetAttribute(k_tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per = 400;
        k_tma2d<<<sms, 64, smem>>>(tm, R, S, 20, (int)(foot / 128));
        cudaEventRecord(e0);
        k_tma2d<<<sms, 64, smem>>>(tm, R, S, per, (int)(foot / 128));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char name[96]; snprintf(name, sizeof(name), "TMA 2-D box 128B x %d rows, %d slots", R, S);
        report(name, (double)sms * per * R * 128, ms);
      }
    }
    // (a') 3-D overlapping windows: x stride 32 B / 128 B, row pitch 1344 B (block_4_0 input) / 5376
    for (int xs : {32, 128}) {
      for (int S : {3, 6}) {
        const int BW = 8, R = 22, pitch = xs == 32 ? 1344 : 5376, nx = 5, ny = (int)(foot / pitch / 16) - 2;
        CUtensorMap tm;
        const cuuint64_t gdim[3] = {64, 40, foot / pitch};
        const cuuint64_t gstr[2] = {(cuuint64_t)xs, (cuuint64_t)pitch};
        const cuuint32_t box[3] = {64, (cuuint32_t)BW, (cuuint32_t)R}, es[3] = {1, 1, 1};
        if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode3d failed\n"); continue; }
        const size_t smem = 1024 + (size_t)S * BW * R * 128 + 512;
        CK(cudaFuncSetAttribute(k_tma3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per = 300;
        k_tma3d<<<sms, 64, smem>>>(tm, BW, R, S, 20, nx, ny);
        cudaEventRecord(e0);
        k_tma3d<<<sms, 64, smem>>>(tm, BW, R, S, per, nx, ny);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char name[96]; snprintf(name, sizeof(name), "TMA 3-D box 128B x %d x %d, xstride %d B, %d slots", BW, R, xs, S);
        report(name, (double)sms * per * BW * R * 128, ms);
      }
    }
    // (b) bulk 1-D
    for (int bytes : {2048, 16384}) {
      for (int S : {2, 4, 8}) {
        const size_t smem = 1024 + (size_t)S * bytes + 512;
        CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per = 16384 / bytes * 400;
        k_bulk<<<sms, 64, smem>>>(d, foot, bytes, S, 20);
        cudaEventRecord(e0);
        k_bulk<<<sms, 64, smem>>>(d, foot, bytes, S, per);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char name[96]; snprintf(name, sizeof(name), "bulk 1-D %d B, %d slots", bytes, S);
        report(name, (double)sms * per * bytes, ms);
      }
    }
    // (c) LDGSTS
    {
      const int per = 400;
      auto run = [&](auto kern, int threads, int lag, const char* nm) {
        const size_t smem = (size_t)(lag + 1) * 16384;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<sms, threads, smem>>>(d, foot, 20);
        cudaEventRecord(e0);
        kern<<<sms, threads, smem>>>(d, foot, per);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        report(nm, (double)sms * per * 16384, ms);
      };
      run(k_ldgsts<128, 2>, 128, 2, "cp.async 16B, 128 thr, 3 stages in flight");
      run(k_ldgsts<128, 5>, 128, 5, "cp.async 16B, 128 thr, 6 stages in flight");
      run(k_ldgsts<256, 5>, 256, 5, "cp.async 16B, 256 thr, 6 stages in flight");
      run(k_ldgsts<512, 5>, 512, 5, "cp.async 16B, 512 thr, 6 stages in flight");
    }
    // (c') conv-gather pattern
    {
      const int per = 400;
      auto run = [&](auto kern, int lag, int stride, const char* nm) {
        const size_t smem = (size_t)(lag + 1) * 16384;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<sms, 128, smem>>>(d, foot, stride, 20);
        cudaEventRecord(e0);
        kern<<<sms, 128, smem>>>(d, foot, stride, per);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        report(nm, (double)sms * per * 16384, ms);
      };
      run(k_gather<2, false, false>, 2, 128, "gather rows@128B (contiguous), lag 2");
      run(k_gather<2, false, false>, 2, 512, "gather rows@512B, lag 2");
      run(k_gather<2, true, false>, 2, 512, "gather rows@512B, zfill form, lag 2");
      run(k_gather<2, true, true>, 2, 512, "gather rows@512B, zfill + fence + mbar, lag 2");
      run(k_gather<5, true, true>, 5, 512, "gather rows@512B, zfill + fence + mbar, lag 5");
      run(k_gather<2, true, true>, 2, 576, "gather rows@576B (64B-misaligned rows), full, lag 2");
    }
    // (d) LDG + STS
    {
      auto run = [&](auto kern, int threads, int unroll, const char* nm) {
        const int per = 16384 * 400 / (threads * unroll * 16);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        kern<<<sms, threads, 65536>>>((const uint4*)d, foot / 16, 20);
        cudaEventRecord(e0);
        kern<<<sms, threads, 65536>>>((const uint4*)d, foot / 16, per);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        report(nm, (double)sms * per * threads * unroll * 16, ms);
      };
      run(k_ldg_sts<128, 8>, 128, 8, "LDG.128->STS.128, 128 thr x 8 in flight");
      run(k_ldg_sts<256, 8>, 256, 8, "LDG.128->STS.128, 256 thr x 8 in flight");
      run(k_ldg_sts<512, 8>, 512, 8, "LDG.128->STS.128, 512 thr x 8 in flight");
      run(k_ldg_sts<256, 16>, 256, 16, "LDG.128->STS.128, 256 thr x 16 in flight");
    }
  }
  return 0;
}
