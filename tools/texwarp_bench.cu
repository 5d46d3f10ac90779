// Prototype + micro-benchmark: homography warp with the four bilinear taps fetched by ONE texture gather (tld4) from a
// 2-D CUDA array that holds the whole batch as zero-separated cells, instead of four shared-memory byte loads from a
// staged window.  Answers, before the library kernel is touched:
//   * tld4 component order and border behaviour,
//   * tld4 throughput per SM,
//   * time of the full warp + concat + pool kernel (POOL 1/2/4) against the 573 440-B-per-pair HBM credit,
//   * cost of filling the cell array from linear frames.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o texwarp_bench tools/texwarp_bench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorString(e_), __FILE__, __LINE__, #x); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int IMG_H = 224, IMG_W = 320, IMG_PIXELS = IMG_H * IMG_W;
constexpr int CELL_W = 336, CELL_H = 240, CELL_X0 = 16, CELL_Y0 = 16, CELL_COLS = 64;
constexpr float FLOOR_MAGIC = 12582912.0f;
constexpr float FAST_EPS = 4.0e-4f;
constexpr float INV255 = 1.0f / 255.0f;
constexpr int THREADS = 256, BAND = 32;

__global__ void fill_cells_kernel(const uint8_t* __restrict__ frames, cudaSurfaceObject_t surf, int n) {
  // one thread per 16 pixels
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = IMG_PIXELS / 16;
  if (idx >= n * per) return;
  const int img = idx / per, r = idx - img * per, y = r / (IMG_W / 16), c = r - y * (IMG_W / 16);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(frames) + idx);
  const int cx = img % CELL_COLS, cy = img / CELL_COLS;
  surf2Dwrite(v, surf, (CELL_X0 + cx * CELL_W + c * 16), CELL_Y0 + cy * CELL_H + y);
}

// coordinates of the fast chain (3 FMA + RCP + 2 MUL); fraction check as in the library kernel (no exact fallback here:
// the prototype only counts how often it would be taken)
__device__ __forceinline__ void fast_coords(const float* h, const float* rowc, float fu, float& ix, float& iy) {
  const float x = fmaf(h[0], fu, rowc[0]), y = fmaf(h[3], fu, rowc[1]), z = fmaf(h[6], fu, rowc[2]);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  ix = __fmul_rn(x, r);
  iy = __fmul_rn(y, r);
}

__device__ __forceinline__ void div2_shared_rcp(float x, float y, float z, float& xn, float& yn) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z));
  const float r = __fmaf_rn(r0, __fmaf_rn(-z, r0, 1.0f), r0);
  const float qx = __fmul_rn(x, r), qy = __fmul_rn(y, r);
  xn = __fmaf_rn(r, __fmaf_rn(-z, qx, x), qx);
  yn = __fmaf_rn(r, __fmaf_rn(-z, qy, y), qy);
}
__device__ __forceinline__ void exact_coords_rcp(const float* h, float fu, float fv, float& ix, float& iy) {
  const float x = __fadd_rn(__fmaf_rn(h[1], fv, __fmul_rn(h[0], fu)), h[2]);
  const float y = __fadd_rn(__fmaf_rn(h[4], fv, __fmul_rn(h[3], fu)), h[5]);
  const float z = __fadd_rn(__fmaf_rn(h[7], fv, __fmul_rn(h[6], fu)), h[8]);
  float xn, yn;
  div2_shared_rcp(x, y, z, xn, yn);
  const float FX = (float)(2.0 / (IMG_W - 1)), FY = (float)(2.0 / (IMG_H - 1));
  const float gx = __fsub_rn(__fmul_rn(xn, FX), 1.f), gy = __fsub_rn(__fmul_rn(yn, FY), 1.f);
  ix = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (IMG_W - 1));
  iy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (IMG_H - 1));
}

template <bool CLAMP, bool REDO>
__device__ __forceinline__ float sample_tex(cudaTextureObject_t tex, float orgx, float orgy, float ix, float iy, int order,
                                            const float* h, float fu, float fv) {
  float tx = __fadd_rd(ix, FLOOR_MAGIC), ty = __fadd_rd(iy, FLOOR_MAGIC);
  float fx0 = __fsub_rn(tx, FLOOR_MAGIC), fy0 = __fsub_rn(ty, FLOOR_MAGIC);
  float w = __fsub_rn(ix, fx0), n = __fsub_rn(iy, fy0);
  if (REDO) {
    if (!(fmaxf(fabsf(w - 0.5f), fabsf(n - 0.5f)) <= 0.5f - FAST_EPS)) {
      exact_coords_rcp(h, fu, fv, ix, iy);
      tx = __fadd_rd(ix, FLOOR_MAGIC); ty = __fadd_rd(iy, FLOOR_MAGIC);
      fx0 = __fsub_rn(tx, FLOOR_MAGIC); fy0 = __fsub_rn(ty, FLOOR_MAGIC);
      w = __fsub_rn(ix, fx0); n = __fsub_rn(iy, fy0);
    }
  }
  if (CLAMP) {
    fx0 = fminf(fmaxf(fx0, -2.f), (float)IMG_W);
    fy0 = fminf(fmaxf(fy0, -2.f), (float)IMG_H);
  }
  // texel-space point shared by the four taps: the corner between (x0, y0) and (x0+1, y0+1)
  const float4 t = tex2Dgather<float4>(tex, fx0 + orgx, fy0 + orgy, 0);
  // order 0: CUDA-documented (w = (x0,y0), z = (x1,y0), x = (x0,y1), y = (x1,y1))
  float m00, m01, m10, m11;
  if (order == 0) { m00 = t.w; m01 = t.z; m10 = t.x; m11 = t.y; }
  else { m00 = t.x; m01 = t.y; m10 = t.z; m11 = t.w; }
  const float top = fmaf(w, m01 - m00, m00), bot = fmaf(w, m11 - m10, m10);
  return fmaf(n, bot - top, top);
}

__device__ __forceinline__ uint32_t pack_pair_bf16(float c0, float c1) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(c0, c1);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <int I>
__device__ __forceinline__ float byte_magic(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + I));
}

// out: [n][H/P][W/P][2] bf16, base offset by 4 bytes to mimic the haloed tensor's alignment
template <int POOL, bool CLAMP, bool REDO, int STORE, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) tex_warp_pool_kernel(const uint8_t* __restrict__ prev, cudaTextureObject_t tex,
                                                                  const float* __restrict__ Hmat,
                                                                  __nv_bfloat16* __restrict__ out, int order) {
  __shared__ float s_h[9];
  const int n = blockIdx.y, v0 = blockIdx.x * BAND;
  const uint8_t* g_prev = prev + (size_t)n * IMG_PIXELS;
  if (threadIdx.x < 9) s_h[threadIdx.x] = Hmat[n * 9 + threadIdx.x];
  __syncthreads();
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = s_h[i];
  const float orgx = (float)(CELL_X0 + (n % CELL_COLS) * CELL_W + 1), orgy = (float)(CELL_Y0 + (n / CELL_COLS) * CELL_H + 1);
  constexpr int SW = IMG_W / 8;
  constexpr float NORM = 1.0f / (float)(POOL * POOL);          // taps arrive already divided by 255
  constexpr float PNORM = INV255 / (float)(POOL * POOL);
  constexpr int DQ = THREADS / POOL, DSX = DQ % SW, DSY = DQ / SW;
  constexpr int TRIPS = (SW * BAND + THREADS - 1) / THREADS;
  const int dy = threadIdx.x % POOL, q0 = threadIdx.x / POOL;
  int sx = q0 % SW, sy = q0 / SW;
  constexpr int OW = IMG_W / POOL, OH = IMG_H / POOL;
  __nv_bfloat16* const obase = out + 10 + ((size_t)n * OH + v0 / POOL) * (OW * 2 + 8);
  constexpr int opitch = OW * 2 + 8;
#pragma unroll 1
  for (int trip = 0; trip < TRIPS; ++trip) {
    const int v = v0 + sy * POOL + dy;
    const uint2 pw2 = __ldg(reinterpret_cast<const uint2*>(g_prev + v * IMG_W + sx * 8));
    const float fv = (float)v;
    const float rowc[3] = {fmaf(h[1], fv, h[2]), fmaf(h[4], fv, h[5]), fmaf(h[7], fv, h[8])};
    __nv_bfloat16* const orow = obase + sy * opitch;
    uint32_t pk[8];
#pragma unroll
    for (int hs = 0; hs < 2; ++hs) {
      const int u0 = sx * 8 + 4 * hs;
      const uint32_t pw = hs ? pw2.y : pw2.x;
      float a1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float ix, iy;
        fast_coords(h, rowc, (float)(u0 + i), ix, iy);
        a1[i] = sample_tex<CLAMP, REDO>(tex, orgx, orgy, ix, iy, order, h, (float)(u0 + i), fv);
      }
      if (POOL == 1) {
        const float p0 = fmaf(byte_magic<0>(pw), PNORM, -8388608.0f * PNORM), p1 = fmaf(byte_magic<1>(pw), PNORM, -8388608.0f * PNORM);
        const float p2 = fmaf(byte_magic<2>(pw), PNORM, -8388608.0f * PNORM), p3 = fmaf(byte_magic<3>(pw), PNORM, -8388608.0f * PNORM);
        pk[4 * hs] = pack_pair_bf16(p0, a1[0]);
        pk[4 * hs + 1] = pack_pair_bf16(p1, a1[1]);
        pk[4 * hs + 2] = pack_pair_bf16(p2, a1[2]);
        pk[4 * hs + 3] = pack_pair_bf16(p3, a1[3]);
      } else {
        float a0[4];
        a0[0] = byte_magic<0>(pw) - 8388608.0f;
        a0[1] = byte_magic<1>(pw) - 8388608.0f;
        a0[2] = byte_magic<2>(pw) - 8388608.0f;
        a0[3] = byte_magic<3>(pw) - 8388608.0f;
        if (POOL == 2) {
          float p0 = a0[0] + a0[1], p1 = a0[2] + a0[3], w0 = a1[0] + a1[1], w1 = a1[2] + a1[3];
          p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
          w0 += __shfl_xor_sync(0xffffffffu, w0, 1); w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
          if (dy == 0) {
            uint32_t* d = reinterpret_cast<uint32_t*>(orow + u0);
            d[0] = pack_pair_bf16(p0 * PNORM, w0 * NORM);
            d[1] = pack_pair_bf16(p1 * PNORM, w1 * NORM);
          }
        } else {
          float p0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), w0 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
          p0 += __shfl_xor_sync(0xffffffffu, p0, 1); w0 += __shfl_xor_sync(0xffffffffu, w0, 1);
          p0 += __shfl_xor_sync(0xffffffffu, p0, 2); w0 += __shfl_xor_sync(0xffffffffu, w0, 2);
          if (dy == 0) *reinterpret_cast<uint32_t*>(orow + (u0 >> 1)) = pack_pair_bf16(p0 * PNORM, w0 * NORM);
        }
      }
    }
    if constexpr (POOL == 1) {
      uint32_t* d = reinterpret_cast<uint32_t*>(orow + sx * 16);
      if (STORE == 0) {
        d[0] = pk[0];
        *reinterpret_cast<uint2*>(d + 1) = make_uint2(pk[1], pk[2]);
        *reinterpret_cast<uint4*>(d + 3) = make_uint4(pk[3], pk[4], pk[5], pk[6]);
        d[7] = pk[7];
      } else {
        // 16-byte aligned chunks {left neighbour's last word, pk0..2}, {pk3..6}; the last word goes with the right neighbour
        const int lane = threadIdx.x & 31;
        uint32_t left = __shfl_up_sync(0xffffffffu, pk[7], 1);
        if (sx == 0) left = 0u;                                   // the halo in front of a row
        if (lane == 0 && sx != 0) {
          d[0] = pk[0];
          *reinterpret_cast<uint2*>(d + 1) = make_uint2(pk[1], pk[2]);
        } else {
          *reinterpret_cast<uint4*>(d - 1) = make_uint4(left, pk[0], pk[1], pk[2]);
        }
        *reinterpret_cast<uint4*>(d + 3) = make_uint4(pk[3], pk[4], pk[5], pk[6]);
        if (lane == 31 || sx == SW - 1) d[7] = pk[7];
      }
    }
    sx += DSX; sy += DSY;
    if (sx >= SW) { sx -= SW; ++sy; }
  }
}

// ===================================================== v2: the design meant for the library ==========================
constexpr int CM_IEEE = 0, CM_RCP = 1, CM_FAST = 2;
constexpr float REDO_C = 0.5f - FAST_EPS;

// NW tap (as floats), fractions; exact fallback per pixel behind ONE branch per 4 pixels
template <int CM, bool CLAMP>
__device__ __forceinline__ void tex_sample4(cudaTextureObject_t cells, const float* h, const float* rowc, float orgx, float orgy,
                                            float fu0, float fv, float* a) {
  float fx0[4], fy0[4], w[4], nn[4];
  bool valid[4];
  if (CM == CM_FAST) {
    bool redo = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float fu = fu0 + (float)i;
      const float x = fmaf(h[0], fu, rowc[0]), y = fmaf(h[3], fu, rowc[1]), z = fmaf(h[6], fu, rowc[2]);
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
      const float ix = __fmul_rn(x, r), iy = __fmul_rn(y, r);
#ifdef USE_FRND
      fx0[i] = floorf(ix);
      fy0[i] = floorf(iy);
#else
      fx0[i] = __fsub_rn(__fadd_rd(ix, FLOOR_MAGIC), FLOOR_MAGIC);
      fy0[i] = __fsub_rn(__fadd_rd(iy, FLOOR_MAGIC), FLOOR_MAGIC);
#endif
      w[i] = __fsub_rn(ix, fx0[i]);
      nn[i] = __fsub_rn(iy, fy0[i]);
      redo = redo || !(fabsf(w[i] - 0.5f) <= REDO_C) || !(fabsf(nn[i] - 0.5f) <= REDO_C);
      valid[i] = true;
    }
    if (redo) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!(fmaxf(fabsf(w[i] - 0.5f), fabsf(nn[i] - 0.5f)) <= REDO_C)) {
          float ix, iy;
          exact_coords_rcp(h, fu0 + (float)i, fv, ix, iy);
          fx0[i] = __fsub_rn(__fadd_rd(ix, FLOOR_MAGIC), FLOOR_MAGIC);
          fy0[i] = __fsub_rn(__fadd_rd(iy, FLOOR_MAGIC), FLOOR_MAGIC);
          w[i] = __fsub_rn(ix, fx0[i]);
          nn[i] = __fsub_rn(iy, fy0[i]);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float fu = fu0 + (float)i;
      float ix, iy;
      if (CM == CM_RCP) {
        exact_coords_rcp(h, fu, fv, ix, iy);
      } else {
        const float x = __fadd_rn(__fmaf_rn(h[1], fv, __fmul_rn(h[0], fu)), h[2]);
        const float y = __fadd_rn(__fmaf_rn(h[4], fv, __fmul_rn(h[3], fu)), h[5]);
        const float z = __fadd_rn(__fmaf_rn(h[7], fv, __fmul_rn(h[6], fu)), h[8]);
        const float xn = __fdiv_rn(x, z), yn = __fdiv_rn(y, z);
        const float FX = (float)(2.0 / (IMG_W - 1)), FY = (float)(2.0 / (IMG_H - 1));
        const float gx = __fsub_rn(__fmul_rn(xn, FX), 1.f), gy = __fsub_rn(__fmul_rn(yn, FY), 1.f);
        ix = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (IMG_W - 1));
        iy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (IMG_H - 1));
      }
      valid[i] = fabsf(ix) < 4.0e6f && fabsf(iy) < 4.0e6f;      // NaN / far outside: the sample is 0
      fx0[i] = __fsub_rn(__fadd_rd(ix, FLOOR_MAGIC), FLOOR_MAGIC);
      fy0[i] = __fsub_rn(__fadd_rd(iy, FLOOR_MAGIC), FLOOR_MAGIC);
      w[i] = __fsub_rn(ix, fx0[i]);
      nn[i] = __fsub_rn(iy, fy0[i]);
    }
  }
  float4 t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float cx = fx0[i], cy = fy0[i];
    if (CLAMP) {
      cx = fminf(fmaxf(cx, -2.f), (float)IMG_W);
      cy = fminf(fmaxf(cy, -2.f), (float)IMG_H);
    }
    t[i] = tex2Dgather<float4>(cells, cx + orgx, cy + orgy, 0);     // w = (x0,y0), z = (x1,y0), x = (x0,y1), y = (x1,y1)
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float top = fmaf(w[i], t[i].z - t[i].w, t[i].w), bot = fmaf(w[i], t[i].y - t[i].x, t[i].x);
    const float r = fmaf(nn[i], bot - top, top);
    a[i] = (CM == CM_FAST || valid[i]) ? r : 0.f;
  }
}

template <int POOL, int CM, bool CLAMP, int STORE>
__device__ __forceinline__ void tex_pool_band(cudaTextureObject_t cells, const float* h, const uint8_t* g_prev, __nv_bfloat16* obase,
                                              int opitch, int n, int v0) {
  constexpr int SW = IMG_W / 8;
  constexpr float NORM = 1.0f / (float)(POOL * POOL);          // taps arrive already divided by 255
  constexpr float PNORM = INV255 / (float)(POOL * POOL);
  constexpr int DQ = THREADS / POOL, DSX = DQ % SW, DSY = DQ / SW;
  constexpr int TRIPS = (SW * BAND + THREADS - 1) / THREADS;
  const float orgx = (float)(CELL_X0 + (n % CELL_COLS) * CELL_W + 1), orgy = (float)(CELL_Y0 + (n / CELL_COLS) * CELL_H + 1);
  const int dy = threadIdx.x % POOL, q0 = threadIdx.x / POOL;
  int sx = q0 % SW, sy = q0 / SW;
#pragma unroll 1
  for (int trip = 0; trip < TRIPS; ++trip) {
    const int v = v0 + sy * POOL + dy;
    const uint2 pw2 = __ldg(reinterpret_cast<const uint2*>(g_prev + v * IMG_W + sx * 8));
    const float fv = (float)v, fu0 = (float)(sx * 8);
    const float rowc[3] = {fmaf(h[1], fv, h[2]), fmaf(h[4], fv, h[5]), fmaf(h[7], fv, h[8])};
    __nv_bfloat16* const orow = obase + sy * opitch;
    float a1[8];
    tex_sample4<CM, CLAMP>(cells, h, rowc, orgx, orgy, fu0, fv, a1);
    tex_sample4<CM, CLAMP>(cells, h, rowc, orgx, orgy, fu0 + 4.f, fv, a1 + 4);
    if constexpr (POOL == 1) {
      uint32_t pk[8];
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        const uint32_t pw = hs ? pw2.y : pw2.x;
        pk[4 * hs] = pack_pair_bf16(fmaf(byte_magic<0>(pw), PNORM, -8388608.0f * PNORM), a1[4 * hs]);
        pk[4 * hs + 1] = pack_pair_bf16(fmaf(byte_magic<1>(pw), PNORM, -8388608.0f * PNORM), a1[4 * hs + 1]);
        pk[4 * hs + 2] = pack_pair_bf16(fmaf(byte_magic<2>(pw), PNORM, -8388608.0f * PNORM), a1[4 * hs + 2]);
        pk[4 * hs + 3] = pack_pair_bf16(fmaf(byte_magic<3>(pw), PNORM, -8388608.0f * PNORM), a1[4 * hs + 3]);
      }
      uint32_t* d = reinterpret_cast<uint32_t*>(orow + sx * 16);
      if (STORE == 0) {
        d[0] = pk[0];
        *reinterpret_cast<uint2*>(d + 1) = make_uint2(pk[1], pk[2]);
        *reinterpret_cast<uint4*>(d + 3) = make_uint4(pk[3], pk[4], pk[5], pk[6]);
        d[7] = pk[7];
      } else if (STORE == 2) {
        *reinterpret_cast<uint4*>(d) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(d + 4) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else if (STORE == 3) {
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                     "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
      } else {
        const int lane = threadIdx.x & 31;
        uint32_t left = __shfl_up_sync(0xffffffffu, pk[7], 1);
        if (sx == 0) left = 0u;
        if (lane == 0 && sx != 0) {
          d[0] = pk[0];
          *reinterpret_cast<uint2*>(d + 1) = make_uint2(pk[1], pk[2]);
        } else {
          *reinterpret_cast<uint4*>(d - 1) = make_uint4(left, pk[0], pk[1], pk[2]);
        }
        *reinterpret_cast<uint4*>(d + 3) = make_uint4(pk[3], pk[4], pk[5], pk[6]);
        if (lane == 31 || sx == SW - 1) d[7] = pk[7];
      }
    } else if constexpr (POOL == 2) {
      // prev: byte sums by dp4a, the two sums of a word packed in one register for the row shuffle
      uint32_t ps0 = __dp4a(pw2.x, 0x00000101u, 0u) | (__dp4a(pw2.x, 0x01010000u, 0u) << 16);
      uint32_t ps1 = __dp4a(pw2.y, 0x00000101u, 0u) | (__dp4a(pw2.y, 0x01010000u, 0u) << 16);
      ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1);
      ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1);
      float w0 = a1[0] + a1[1], w1 = a1[2] + a1[3], w2 = a1[4] + a1[5], w3 = a1[6] + a1[7];
      w0 += __shfl_xor_sync(0xffffffffu, w0, 1); w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
      w2 += __shfl_xor_sync(0xffffffffu, w2, 1); w3 += __shfl_xor_sync(0xffffffffu, w3, 1);
      if (dy == 0) {
        // 4 pooled pixels = 16 bytes, 16-byte aligned + 4 (halo): 4 + 8 + 4
        uint32_t* d = reinterpret_cast<uint32_t*>(orow + sx * 8);
        const uint32_t o0 = pack_pair_bf16((float)(ps0 & 0xffffu) * PNORM, w0 * NORM), o1 = pack_pair_bf16((float)(ps0 >> 16) * PNORM, w1 * NORM);
        const uint32_t o2 = pack_pair_bf16((float)(ps1 & 0xffffu) * PNORM, w2 * NORM), o3 = pack_pair_bf16((float)(ps1 >> 16) * PNORM, w3 * NORM);
        d[0] = o0;
        *reinterpret_cast<uint2*>(d + 1) = make_uint2(o1, o2);
        d[3] = o3;
      }
    } else {
      uint32_t ps = __dp4a(pw2.x, 0x01010101u, 0u) | (__dp4a(pw2.y, 0x01010101u, 0u) << 16);
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      float w0 = (a1[0] + a1[1]) + (a1[2] + a1[3]), w1 = (a1[4] + a1[5]) + (a1[6] + a1[7]);
      w0 += __shfl_xor_sync(0xffffffffu, w0, 1); w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
      w0 += __shfl_xor_sync(0xffffffffu, w0, 2); w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
      if (dy == 0) {
        uint32_t* d = reinterpret_cast<uint32_t*>(orow + sx * 4);
        d[0] = pack_pair_bf16((float)(ps & 0xffffu) * PNORM, w0 * NORM);
        d[1] = pack_pair_bf16((float)(ps >> 16) * PNORM, w1 * NORM);
      }
    }
    sx += DSX; sy += DSY;
    if (sx >= SW) { sx -= SW; ++sy; }
  }
}

// failure bits of one band corner (threads 0..3), OR-reduced over the CTA
__device__ __forceinline__ int corner_flags(const float* h, int v0, int v1, int c) {
  const float fu = (c & 1) ? (float)(IMG_W - 1) : 0.f, fv = (c & 2) ? (float)v1 : (float)v0;
  const float y = h[3] * fu + h[4] * fv + h[5], z = h[6] * fu + h[7] * fv + h[8], x = h[0] * fu + h[1] * fv + h[2];
  int f = 0;
  if (!(z >= 0.25f && z <= 4.0f && fabsf(x) <= 1048576.f && fabsf(y) <= 1048576.f)) f |= 1;     // shared-reciprocal division not exact
  if (!(fabsf(x) <= 262144.f && fabsf(y) <= 262144.f)) f |= 2;                                 // fast chain not within 1 px
  const float yy = y / z, xx = x / z;
  if (fabsf(xx - rintf(xx)) > 4.0f * FAST_EPS || fabsf(yy - rintf(yy)) > 4.0f * FAST_EPS) f |= 4;   // corner off the integer grid
  if (!(xx >= -14.f && xx <= (float)(IMG_W + 13) && yy >= -14.f && yy <= (float)(IMG_H + 13))) f |= 8;   // leaves the zero margin
  return f;
}

template <int POOL, int STORE>
__global__ void __launch_bounds__(THREADS) texv2_kernel(const uint8_t* __restrict__ prev, cudaTextureObject_t cells,
                                                         const float* __restrict__ Hmat, __nv_bfloat16* __restrict__ out) {
  const int n = blockIdx.y, v0 = blockIdx.x * BAND;
  float h[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = __ldg(Hmat + n * 9 + i);
  const int lane_ = threadIdx.x & 31;
  const int f = __reduce_or_sync(0xffffffffu, lane_ < 4 ? corner_flags(h, v0, v0 + BAND - 1, lane_) : 0);   // every warp for itself: no CTA barrier
  constexpr int OW = IMG_W / POOL, OH = IMG_H / POOL;
  constexpr int opitch = STORE >= 2 ? OW * 2 + 16 : OW * 2 + 8;
  __nv_bfloat16* const obase = out + (STORE >= 2 ? 16 : 10) + ((size_t)n * OH + v0 / POOL) * opitch;
  const uint8_t* g_prev = prev + (size_t)n * IMG_PIXELS;
  if (f & 1) tex_pool_band<POOL, CM_IEEE, true, STORE>(cells, h, g_prev, obase, opitch, n, v0);
  else if ((f & 2) || !(f & 4)) tex_pool_band<POOL, CM_RCP, true, STORE>(cells, h, g_prev, obase, opitch, n, v0);
  else if (f & 8) tex_pool_band<POOL, CM_FAST, true, STORE>(cells, h, g_prev, obase, opitch, n, v0);
  else tex_pool_band<POOL, CM_FAST, false, STORE>(cells, h, g_prev, obase, opitch, n, v0);
}

// reference: same coordinates, taps by plain global loads with explicit zero padding; one thread per pixel, fp32 out
__global__ void ref_warp_kernel(const uint8_t* __restrict__ curr, const float* __restrict__ Hmat, float* __restrict__ out, int n_img,
                                int* n_redo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * IMG_PIXELS) return;
  const int n = idx / IMG_PIXELS, r = idx - n * IMG_PIXELS, v = r / IMG_W, u = r - v * IMG_W;
  const float* h = Hmat + n * 9;
  const float fv = (float)v;
  const float rowc[3] = {fmaf(h[1], fv, h[2]), fmaf(h[4], fv, h[5]), fmaf(h[7], fv, h[8])};
  float ix, iy;
  fast_coords(h, rowc, (float)u, ix, iy);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float w = ix - fx0, nn = iy - fy0;
  if (!(fmaxf(fabsf(w - 0.5f), fabsf(nn - 0.5f)) <= 0.5f - FAST_EPS)) atomicAdd(n_redo, 1);
  const int x0 = (int)fx0, y0 = (int)fy0;
  float m[4];
  for (int t = 0; t < 4; ++t) {
    const int x = x0 + (t & 1), y = y0 + (t >> 1);
    m[t] = ((unsigned)x < (unsigned)IMG_W && (unsigned)y < (unsigned)IMG_H) ? (float)curr[(size_t)n * IMG_PIXELS + y * IMG_W + x] : 0.f;
  }
  const float top = fmaf(w, m[1] - m[0], m[0]), bot = fmaf(w, m[3] - m[2], m[2]);
  out[idx] = fmaf(nn, bot - top, top) * INV255;
}

// pure gather throughput: 8 gathers per thread per trip, trivially computed coordinates
// lanes on ADJACENT pixels (does the texture unit's rate depend on the locality inside a quad?)
__global__ void __launch_bounds__(THREADS) gather_adjacent_kernel(cudaTextureObject_t tex, float* __restrict__ out, float shift) {
  const int n = blockIdx.y, v0 = blockIdx.x * BAND;
  const float orgx = (float)(CELL_X0 + (n % CELL_COLS) * CELL_W + 1) + shift, orgy = (float)(CELL_Y0 + (n / CELL_COLS) * CELL_H + 1) + shift;
  float acc = 0.f;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;     // warp w: rows 4w..4w+3 of the band
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const float4 t = tex2Dgather<float4>(tex, orgx + (float)(i * 32 + lane), orgy + (float)(v0 + wrp * 4 + r), 0);
      acc += (t.x + t.y) + (t.z + t.w);
    }
  }
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * THREADS + threadIdx.x] = acc;
}
// 2x2 pixel blocks per quad of lanes
__global__ void __launch_bounds__(THREADS) gather_quad_kernel(cudaTextureObject_t tex, float* __restrict__ out, float shift) {
  const int n = blockIdx.y, v0 = blockIdx.x * BAND;
  const float orgx = (float)(CELL_X0 + (n % CELL_COLS) * CELL_W + 1) + shift, orgy = (float)(CELL_Y0 + (n / CELL_COLS) * CELL_H + 1) + shift;
  float acc = 0.f;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int qx = (lane & 1) + 2 * (lane >> 2), qy = (lane >> 1) & 1;    // 16 x 2 pixels per warp instruction
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int i = 0; i < 20; ++i) {
      const float4 t = tex2Dgather<float4>(tex, orgx + (float)(i * 16 + qx), orgy + (float)(v0 + wrp * 4 + r * 2 + qy), 0);
      acc += (t.x + t.y) + (t.z + t.w);
    }
  }
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * THREADS + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(THREADS) gather_only_kernel(cudaTextureObject_t tex, float* __restrict__ out, float shift) {
  const int n = blockIdx.y, v0 = blockIdx.x * BAND;
  const float orgx = (float)(CELL_X0 + (n % CELL_COLS) * CELL_W + 1) + shift, orgy = (float)(CELL_Y0 + (n / CELL_COLS) * CELL_H + 1) + shift;
  float acc = 0.f;
  for (int trip = 0; trip < 5; ++trip) {
    const int q = threadIdx.x + trip * THREADS, sx = q % 40, sy = q / 40;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = tex2Dgather<float4>(tex, orgx + (float)(sx * 8 + i), orgy + (float)(v0 + sy), 0);
      acc += (t.x + t.y) + (t.z + t.w);
    }
  }
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * THREADS + threadIdx.x] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  const bool prof = argc > 2;      // one launch of each final variant, for ncu
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s, %d SMs, maxTexture2DGather %d x %d, maxTexture2DLayered %d x %d x %d\n", prop.name, prop.multiProcessorCount,
         prop.maxTexture2DGather[0], prop.maxTexture2DGather[1], prop.maxTexture2DLayered[0], prop.maxTexture2DLayered[1],
         prop.maxTexture2DLayered[2]);
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));

  // layered + gather: does the runtime accept it?
  {
    cudaArray_t la = nullptr;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
    cudaError_t e = cudaMalloc3DArray(&la, &cd, make_cudaExtent(IMG_W, IMG_H, 16), cudaArrayLayered | cudaArrayTextureGather);
    printf("layered + gather array: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) cudaFreeArray(la);
    cudaGetLastError();
  }

  const int rows = (n + CELL_COLS - 1) / CELL_COLS;
  const int AW = CELL_X0 + CELL_COLS * CELL_W, AH = CELL_Y0 + rows * CELL_H;
  cudaArray_t arr;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
  CK(cudaMallocArray(&arr, &cd, AW, AH, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
  printf("cell array %d x %d (%.1f MB)\n", AW, AH, AW * (double)AH / 1e6);
  {  // zero it
    std::vector<uint8_t> z((size_t)AW * 64, 0);
    for (int y = 0; y < AH; y += 64) {
      const int hh = (AH - y) < 64 ? (AH - y) : 64;
      CK(cudaMemcpy2DToArray(arr, 0, y, z.data(), AW, AW, hh, cudaMemcpyHostToDevice));
    }
  }
  cudaResourceDesc rd{};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td{};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  cudaSurfaceObject_t surf;
  CK(cudaCreateSurfaceObject(&surf, &rd));

  // frames + homographies
  std::vector<uint8_t> hprev((size_t)n * IMG_PIXELS), hcurr((size_t)n * IMG_PIXELS);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
  for (size_t i = 0; i < hprev.size(); ++i) { hprev[i] = rnd() & 255; hcurr[i] = rnd() & 255; }
  std::vector<float> hH((size_t)n * 9);
  for (int i = 0; i < n; ++i) {
    auto u = [&]() { return (rnd() & 0xffff) / 65536.0f * 2.f - 1.f; };
    float* h = &hH[i * 9];
    h[0] = 1.f + 0.02f * u(); h[1] = 0.02f * u(); h[2] = 8.f * u();
    h[3] = 0.02f * u(); h[4] = 1.f + 0.02f * u(); h[5] = 8.f * u();
    h[6] = 2e-5f * u(); h[7] = 2e-5f * u(); h[8] = 1.f;
  }
  uint8_t *dprev, *dcurr;
  float* dH;
  CK(cudaMalloc(&dprev, hprev.size()));
  CK(cudaMalloc(&dcurr, hcurr.size()));
  CK(cudaMalloc(&dH, hH.size() * 4));
  CK(cudaMemcpy(dprev, hprev.data(), hprev.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcurr, hcurr.data(), hcurr.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dH, hH.data(), hH.size() * 4, cudaMemcpyHostToDevice));

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int fill_threads = n * (IMG_PIXELS / 16);
  for (int it = 0; it < 3; ++it) fill_cells_kernel<<<(fill_threads + 255) / 256, 256>>>(dcurr, surf, n);
  CK(cudaEventRecord(e0));
  for (int it = 0; it < 10; ++it) fill_cells_kernel<<<(fill_threads + 255) / 256, 256>>>(dcurr, surf, n);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  printf("fill_cells: %.1f us per %d frames\n", time_ms(e0, e1) * 100.f, n);

  // reference
  float* dref;
  int* dredo;
  const int nref = n < 64 ? n : 64;
  CK(cudaMalloc(&dref, (size_t)nref * IMG_PIXELS * 4));
  CK(cudaMalloc(&dredo, 4));
  CK(cudaMemset(dredo, 0, 4));
  ref_warp_kernel<<<(nref * IMG_PIXELS + 255) / 256, 256>>>(dcurr, dH, dref, nref, dredo);
  CK(cudaDeviceSynchronize());
  std::vector<float> href((size_t)nref * IMG_PIXELS);
  int hredo = 0;
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hredo, dredo, 4, cudaMemcpyDeviceToHost));
  printf("fallback fraction (fast coords within %g of an integer): %.3f %%\n", FAST_EPS, 100.0 * hredo / ((double)nref * IMG_PIXELS));

  __nv_bfloat16* dout;
  const size_t pitch1 = IMG_W * 2 + 16;
  const size_t out_el = (size_t)n * IMG_H * pitch1 + 64;
  CK(cudaMalloc(&dout, out_el * 2));
  std::vector<__nv_bfloat16> hout(out_el);
  auto check = [&](const char* what, int pool, int store = 0) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hout.data(), dout, out_el * 2, cudaMemcpyDeviceToHost));
    const int OW = IMG_W / pool, OH = IMG_H / pool;
    const size_t pitch = store >= 2 ? OW * 2 + 16 : OW * 2 + 8;
    const size_t off0 = store >= 2 ? 16 : 10;
    double maxerr = 0, maxerr_prev = 0;
    for (int i = 0; i < nref; ++i)
      for (int oy = 0; oy < OH; ++oy)
        for (int ox = 0; ox < OW; ++ox) {
          double sw = 0, sp = 0;
          for (int dy = 0; dy < pool; ++dy)
            for (int dx = 0; dx < pool; ++dx) {
              const size_t p = (size_t)i * IMG_PIXELS + (oy * pool + dy) * IMG_W + ox * pool + dx;
              sw += href[p];
              sp += hprev[p] * (double)INV255;
            }
          sw /= pool * pool; sp /= pool * pool;
          const size_t o = off0 + ((size_t)i * OH + oy) * pitch + ox * 2;
          const double e = fabs((double)__bfloat162float(hout[o + 1]) - sw), ep = fabs((double)__bfloat162float(hout[o]) - sp);
          if (e > maxerr) maxerr = e;
          if (ep > maxerr_prev) maxerr_prev = ep;
        }
    printf("%s: max |tex - ref| = %.3e (bf16 half-ulp at 1.0 = 1.95e-3), prev channel %.3e\n", what, maxerr, maxerr_prev);
  };
  dim3 grid(IMG_H / BAND, n);
  if (prof) {
    texv2_kernel<1, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
    texv2_kernel<1, 1><<<grid, THREADS>>>(dprev, tex, dH, dout);
    texv2_kernel<2, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
    texv2_kernel<4, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
    CK(cudaDeviceSynchronize());
    return 0;
  }
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<1, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 1, 4 stores", 1);
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<1, 1><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 1, shifted 16-byte stores", 1);
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<2, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 2", 2);
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<4, 0><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 4", 4);

#define RUN2(POOL, STORE)                                                                                            \
  {                                                                                                                  \
    for (int it = 0; it < 30; ++it) texv2_kernel<POOL, STORE><<<grid, THREADS>>>(dprev, tex, dH, dout);              \
    CK(cudaEventRecord(e0));                                                                                         \
    for (int it = 0; it < 20; ++it) texv2_kernel<POOL, STORE><<<grid, THREADS>>>(dprev, tex, dH, dout);              \
    CK(cudaEventRecord(e1));                                                                                         \
    CK(cudaDeviceSynchronize());                                                                                     \
    const float us = time_ms(e0, e1) * 50.f;                                                                         \
    printf("texv2<POOL=%d, store=%d>: %.1f us per %d pairs = %.2f TB/s on the 573 440-B credit\n", POOL, STORE, us, n, \
           573440.0 * n / us / 1e6);                                                                                 \
  }
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<1, 2><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 1, aligned 2 x 16-byte stores", 1, 2);
  CK(cudaMemset(dout, 0, out_el * 2));
  texv2_kernel<1, 3><<<grid, THREADS>>>(dprev, tex, dH, dout);
  check("v2 POOL 1, aligned 32-byte stores", 1, 3);
  RUN2(1, 0)
  RUN2(1, 1)
  RUN2(1, 2)
  RUN2(1, 3)
  RUN2(2, 0)
  RUN2(4, 0)
  RUN2(1, 0)
  RUN2(1, 1)

  float* dacc;
  CK(cudaMalloc(&dacc, (size_t)n * 7 * THREADS * 4));
  for (int var = 0; var < 3; ++var) {
    const float shift = 0.f;
    auto go = [&]() {
      if (var == 0) gather_only_kernel<<<dim3(7, n), THREADS>>>(tex, dacc, shift);
      if (var == 1) gather_adjacent_kernel<<<dim3(7, n), THREADS>>>(tex, dacc, shift);
      if (var == 2) gather_quad_kernel<<<dim3(7, n), THREADS>>>(tex, dacc, shift);
    };
    for (int it = 0; it < 3; ++it) go();
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 20; ++it) go();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    const float us = time_ms(e0, e1) * 50.f;
    printf("gather only (variant %d: 0 strips, 1 adjacent lanes, 2 2x2 quads; shift %.2f): %.1f us per %d x 71680 gathers = %.1f G gathers/s = %.2f per SM per clock at %d MHz nominal\n", var, shift, us, n,
           (double)n * IMG_PIXELS / us / 1e3, (double)n * IMG_PIXELS / us / 1e3 / prop.multiProcessorCount / (clk_khz / 1e6), clk_khz / 1000);
  }
  CK(cudaGetLastError());
  printf("done\n");
  return 0;
}
