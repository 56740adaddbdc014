// TMEM <-> register bandwidth and tcgen05.mma issue-rate probe (sm_100a).  The planned kernel parks 10 words per
// (cell, marker) in tensor memory between its forward and backward halves, so what matters is the SUSTAINED rate of
// tcgen05.st / tcgen05.ld per SM with several warps per quadrant and several operations in flight, per shape.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_bw_probe tools/tmem_bw_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Ops;
#define LD_LIST8(r, o) "=r"(r[o]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7])
#define ST_LIST8(r, o) "r"(r[o]), "r"(r[o+1]), "r"(r[o+2]), "r"(r[o+3]), "r"(r[o+4]), "r"(r[o+5]), "r"(r[o+6]), "r"(r[o+7])
__device__ __forceinline__ void ld8(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : LD_LIST8(r, 0) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld2(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : LD_LIST8(r, 0), LD_LIST8(r, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : LD_LIST8(r, 0), LD_LIST8(r, 8), LD_LIST8(r, 16), LD_LIST8(r, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "r"(a), ST_LIST8(r, 0) : "memory");
}
__device__ __forceinline__ void st2(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(a), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(a), ST_LIST8(r, 0), ST_LIST8(r, 8) : "memory");
}
__device__ __forceinline__ void st32(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               :: "r"(a), ST_LIST8(r, 0), ST_LIST8(r, 8), ST_LIST8(r, 16), ST_LIST8(r, 24) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode: 0 st(x8+x2)  1 st x16  2 st x32  3 ld(x8+x2) wait/1  4 ld(x8+x2) wait/4  5 ld x16 wait/2  6 ld x32 wait/2
//       7 st(x8+x2) + ld(x8+x2) wait/4 (the kernel's pattern)  8: 7 with 40 FFMA of filler per iteration (overlap test)
__global__ void __launch_bounds__(512, 1) bw(int mode, int iters, long long* clk, float* sink) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int nw = blockDim.x >> 5, per_q = nw >> 2;              // warps per quadrant
  const uint32_t cols = 512 / per_q;                            // column range owned by this warp
  const uint32_t base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + cols * (uint32_t)(warp >> 2);
  uint32_t r[32];
  for (int q = 0; q < 32; ++q) r[q] = tid + q;
  float fa = tid * 1e-3f, fb = 1.0001f, fc = 0.5f, fd = 0.25f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (uint32_t)((it * 32) % (cols - 31) & ~1u);
    const uint32_t a10 = base + (uint32_t)((it % ((cols - 10) / 10)) * 10);
    if (mode == 0) { st8(a10, r); st2(a10 + 8, r + 8); }
    if (mode == 1) st16(a, r);
    if (mode == 2) st32(a, r);
    if (mode == 3) { ld8(a10, r); ld2(a10 + 8, r + 8); wait_ld(); acc += r[0] + r[9]; }
    if (mode == 4) { ld8(a10, r + 10 * (it & 1)); ld2(a10 + 8, r + 10 * (it & 1) + 8); if ((it & 3) == 3) { wait_ld(); acc += r[0] + r[19]; } }
    if (mode == 5) { ld16(a, r + 16 * (it & 1)); if (it & 1) { wait_ld(); acc += r[0] + r[31]; } }
    if (mode == 6) { ld32(a, r); if (it & 1) { wait_ld(); acc += r[0] + r[31]; } }
    if (mode == 7 || mode == 8) {
      st8(a10, r); st2(a10 + 8, r + 8);
      const uint32_t b10 = base + (uint32_t)(((it + 5) % ((cols - 10) / 10)) * 10);
      ld8(b10, r + 16); ld2(b10 + 8, r + 24);
      if (mode == 8) {
#pragma unroll
        for (int u = 0; u < 10; ++u) { fa = fmaf(fa, fb, fc); fb = fmaf(fb, fc, fd); fc = fmaf(fc, fd, fa); fd = fmaf(fd, fa, fb); }
      }
      if ((it & 3) == 3) { wait_ld(); acc += r[16] + r[25]; }
    }
  }
  wait_ld(); wait_st();
  __syncthreads();
  if (tid == 0) clk[0] = clock64() - t0;
  if (acc == 0x1234567u || fa == 1.2345f) sink[0] = fa + fb + fc + fd + acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

// ---- MMA issue rate: n MMAs (M128, N, K8, tf32, K-major SW128, operands = zeros) from ONE thread, descriptors in registers
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(128, 1) mma_rate(int n, int N, int same_acc, long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 32768 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = smem_u32(smem);
    const long long t0 = clock64();
    for (int o = 0; o < n; ++o) {
      const uint64_t da = desc_sw128(sa + 32u * (o & 3)), db = desc_sw128(sa + 16384u + 32u * (o & 3));
      const uint32_t d = tmem_base + (same_acc ? 0u : (uint32_t)((o & 3) * N) % 256u);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(o > 3)) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    clk[0] = t1 - t0;
    clk[1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

int main() {
  long long* clk; float* sink;
  cudaMalloc(&clk, 16); cudaMalloc(&sink, 4);
  const char* names[] = {"st x8+x2", "st x16", "st x32", "ld x8+x2 wait/1", "ld x8+x2 wait/4", "ld x16 wait/2", "ld x32 wait/2",
                         "st+ld x8+x2 wait/4", "st+ld x8+x2 wait/4 + 40 FFMA"};
  const int words[] = {10, 16, 32, 10, 10, 16, 32, 10, 10};
  const int iters = 8192;
  for (int nwarps : {4, 8, 16}) {
    for (int mode = 0; mode < 9; ++mode) {
      bw<<<1, nwarps * 32>>>(mode, iters, clk, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)nwarps * 32 * 4 * words[mode] * iters;
      printf("%2d warps  %-30s %s  %7.1f cyc/iter  %7.1f B/cyc/SM %s\n", nwarps, names[mode], cudaGetErrorString(e), (double)c / iters,
             bytes / c, mode >= 7 ? "(each way)" : "");
    }
  }
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int same : {1, 0})
    for (int N : {32, 64, 128})
      for (int n : {1, 8, 32, 128}) {
        mma_rate<<<1, 128, 32768>>>(n, N, same, clk);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[2]; cudaMemcpy(c, clk, 16, cudaMemcpyDeviceToHost);
        printf("mma M128 N%-3d K8 x%-3d %s: %s issue %6lld cyc (%.1f / MMA), issue->done %6lld cyc (%.1f / MMA)\n", N, n,
               same ? "one accumulator " : "four accumulators", cudaGetErrorString(e), c[0], (double)c[0] / n, c[1], (double)c[1] / n);
      }
  return 0;
}
