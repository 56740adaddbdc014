// Pipe-throughput microbenchmarks for sm_100a design decisions:
// FFMA, FFMA2 (fma.rn.f32x2), FADD2, MUFU.EX2, SHFL, CREDUX.MAX.F32, REDUX.SUM, LDS.128 broadcast.
// Prints warp-instructions per cycle per SM (all 4 SMSPs busy, 32 warps/SM).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 8
template <int OP>
__global__ void k(float* out, long long* cyc) {
  __shared__ float4 sm[64];
  float a[UNROLL], b = threadIdx.x * 1e-3f + 1.0f, c = 0.999f;
  unsigned long long p[UNROLL];
  for (int u = 0; u < UNROLL; ++u) { a[u] = u + threadIdx.x * 0.01f; p[u] = ((unsigned long long)__float_as_uint(a[u]) << 32) | __float_as_uint(b); }
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(1, 2, 3, 4);
  __syncthreads();
  unsigned long long bb = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (OP == 0) a[u] = fmaf(a[u], c, b);
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[u]) : "l"(bb));
      if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
      if (OP == 3) a[u] = __shfl_xor_sync(0xffffffffu, a[u], 1);
      if (OP == 4) asm volatile("redux.sync.max.f32 %0, %0, 0xffffffff;" : "+f"(a[u]));
      if (OP == 5) { unsigned r = __float_as_uint(a[u]); asm volatile("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(r)); a[u] = __uint_as_float(r); }
      if (OP == 6) { float4 v = sm[(i + u) & 63]; a[u] += v.x + v.y + v.z + v.w; }
      if (OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[u]) : "l"(bb));
      if (OP == 8) a[u] = fmaxf(a[u], b) + 0.f;
      if (OP == 9) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
      if (OP == 10) {   // mma.sync m16n8k8 tf32, UNROLL independent accumulators (rate)
        float d1 = a[(u + 1) & (UNROLL - 1)], d2 = b, d3 = c;
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(a[u]), "+f"(d1), "+f"(d2), "+f"(d3)
                     : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)),
                       "r"(__float_as_uint(c)), "r"(__float_as_uint(b)));
      }
      if (OP == 11) {   // mma.sync m16n8k16 bf16
        float d1 = b, d2 = b, d3 = c;
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(a[u]), "+f"(d1), "+f"(d2), "+f"(d3)
                     : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)),
                       "r"(__float_as_uint(c)), "r"(__float_as_uint(b)));
      }
      if (OP == 12) {   // dependent chain of tf32 MMAs on ONE accumulator (latency): uses a[0] only
        float d1 = b, d2 = b, d3 = c;
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(a[0]), "+f"(d1), "+f"(d2), "+f"(d3)
                     : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)),
                       "r"(__float_as_uint(c)), "r"(__float_as_uint(b)));
      }
      if (OP == 13) {   // ldmatrix.x4 (b16 8x8 == 8 rows x 4 fp32)
        unsigned r0, r1, r2, r3;
        unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + i + u) & 31]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
        a[u] += __uint_as_float(r0 ^ r1 ^ r2 ^ r3);
      }
      if (OP == 14) { unsigned r; asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a[u])); a[u] = __uint_as_float(r) + 1e-3f; }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int u = 0; u < UNROLL; ++u) s += a[u] + __uint_as_float((unsigned)p[u]) + __uint_as_float((unsigned)(p[u] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int extra, int threads = 1024) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, threads>>>(out, cyc);
  k<OP><<<148, threads>>>(out, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  double winst = (threads / 32.0) * ITERS * UNROLL * (1 + extra);
  printf("%-28s %8.3f warp-inst/cycle/SM   (%.0f cycles)  err=%s\n", name, winst / avg, avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA", 0); run<1>("FFMA2 (fma.rn.f32x2)", 0); run<7>("FADD2 (add.rn.f32x2)", 0); run<2>("MUFU.EX2", 0);
  run<9>("MUFU.LG2", 0); run<3>("SHFL.BFLY", 0); run<4>("CREDUX.MAX.F32", 0); run<5>("REDUX.SUM.U32", 0);
  run<6>("LDS.128 bcast (+4 FADD)", 4); run<8>("FMNMX+FADD", 1);
  run<10>("mma.sync m16n8k8 tf32", 0); run<11>("mma.sync m16n8k16 bf16", 0); run<12>("mma tf32 dependent chain", 0);
  run<10>("mma tf32, 8 warps/SM", 0, 256); run<12>("mma tf32 dep chain, 4 warps/SM", 0, 128);
  run<0>("FFMA, 8 warps/SM", 0, 256); run<2>("MUFU.EX2, 8 warps/SM", 0, 256);
  run<13>("ldmatrix.x4 (+4 ops)", 0); run<14>("cvt.rna.tf32 + FADD", 1);
  return 0;
}
