// Probe for the tcgen05 / TMEM building block planned for the dZ accumulation (profiles/README.md
// "next steps"): ONE tcgen05.mma.kind::tf32, M = 64, N = 64, K = 8, both operands MN-major in the
// no-swizzle ("interleave") canonical layout, accumulator in TMEM, read back with tcgen05.ld.
//   A[m][k] = operand tile stored as [m / 4][k][m % 4] floats (core matrix = 8 k-rows x 16 bytes),
//   B[n][k] likewise, D[m][n] = sum_k A[m][k] B[n][k].
// Inputs are small integers (exact in TF32), so the check against the CPU product is exact.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 64, N = 64, K = 8;
constexpr int SBO_BYTES = 144;   // stride between groups of 4 rows (128 B of data + 16 B pad: conflict-free stores)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  return d;                        // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}

// mn_major = 1: [m / 4][k][m % 4] (SBO between groups of 4 rows);  mn_major = 0 (K-major): core matrix = 8 rows x 16 B
// (4 k), [m / 8][k / 4][m % 8][k % 4] with LBO between the two k halves and SBO between groups of 8 rows
__global__ void probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int lbo_bytes,
                      int sbo_bytes, int mn_major) {
  __shared__ __align__(1024) unsigned char sa[4096];
  __shared__ __align__(1024) unsigned char sb[4096];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 1024; e += blockDim.x) { reinterpret_cast<float*>(sa)[e] = 0.f; reinterpret_cast<float*>(sb)[e] = 0.f; }
  __syncthreads();
  for (int e = tid; e < M * K; e += blockDim.x) {
    const int m = e / K, k = e % K;
    const int off = mn_major ? (m / 4) * sbo_bytes + k * 16 + (m % 4) * 4
                             : (m / 8) * sbo_bytes + (k / 4) * lbo_bytes + (m % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float*>(sa + off) = A[m * K + k];
    *reinterpret_cast<float*>(sb + off) = B[m * K + k];      // M == N
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    // c_format F32 (1 << 4), a/b format TF32 (2 << 7, 2 << 10), a/b MN-major (bits 15, 16), N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    const uint32_t idesc2 = mn_major ? idesc : (idesc & ~((1u << 15) | (1u << 16)));
    const uint64_t da = make_desc(smem_u32(sa), (uint32_t)lbo_bytes, (uint32_t)sbo_bytes);
    const uint64_t db = make_desc(smem_u32(sb), (uint32_t)lbo_bytes, (uint32_t)sbo_bytes);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tb), "l"(da), "l"(db), "r"(idesc2), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMA (phase 0), bounded
  {
    uint32_t done = 0;
    for (int it = 0; it < 2000000 && !done; ++it) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    if (!done && lane == 0) printf("warp %d: mbarrier wait timed out\n", warp);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // M = 64: rows 16 q .. 16 q + 15 live in TMEM lanes 32 q .. 32 q + 15 (q = warp % 4)
  uint32_t r[64];
  const uint32_t taddr = tb + ((uint32_t)(32 * (warp & 3)) << 16);
#pragma unroll
  for (int c = 0; c < 64; c += 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[c + 0]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]),
          "=r"(r[c + 7]), "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]), "=r"(r[c + 12]),
          "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
        : "r"(taddr + c));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 64; ++c) D[(32 * (warp & 3) + lane) * N + c] = __uint_as_float(r[c]);   // raw dump: [TMEM lane][column]
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tb) : "memory");
}

int main() {
  float hA[M * K], hB[N * K], hD[128 * N], ref[M * N];
  for (int i = 0; i < M * K; ++i) hA[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < N * K; ++i) hB[i] = (float)((i * 5 + 1) % 13 - 6);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float a = 0;
      for (int k = 0; k < K; ++k) a += hA[m * K + k] * hB[n * K + k];
      ref[m * N + n] = a;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const int cfgs[][3] = {{1, 0, 128}, {1, 0, 144}, {1, 128, 128}, {0, 128, 256}, {0, 256, 128}, {1, 128, 0}, {1, 144, 0}};
  for (auto& c : cfgs) {
    cudaMemset(dD, 0xff, sizeof(hD));
    probe<<<1, 128>>>(dA, dB, dD, c[1], c[2], c[0]);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    // expected placement for M = 64: row m -> TMEM lane (m % 16) + 32 (m / 16)
    int bad = 0, nz = 0, nzlane[128] = {0};
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) bad += hD[((m % 16) + 32 * (m / 16)) * N + n] != ref[m * N + n];
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < N; ++n) if (hD[l * N + n] != 0.f) { ++nz; ++nzlane[l]; }
    printf("mn_major %d lbo %3d sbo %3d: %s, mismatches %d / %d, nonzeros %d, lanes with data:", c[0], c[1], c[2],
           cudaGetErrorString(e), bad, M * N, nz);
    for (int l = 0; l < 128; ++l) if (nzlane[l]) printf(" %d", l);
    printf("\n   D[lane0][0..3] = %g %g %g %g, ref %g %g %g %g\n", hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
