// Probe #2 for the CTA-wide tcgen05 kernel (csrc/elbo_kernel_umma.cuh): pins, on the real machine, every operand form
// that kernel relies on before it is written against them.
//   * every operand tile is "row = one cell, 128 bytes = 32 fp32, 128-byte swizzle" (16-byte chunk index XOR row % 8):
//       K-major  use (contraction over the 32 floats of a row):   f = A . Z,  c1^T = Zrep . G^T
//       MN-major use (contraction over rows = cells, 8 per MMA):  dZ += A^T . G
//   * M = 128, N = 32, kind::tf32, accumulation over several MMAs, k-steps by advancing the start address
//   * tcgen05.st / tcgen05.ld (32x32b .x8 + .x2) round trips at 10-column strides (the responsibility stash)
//   * timings: TMEM store / load throughput with 8 warps, latency of a 16-MMA / 8-MMA batch up to its mbarrier
// The device side is generic: the host lays the smem images out, lists the MMAs (descriptors as numbers) and checks
// the dumped TMEM columns, so layout hypotheses are plain host code.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/tcgen05_probe2 tools/tcgen05_probe2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int kABytes = 65536, kBBytes = 32768;

struct MmaOp {
  uint32_t a_off, b_off;        // byte offsets of the start addresses inside the A / B regions
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
  uint32_t a_layout, b_layout;  // descriptor layout type (0 none, 2 = 128-byte swizzle)
  uint32_t idesc, accumulate, tmem_col;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* mbar, uint32_t par) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 20) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(par) : "memory");
  return done != 0;
}

// out: [128 lanes][ncols] raw dump of TMEM columns [0, ncols); clk[0] = cycles from first issue to barrier completion
__global__ void __launch_bounds__(128, 1) probe_mma(const unsigned char* __restrict__ imgA, const unsigned char* __restrict__ imgB,
                                                    const MmaOp* __restrict__ ops, int nops, int ncols, float* __restrict__ out,
                                                    long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + kABytes;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < kABytes / 16; e += blockDim.x) reinterpret_cast<uint4*>(sA)[e] = reinterpret_cast<const uint4*>(imgA)[e];
  for (int e = tid; e < kBBytes / 16; e += blockDim.x) reinterpret_cast<uint4*>(sB)[e] = reinterpret_cast<const uint4*>(imgB)[e];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  // zero the columns first (so "never written" shows as 0, and accumulate = 1 chains have a defined start)
  {
    const uint32_t taddr = tb + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < 128; c += 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr + c), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int o = 0; o < nops; ++o) {
      const MmaOp op = ops[o];
      const uint64_t da = make_desc(smem_u32(sA) + op.a_off, op.a_lbo, op.a_sbo, op.a_layout);
      const uint64_t db = make_desc(smem_u32(sB) + op.b_off, op.b_lbo, op.b_sbo, op.b_layout);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tb + op.tmem_col), "l"(da), "l"(db), "r"(op.idesc), "r"(op.accumulate) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  const bool ok = mbar_wait(&mbar, 0);
  if (tid == 0) clk[0] = clock64() - t0;
  if (!ok && lane == 0) printf("warp %d: mbarrier wait timed out\n", warp);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tb + ((uint32_t)(32 * warp) << 16);
  for (int c = 0; c < ncols; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 8; ++q) out[(32 * warp + lane) * ncols + c + q] = __uint_as_float(r[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tb) : "memory");
}

// ---- stash round trip + TMEM store / load timing: 8 warps (2 per quadrant), each owns 160 columns = 16 slots x 10
__global__ void __launch_bounds__(256, 1) probe_stash(int iters, int* __restrict__ bad, long long* __restrict__ clk) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 192u + 160u * (uint32_t)(warp >> 2);
  int nbad = 0;
  // correctness: slot s holds value (warp, lane, s, c)
  for (int s = 0; s < 16; ++s) {
    uint32_t v[10];
    for (int c = 0; c < 10; ++c) v[c] = (uint32_t)(warp << 24 | lane << 16 | s << 8 | c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(base + 10 * s), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(base + 10 * s + 8), "r"(v[8]), "r"(v[9]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int s = 15; s >= 0; --s) {
    uint32_t r[10];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base + 10 * s));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[8]), "=r"(r[9]) : "r"(base + 10 * s + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 10; ++c) nbad += r[c] != (uint32_t)(warp << 24 | lane << 16 | s << 8 | c);
  }
  atomicAdd(bad, nbad);
  __syncthreads();
  // timing 1: stores only;  2: loads only (one wait per load pair);  3: st + ld interleaved like the kernel would
  uint32_t acc = 0;
  for (int mode = 0; mode < 3; ++mode) {
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + 10 * (it & 15);
      if (mode != 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(a), "r"(acc) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%1};" :: "r"(a + 8), "r"(acc) : "memory");
      }
      if (mode != 0) {
        uint32_t r[10];
        const uint32_t b = base + 10 * ((it + 8) & 15);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(b));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[8]), "=r"(r[9]) : "r"(b + 8));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 10; ++c) acc += r[c];
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    if (tid == 0) clk[mode] = clock64() - t0;
  }
  if (acc == 0x12345678u) bad[1] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------
static inline uint32_t sw128(int row, int col) { return (uint32_t)(row * 128 + (((col >> 2) ^ (row & 7)) << 4) + (col & 3) * 4); }
static inline uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static float ival(int a, int b, int salt) { return (float)(((a * 7 + b * 13 + salt * 5) % 9) - 4); }

struct Dev {
  unsigned char *A, *B; MmaOp* ops; float* out; long long* clk;
};

static int run_case(const char* name, Dev& d, const std::vector<unsigned char>& imgA, const std::vector<unsigned char>& imgB,
                    const std::vector<MmaOp>& ops, int ncols, const std::vector<float>& expect /* [128][ncols] */) {
  cudaMemcpy(d.A, imgA.data(), kABytes, cudaMemcpyHostToDevice);
  cudaMemcpy(d.B, imgB.data(), kBBytes, cudaMemcpyHostToDevice);
  cudaMemcpy(d.ops, ops.data(), ops.size() * sizeof(MmaOp), cudaMemcpyHostToDevice);
  cudaMemset(d.out, 0xff, 128 * 128 * sizeof(float));
  probe_mma<<<1, 128, kABytes + kBBytes>>>(d.A, d.B, d.ops, (int)ops.size(), ncols, d.out, d.clk);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> got(128 * ncols);
  long long clk = 0;
  cudaMemcpy(got.data(), d.out, got.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(&clk, d.clk, sizeof(clk), cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (size_t q = 0; q < got.size(); ++q)
    if (got[q] != expect[q]) { if (first < 0) first = (int)q; ++bad; }
  printf("%-44s %s  mismatches %5d / %5d  (%d MMAs, %lld cycles issue->barrier)", name, cudaGetErrorString(e), bad, (int)got.size(),
         (int)ops.size(), clk);
  if (bad) printf("  first at lane %d col %d: got %g want %g", first / ncols, first % ncols, got[first], expect[first]);
  printf("\n");
  if (e != cudaSuccess) exit(1);
  return bad;
}

int main() {
  Dev d;
  cudaMalloc(&d.A, kABytes); cudaMalloc(&d.B, kBBytes); cudaMalloc(&d.ops, 256 * sizeof(MmaOp));
  cudaMalloc(&d.out, 128 * 128 * sizeof(float)); cudaMalloc(&d.clk, 8 * sizeof(long long));
  cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kABytes + kBBytes);
  int fails = 0;

  // operand values: 4 A tiles T[t][cell][j] (t = A1hi, A1lo, N0hi, N0lo), 2 G tiles Gt[t][cell][k], Z[j][k]
  static float T[4][128][32], Gt[2][128][32], Z[32][32];
  for (int t = 0; t < 4; ++t) for (int n = 0; n < 128; ++n) for (int j = 0; j < 32; ++j) T[t][n][j] = ival(n, j, t);
  for (int t = 0; t < 2; ++t) for (int n = 0; n < 128; ++n) for (int k = 0; k < 32; ++k) Gt[t][n][k] = ival(n, k, 11 + t);
  for (int j = 0; j < 32; ++j) for (int k = 0; k < 32; ++k) Z[j][k] = (float)(((j * 5 + k * 3) % 7) < 3);
  auto put = [](std::vector<unsigned char>& img, uint32_t off, float v) { memcpy(&img[off], &v, 4); };

  // ---- test 1: f[n][k] = sum_t sum_j T[t][n][j] * W[t][k][j]   (A K-major SW128, 4 k-steps x 4 tiles; B K-major SW128)
  // B region holds 4 weight tiles W[t][k][j] = (t < 2 ? Z[j][k] : Z[j][k] - 1) at 4 KB each
  {
    std::vector<unsigned char> imgA(kABytes, 0), imgB(kBBytes, 0);
    for (int t = 0; t < 4; ++t) for (int n = 0; n < 128; ++n) for (int j = 0; j < 32; ++j) put(imgA, t * 16384 + sw128(n, j), T[t][n][j]);
    for (int t = 0; t < 4; ++t) for (int k = 0; k < 32; ++k) for (int j = 0; j < 32; ++j)
      put(imgB, t * 4096 + sw128(k, j), t < 2 ? Z[j][k] : Z[j][k] - 1.f);
    std::vector<float> expect(128 * 32);
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 32; ++k) {
      float a = 0;
      for (int t = 0; t < 4; ++t) for (int j = 0; j < 32; ++j) a += T[t][n][j] * (t < 2 ? Z[j][k] : Z[j][k] - 1.f);
      expect[n * 32 + k] = a;
    }
    for (int lbo : {16, 0}) {
      std::vector<MmaOp> ops;
      for (int t = 0; t < 4; ++t) for (int ks = 0; ks < 4; ++ks)
        ops.push_back({(uint32_t)(t * 16384 + ks * 32), (uint32_t)(t * 4096 + ks * 32), (uint32_t)lbo, 1024, (uint32_t)lbo, 1024, 2, 2,
                       idesc_tf32(128, 32, 0, 0), (uint32_t)(ops.size() > 0), 0});
      char nm[96]; snprintf(nm, sizeof nm, "f: K-major SW128 M128 N32 (lbo %d sbo 1024)", lbo);
      fails += run_case(nm, d, imgA, imgB, ops, 32, expect) != 0;
    }
  }
  // ---- test 2: c1T[m][n] = sum_k Zrep[m][k] * G[32 q + n][k] for one quarter q (A = Zrep K-major, B = G rows K-major)
  {
    std::vector<unsigned char> imgA(kABytes, 0), imgB(kBBytes, 0);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) put(imgA, sw128(m, k), Z[m & 31][k]);
    for (int t = 0; t < 2; ++t) for (int n = 0; n < 128; ++n) for (int k = 0; k < 32; ++k) put(imgB, t * 16384 + sw128(n, k), Gt[t][n][k]);
    for (int q : {0, 2, 3}) {
      std::vector<float> expect(128 * 32);
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
        float a = 0;
        for (int t = 0; t < 2; ++t) for (int k = 0; k < 32; ++k) a += Z[m & 31][k] * Gt[t][32 * q + n][k];
        expect[m * 32 + n] = a;
      }
      std::vector<MmaOp> ops;
      for (int t = 0; t < 2; ++t) for (int ks = 0; ks < 4; ++ks)
        ops.push_back({(uint32_t)(ks * 32), (uint32_t)(t * 16384 + q * 4096 + ks * 32), 16, 1024, 16, 1024, 2, 2,
                       idesc_tf32(128, 32, 0, 0), (uint32_t)(ops.size() > 0), 0});
      char nm[96]; snprintf(nm, sizeof nm, "c1T: Zrep x G rows of quarter %d (K-major)", q);
      fails += run_case(nm, d, imgA, imgB, ops, 32, expect) != 0;
    }
  }
  // ---- test 3: dZ[32 t + j][k] = sum_{cells of quarter q} T[t][cell][j] * (G0 + G1)[cell][k]   (both MN-major SW128)
  {
    std::vector<unsigned char> imgA(kABytes, 0), imgB(kBBytes, 0);
    for (int t = 0; t < 4; ++t) for (int n = 0; n < 128; ++n) for (int j = 0; j < 32; ++j) put(imgA, t * 16384 + sw128(n, j), T[t][n][j]);
    for (int t = 0; t < 2; ++t) for (int n = 0; n < 128; ++n) for (int k = 0; k < 32; ++k) put(imgB, t * 16384 + sw128(n, k), Gt[t][n][k]);
    const int q = 1;
    std::vector<float> expect(128 * 32);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) {
      float a = 0;
      for (int c = 32 * q; c < 32 * q + 32; ++c) a += T[m >> 5][c][m & 31] * (Gt[0][c][k] + Gt[1][c][k]);
      expect[m * 32 + k] = a;
    }
    struct Cand { uint32_t a_lbo, a_sbo, b_lbo, b_sbo; } cands[] = {{16384, 1024, 16384, 1024}, {1024, 16384, 1024, 16384},
                                                                      {16384, 128, 16384, 128}, {16384, 0, 0, 0}};
    for (auto& cd : cands) {
      std::vector<MmaOp> ops;
      for (int s = 4 * q; s < 4 * q + 4; ++s) for (int t = 0; t < 2; ++t)
        ops.push_back({(uint32_t)(s * 1024), (uint32_t)(t * 16384 + s * 1024), cd.a_lbo, cd.a_sbo, cd.b_lbo, cd.b_sbo, 2, 2,
                       idesc_tf32(128, 32, 1, 1), (uint32_t)(ops.size() > 0), 32});
      std::vector<float> ex2(128 * 64, 0.f);
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) ex2[m * 64 + 32 + k] = expect[m * 32 + k];
      char nm[96]; snprintf(nm, sizeof nm, "dZ: MN-major SW128 a(lbo %u sbo %u) b(%u %u)", cd.a_lbo, cd.a_sbo, cd.b_lbo, cd.b_sbo);
      fails += run_case(nm, d, imgA, imgB, ops, 64, ex2) != 0;
    }
  }
  // ---- timing of MMA batches (values irrelevant): 1, 8, 16 MMAs of M128 N32; 8 of M128 N128
  {
    std::vector<unsigned char> imgA(kABytes, 0), imgB(kBBytes, 0);
    std::vector<float> z32(128 * 32, 0.f), z128(128 * 128, 0.f);
    for (int n : {1, 8, 16, 32}) {
      std::vector<MmaOp> ops;
      for (int o = 0; o < n; ++o) ops.push_back({(uint32_t)((o & 3) * 32), (uint32_t)((o & 3) * 32), 16, 1024, 16, 1024, 2, 2, idesc_tf32(128, 32, 0, 0), (uint32_t)(o > 0), 0});
      char nm[96]; snprintf(nm, sizeof nm, "timing: %d x (M128 N32 K8)", n);
      run_case(nm, d, imgA, imgB, ops, 32, z32);
    }
    std::vector<MmaOp> ops;
    for (int o = 0; o < 8; ++o) ops.push_back({(uint32_t)((o & 3) * 32), (uint32_t)((o & 3) * 32), 16, 1024, 16, 1024, 2, 2, idesc_tf32(128, 128, 0, 0), (uint32_t)(o > 0), 0});
    run_case("timing: 8 x (M128 N128 K8)", d, imgA, imgB, ops, 128, z128);
  }
  // ---- stash
  {
    int* bad; long long* clk;
    cudaMalloc(&bad, 2 * sizeof(int)); cudaMalloc(&clk, 4 * sizeof(long long));
    cudaMemset(bad, 0, 2 * sizeof(int));
    const int iters = 4096;
    probe_stash<<<1, 256>>>(iters, bad, clk);
    cudaError_t e = cudaDeviceSynchronize();
    int hb[2]; long long hc[4];
    cudaMemcpy(hb, bad, sizeof hb, cudaMemcpyDeviceToHost); cudaMemcpy(hc, clk, sizeof hc, cudaMemcpyDeviceToHost);
    printf("stash round trip (x8 + x2 at 10-column stride, 8 warps): %s, mismatches %d\n", cudaGetErrorString(e), hb[0]);
    // per iteration every warp moves 32 lanes x 10 columns x 4 B = 1280 B; 8 warps
    printf("  st only : %.1f cycles / iteration / warp-set  -> %.1f B/cycle/SM\n", (double)hc[0] / iters, 8 * 1280.0 * iters / hc[0]);
    printf("  ld only : %.1f cycles / iteration             -> %.1f B/cycle/SM\n", (double)hc[1] / iters, 8 * 1280.0 * iters / hc[1]);
    printf("  st + ld : %.1f cycles / iteration             -> %.1f B/cycle/SM each way\n", (double)hc[2] / iters, 8 * 1280.0 * iters / hc[2]);
    fails += hb[0] != 0;
  }
  printf(fails ? "PROBE: %d FAILED\n" : "PROBE: all layout hypotheses hold\n", fails);
  return 0;
}
