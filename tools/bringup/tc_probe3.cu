// Bring-up probe #3: layout DISCOVERY for MN-major tcgen05.mma kind::tf32 operands (SWIZZLE_NONE).  Probe #2 found that the
// INTERLEAVE formula of CUTLASS's comments yields zeros for MN-major operands.  Here the operand region is filled with
// position codes (word index, split into two exactly-representable halves) and the other operand is a K-selector, so D
// reports WHICH shared-memory word the hardware reads for every (row, k): the layout rule can be read off the printout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bringup/tc_probe3.bin tools/bringup/tc_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t mkidesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// mode 0: B is the coded MN-major operand (N = 16), A = K-major selector: A[m][k] = (k == m % 8)  ->  D[m][n] = B(n, m % 8)
// mode 1: A is the coded MN-major operand, B = K-major selector: B[n][k] = (k == n % 8)           ->  D[m][n] = A(m, n % 8)
__global__ void __launch_bounds__(128) probe(int mode, int pass, uint32_t lbo, uint32_t sbo, uint32_t start, float* D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t slot;
  float* coded = (float*)sm;                  // 64 KB = 16384 words
  float* sel = (float*)(sm + 64 * 1024);      // K-major selector, 8 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384; i += 128) coded[i] = pass == 0 ? (float)(i & 1023) : (float)(i >> 10);
  for (int i = tid; i < 2048; i += 128) sel[i] = 0.f;
  __syncthreads();
  // K-major selector rows r: off(r,k) = (r/8)*256 + (r%8)*16 + (k/4)*128 + (k%4)*4   (LBO 128, SBO 256)
  for (int r = tid; r < 128; r += 128) { const int k = r % 8; sel[((r / 8) * 256 + (r % 8) * 16 + (k / 4) * 128 + (k % 4) * 4) / 4] = 1.f; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&mbar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t dc = mkdesc(s32(coded) + start, lbo, sbo), ds = mkdesc(s32(sel), 128u, 256u);
    const uint32_t idesc = mode == 0 ? mkidesc(16, 0, 1) : mkidesc(16, 1, 0);
    const uint64_t da = mode == 0 ? ds : dc, db = mode == 0 ? dc : ds;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                 "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
  }
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 22)) break;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(s32(&mbar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 16; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 16 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  float* dD; cudaMalloc(&dD, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  struct Case { int mode; uint32_t lbo, sbo, start; };
  const Case cases[] = {{0, 2048, 4096, 0}, {0, 4096, 2048, 0}, {0, 2048, 4096, 16}, {1, 2048, 4096, 0}, {1, 4096, 2048, 0}, {0, 256, 128, 0}, {1, 256, 128, 0}};
  for (const Case& c : cases) {
    std::vector<float> lo(128 * 16), hi(128 * 16);
    cudaError_t e = cudaSuccess;
    for (int pass = 0; pass < 2; ++pass) {
      cudaMemset(dD, 0xff, 128 * 16 * 4);
      probe<<<1, 128, 72 * 1024>>>(c.mode, pass, c.lbo, c.sbo, c.start, dD);
      e = cudaDeviceSynchronize();
      cudaMemcpy((pass ? hi : lo).data(), dD, 128 * 16 * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) break;
    }
    printf("mode %d (%s coded, MN-major) lbo %u sbo %u start %u: %s\n", c.mode, c.mode ? "A" : "B", c.lbo, c.sbo, c.start, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    // byte offset fetched for (row, k): mode 0: row = n (0..15), k = m % 8 -> D[k][n]; mode 1: row = m, k = n % 8 -> D[m][k]
    const int rows = c.mode == 0 ? 16 : 24;
    for (int r = 0; r < rows; ++r) {
      printf("  row %2d:", r);
      for (int k = 0; k < 8; ++k) {
        const int idx = c.mode == 0 ? k * 16 + r : r * 16 + k;
        const long word = (long)hi[idx] * 1024 + (long)lo[idx];
        printf(" %6ld", word * 4);
      }
      printf("\n");
    }
    if (c.mode == 1) {
      printf("  rows 32, 64, 96, 127 k=0: %ld %ld %ld %ld\n", ((long)hi[32 * 16] * 1024 + (long)lo[32 * 16]) * 4, ((long)hi[64 * 16] * 1024 + (long)lo[64 * 16]) * 4,
             ((long)hi[96 * 16] * 1024 + (long)lo[96 * 16]) * 4, ((long)hi[127 * 16] * 1024 + (long)lo[127 * 16]) * 4);
    }
    fflush(stdout);
  }
  return 0;
}
