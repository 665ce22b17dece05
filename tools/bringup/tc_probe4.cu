// Bring-up probe #4: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (the ".ts" form), B K-major SWIZZLE_NONE in
// shared memory.  Question: is row m of A = TMEM lane m and element k of a k-step = column a_col + k (one 32-bit column per
// tf32 element), so that a "thread = voxel" register tile can be handed to the tensor core with tcgen05.st and no
// shared-memory transposition?  K = 16 (two k-steps, the second at a_col + 8), N = 16, M = 128; A and B are small integers
// (exact in tf32), D is compared with the integer product on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bringup/tc_probe4.bin tools/bringup/tc_probe4.cu
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
__device__ __forceinline__ uint32_t mkidesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int K = 16, N = 16;

// a_col: first TMEM column of A (D sits at column 0..N-1); nthreads 128 or 256 (with 256 the second warpgroup's warps store
// the second k-step: warp w and w + 4 address the same lane quadrant)
__global__ void __launch_bounds__(256) probe(const float* A, const float* B, float* D, int a_col) {
  __shared__ __align__(128) float Bs[K * N];          // per k-step g: [n/8][k half][n%8][k%4]: off = g*N*8 + (n/8)*64 + (k/4)*32 + (n%8)*4 + k%4
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int i = tid; i < K * N; i += blockDim.x) {
    const int n = i / K, k = i % K, g = k >> 3, kk = k & 7;
    Bs[g * N * 8 + (n >> 3) * 64 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3)] = B[n * K + k];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&mbar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  // ---- A: thread = row m = (warp % 4) * 32 + lane; the k-steps are split over the warpgroups when there are two
  {
    const int m = (warp & 3) * 32 + lane;
    const int gbeg = nwarp == 8 ? (warp >> 2) : 0, gend = nwarp == 8 ? gbeg + 1 : 2;
    for (int g = gbeg; g < gend; ++g) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(A[m * K + g * 8 + j]);
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a_col + g * 8);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                   "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = mkidesc(N);
    for (int g = 0; g < 2; ++g) {
      const uint64_t db = mkdesc(s32(Bs) + (uint32_t)g * N * 8 * 4, 128u, 256u);
      const uint32_t ta = tmem + (uint32_t)(a_col + g * 8);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                   "r"(ta), "l"(db), "r"(idesc), "r"(g > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
  }
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 22)) break;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(s32(&mbar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
  srand(7);
  for (auto& v : A) v = (float)(rand() % 17 - 8);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  int rc = 0;
  for (int nt : {128, 256})
    for (int a_col : {16, 32, 40}) {
      cudaMemset(dD, 0xff, D.size() * 4);
      probe<<<1, nt>>>(dA, dB, dD, a_col);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (size_t i = 0; i < D.size(); ++i) bad += !(D[i] == R[i]);
      printf("threads %d a_col %d: %s, mismatches %d / %zu   D[0][0..3] = %g %g %g %g (want %g %g %g %g)  D[77][5] = %g (want %g)\n", nt, a_col,
             cudaGetErrorString(e), bad, D.size(), D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], D[77 * N + 5], R[77 * N + 5]);
      if (e != cudaSuccess) return 1;
      rc |= bad != 0;
    }
  printf(rc ? "TMEM-A: MISMATCH\n" : "TMEM-A: OK (row m = lane m, k = consecutive 32-bit columns)\n");
  return rc;
}
