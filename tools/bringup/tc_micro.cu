// Bring-up probe for tcgen05.mma kind::tf32 shared-memory descriptors (SWIZZLE_NONE).  One CTA, one MMA (M=128, N, K=8).
// The host supplies, for every logical element of A (128 x 8) and B (N x 8), the byte offset it should be stored at, plus
// the descriptor LBO/SBO and the instruction-descriptor major bits; the kernel reports D.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

struct Cfg { int N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int a_mn, b_mn; int a_bytes, b_bytes; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(128) probe(Cfg c, const float* A, const int* aoff, const float* B, const int* boff, float* D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t slot;
  unsigned char* sa = sm;
  unsigned char* sb = sm + c.a_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (c.a_bytes + c.b_bytes) / 4; i += 128) ((float*)sm)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < 128 * 8; i += 128) *(float*)(sa + aoff[i]) = A[i];
  for (int i = tid; i < c.N * 8; i += 128) *(float*)(sb + boff[i]) = B[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(256u) : "memory");
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = mkdesc(s32(sa), c.a_lbo, c.a_sbo), db = mkdesc(s32(sb), c.b_lbo, c.b_sbo);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
  }
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 22)) { if (tid == 0) printf("TIMEOUT\n"); break; }
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(s32(&mbar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  const int N = 16;
  std::vector<float> A(128 * 8), B(N * 8), ref(128 * N);
  srand(1);
  auto rnd = [] { return (float)((rand() % 17) - 8) / 4.f; };     // exactly representable in tf32
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 8; ++k) s += A[m * 8 + k] * B[n * 8 + k]; ref[m * N + n] = s; }
  float *dA, *dB, *dD; int *dao, *dbo;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, ref.size() * 4);
  cudaMalloc(&dao, A.size() * 4); cudaMalloc(&dbo, B.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  // candidate layouts. K-major: off(r,k) = (r/8)*G + (r%8)*16 + (k/4)*H + (k%4)*4.  MN-major: off(r,k) = (r/4)*G + (r%4)*4 + k*16.
  struct Cand { const char* name; int a_mn, b_mn; int aG, aH, bG, bH; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; };
  std::vector<Cand> cands;
  // both K-major: G = stride between 8-row groups, H = stride between the two 16-byte k halves
  cands.push_back({"A:K(lbo=H,sbo=G) B:K(lbo=H,sbo=G)", 0, 0, 256, 128, 256, 128, 128, 256, 128, 256});
  cands.push_back({"A:K(lbo=G,sbo=H) B:K(lbo=G,sbo=H)", 0, 0, 256, 128, 256, 128, 256, 128, 256, 128});
  // A MN-major (chunks of 4 rows 128 B apart), B K-major
  cands.push_back({"A:MN(sbo=G) B:K(lbo=H,sbo=G)", 1, 0, 128, 0, 256, 128, 4096, 128, 128, 256});
  cands.push_back({"A:MN(lbo=G) B:K(lbo=H,sbo=G)", 1, 0, 128, 0, 256, 128, 128, 4096, 128, 256});
  cands.push_back({"A:MN(sbo=G) B:K(lbo=G,sbo=H)", 1, 0, 128, 0, 256, 128, 4096, 128, 256, 128});
  cands.push_back({"A:MN(lbo=G) B:K(lbo=G,sbo=H)", 1, 0, 128, 0, 256, 128, 128, 4096, 256, 128});
  for (auto& cd : cands) {
    std::vector<int> ao(128 * 8), bo(N * 8);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 8; ++k)
      ao[r * 8 + k] = cd.a_mn ? (r / 4) * cd.aG + (r % 4) * 4 + k * 16 : (r / 8) * cd.aG + (r % 8) * 16 + (k / 4) * cd.aH + (k % 4) * 4;
    for (int r = 0; r < N; ++r) for (int k = 0; k < 8; ++k)
      bo[r * 8 + k] = cd.b_mn ? (r / 4) * cd.bG + (r % 4) * 4 + k * 16 : (r / 8) * cd.bG + (r % 8) * 16 + (k / 4) * cd.bH + (k % 4) * 4;
    cudaMemcpy(dao, ao.data(), ao.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbo, bo.data(), bo.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, ref.size() * 4);
    Cfg c{N, cd.a_lbo, cd.a_sbo, cd.b_lbo, cd.b_sbo, cd.a_mn, cd.b_mn, 8192, 4096};
    probe<<<1, 128, 8192 + 4096>>>(c, dA, dao, dB, dbo, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(ref.size());
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); ++i) { err += (D[i] - ref[i]) * (double)(D[i] - ref[i]); nrm += ref[i] * (double)ref[i]; nz += D[i] != 0.f; }
    printf("%-40s  %s  rel err %.3e  nonzero %d/%zu  D[0..3] = %g %g %g %g (ref %g %g %g %g)\n", cd.name, cudaGetErrorString(e),
           sqrt(err / nrm), nz, D.size(), D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
