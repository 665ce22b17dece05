// Bring-up probe #2 for tcgen05.mma kind::tf32 (SWIZZLE_NONE descriptors), round 2.  Answers, on a B200, the layout questions
// the implicit-GEMM convolution kernels (conv_dense_tc.cu, jlc_tc.cu) rest on, before any of them is debugged:
//   (1) K-major A whose descriptor start address is only 16-byte aligned (the "shifted descriptor" tap addressing),
//   (2) MN-major B / MN-major A operands (idesc bits 16 / 15) in the INTERLEAVE layout: 8 k-rows x 16 B core matrices,
//       MN groups SBO apart, also with 16-byte-granular start addresses (weight-gradient kernel),
//   (3) issue rate of back-to-back MMAs at N = 16 / 32 / 64 / 128 / 256 (cycles per MMA, one issuing thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bringup/tc_probe2.bin tools/bringup/tc_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

struct Cfg { int N; uint32_t a_start, a_lbo, a_sbo, b_start, b_lbo, b_sbo; int a_mn, b_mn; int a_bytes, b_bytes; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t mkidesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
               "l"(da), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 22)) return false;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
  return true;
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
    mma(tmem, mkdesc(s32(sa) + c.a_start, c.a_lbo, c.a_sbo), mkdesc(s32(sb) + c.b_start, c.b_lbo, c.b_sbo), mkidesc(c.N, c.a_mn, c.b_mn), 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
  }
  if (!wait_bar(s32(&mbar), 0u) && tid == 0) printf("TIMEOUT\n");
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

// issue-rate probe: `reps` MMAs (M = 128, N, K = 8) back to back from one thread, operands K-major; a_stride bytes between the
// A operands of consecutive MMAs (0 = same tile, 16 = the shifted-tap pattern, 4096 = distinct tiles)
__global__ void __launch_bounds__(128) rate(int N, int reps, int a_stride, int b_mn, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) ((float*)sm)[i] = 1.0f;
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
    const uint32_t idesc = mkidesc(N, 0, b_mn);
    const uint32_t ab = s32(sm), bb = s32(sm) + 96 * 1024;
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint32_t ao = (uint32_t)((i & 15) * a_stride);
      mma(tmem, mkdesc(ab + ao, 8192u, 128u), b_mn ? mkdesc(bb, 128u, 512u) : mkdesc(bb, 128u, 256u), idesc, i ? 1u : 0u);
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
    wait_bar(s32(&mbar), 0u);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
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
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  // data placement -- K-major:  off(r,k) = start + (r/8)*grp + (r%8)*16 + (k/4)*kh + (k%4)*4
  //                   MN-major: off(r,k) = start + (r/4)*grp + (r%4)*4 + (k%8)*16            (one k-step: K = 8)
  // descriptor fields (lbo, sbo) are given separately so both assignments of (grp, kh) to (SBO, LBO) can be tried
  struct Op { int mn; uint32_t start, grp, kh, lbo, sbo; };
  struct Cand { const char* name; Op a, b; };
  const Op AK{0, 0, 128, 8192, 8192, 128}, BK{0, 0, 256, 128, 128, 256};
  std::vector<Cand> cands = {
      {"A:K aligned | B:K", AK, BK},
      {"A:K start+16 | B:K", {0, 16, 128, 8192, 8192, 128}, BK},
      {"A:K start+48 kh 8208 | B:K", {0, 48, 128, 8208, 8208, 128}, BK},
      {"A:K start+1040 | B:K start+16", {0, 1040, 128, 8192, 8192, 128}, {0, 16, 256, 128, 128, 256}},
      {"A:K | B:MN grp 4096 as SBO", AK, {1, 0, 4096, 0, 128, 4096}},
      {"A:K | B:MN grp 4096 as LBO", AK, {1, 0, 4096, 0, 4096, 128}},
      {"A:K | B:MN start+16 grp 4112 as SBO", AK, {1, 16, 4112, 0, 128, 4112}},
      {"A:K | B:MN start+16 grp 4112 as LBO", AK, {1, 16, 4112, 0, 4112, 128}},
      {"A:K start+16 | B:MN start+272 grp 4112 as SBO", {0, 16, 128, 8192, 8192, 128}, {1, 272, 4112, 0, 128, 4112}},
      {"A:MN grp 128 as SBO | B:K", {1, 0, 128, 0, 8192, 128}, BK},
      {"A:MN grp 128 as LBO | B:K", {1, 0, 128, 0, 128, 8192}, BK},
      {"A:MN start+16 grp 144 as SBO | B:K", {1, 16, 144, 0, 8192, 144}, BK},
      {"A:MN start+16 grp 144 as LBO | B:K", {1, 16, 144, 0, 144, 8192}, BK},
  };
  auto off = [](const Op& o, int r, int k) {
    return (int)(o.start + (o.mn ? (r / 4) * o.grp + (r % 4) * 4 + k * 16 : (r / 8) * o.grp + (r % 8) * 16 + (k / 4) * o.kh + (k % 4) * 4));
  };
  for (auto& cd : cands) {
    std::vector<int> ao(128 * 8), bo(N * 8);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 8; ++k) ao[r * 8 + k] = off(cd.a, r, k);
    for (int r = 0; r < N; ++r) for (int k = 0; k < 8; ++k) bo[r * 8 + k] = off(cd.b, r, k);
    cudaMemcpy(dao, ao.data(), ao.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbo, bo.data(), bo.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, ref.size() * 4);
    Cfg c{N, cd.a.start, cd.a.lbo, cd.a.sbo, cd.b.start, cd.b.lbo, cd.b.sbo, cd.a.mn, cd.b.mn, 48 * 1024, 32 * 1024};
    probe<<<1, 128, 80 * 1024>>>(c, dA, dao, dB, dbo, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(ref.size());
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); ++i) { err += (D[i] - ref[i]) * (double)(D[i] - ref[i]); nrm += ref[i] * (double)ref[i]; nz += D[i] != 0.f; }
    printf("%-58s %s  rel err %.3e  nonzero %d/%zu\n", cd.name, cudaGetErrorString(e), sqrt(err / nrm), nz, D.size());
    fflush(stdout);
    if (e != cudaSuccess) return 1;
  }
  long long* dout; cudaMalloc(&dout, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int bmn = 0; bmn < 2; ++bmn)
    for (int a_stride : {0, 16, 4096})
      for (int n : Ns) {
        const int reps = 2048;
        rate<<<1, 128, 160 * 1024>>>(n, reps, a_stride, bmn, dout);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        printf("rate N=%3d B:%s a_stride %4d: %s issue %.1f cyc/mma, complete %.1f cyc/mma (%.0f MAC/cyc)\n", n, bmn ? "MN" : "K ", a_stride,
               cudaGetErrorString(e), (double)h[0] / reps, (double)h[1] / reps, 128.0 * n * 8 * reps / (double)h[1]);
        fflush(stdout);
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
