// Microbenchmark: throughput of 1-D bulk copies (cp.async.bulk global -> shared, "TMA 1-D") as a function of the
// request size, at the occupancy of the staged flux kernel (128 threads, ~55 KB of shared memory, 4 CTAs per SM).
// Each CTA copies `total` bytes as total/S requests of S bytes (issued by the lanes of warp 0), waits on one mbarrier,
// and exits — the copy phase of flux_rk_tma_kernel without anything else.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_rate bulk_copy_rate.cu && ./bulk_copy_rate
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 4) copy_kernel(const double *src, size_t tile_doubles, unsigned req_bytes,
                                                     unsigned total_bytes, double *sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned bar = smem_addr(smem + total_bytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total_bytes) : "memory");
    __syncwarp();
    const unsigned nreq = total_bytes / req_bytes;
    const char *base = reinterpret_cast<const char *>(src + (size_t)blockIdx.x * tile_doubles);
    for (unsigned i = tid; i < nreq; i += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(smem) + i * req_bytes),
                   "l"(base + (size_t)i * req_bytes), "r"(req_bytes), "r"(bar)
                   : "memory");
  }
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(bar), "r"(0u)
                 : "memory");
  } while (!done);
  const double v = reinterpret_cast<const double *>(smem)[tid];
  if (v == 1.2345e300) sink[0] = v;
}

int main() {
  const unsigned total = 53248;  // 52 KB per tile
  const int tiles = 65536;
  const size_t tile_doubles = total / 8;
  double *src, *sink;
  cudaMalloc(&src, (size_t)tiles * total);
  cudaMalloc(&sink, 8);
  cudaMemset(src, 0, (size_t)tiles * total);
  cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total + 16);
  cudaFuncSetAttribute(copy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, copy_kernel, 128, total + 16);
  printf("{\"ctas_per_sm\": %d, \"tile_bytes\": %u, \"tiles\": %d}\n", occ, total, tiles);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const unsigned sizes[] = {512, 1024, 2048, 4096, 13312, 26624, 53248};
  for (unsigned S : sizes) {
    for (int rep = 0; rep < 2; ++rep) copy_kernel<<<tiles, 128, total + 16>>>(src, tile_doubles, S, total, sink);
    cudaEventRecord(a);
    for (int rep = 0; rep < 5; ++rep) copy_kernel<<<tiles, 128, total + 16>>>(src, tile_doubles, S, total, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    printf("{\"request_bytes\": %u, \"requests_per_tile\": %u, \"ms\": %.4f, \"GBps\": %.1f, \"err\": \"%s\"}\n", S,
           total / S, ms, (double)tiles * total / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
