// microbench.cu — three measurements that size the persistent decode-step kernel planned in DESIGN.md §8 (not part of
// libvgslm.so; built by tools/microbench/Makefile into tools/microbench/libmb.so and driven by tools/microbench/run.py):
//   mb_grid_barrier      latency of a device-wide barrier between co-resident CTAs (one per SM), three protocols
//   mb_atomic_contention cost of every CTA reducing a [rows x 1024] fp32 tile into the SAME global addresses
//   mb_stream            HBM → shared-memory streaming rate of n CTAs with 1-D bulk TMA copies (cp.async.bulk)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cg = cooperative_groups;

#define MB_CUDA(x)                                                                 \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));              \
      return -1;                                                                   \
    }                                                                              \
  } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// variant 0: one monotonic counter, everybody polls it.
// variant 1: arrivals on one counter (atom returning the old value), the LAST arriver publishes the epoch in a separate
//            flag word that the others poll (the polled line is written once per barrier).
// variant 2: cooperative_groups grid.sync() (the runtime's implementation), for reference.
template <int VARIANT>
__global__ void grid_barrier_kernel(int iters, unsigned* counter, unsigned* flag, long long* cycles, float* sink) {
  cg::grid_group grid = cg::this_grid();
  const unsigned G = gridDim.x;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    acc += (float)i * 1e-9f;                       // stand-in for the phase's work
    if (VARIANT == 2) {
      grid.sync();
    } else {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (VARIANT == 0) {
          red_release_add(counter, 1u);
          const unsigned target = (unsigned)(i + 1) * G;
          while (ld_acquire(counter) < target) {
          }
        } else {
          __threadfence();
          const unsigned old = atomicAdd(counter, 1u);
          if (old == (unsigned)(i + 1) * G - 1u) {
            st_release(flag, (unsigned)(i + 1));
          } else {
            while (ld_acquire(flag) < (unsigned)(i + 1)) {
            }
          }
        }
      }
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.f) sink[0] = acc;
}

extern "C" int mb_grid_barrier(int variant, int iters, int threads, unsigned* scratch /* 2 words, zeroed */,
                               long long* cycles /* one per SM */, float* ms_out) {
  int dev = 0, sms = 0;
  MB_CUDA(cudaGetDevice(&dev));
  MB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MB_CUDA(cudaMemset(scratch, 0, 2 * sizeof(unsigned)));
  unsigned* counter = scratch;
  unsigned* flag = scratch + 1;
  float* sink = (float*)(scratch + 1);
  void* args[] = {&iters, &counter, &flag, &cycles, &sink};
  const void* fn = variant == 0   ? (const void*)grid_barrier_kernel<0>
                   : variant == 1 ? (const void*)grid_barrier_kernel<1>
                                  : (const void*)grid_barrier_kernel<2>;
  cudaEvent_t e0, e1;
  MB_CUDA(cudaEventCreate(&e0));
  MB_CUDA(cudaEventCreate(&e1));
  MB_CUDA(cudaEventRecord(e0));
  MB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(threads), args, 0, 0));
  MB_CUDA(cudaEventRecord(e1));
  MB_CUDA(cudaEventSynchronize(e1));
  MB_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
  return sms;
}

// every CTA adds its own [rows x 1024] fp32 tile into the same global tile: 256 threads x float4 per row
__global__ void atomic_contention_kernel(float* __restrict__ dst, int rows, int rounds, long long* cycles) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r)
    for (int row = 0; row < rows; ++row) atomicAdd(reinterpret_cast<float4*>(dst + row * 1024) + threadIdx.x, v);
  __threadfence();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

extern "C" int mb_atomic_contention(int ctas, int rows, int rounds, float* dst, long long* cycles, float* ms_out) {
  MB_CUDA(cudaMemset(dst, 0, (size_t)rows * 1024 * sizeof(float)));
  cudaEvent_t e0, e1;
  MB_CUDA(cudaEventCreate(&e0));
  MB_CUDA(cudaEventCreate(&e1));
  MB_CUDA(cudaEventRecord(e0));
  atomic_contention_kernel<<<ctas, 256>>>(dst, rows, rounds, cycles);
  MB_CUDA(cudaEventRecord(e1));
  MB_CUDA(cudaEventSynchronize(e1));
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
  return 0;
}

// each CTA streams `chunks` x CHUNK bytes of its own region of `src` into a STAGES-deep shared-memory ring with
// cp.async.bulk (1-D TMA) completing on mbarriers; a consumer warp only waits and releases the stages.
constexpr int CHUNK = 16384;
constexpr int STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* __restrict__ src, int chunks, long long* cycles) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ __align__(8) uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* mine = src + (size_t)blockIdx.x * chunks * CHUNK;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int c = 0; c < chunks + STAGES; ++c) {
      if (c >= STAGES) {                                 // wait for chunk c - STAGES, which frees its stage
        const int s = c % STAGES;
        const uint32_t parity = ((c / STAGES) - 1) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok)
                       : "r"(smem_u32(&full[s])), "r"(parity)
                       : "memory");
      }
      if (c < chunks) {
        const int s = c % STAGES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(CHUNK)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(ring + s * CHUNK)),
            "l"(mine + (size_t)c * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s]))
            : "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

extern "C" int mb_stream(int ctas, int chunks, const uint8_t* src /* ctas*chunks*16 KB */, long long* cycles,
                         float* ms_out) {
  MB_CUDA(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
  cudaEvent_t e0, e1;
  MB_CUDA(cudaEventCreate(&e0));
  MB_CUDA(cudaEventCreate(&e1));
  MB_CUDA(cudaEventRecord(e0));
  stream_kernel<<<ctas, 64, STAGES * CHUNK>>>(src, chunks, cycles);
  MB_CUDA(cudaEventRecord(e1));
  MB_CUDA(cudaEventSynchronize(e1));
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
  return 0;
}
