// Micro-benchmark behind the cluster GRU design (gru_cluster.cu): how fast can the CTAs of one thread-block
// cluster exchange their slice of the recurrent state through distributed shared memory?
//
//   mode 0: all-to-all with cp.async.bulk shared::cta -> shared::cluster (+ complete_tx on the peer's mbarrier)
//   mode 1: all-to-all with st.shared::cluster.v4 from 128 threads + one release.cluster arrive per peer
//   mode 2: ping-pong CTA0 <-> CTA1 with one bulk copy (one-way latency)
//   mode 3: global-memory exchange as gru_wave does it (st.global + red.release.gpu + ld.acquire.gpu poll + ld)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dsmem tools/ubench_dsmem.cu
// Run:   tools/ubench_dsmem            (prints one JSON line per configuration)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred P;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait_cluster(bar, parity); ++spin)
    if (spin > (1u << 24)) { printf("ubench: mbarrier timeout block %d\n", blockIdx.x); __trap(); }
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Res { long long cycles; int errors; };

// layout of dynamic smem: recv[2][CS*bytes] | send[2][bytes] | bars[2]
__global__ void __launch_bounds__(160, 1) a2a_kernel(int mode, int iters, int bytes, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t CS;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(CS));
  const uint32_t me = cluster_rank();
  uint8_t* recv = smem;
  uint8_t* send = smem + 2 * CS * bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(send + 2 * bytes);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], mode == 1 ? CS : 1);
    mbar_init(&bars[1], mode == 1 ? CS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const long long t0 = clock64();
  int errors = 0;
  if (mode == 0 || mode == 1) {
    for (int it = 0; it < iters; ++it) {
      const int par = it & 1;
      uint8_t* sb = send + par * bytes;
      // "epilogue": 128 threads produce the slice
      if (threadIdx.x < 128) {
        if (mode == 0) {
          for (int o = threadIdx.x * 16; o < bytes; o += 128 * 16)
            *reinterpret_cast<uint4*>(sb + o) = make_uint4(it, me, o, 0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
          for (int o = threadIdx.x * 16; o < bytes; o += 128 * 16) {
            const uint4 v = make_uint4(it, me, o, 0);
            const uint32_t local = smem_u32(recv + par * CS * bytes + me * bytes + o);
            for (uint32_t p = 0; p < CS; ++p) st_cluster_v4(mapa(local, (p + me) % CS), v);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (mode == 0) {
          if (threadIdx.x == 0) mbar_expect_tx(&bars[par], CS * bytes);
          if (threadIdx.x < CS) {
            const uint32_t p = (threadIdx.x + me) % CS;
            bulk_s2c(mapa(smem_u32(recv + par * CS * bytes + me * bytes), p), smem_u32(sb), bytes,
                     mapa(smem_u32(&bars[par]), p));
          }
        } else {
          if (threadIdx.x < CS) remote_arrive(mapa(smem_u32(&bars[par]), (threadIdx.x + me) % CS));
        }
      }
      // "MMA thread": waits for every peer's slice
      if (threadIdx.x == 128) {
        mbar_wait(&bars[par], (it >> 1) & 1);
        if (it == iters - 1 || it == 0) {
          for (uint32_t p = 0; p < CS; ++p) {
            const uint4 v = *reinterpret_cast<const uint4*>(recv + par * CS * bytes + p * bytes + (bytes - 16));
            if (v.x != (uint32_t)it || v.y != p || v.z != (uint32_t)(bytes - 16)) ++errors;
          }
        }
      }
      // the epilogue of step it+1 may not start before the MMA of it+1 has all of h_it: model with a CTA barrier
      __syncthreads();
    }
  } else if (mode == 2) {
    // ping-pong between rank 0 and 1
    if (threadIdx.x == 0 && me < 2) {
      const uint32_t peer = me ^ 1;
      for (int it = 0; it < iters; ++it) {
        const int par = it & 1;
        if ((it & 1) == (int)me) {   // my turn to send
          // (nothing to wait for on the first send)
          bulk_s2c(mapa(smem_u32(recv), peer), smem_u32(send), bytes, mapa(smem_u32(&bars[0]), peer));
        } else {
          mbar_expect_tx(&bars[0], bytes);
          mbar_wait(&bars[0], (it >> 1) & 1);
        }
        (void)par;
      }
    }
  }
  const long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 128 || (mode == 2 && threadIdx.x == 0)) {
    out[blockIdx.x].cycles = t1 - t0;
    out[blockIdx.x].errors = errors;
  }
}

// mode 3: the global-memory exchange of gru_wave: every CTA writes its slice, release-adds a counter, then one
// thread polls until all CS arrived and every thread reads the whole row block back (plain loads, stand-in for TMA)
__global__ void __launch_bounds__(160, 1) gmem_kernel(int iters, int bytes, uint8_t* buf, int* counter, Res* out) {
  const int CS = gridDim.x, me = blockIdx.x;
  const long long t0 = clock64();
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
    uint8_t* b = buf + (size_t)(it & 1) * CS * bytes;
    if (threadIdx.x < 128) {
      for (int o = threadIdx.x * 16; o < bytes; o += 128 * 16)
        *reinterpret_cast<uint4*>(b + (size_t)me * bytes + o) = make_uint4(it, me, o, 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
    }
    if (threadIdx.x == 128) {
      int v;
      do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < (it + 1) * CS);
    }
    __syncthreads();
    for (int o = threadIdx.x * 16; o < CS * bytes; o += 160 * 16) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(b + o));
      acc.x += v.x; acc.y ^= v.y;
    }
  }
  const long long t1 = clock64();
  if (acc.x == 0xdeadbeefu && acc.y == 1u) out[blockIdx.x].errors = 1;   // keeps the loads alive
  if (threadIdx.x == 128) { out[blockIdx.x].cycles = t1 - t0; }
}

static void run(int mode, int CS, int bytes, int iters, int nclusters, int extra_smem) {
  Res* d_out;
  CK(cudaMalloc(&d_out, sizeof(Res) * CS * nclusters));
  CK(cudaMemset(d_out, 0, sizeof(Res) * CS * nclusters));
  const int smem = 2 * CS * bytes + 2 * bytes + 64 + extra_smem;
  CK(cudaFuncSetAttribute(a2a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(a2a_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * nclusters);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxc = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, a2a_kernel, &cfg);
  if (e != cudaSuccess) { printf("{\"mode\": %d, \"cluster\": %d, \"smem\": %d, \"error\": \"%s\"}\n", mode, CS, smem, cudaGetErrorString(e)); cudaGetLastError(); return; }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a));
  e = cudaLaunchKernelEx(&cfg, a2a_kernel, mode, iters, bytes, d_out);
  if (e != cudaSuccess) { printf("{\"mode\": %d, \"cluster\": %d, \"smem\": %d, \"max_clusters\": %d, \"launch_error\": \"%s\"}\n", mode, CS, smem, maxc, cudaGetErrorString(e)); cudaGetLastError(); return; }
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  Res* h = (Res*)malloc(sizeof(Res) * CS * nclusters);
  CK(cudaMemcpy(h, d_out, sizeof(Res) * CS * nclusters, cudaMemcpyDeviceToHost));
  long long mx = 0; int err = 0;
  for (int i = 0; i < CS * nclusters; ++i) { if (h[i].cycles > mx) mx = h[i].cycles; err += h[i].errors; }
  printf("{\"mode\": %d, \"cluster\": %d, \"nclusters\": %d, \"bytes_per_peer\": %d, \"smem\": %d, \"max_active_clusters\": %d, "
         "\"iters\": %d, \"cycles_per_iter\": %.1f, \"us_per_iter_evt\": %.3f, \"errors\": %d}\n",
         mode, CS, nclusters, bytes, smem, maxc, iters, (double)mx / iters, ms * 1e3 / iters, err);
  free(h); CK(cudaFree(d_out));
}

static void run_gmem(int CS, int bytes, int iters) {
  Res* d_out; uint8_t* buf; int* ctr;
  CK(cudaMalloc(&d_out, sizeof(Res) * CS));
  CK(cudaMalloc(&buf, (size_t)2 * CS * bytes));
  CK(cudaMalloc(&ctr, 4)); CK(cudaMemset(ctr, 0, 4));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a));
  void* args[] = {&iters, &bytes, &buf, &ctr, &d_out};
  CK(cudaLaunchCooperativeKernel((const void*)gmem_kernel, dim3(CS), dim3(160), args, 0, 0));
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  Res* h = (Res*)malloc(sizeof(Res) * CS);
  CK(cudaMemcpy(h, d_out, sizeof(Res) * CS, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < CS; ++i) if (h[i].cycles > mx) mx = h[i].cycles;
  printf("{\"mode\": 3, \"ctas\": %d, \"bytes_per_peer\": %d, \"iters\": %d, \"cycles_per_iter\": %.1f, \"us_per_iter_evt\": %.3f}\n",
         CS, bytes, iters, (double)mx / iters, ms * 1e3 / iters);
  free(h); CK(cudaFree(d_out)); CK(cudaFree(buf)); CK(cudaFree(ctr));
}

int main() {
  const int iters = 2000;
  // warm-up launch
  run(0, 8, 1024, 10, 1, 0);
  for (int CS : {4, 8, 16}) {
    for (int bytes : {512, 1024, 2048, 3072, 6144}) {
      if (2 * CS * bytes + 2 * bytes > 200 * 1024) continue;
      run(0, CS, bytes, iters, 1, 0);
      run(1, CS, bytes, iters, 1, 0);
    }
    run(2, CS, 1024, iters, 1, 0);
  }
  // co-residency of several fat clusters (what gru_cluster needs): 3 and 6 clusters of 16 CTAs with ~200 KB each
  run(0, 16, 1024, iters, 3, 160 * 1024);
  run(0, 16, 1024, iters, 6, 160 * 1024);
  run(0, 16, 3072, iters, 6, 100 * 1024);
  run(0, 8, 1024, iters, 12, 180 * 1024);
  for (int bytes : {1024, 3072}) run_gmem(16, bytes, iters);
  run_gmem(32, 512, iters);
  return 0;
}
