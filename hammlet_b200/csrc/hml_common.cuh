// hammlet_b200 — shared device/host helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace hml {

// Launch with the programmatic-stream-serialization attribute (HML_PDL=0 in the environment: plain stream order).
inline bool pdl_enabled() {
  static const bool on = !(getenv("HML_PDL") && getenv("HML_PDL")[0] == '0');
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kMaxStates = 32;        // HML_MAX_STATES
constexpr int kCellLog2 = 12;         // integral-array cell = 4096 observations
constexpr int kCell = 1 << kCellLog2;
constexpr int kTileLog2 = 12;         // maxlet / detect tile = 4096 observations
constexpr int kTile = 1 << kTileLog2;

// ---- streaming loads: read-once data bypasses L1 allocation
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- programmatic dependent launch: a kernel launched with launch_k may be scheduled while its predecessor in the
// stream is still running; its CTAs then park in pdl_enter() until that grid has completed and its writes are visible.
// pdl_enter() is the FIRST statement of every kernel launched this way (a CTA that left without waiting could let the
// grid — and the stream — get ahead of its predecessor); it also lets the successor of this kernel be scheduled.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// volatile 64-bit accessors for the decoupled look-back descriptors (value and flag travel in one word)
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- bulk asynchronous copies (TMA, 1-D) completing on an mbarrier in shared memory.  Addresses and sizes
// are multiples of 16 bytes.  One thread arms the barrier with the byte count and issues the copies; everybody
// who reads the data waits on the barrier's phase parity.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the shared-memory source of every committed bulk store has been read
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double shfl_double(double v, int src, unsigned mask = 0xffffffffu) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(mask, lo, src);
  hi = __shfl_sync(mask, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_double(double v, int delta, unsigned mask = 0xffffffffu) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(mask, lo, delta);
  hi = __shfl_up_sync(mask, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_double(double v, int m, unsigned mask = 0xffffffffu) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(mask, lo, m);
  hi = __shfl_xor_sync(mask, hi, m);
  return __hiloint2double(hi, lo);
}

// ---- exact power-of-two scaling of positive doubles (no rounding): used to keep products of
// forward operators in range without touching their mantissas.
__device__ __forceinline__ int exponent_of(double m) {  // m > 0, normal
  return ((__double2hiint(m) >> 20) & 0x7ff) - 1023;
}
__device__ __forceinline__ double pow2i(int e) {  // 2^e, e in [-1022, 1023]; below -> 0
  if (e < -1022) return 0.0;
  if (e > 1023) e = 1023;
  return __hiloint2double((e + 1023) << 20, 0);
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: uniforms are a pure function of
// (seed, sweep, block) and therefore independent of launch geometry.
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ static inline void run(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(M0, c[0], h0, l0);
      mulhilo(M1, c[2], h1, l1);
      uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += W0; k1 += W1;
    }
  }
  // 53-bit uniform in [0,1) for (seed, sweep, stream, index)
  __host__ __device__ static inline double uniform(uint64_t seed, uint64_t sweep, uint32_t stream, uint64_t index) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)sweep, (uint32_t)(sweep >> 32) ^ (stream << 24)};
    run(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = ((uint64_t)c[0] << 32) | c[1];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
  }
};

}  // namespace hml
