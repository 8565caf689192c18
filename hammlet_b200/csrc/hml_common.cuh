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

// Called by all threads of every CTA of a grid when the CTA's part of a step is written: true in the one CTA that
// arrives last, which can then finish the step (a scan, a final sum, an exchange) in the same launch instead of a
// follow-up kernel.  What the other CTAs wrote before the call is visible to it through L2 (__ldcg).  The counter is
// back at zero when the kernel ends.
__device__ __forceinline__ bool last_cta_arrives(unsigned* ticket) {
  __shared__ int s_last_cta;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    s_last_cta = (t == gridDim.x - 1u) ? 1 : 0;
    if (s_last_cta) *ticket = 0u;
  }
  __syncthreads();
  const bool last = s_last_cta != 0;
  if (last) __threadfence();
  return last;
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

// ---- exp(x) for x <= 0 (emission terms exp(E_s - max E), self-transition factors A_ss^(N-1)): 2^(n/64) from a
// 64-entry table (correctly rounded entries, staged in shared memory by the caller) times a degree-5 polynomial in
// the remainder |r| <= ln2/128, scaled by an integer added to the exponent field.  18 instructions, 11 of them on the
// fp64 pipe, against ~50 of the library routine; at most 1.3 ulp off (measured against expl over [-708, 0]).  Values
// below exp(-708) = 3e-308 are flushed to zero (the library returns denormals there); NaN propagates.
static __device__ const double g_exp2_64ths[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};
__device__ __forceinline__ void exp_table_load(double* s_tab) {  // 64 doubles in shared memory; barrier by the caller
  if (threadIdx.x < 64) s_tab[threadIdx.x] = g_exp2_64ths[threadIdx.x];
}
__device__ __forceinline__ double exp_nonpos(double x, const double* s_tab) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the low word of t holds round(x * 64 / ln 2)
  const double t = fma(x, 0x1.71547652b82fep+6, kMagic);
  const int n = __double2loint(t);
  const double nf = t - kMagic;
  double r = fma(nf, -0x1.62e42fee00000p-7, x);  // ln2/64, high part with 21 trailing zero bits: n * hi is exact
  r = fma(nf, -0x1.a39ef35793c76p-39, r);
  double p = fma(1.0 / 120.0, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p *= r;  // e^r - 1
  const double tj = s_tab[n & 63];
  double res = fma(tj, p, tj);
  res = __hiloint2double(__double2hiint(res) + ((n >> 6) << 20), __double2loint(res));
  res = x < -708.0 ? 0.0 : res;
  return x != x ? x : res;
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
