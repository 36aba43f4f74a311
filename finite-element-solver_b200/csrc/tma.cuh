// sm_100a async-copy plumbing: 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on
// a shared-memory mbarrier.  Used to stage the contiguous per-tile plan streams of the assembly
// kernel into shared memory while the geometry phase runs.
#pragma once
#include <stdint.h>

namespace fes {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "FES_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra FES_DONE;\n\t"
      "bra FES_WAIT;\n\t"
      "FES_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// shared -> global bulk copy (SASS UBLKCP.G.S); both 16-byte aligned, bytes a multiple of 16.  The
// generic-proxy writes to shared memory must be fenced into the async proxy first.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Explicit shared-memory accesses by 32-bit shared address: the hot loops of the assembly kernel
// keep their base addresses in registers instead of re-deriving the shared window from a generic
// pointer at every access.
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double a) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory");
}

// A stream of `count` elements of size ES starting at element i0 of `base`, staged with one bulk
// copy from the enclosing 16-byte aligned window.  Returns the byte skew of element i0 inside
// the window; the caller reads dst + skew.  Arrays are allocated with 32 bytes of tail padding.
template <int ES>
__device__ __forceinline__ uint32_t stage_window(void* dst, const void* base, uint32_t i0, uint32_t count,
                                                 uint64_t* bar, uint32_t& bytes_out) {
  const uint64_t begin = (uint64_t)i0 * ES;
  const uint64_t aligned = begin & ~(uint64_t)15;
  const uint32_t skew = (uint32_t)(begin - aligned);
  const uint32_t bytes = (skew + count * ES + 15u) & ~15u;
  bytes_out = bytes;
  if (bytes) bulk_g2s(dst, reinterpret_cast<const char*>(base) + aligned, bytes, bar);
  return skew;
}

}  // namespace fes
