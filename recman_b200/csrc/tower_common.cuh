// Helpers shared by the fused DeepFM "tower" kernels (tower_fwd.cu, tower_bwd.cu): cp.async gathers into
// SWIZZLE_128B tiles, UMMA descriptors for both operand majors, TMEM allocation of a chosen width.
#pragma once
#include "tc_common.cuh"

namespace rm {

// 16-byte async copy global -> shared; src_bytes = 0 writes zeros (ids outside their table, ragged tiles)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Shared-memory operand descriptor with explicit layout type and leading / stride byte offsets
// (cute::UMMA::SmemDescriptor bits).  layout_type: 2 = SWIZZLE_128B (K-major operands: rows of 128 B, 16-byte chunks
// XOR row & 7, 8-row atoms `sbo` = 1024 B apart), 1 = SWIZZLE_128B_BASE32B - the only swizzled layout the tensor core
// accepts for MN-major tf32 operands: rows (= K index) of 128 B holding 32 consecutive M/N elements, 32-byte chunks
// XOR row & 3, atoms of 4 rows `sbo` = 512 B apart, the 32-element M/N atoms `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc_make(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
// byte offset of 16-byte chunk q (0..7) inside row `r` of a SWIZZLE_128B_BASE32B tile
__device__ __forceinline__ uint32_t sw32b_chunk(uint32_t q, uint32_t r) {
  return ((((q >> 1) ^ (r & 3u)) << 5) | ((q & 1u) << 4));
}
// kind::tf32 instruction descriptor with operand majors (bit 15: A is MN-major, bit 16: B is MN-major)
__host__ __device__ constexpr uint32_t umma_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc_cols(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_cols(uint32_t tmem_base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// exact fp32 remainder of the truncation the tensor core applies to a tf32 operand
__device__ __forceinline__ float4 trunc_lo4(const float4 v) {
  float4 lo;
  lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  return lo;
}

constexpr uint32_t TW_NONE = 0xFFFFFFFFu;  // "no row": id outside its table / position past the end

}  // namespace rm
