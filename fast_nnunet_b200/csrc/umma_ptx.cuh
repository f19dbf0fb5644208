// Inline-PTX building blocks shared by the tcgen05 kernels (sm_100a): mbarrier, bulk async copy (TMA
// engine), UMMA shared-memory descriptors, tcgen05.mma / commit / ld, fences, named barriers.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace fnnu {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// elect.sync: true in exactly one lane of a converged warp.  Guarding tcgen05.mma / tcgen05.commit with THIS predicate
// (rather than lane == 0) lets ptxas see that a single thread executes them; otherwise every UTCHMMA / UTCBAR is
// wrapped in an ELECT + R2UR + BRA.U.ANY serialisation loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP (barrier count = number of warps): a warp-wide arrive on one mbarrier is 32 serialised
// shared-memory atomics.  __syncwarp orders the lanes' earlier writes / TMEM reads before the elected arrival.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100); SWIZZLE_NONE, base offset 0
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with compile-time offsets added to both descriptors INSIDE the asm block: the front end cannot hoist
// "base + constant" out of the issuing loop into one live 64-bit register per MMA (the 40-register MMA warp spilled 18
// weight descriptors to local memory); ptxas turns each add into one UIADD3.64 on the uniform datapath.
template <uint32_t kAOff, uint32_t kBOff>
__device__ __forceinline__ void umma_f16_off(uint32_t tmem_d, uint64_t da_base, uint64_t db_base, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\tadd.u64 ta, %1, %5;\n\tadd.u64 tb, %2, %6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da_base), "l"(db_base), "r"(idesc), "r"(accumulate), "n"(kAOff), "n"(kBOff)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// Ampere-style asynchronous 16-byte copy global -> shared; src_bytes < 16 zero-fills the remainder.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Normalise-on-load of 8 fp16 channels: y = lrelu(x * scale + shift) in packed half2 arithmetic
// (one HFMA2 + HMUL2 + HMNMX2 per channel pair).  slope in (0, 1]: max(v, slope * v) is LeakyReLU, slope 1 = identity.
__device__ __forceinline__ uint4 xform8_h2(const uint4 raw, const __half2* s2, const __half2* t2, const __half2* l2) {
  const __half2* x = reinterpret_cast<const __half2*>(&raw);
  uint4 o;
  __half2* y = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 v = __hfma2(x[e], s2[e], t2[e]);
    y[e] = __hmax2(v, __hmul2(v, l2[e]));
  }
  return o;
}

// 32 bytes per lane to addresses that differ between neighbouring lanes (NDHWC rows: one voxel per lane).  Two plain
// 16-byte stores per lane write HALF a 32-byte sector per lane and instruction, and L1 forwards every partial sector to
// L2 as a whole one (ncu on the transposed conv: l1tex2xbar write bytes = 2x the data, l1tex throughput 88 %).  Here
// lanes 2i and 2i+1 swap halves so that each instruction writes whole sectors: first the even lane's voxel (even lane
// the low half, odd lane the high half), then the odd lane's.  All 32 lanes must call; `partner_delta` is the byte
// distance from a lane's own address to its partner's (+stride for even lanes, -stride for odd lanes); `ok` predicates
// the lane's OWN voxel.
__device__ __forceinline__ void stg32_paired(void* q_own, long long partner_delta, const uint4 lo, const uint4 hi, bool ok, int lane) {
  const bool odd = lane & 1;
  const uint4 send = odd ? lo : hi;
  uint4 recv;
  recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
  recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
  recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
  recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
  const bool partner_ok = __shfl_xor_sync(0xffffffffu, (int)ok, 1) != 0;
  char* own = reinterpret_cast<char*>(q_own);
  char* partner = own + partner_delta;
  // instruction 1: the even lane's voxel; instruction 2: the odd lane's voxel
  char* a1 = odd ? partner + 16 : own;
  char* a2 = odd ? own + 16 : partner;
  const uint4 d1 = odd ? recv : lo;
  const uint4 d2 = odd ? hi : recv;
  const bool ok1 = odd ? partner_ok : ok;
  const bool ok2 = odd ? ok : partner_ok;
  if (ok1) *reinterpret_cast<uint4*>(a1) = d1;
  if (ok2) *reinterpret_cast<uint4*>(a2) = d2;
}

}  // namespace fnnu
