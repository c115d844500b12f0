// Thin inline-PTX layer for the Blackwell (sm_100a) tensor path: tcgen05.mma kind::tf32 with
// shared-memory descriptors, TMEM allocation / loads, mbarriers and 1-D bulk (TMA) copies.
// Bit layouts follow the PTX ISA "tcgen05 shared memory descriptor / instruction descriptor"
// tables (cross-checked against the CuTe headers shipped in the image, pattern reference only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qfa {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error flag, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    if (err_flag) atomicExch(err_flag, 1);
    return false;
}

// ------------------------------------------------------------------ async-proxy plumbing
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Where the generic-proxy writes of an operand tile are fenced for the async proxy (tcgen05.mma).  Default: by every WRITER
// before its mbarrier arrive (the documented pattern).  -DQFA_CONSUMER_FENCE: by the ISSUING thread after its mbarrier wait
// (the writers' stores are ordered before the fence by the arrive/wait pair) -- the writers then execute no MEMBAR, which
// otherwise waits for every global load they have in flight.
#ifdef QFA_CONSUMER_FENCE
__device__ __forceinline__ void fence_proxy_async_writer() {}
__device__ __forceinline__ void fence_proxy_async_issuer() { fence_proxy_async(); }
#else
__device__ __forceinline__ void fence_proxy_async_writer() { fence_proxy_async(); }
__device__ __forceinline__ void fence_proxy_async_issuer() {}
#endif

// 1-D bulk copy global -> shared (TMA, SASS UBLKCP); 16-byte aligned src/dst/size
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 1-D bulk copy shared -> global (bulk-group completion); 16-byte aligned src/dst/size
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk groups of this thread are complete (writes performed)
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// thread (warp w, lane l) receives columns [col, col+8) of TMEM lane 32*(w%4)+l
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ descriptors + MMA
// K-major operand tile, 128-byte rows (32 tf32), SWIZZLE_128B, 8-row groups 1024 B apart.
// Tile base must be 1024-byte aligned; advancing K by 8 tf32 (= 32 B) adds 2 to the low word.
__device__ __forceinline__ uint64_t desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset      bits [16,30) (unused: swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B bits [61,64)
    return d;
}
// byte offset of element (row r, k) inside such a tile (k in [0,32))
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}

inline uint32_t sw128_offset_host(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4)                        // D format  : F32
           | (2u << 7)                      // A format  : TF32
           | (2u << 10)                     // B format  : TF32
           | ((uint32_t)(N >> 3) << 17)     // N / 8     (A, B both K-major: bits 15,16 = 0)
           | ((uint32_t)(M >> 4) << 24);    // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M x 8 tf32 per instruction) is read from TENSOR MEMORY -- lane = row,
// 8 consecutive 32-bit columns = the K elements (fp32 accumulator bits are taken as tf32: the low 13 mantissa bits are ignored)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all previously issued MMAs of this thread -> arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// float -> tf32 (round to nearest, ties away) and the residual for the 3xTF32 split
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc
}  // namespace qfa
