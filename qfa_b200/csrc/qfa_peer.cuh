// One-shot all-reduce(sum) of the packed gradient accumulator over PEER-MAPPED memory (NVLink 5 / NVSwitch), sm_100a.
//
// The exchange step of the data-parallel train step (one all-reduce of `acc`, 87 KB on the SDSS grid, 138 KB on
// L32; 8(f) row 4) is latency-bound: NCCL's ring/tree launch + protocol costs ~50 us inside the captured step graph, 16 % of the
// 8 192-spectra step.  Here every rank PUBLISHES its accumulator in a buffer that all peers have mapped (torch symmetric memory:
// cuMem + NVLink peer mappings; the library only sees raw pointers), raises a flag in every peer's memory, waits for every peer's
// flag in its own, and then sums the world's buffers straight out of peer memory in rank order -- one kernel, no staging copies,
// no host involvement, the same bits on every rank (fixed summation order), so the replicated Adam update stays bit-identical.
//
//   peer buffer of rank r (same layout on every rank, qfa_peer_buffer_bytes):  [ flags[32] | pub[0] | pub[1] ]
//     pub[s & 1]   the accumulator of step s (double-buffered: a rank that is one step ahead writes the OTHER half; it cannot be
//                  two ahead, because finishing step s+1 needs every peer's flag of step s+1, which that peer raises after it has
//                  finished READING step s)
//     flags[q]     number of steps rank q has published (written by rank q with st.release.sys, polled here with ld.acquire.sys)
//   state (local, device): {steps done, CTA ticket}: the step number lives in device memory, so a captured CUDA graph of the
//                  train step replays without host arguments.
//
// The grid is at most one CTA per SM (all co-resident: CTAs spin on flags), phase 1 (publish) and phase 2 (sum) are separated by a
// ticket: the last CTA to finish publishing raises the flags.  A rank may wait for a slow peer (one that writes a checkpoint,
// say) for as long as NCCL's watchdog would: the wait backs off with nanosleep and only traps after QFA_PEER_TIMEOUT_S seconds
// (default 600) instead of hanging the process for ever.
#pragma once
#include <cstddef>
#include <cstdint>

namespace qfa {
namespace peer {

__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <typename T>
__device__ __forceinline__ T ld_volatile(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void add(uint4& a, const uint4& b) {
        a.x = __float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x));
        a.y = __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y));
        a.z = __float_as_uint(__uint_as_float(a.z) + __uint_as_float(b.z));
        a.w = __float_as_uint(__uint_as_float(a.w) + __uint_as_float(b.w));
    }
};
template <> struct Vec16<double> {
    static constexpr int N = 2;
    static __device__ __forceinline__ void add(uint4& a, const uint4& b) {
        const double lo = __hiloint2double((int)a.y, (int)a.x) + __hiloint2double((int)b.y, (int)b.x);
        const double hi = __hiloint2double((int)a.w, (int)a.z) + __hiloint2double((int)b.w, (int)b.z);
        a.x = (unsigned)__double2loint(lo); a.y = (unsigned)__double2hiint(lo);
        a.z = (unsigned)__double2loint(hi); a.w = (unsigned)__double2hiint(hi);
    }
};

constexpr int kMaxWorld = 32;          // one polling thread per rank (first warp of every CTA)
constexpr size_t kFlagBytes = 128;     // flags first: their place does not depend on the element type of the accumulator

__host__ __device__ constexpr size_t pub_bytes(size_t n, size_t elem) { return (n * elem + 15) / 16 * 16; }

// acc (n elements, 16-byte aligned) <- sum over ranks of acc, in rank order.  peer_base[q] = rank q's peer buffer as mapped HERE.
template <typename T>
__global__ void __launch_bounds__(256) k_peer_allreduce(T* __restrict__ acc, size_t n, char* const* __restrict__ peer_base,
                                                        unsigned* __restrict__ state, int world, int rank,
                                                        unsigned long long timeout_ns) {
    constexpr int VN = Vec16<T>::N;
    __shared__ int s_last;
    const size_t pb = pub_bytes(n, sizeof(T));
    const unsigned s = ld_volatile(state);                           // steps done so far: every thread reads it BEFORE the ticket
    const size_t pub_off = kFlagBytes + (size_t)(s & 1u) * pb, flags_off = 0;
    const size_t nvec = n / VN;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    // ---- phase 1: publish
    {
        char* mine = peer_base[rank] + pub_off;
        const uint4* a4 = reinterpret_cast<const uint4*>(acc);
        uint4* p4 = reinterpret_cast<uint4*>(mine);
        for (size_t i = gid; i < nvec; i += gsz) p4[i] = a4[i];
        T* pt = reinterpret_cast<T*>(mine);
        for (size_t i = nvec * VN + gid; i < n; i += gsz) pt[i] = acc[i];
    }
    // ordering of the published data before the flags: bar.sync -> one device-scope fence per CTA -> ticket -> (last CTA) system-
    // scope fence + release store.  Fences are cumulative, so the per-thread system fences this started with (several us of
    // membar.sys) are not needed.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(state + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        // every CTA has published (and read `s`): raise this rank's flag in every peer's memory, then advance the step counter
        if (threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned*>(peer_base[threadIdx.x] + flags_off) + rank, s + 1u);
        }
        if (threadIdx.x == 0) { state[1] = 0u; state[0] = s + 1u; }
    }
    // ---- wait until every rank has published step s
    if (threadIdx.x < world) {
        const unsigned* f = reinterpret_cast<const unsigned*>(peer_base[rank] + flags_off) + threadIdx.x;
        unsigned long long t0 = 0ull;
        int spins = 0;
        while ((int)(ld_acquire_sys(f) - (s + 1u)) < 0) {
            if (++spins > 64) {
                __nanosleep(spins > 4096 ? 1000 : 64);
                const unsigned long long t = globaltimer_ns();
                if (t0 == 0ull) t0 = t;
                else if (t - t0 > timeout_ns) __trap();               // a peer never arrived
            }
        }
    }
    __syncthreads();
    // ---- phase 2: sum the world's buffers out of peer memory, in rank order
    {
        uint4* a4 = reinterpret_cast<uint4*>(acc);
        for (size_t i = gid; i < nvec; i += gsz) {
            // the loads of eight ranks are issued together (one NVLink round trip per group, not per rank); the adds keep rank order
            uint4 sum = make_uint4(0u, 0u, 0u, 0u);
            for (int q0 = 0; q0 < world; q0 += 8) {
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (q0 + j < world) v[j] = ld_volatile_v4(reinterpret_cast<const uint4*>(peer_base[q0 + j] + pub_off) + i);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (q0 + j < world) { if (q0 + j == 0) sum = v[0]; else Vec16<T>::add(sum, v[j]); }
            }
            a4[i] = sum;
        }
        for (size_t i = nvec * VN + gid; i < n; i += gsz) {
            T sum = ld_volatile(reinterpret_cast<const T*>(peer_base[0] + pub_off) + i);
            for (int q = 1; q < world; ++q) sum += ld_volatile(reinterpret_cast<const T*>(peer_base[q] + pub_off) + i);
            acc[i] = sum;
        }
    }
}

}  // namespace peer
}  // namespace qfa
