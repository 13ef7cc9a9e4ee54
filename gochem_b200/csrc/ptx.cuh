// ptx.cuh -- inline-PTX helpers shared by the TMA-staged streaming kernels (sm_100a): mbarriers, 1-D bulk copies
// (cp.async.bulk ... mbarrier::complete_tx), L2 cache policies, cluster addressing and DSMEM stores, named barriers.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace gcb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// shared-memory counter bump that releases the caller's earlier writes and acquires those of the previous bumpers
// (with __syncwarp() on both sides it hands a warp's shared-memory stores to the warp that arrives last; no MEMBAR)
__device__ __forceinline__ uint32_t atom_add_acq_rel_cta(unsigned int* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// 2-D tiled TMA load (cp.async.bulk.tensor): box at element coordinates (c0 = innermost, c1 = row) of the tensor described by
// tmap -> shared memory, completion counted in bytes on the mbarrier.  Out-of-bounds elements arrive as zeros.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// sqrt of a non-negative double from the hardware reciprocal-square-root seed and one correction of cubic order (relative
// error < 2^-58; derivation at fast_sqrt_pos in super_cluster.cu: five operations).  Zero and denormal arguments return 0.
__device__ __forceinline__ double sqrt_cubic(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y;
    const double e = fma(-g, y, 1.0);
    const double c = fma(e, 0.375, 0.5);
    const double out = fma(g, e * c, g);
    return (__double2hiint(x) < 0x00200000) ? 0.0 : out;
}

// 12 per-lane values -> warp totals with 15 shuffles (halving butterfly).  On return, lanes with
// (lane & 3) == 0 hold: t0 = total of value 6*b4 + 3*b3 + 2*b2, t1 = total of value 6*b4 + 3*b3 + 1,
// where b4, b3, b2 are lane bits 4, 3, 2.  Fixed tree -> deterministic.
__device__ __forceinline__ void warp_reduce12(const double (&v)[12], int lane, double& t0, double& t1) {
    double w[6], u[3];
    const bool h16 = lane & 16;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const double send = h16 ? v[j] : v[j + 6];
        const double keep = h16 ? v[j + 6] : v[j];
        w[j] = keep + shx(send, 16);
    }
    const bool h8 = lane & 8;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double send = h8 ? w[j] : w[j + 3];
        const double keep = h8 ? w[j + 3] : w[j];
        u[j] = keep + shx(send, 8);
    }
    const bool h4 = lane & 4;
    {
        const double send = h4 ? u[0] : u[2];
        const double keep = h4 ? u[2] : u[0];
        t0 = keep + shx(send, 4);
        t1 = u[1] + shx(u[1], 4);
    }
    t0 += shx(t0, 2); t1 += shx(t1, 2);
    t0 += shx(t0, 1); t1 += shx(t1, 1);
}

}  // namespace ptx
}  // namespace gcb
