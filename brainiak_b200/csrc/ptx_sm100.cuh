// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA, tcgen05 (MMA / TMEM), legacy mma.sync.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fcma {

#ifndef FCMA_WATCHDOG_SPINS
// Bounded mbarrier waits: a protocol bug traps (kernel aborts with an error) instead of hanging the
// GPU.  ~2^28 polls is tens of seconds of spinning — far beyond any legitimate wait.
#define FCMA_WATCHDOG_SPINS (1u << 28)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lane_id()
{
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > FCMA_WATCHDOG_SPINS) __trap();
    }
}

// One lane of a fully converged warp (elect.sync).  Together with a warp index obtained through
// __shfl_sync(.., 0) this lets the compiler keep the single-thread TMA / tcgen05 issue code on the
// UNIFORM datapath (UR operands) instead of wrapping every UTCHMMA / UTMALDG in an ELECT +
// R2UR.BROADCAST waterfall loop (measured: 170 instead of 128 clk per MMA issued).
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *m, uint64_t *bar, void *smem_dst, int c0,
                                            int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
          "r"(c2)
        : "memory");
}

// 5-D tiled load global -> shared (column-direction pass: one bulk copy per [32 epochs][16 rows][32 columns] brick)
__device__ __forceinline__ void tma_load_5d(const CUtensorMap *m, uint64_t *bar, uint32_t smem_dst, int c0, int c1,
                                            int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "r"(c4)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; KIND 0: kind::f16 (bf16 operands), 1: kind::tf32
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate)
{
    if constexpr (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            :
            : "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :
            : "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of exactly 128 bytes.
//   start address >> 4 | LBO (ignored for swizzled K-major) | SBO = 8 rows * 128 B = 1024 B |
//   version 1 (sm_100) | layout type 2 (SWIZZLE_128B).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // bits [0,14)
    d |= (uint64_t)0 << 16;                            // LBO
    d |= (uint64_t)(1024 >> 4) << 32;                  // SBO, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}


// shared-memory matrix descriptor: MN-major tile, 128-byte swizzle (canonical layout, in 16-byte units,
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)): a row of 64 MN-elements (128 B) per k, 8 k-rows per 1024-byte swizzle atom,
// atoms `sbo` bytes apart along K and `lbo` bytes apart along MN).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // bits [0,14)
    d |= (uint64_t)(lbo >> 4) << 16;                   // LBO, bits [16,30)
    d |= (uint64_t)(sbo >> 4) << 32;                   // SBO, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// instruction descriptor, fp16 operands, fp32 accumulate, both operands MN-major ([15] A major, [16] B major = 1)
__device__ __forceinline__ uint32_t make_idesc_f16_mn(uint32_t M, uint32_t N)
{
    return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, both operands K-major
//   [4,6) D format (1 = f32) | [7,10) A format | [10,13) B format (1 = bf16, 2 = tf32)
//   [15] A major (0 = K) | [16] B major | [17,23) N >> 3 | [24,29) M >> 4
__device__ __forceinline__ uint32_t make_idesc(int operand_fmt, uint32_t M, uint32_t N)
{
    uint32_t fmt = (uint32_t)operand_fmt;  // 0 = f16, 1 = bf16, 2 = tf32
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// 32 lanes x 32 consecutive TMEM columns -> 32 registers per thread (thread = lane = D row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- packed fp32x2 arithmetic (sm_100: FFMA2)
// d = a * b + c on two fp32 lanes in one instruction: halves the issue slots (and the instruction energy)
// of the fp32 epilogue math, which matters on a kernel that runs at the board's power cap.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

// ---------------------------------------------------------------- cp.async (LDGSTS) with zero fill
// copies src_bytes (0..16) from global and zero-fills the rest of the 16-byte shared destination
__device__ __forceinline__ void cp_async_16_zfill(void *smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
// same with a shared-state-space destination address
__device__ __forceinline__ void cp_async_16_zfill_s(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// unconditional 16-byte copy
__device__ __forceinline__ void cp_async_16_s(uint32_t smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// same with a 256-byte L2 prefetch hint: the miss also brings the other 128-byte half of the 256-byte granule into L2
// (the column pass reads 128-byte lines 1 KB apart; the neighbouring strip's CTA wants the other half at about the same time)
__device__ __forceinline__ void cp_async_16_zfill_s_l2_256(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// 4-byte copy (cp.async.ca; zero-filled when src_bytes == 0): scatters single words, used where the shared layout is a
// transpose of the global one (k_norm_syrk_cols64)
__device__ __forceinline__ void cp_async_4_zfill_s(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// unconditional 4-byte copy
__device__ __forceinline__ void cp_async_4_s(uint32_t smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- legacy warp MMA (tf32, m16n8k8)
__device__ __forceinline__ uint32_t f32_to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// explicit shared-state-space 16-byte accesses (a pointer derived from the dynamic smem base plus a runtime
// offset is generic for the compiler, which then emits the slower generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t a)
{
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
{
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr)
{
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}

// ---------------------------------------------------------------- TMA store (shared -> global)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, uint32_t smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read0()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// four 8x8 b16 matrices, transposed on the way: lane l supplies the address of row (l % 8) of matrix (l / 8); register k
// of lane (g = l / 4, t = l % 4) receives elements [2t][g], [2t+1][g] of matrix k as a half2
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(saddr)
                 : "memory");
}
// packed fp32x2 add
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr)
{
    uint2 r;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr) : "memory");
    return r;
}
// fp16 operands (same 11-bit significand as tf32), twice the k extent per instruction
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2_rn(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// arrive (count only) on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        :
        : "r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// TMA load issued by either CTA of a pair; the complete_tx goes to the barrier of the EVEN CTA
// (peer bit of the shared::cluster address cleared), the data to this CTA's shared memory.
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap *m, uint64_t *bar, void *smem_dst, int c0,
                                                int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
          "r"(c1), "r"(c2)
        : "memory");
}
// same with an L2 cache-policy operand (createpolicy): the operand tiles are re-read by every tile of a column
// group while 26 GB of streaming stores pass through L2
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_2sm_hint(const CUtensorMap *m, uint64_t *bar, void *smem_dst, int c0,
                                                     int c1, int c2, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
          "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this smem offset in BOTH CTAs (mask 0b11)
__device__ __forceinline__ void tc_commit_2sm(uint64_t *bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate)
{
    if constexpr (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            :
            : "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :
            : "r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

}  // namespace fcma
