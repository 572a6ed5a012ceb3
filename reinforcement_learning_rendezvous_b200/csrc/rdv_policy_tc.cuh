// rdv_policy_tc.cuh -- the stand-alone batched actor forward (rdv_policy_forward: model.predict of the SB3
// MlpPolicy, monte_carlo.py:128-133) on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators.
//
// One CTA of 128 threads owns tiles of 128 environments (UMMA M = 128, one env per thread / TMEM lane):
//
//   obs tile -> shared (A operand, K-major, hi / lo TF32 split)
//   layer 1: D1[128x64] = A0[128x24] W0^T   tcgen05.mma kind::tf32, 3 K-steps x 3 products (3xTF32)
//   tcgen05.ld D1 -> registers, + bias, tanh, split -> shared A1[128x64]
//   layer 2: D2[128x64] = A1 W1^T           8 K-steps x 3
//   tcgen05.ld D2 -> registers, + bias, tanh, split -> shared A2
//   layer 3: D3[128x16] = A2 W2^T           8 K-steps x 3 (6 outputs padded to N = 16)
//   tcgen05.ld D3 -> registers, + bias, clip -> actions
//
// A CTA runs TWO such 128-thread groups side by side (256 threads, shared weights, private activation tiles,
// mbarriers, named barriers and TMEM columns), so the tensor-core phase of one tile overlaps the tanh epilogue of
// the other.  Per group, a single thread issues the MMAs and commits them to an mbarrier; the 128 threads wait on it, pull their TMEM
// lane with tcgen05.ld (32x32b: thread r of warp w <-> lane 32 w + r) and run the epilogue.  Every product is
// 3xTF32 (a_hi b_hi + a_lo b_hi + a_hi b_lo) so the result has fp32-level accuracy, like the reference's torch
// policy; weights are split once per CTA.  Shared-memory operands use the canonical no-swizzle K-major UMMA
// layout: 8-row x 16-byte core matrices, rows of a core matrix 16 B apart, 8-row groups SBO = 128 B apart, the
// two 16-byte K-chunks of a K = 8 step LBO = rows*16 B apart, i.e. element (r, k) lives at
// ((k / 4) * rows + r) * 16 + (k % 4) * 4 bytes.  (Descriptor bit layouts: CUTLASS cute/arch/mma_sm100_desc.hpp.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rdv_b200.h"
#include "rdv_policy.cuh"

namespace rdv {
namespace tc {

constexpr int TM = 128;                 // envs per tile = UMMA M = threads per CTA
constexpr int H = 64;                   // hidden width
constexpr int K0 = 24;                  // 17 inputs padded to 3 K-steps of 8
constexpr int N3 = 16;                  // 6 outputs padded to the smallest legal UMMA N for M = 128
constexpr uint32_t TMEM_COLS = 256;     // per group 128 columns: D1 @ +0..63, D2 @ +64..127, D3 re-uses +0..15

constexpr int GROUPS = 2;               // independent 128-thread tile pipelines per CTA
struct Smem {
    float ah[GROUPS][(H / 4) * TM * 4], al[GROUPS][(H / 4) * TM * 4];   // activations, [group][chunk][row][4]
    float w0h[(K0 / 4) * H * 4], w0l[(K0 / 4) * H * 4];          // [chunk][n][4]
    float w1h[(H / 4) * H * 4], w1l[(H / 4) * H * 4];
    float w2h[(H / 4) * N3 * 4], w2l[(H / 4) * N3 * 4];
    float b0[H], b1[H], b2[N3];
    uint64_t mbar[GROUPS];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, K-major), see header comment.
__device__ __forceinline__ uint64_t umma_desc(const void *p, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((smem_u32(p) >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                                        // descriptor version of sm_100
    return d;
}
// 32-bit instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    // the registers are only valid after the wait: tie them to it so no use can be scheduled above
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// weights [n_valid][k_valid] row-major -> chunked K-major hi / lo (zero padded to n_pad x k_pad)
__device__ __forceinline__ void load_weights(const float *w, int n_valid, int k_valid, int n_pad, int k_pad, float *wh,
                                             float *wl)
{
    for (int idx = threadIdx.x; idx < n_pad * k_pad; idx += GROUPS * TM) {
        const int nn = idx / k_pad, k = idx % k_pad;
        const float x = (nn < n_valid && k < k_valid) ? w[nn * k_valid + k] : 0.0f;
        const float hi = tf32_hi(x);
        const int o = ((k >> 2) * n_pad + nn) * 4 + (k & 3);
        wh[o] = hi;
        wl[o] = x - hi;
    }
}

// one layer: D[128 x n] (TMEM column d_col) = A[128 x 8*ksteps] W^T, 3xTF32, issued by the calling thread
__device__ __forceinline__ void issue_layer(const float *ah, const float *al, uint32_t tmem_d, const float *wh,
                                            const float *wl, int n, int ksteps, uint64_t *bar)
{
    const uint32_t idesc = umma_idesc(n);
    const uint32_t lbo_a = TM * 16, lbo_b = (uint32_t)n * 16;
    for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t a_hi = umma_desc(ah + (size_t)kk * 2 * TM * 4, lbo_a, 128);
        const uint64_t a_lo = umma_desc(al + (size_t)kk * 2 * TM * 4, lbo_a, 128);
        const uint64_t b_hi = umma_desc(wh + (size_t)kk * 2 * n * 4, lbo_b, 128);
        const uint64_t b_lo = umma_desc(wl + (size_t)kk * 2 * n * 4, lbo_b, 128);
        umma_tf32(tmem_d, a_lo, b_hi, idesc, kk > 0 ? 1u : 0u);
        umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
        umma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
    }
    umma_commit(bar);
}

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" :: "r"(1 + g) : "memory"); }

__global__ void __launch_bounds__(GROUPS * TM, 1) policy_tc_kernel(const RdvPolicy pi, const float *obs, float *actions, int64_t n)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw);
    const int g = threadIdx.x / TM, r = threadIdx.x % TM, warp = r >> 5;      // group, row in tile, warp in group
    float *ah = s.ah[g], *al = s.al[g];
    uint64_t *mbar = &s.mbar[g];

    // ---- one-time set-up: weights (hi / lo), biases, mbarrier, TMEM ----
    load_weights(pi.w0, H, RDV_OBS_DIM, H, K0, s.w0h, s.w0l);
    load_weights(pi.w1, H, H, H, H, s.w1h, s.w1l);
    load_weights(pi.w2, RDV_ACT_DIM, H, N3, H, s.w2h, s.w2l);
    if (threadIdx.x < H) { s.b0[threadIdx.x] = pi.b0[threadIdx.x]; s.b1[threadIdx.x] = pi.b1[threadIdx.x]; }
    if (threadIdx.x < N3) s.b2[threadIdx.x] = threadIdx.x < RDV_ACT_DIM ? pi.b2[threadIdx.x] : 0.0f;
    if (r == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_all = s.tmem_base;
    const uint32_t tmem = tmem_all + (uint32_t)g * 128;                    // this group's columns: D1 / D3 @ +0, D2 @ +64
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);       // this warp's 32 TMEM lanes
    uint32_t phase = 0;

    const int64_t tiles = (n + TM - 1) / TM;
    for (int64_t tile = (int64_t)blockIdx.x * GROUPS + g; tile < tiles; tile += (int64_t)gridDim.x * GROUPS) {
        const int64_t env = tile * TM + r;
        const bool valid = env < n;
        // ---- A0: this thread's observation row, hi / lo, one float4 per 16-byte chunk ----
        {
            float x[K0];
#pragma unroll
            for (int k = 0; k < K0; ++k) x[k] = (valid && k < RDV_OBS_DIM) ? obs[env * RDV_OBS_DIM + k] : 0.0f;
#pragma unroll
            for (int c = 0; c < K0 / 4; ++c) {
                float4 hi, lo;
                hi.x = tf32_hi(x[4 * c]); hi.y = tf32_hi(x[4 * c + 1]); hi.z = tf32_hi(x[4 * c + 2]); hi.w = tf32_hi(x[4 * c + 3]);
                lo.x = x[4 * c] - hi.x; lo.y = x[4 * c + 1] - hi.y; lo.z = x[4 * c + 2] - hi.z; lo.w = x[4 * c + 3] - hi.w;
                reinterpret_cast<float4 *>(ah)[c * TM + r] = hi;
                reinterpret_cast<float4 *>(al)[c * TM + r] = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
        // ---- layer 1 ----
        if (r == 0) {
            tc_fence_after();
            issue_layer(ah, al, tmem + 0, s.w0h, s.w0l, H, K0 / 8, mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                  // 64 columns, 16 at a time
            float v[16];
            tmem_ld16(lane_addr + 0 + 16 * q, v);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 hi, lo;
                float t0 = tanh_fast(v[4 * c] + s.b0[16 * q + 4 * c]), t1 = tanh_fast(v[4 * c + 1] + s.b0[16 * q + 4 * c + 1]);
                float t2 = tanh_fast(v[4 * c + 2] + s.b0[16 * q + 4 * c + 2]), t3 = tanh_fast(v[4 * c + 3] + s.b0[16 * q + 4 * c + 3]);
                hi.x = tf32_hi(t0); hi.y = tf32_hi(t1); hi.z = tf32_hi(t2); hi.w = tf32_hi(t3);
                lo.x = t0 - hi.x; lo.y = t1 - hi.y; lo.z = t2 - hi.z; lo.w = t3 - hi.w;
                reinterpret_cast<float4 *>(ah)[(4 * q + c) * TM + r] = hi;
                reinterpret_cast<float4 *>(al)[(4 * q + c) * TM + r] = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
        // ---- layer 2 ----
        if (r == 0) {
            tc_fence_after();
            issue_layer(ah, al, tmem + 64, s.w1h, s.w1l, H, H / 8, mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v[16];
            tmem_ld16(lane_addr + 64 + 16 * q, v);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 hi, lo;
                float t0 = tanh_fast(v[4 * c] + s.b1[16 * q + 4 * c]), t1 = tanh_fast(v[4 * c + 1] + s.b1[16 * q + 4 * c + 1]);
                float t2 = tanh_fast(v[4 * c + 2] + s.b1[16 * q + 4 * c + 2]), t3 = tanh_fast(v[4 * c + 3] + s.b1[16 * q + 4 * c + 3]);
                hi.x = tf32_hi(t0); hi.y = tf32_hi(t1); hi.z = tf32_hi(t2); hi.w = tf32_hi(t3);
                lo.x = t0 - hi.x; lo.y = t1 - hi.y; lo.z = t2 - hi.z; lo.w = t3 - hi.w;
                reinterpret_cast<float4 *>(ah)[(4 * q + c) * TM + r] = hi;
                reinterpret_cast<float4 *>(al)[(4 * q + c) * TM + r] = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
        // ---- layer 3 (D3 re-uses the columns of D1, which the layer-1 epilogue has drained) ----
        if (r == 0) {
            tc_fence_after();
            issue_layer(ah, al, tmem + 0, s.w2h, s.w2l, N3, H / 8, mbar);
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();
        {
            float v[16];
            tmem_ld16(lane_addr + 0, v);
            if (valid) {
#pragma unroll
                for (int j = 0; j < RDV_ACT_DIM; ++j)
                    actions[env * RDV_ACT_DIM + j] = fminf(1.0f, fmaxf(-1.0f, v[j] + s.b2[j]));     // np.clip
            }
        }
        tc_fence_before();
        group_sync(g);                         // A tiles and TMEM columns are reused by the group's next tile
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_all), "r"(TMEM_COLS) : "memory");
}

}  // namespace tc
}  // namespace rdv
