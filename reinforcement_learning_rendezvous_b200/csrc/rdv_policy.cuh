// rdv_policy.cuh -- the SB3 MlpPolicy actor (17 -> 64 -> 64 -> 6, tanh; main.py:39-48) evaluated by a WARP for
// its 32 environments on the tensor cores, inside the rollout kernel (no round trip through HBM between the
// observation, the policy and the next env step).
//
// Per warp and step: X[32x17] W0^T -> tanh -> H1[32x64] W1^T -> tanh -> H2[32x64] W2^T -> clip -> A[32x6], i.e.
// (48 + 128 + 16) m16n8k8 TF32 MMAs per pass.  TF32 keeps 10 mantissa bits, which would move actions by ~1e-3 and
// flip episodes against the fp32 policy of the reference (hard part 6), so every product is
// formed as 3xTF32 (a_hi b_hi + a_lo b_hi + a_hi b_lo with x = hi + lo): fp32-level accuracy (~1e-6 on the
// actions) at three passes of an otherwise idle tensor pipe.  The weights are split into hi / lo once per
// launch into shared memory.
//
// The accumulator layout of one layer is re-used as the A operand of the next WITHOUT moving data between lanes:
// the K index of a product can be enumerated in any order, so the lane that holds columns (2t, 2t+1) of an 8-wide
// output tile declares them to be K-slots (t, t+4) of the next MMA's 8-deep K tile, and the weight fragments are
// fetched with the same permutation (they come from shared memory with arbitrary indexing anyway).
//
// Legacy warp-level mma.sync is used on purpose: the batch per warp is 32 rows, the MLP is 11 kflop per env
// and <4 % of a step's time; a tcgen05 / TMEM formulation needs 128-row CTA tiles and buys nothing here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rdv_b200.h"
#include "rdv_math.cuh"

namespace rdv {

constexpr int PI_H = 64;             // hidden width (SB3 default net_arch [64, 64])
constexpr int PI_S0 = 24;            // row stride of W0 in shared memory (17 inputs padded to 3 K-tiles of 8)
constexpr int PI_S1 = 72;            // row stride of W1 / W2 (64 + 8: conflict-free 64-bit fragment loads)

struct PolicyShared {
    float w0h[PI_H * PI_S0], w0l[PI_H * PI_S0];
    float w1h[PI_H * PI_S1], w1l[PI_H * PI_S1];
    float w2h[8 * PI_S1], w2l[8 * PI_S1];
    float b0[PI_H], b1[PI_H], b2[8];
    float std[8];                    // exp(log_std) of the Gaussian head (sampling mode)
};

__device__ __forceinline__ float tf32_hi(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
          "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// c += a * b with a = ah + al, b = bh + bl (the al*bl term, 2^-22 relative, is dropped)
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const float (&ah)[4], const float (&al)[4], float2 bh, float2 bl)
{
    mma_tf32(c, al, bh.x, bh.y);
    mma_tf32(c, ah, bl.x, bl.y);
    mma_tf32(c, ah, bh.x, bh.y);
}

// tanh to fp32 accuracy in ABSOLUTE terms (the next layer is linear in it): 1 - 2 / (1 + exp(2x)), five
// instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA).  exp(2x) -> 0 gives -1, -> inf gives +1; no range split.
__device__ __forceinline__ float tanh_fast(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));      // 2^(2x log2 e) = exp(2x)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// x = hi + lo with hi = the TF32 part (mantissa truncated to 10 bits; the tensor core ignores the 13 low bits of
// an operand, so lo may stay a plain fp32 number)
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// Whole CTA: split the weights into TF32 hi / lo parts in shared memory (once per launch).
__device__ __forceinline__ void policy_load(const RdvPolicy &pi, PolicyShared &s)
{
    for (int idx = threadIdx.x; idx < PI_H * PI_S0; idx += blockDim.x) {
        const int n = idx / PI_S0, k = idx % PI_S0;
        const float w = k < RDV_OBS_DIM ? pi.w0[n * RDV_OBS_DIM + k] : 0.0f;
        const float h = tf32_hi(w);
        s.w0h[idx] = h; s.w0l[idx] = tf32_hi(w - h);
    }
    for (int idx = threadIdx.x; idx < PI_H * PI_S1; idx += blockDim.x) {
        const int n = idx / PI_S1, k = idx % PI_S1;
        const float w = k < PI_H ? pi.w1[n * PI_H + k] : 0.0f;
        const float h = tf32_hi(w);
        s.w1h[idx] = h; s.w1l[idx] = tf32_hi(w - h);
    }
    for (int idx = threadIdx.x; idx < 8 * PI_S1; idx += blockDim.x) {
        const int n = idx / PI_S1, k = idx % PI_S1;
        const float w = (n < RDV_ACT_DIM && k < PI_H) ? pi.w2[n * PI_H + k] : 0.0f;
        const float h = tf32_hi(w);
        s.w2h[idx] = h; s.w2l[idx] = tf32_hi(w - h);
    }
    for (int idx = threadIdx.x; idx < PI_H; idx += blockDim.x) { s.b0[idx] = pi.b0[idx]; s.b1[idx] = pi.b1[idx]; }
    if (threadIdx.x < 8) {
        s.b2[threadIdx.x] = threadIdx.x < RDV_ACT_DIM ? pi.b2[threadIdx.x] : 0.0f;
        s.std[threadIdx.x] = (threadIdx.x < RDV_ACT_DIM && pi.log_std) ? expf(pi.log_std[threadIdx.x]) : 0.0f;
    }
}

// Whole warp.  x_stage: the warp's 32 observation rows [32][17] in shared memory (written by the caller, who
// also owns the __syncwarp before the call); act_stage: warp-private [32][8] floats.  On return lane L holds
// the actor's mean action (NOT clipped) of its env in act[0..5].
__device__ __forceinline__ void policy_forward_warp(const PolicyShared &s, const float *x_stage, float *act_stage,
                                                    float (&act)[RDV_ACT_DIM])
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float h[2][8][4];                                   // [m-tile][n-tile][c0..c3]: 32 x 64 hidden activations
    // ---- layer 1: accumulators start at the bias ----
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
        const float bb0 = s.b0[8 * jn + 2 * t], bb1 = s.b0[8 * jn + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { h[mt][jn][0] = bb0; h[mt][jn][1] = bb1; h[mt][jn][2] = bb0; h[mt][jn][3] = bb1; }
    }
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        float ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int k0 = 8 * kt + 2 * t;              // K-slots (t, t+4) <-> inputs (k0, k0+1)
            const float *r0 = x_stage + (16 * mt + g) * RDV_OBS_DIM, *r1 = r0 + 8 * RDV_OBS_DIM;
            const float a[4] = {k0 < RDV_OBS_DIM ? r0[k0] : 0.0f, k0 < RDV_OBS_DIM ? r1[k0] : 0.0f,
                                k0 + 1 < RDV_OBS_DIM ? r0[k0 + 1] : 0.0f, k0 + 1 < RDV_OBS_DIM ? r1[k0 + 1] : 0.0f};
#pragma unroll
            for (int j = 0; j < 4; ++j) tf32_split(a[j], ah[mt][j], al[mt][j]);
        }
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) {
            const int off = (8 * jn + g) * PI_S0 + 8 * kt + 2 * t;
            const float2 bh = *reinterpret_cast<const float2 *>(s.w0h + off);
            const float2 bl = *reinterpret_cast<const float2 *>(s.w0l + off);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_3xtf32(h[mt][jn], ah[mt], al[mt], bh, bl);
        }
    }
    // ---- layer 2: tanh(H1) is the A operand, K-tile jk = output tile jk of layer 1 ----
    float h2[2][8][4];
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
        const float bb0 = s.b1[8 * jn + 2 * t], bb1 = s.b1[8 * jn + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { h2[mt][jn][0] = bb0; h2[mt][jn][1] = bb1; h2[mt][jn][2] = bb0; h2[mt][jn][3] = bb1; }
    }
#pragma unroll
    for (int jk = 0; jk < 8; ++jk) {
        float ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            // accumulator (row g | g+8, cols 2t | 2t+1) -> A fragment (a0, a1, a2, a3) = (c0, c2, c1, c3)
            const float a[4] = {tanh_fast(h[mt][jk][0]), tanh_fast(h[mt][jk][2]), tanh_fast(h[mt][jk][1]),
                                tanh_fast(h[mt][jk][3])};
#pragma unroll
            for (int j = 0; j < 4; ++j) tf32_split(a[j], ah[mt][j], al[mt][j]);
        }
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) {
            const int off = (8 * jn + g) * PI_S1 + 8 * jk + 2 * t;
            const float2 bh = *reinterpret_cast<const float2 *>(s.w1h + off);
            const float2 bl = *reinterpret_cast<const float2 *>(s.w1l + off);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_3xtf32(h2[mt][jn], ah[mt], al[mt], bh, bl);
        }
    }
    // ---- layer 3: 64 -> 6 (one 8-wide tile) ----
    float o[2][4];
    {
        const float bb0 = s.b2[2 * t], bb1 = s.b2[2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { o[mt][0] = bb0; o[mt][1] = bb1; o[mt][2] = bb0; o[mt][3] = bb1; }
    }
#pragma unroll
    for (int jk = 0; jk < 8; ++jk) {
        const int off = g * PI_S1 + 8 * jk + 2 * t;
        const float2 bh = *reinterpret_cast<const float2 *>(s.w2h + off);
        const float2 bl = *reinterpret_cast<const float2 *>(s.w2l + off);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float a[4] = {tanh_fast(h2[mt][jk][0]), tanh_fast(h2[mt][jk][2]), tanh_fast(h2[mt][jk][1]),
                                tanh_fast(h2[mt][jk][3])};
            float ah[4], al[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tf32_split(a[j], ah[j], al[j]);
            mma_3xtf32(o[mt], ah, al, bh, bl);
        }
    }
    // ---- accumulator layout -> one row of 6 actions per lane, through the warp's staging tile ----
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        act_stage[(16 * mt + g) * 8 + 2 * t] = o[mt][0];
        act_stage[(16 * mt + g) * 8 + 2 * t + 1] = o[mt][1];
        act_stage[(16 * mt + g + 8) * 8 + 2 * t] = o[mt][2];
        act_stage[(16 * mt + g + 8) * 8 + 2 * t + 1] = o[mt][3];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < RDV_ACT_DIM; ++j) act[j] = act_stage[lane * 8 + j];
    __syncwarp();
}

// Six N(0,1) float32 draws for (noise seed; env id, step index): Philox blocks 0x20000000 | {0,1,2}, one
// Box-Muller pair per block (the Gaussian head of SB3's DiagGaussianDistribution.sample()).
__device__ __forceinline__ void philox_normals(uint64_t seed, int64_t env_id, int64_t step_index, float (&z)[6])
{
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        uint32_t c[4] = {(uint32_t)env_id, (uint32_t)((uint64_t)env_id >> 32), (uint32_t)step_index,
                         0x20000000u | blk | (((uint32_t)((uint64_t)step_index >> 32) & 0x00FFFFFFu) << 4)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
        const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * blk] = r * cs;
        z[2 * blk + 1] = r * sn;
    }
}

}  // namespace rdv
