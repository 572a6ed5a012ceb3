// rdv_b200.cu -- kernels and C ABI of librdv_b200.so (sm_100a only).
//
// Layout in HBM: structure-of-arrays fp64 state [RDV_NF64][ld] + int32 [RDV_NI32][ld]; one
// thread per environment, so every state load/store is a fully coalesced 256 B warp access.
// Row-major [n][6] actions are read with 8/16-byte vector loads; the [n][17] float32
// observation is staged in shared memory and written out as contiguous float4 rows.
// Per-configuration constants (CW transition matrix, thresholds, reward coefficients,
// inertia) travel in the __grid_constant__ RdvParams kernel argument, i.e. the constant bank.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include "rdv_env.cuh"

namespace rdv {

constexpr int TPB = 64;   // 65,536 envs -> 1024 CTAs = 6.9 per SM on 148 SMs (1 % tail), see DESIGN.md

// ---------------------------------------------------------------------------------
// warp / block reduction of the statistics vector
// ---------------------------------------------------------------------------------
RDV_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct StepStats {
    // per-thread contributions; integers go through REDUX (__reduce_add_sync), doubles through shuffles
    unsigned steps, episodes, succeeded, collided, end[4], rk_acc, rk_rej, fail;
    double ep_return, ep_length, delta_v, delta_w, reward;
};

template <int NWARPS>
RDV_DEV void reduce_stats(const StepStats &st, double *g_stats, double (*s_stats)[RDV_NSTATS])
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[RDV_NSTATS];
    v[RDV_S_STEPS] = (double)__reduce_add_sync(full, st.steps);
    v[RDV_S_EPISODES] = (double)__reduce_add_sync(full, st.episodes);
    v[RDV_S_RK_ACCEPTED] = (double)__reduce_add_sync(full, st.rk_acc);
    v[RDV_S_RK_REJECTED] = (double)__reduce_add_sync(full, st.rk_rej);
    v[RDV_S_FAILURES] = (double)__reduce_add_sync(full, st.fail);
    v[RDV_S_REWARD] = warp_sum(st.reward);
    const bool any_done = v[RDV_S_EPISODES] > 0.0;       // warp-uniform
    if (any_done) {
        v[RDV_S_SUCCEEDED] = (double)__reduce_add_sync(full, st.succeeded);
        v[RDV_S_COLLIDED] = (double)__reduce_add_sync(full, st.collided);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[RDV_S_END_OBS + k] = (double)__reduce_add_sync(full, st.end[k]);
        v[RDV_S_RETURN] = warp_sum(st.ep_return);
        v[RDV_S_LENGTH] = warp_sum(st.ep_length);
        v[RDV_S_DELTA_V] = warp_sum(st.delta_v);
        v[RDV_S_DELTA_W] = warp_sum(st.delta_w);
    } else {
        v[RDV_S_SUCCEEDED] = v[RDV_S_COLLIDED] = v[RDV_S_RETURN] = v[RDV_S_LENGTH] = 0.0;
        v[RDV_S_DELTA_V] = v[RDV_S_DELTA_W] = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[RDV_S_END_OBS + k] = 0.0;
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < RDV_NSTATS; ++k) s_stats[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < RDV_NSTATS) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) acc += s_stats[w][threadIdx.x];
        if (acc != 0.0) atomicAdd(g_stats + threadIdx.x, acc);
    }
}

// ---------------------------------------------------------------------------------
// step kernel: RendezvousEnv.step (rendezvous_env.py:160-221), one thread per env.
// Finished envs are appended to reset_list (count in reset_list[0]) for the compacted
// reset kernel below, so the rare, long reset path never diverges a stepping warp.
// ---------------------------------------------------------------------------------
template <bool ISO, bool ACT_F64, bool CLOSED>
__global__ void __launch_bounds__(TPB) step_kernel(const __grid_constant__ RdvParams P, const RdvState S,
                                                   const RdvStepIO io, const int64_t n, int32_t *reset_list)
{
    __shared__ __align__(16) float s_obs[TPB * RDV_OBS_DIM];
    __shared__ double s_stats[TPB / 32][RDV_NSTATS];

    const int64_t base = (int64_t)blockIdx.x * TPB;
    const int64_t i = base + threadIdx.x;
    const bool active = i < n;
    StepStats st;
    memset(&st, 0, sizeof(st));

    if (active) {
        EnvRegs e;
        load_env(S, i, e);
        const int64_t ld = S.ld;
        double tdv = S.f64[RDV_TDV * ld + i], tdw = S.f64[RDV_TDW * ld + i], ep_ret = S.f64[RDV_EPRET * ld + i];
        int step = S.i32[RDV_I_STEP * ld + i], success = S.i32[RDV_I_SUCCESS * ld + i];
        int collided = S.i32[RDV_I_COLLIDED * ld + i];

        // ---- action ingest (:168-173, :201-202, :333) ----
        double dvb[3], dw[3], fuel;
        if (ACT_F64) {
            const double2 *ap = reinterpret_cast<const double2 *>(static_cast<const double *>(io.actions) + 6 * i);
            double2 a01 = ap[0], a23 = ap[1], a45 = ap[2];
            dvb[0] = a01.x * P.max_delta_v; dvb[1] = a01.y * P.max_delta_v; dvb[2] = a23.x * P.max_delta_v;
            dw[0] = a23.y * P.max_delta_w; dw[1] = a45.x * P.max_delta_w; dw[2] = a45.y * P.max_delta_w;
            double sv = fabs(a01.x) + fabs(a01.y) + fabs(a23.x);
            double sw = fabs(a23.y) + fabs(a45.x) + fabs(a45.y);
            tdv += sv * P.max_delta_v;
            tdw += sw * P.max_delta_w;
            fuel = __ddiv_rn(P.dt * P.fuel_coef * sv, 3.0 * P.max_delta_v);
        } else {
            // float32 actions follow NumPy-2 promotion: delta_v, total_delta_v
            // and the fuel term are rounded in fp32; delta_w and total_delta_w are fp64.
            const float2 *ap = reinterpret_cast<const float2 *>(static_cast<const float *>(io.actions) + 6 * i);
            float2 a01 = ap[0], a23 = ap[1], a45 = ap[2];
            dvb[0] = (double)__fmul_rn(a01.x, P.max_delta_v_f32);
            dvb[1] = (double)__fmul_rn(a01.y, P.max_delta_v_f32);
            dvb[2] = (double)__fmul_rn(a23.x, P.max_delta_v_f32);
            dw[0] = (double)a23.y * P.max_delta_w; dw[1] = (double)a45.x * P.max_delta_w;
            dw[2] = (double)a45.y * P.max_delta_w;
            float sv = __fadd_rn(__fadd_rn(fabsf(a01.x), fabsf(a01.y)), fabsf(a23.x));
            float sw = __fadd_rn(__fadd_rn(fabsf(a23.y), fabsf(a45.x)), fabsf(a45.y));
            tdv = (double)__fadd_rn((float)tdv, __fmul_rn(sv, P.max_delta_v_f32));
            tdw += (double)sw * P.max_delta_w;
            fuel = (double)__fdiv_rn(__fmul_rn(P.fuel_num_f32, sv), P.fuel_den_f32);
        }

        // ---- translation: impulse in LVLH, then the CW transition (:172-177, dynamics.py:24-55) ----
        {
            Rot Rc_old = rot_from_quat(e.qc);
            double dv[3];
            rot_apply(Rc_old, dvb, dv);
            double r0 = e.rc[0], r1 = e.rc[1], r2 = e.rc[2];
            double v0 = e.vc[0] + dv[0], v1 = e.vc[1] + dv[1], v2 = e.vc[2] + dv[2];
            const double *c = P.cw;
            e.rc[0] = fma(c[2], v1, fma(c[1], v0, c[0] * r0));
            e.rc[1] = fma(c[6], v1, fma(c[5], v0, fma(c[3], r0, c[4] * r1)));
            e.rc[2] = fma(c[8], v2, c[7] * r2);
            e.vc[0] = fma(c[11], v1, fma(c[10], v0, c[9] * r0));
            e.vc[1] = fma(c[14], v1, fma(c[13], v0, c[12] * r0));
            e.vc[2] = fma(c[16], v2, c[15] * r2);
        }

        // ---- attitude: impulsive rate change, then torque-free propagation of both bodies (:180-184) ----
        int rk_acc = 0, rk_rej = 0, fail = 0;
        {
            double y[7] = {e.qc[0], e.qc[1], e.qc[2], e.qc[3], e.wc[0] + dw[0], e.wc[1] + dw[1], e.wc[2] + dw[2]};
            double z[7] = {e.qt[0], e.qt[1], e.qt[2], e.qt[3], e.wt[0], e.wt[1], e.wt[2]};
#pragma unroll 1
            for (int body = 0; body < 2; ++body) {
                if (CLOSED) {
                    closed_form_attitude(y, P.dt);
                } else {
                    BodyConst bc;
                    bc.I = body ? P.inertia_t : P.inertia_c;
                    bc.Iinv = body ? P.inv_inertia_t : P.inv_inertia_c;
                    const double zero3[3] = {0.0, 0.0, 0.0};
                    bc.tau = body ? zero3 : P.torque_c;
                    int k = rk45_attitude<ISO>(y, P.dt, bc, rk_rej);
                    if (k < 0) fail = 1; else rk_acc += k;
                }
                double r = fast_rsqrt(dot4(y, y));                 // q / |q|  (:574-575, :601-602)
#pragma unroll
                for (int k = 0; k < 4; ++k) y[k] *= r;
#pragma unroll
                for (int k = 0; k < 7; ++k) { double t = y[k]; y[k] = z[k]; z[k] = t; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { e.qc[k] = y[k]; e.qt[k] = z[k]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { e.wc[k] = y[4 + k]; e.wt[k] = z[4 + k]; }
        }

        // ---- collision / success latch (:186-190) ----
        const Rot Rc = rot_from_quat(e.qc), Rt = rot_from_quat(e.qt);
        const double rc_sq = dot3(e.rc, e.rc), rc_n = sqrt(rc_sq);
        const double att = attitude_error(P, e, Rc, rc_sq);
        const bool col_now = collision_now(P, e, Rt, rc_sq, rc_n);
        ErrSq es = errors_sq(P, e, Rc, Rt);
        if (!collided) {
            collided = col_now ? 1 : 0;
            if (!collided && sqrt(es.pos) <= P.max_rd_error && sqrt(es.vel) <= P.max_vd_error &&
                att <= P.max_qd_error && sqrt(es.rot) <= P.max_wd_error)
                success += 1;
        }
        // ---- time and bubble (:193-198), derived from the step counter ----
        step += 1;
        const double t = __ddiv_rn(rint((double)step * P.dt * 1000.0), 1000.0);
        const double bubble = fmax(fma(-(double)step, P.bubble_rate, P.bubble0), P.bubble_min);

        // ---- observation (:205) into the shared staging row ----
        float *o = s_obs + threadIdx.x * RDV_OBS_DIM;
        float ov[RDV_OBS_DIM];
        make_obs(e, obs_scale(P), ov);
#pragma unroll
        for (int k = 0; k < RDV_OBS_DIM; ++k) o[k] = ov[k];

        // ---- done (:355-386): first true condition is the end reason ----
        const bool c0 = !obs_in_box(ov), c1 = t >= P.t_max, c2 = rc_n > bubble, c3 = att > P.max_attitude_error;
        const bool done = c0 || c1 || c2 || c3;
        const int reason = c0 ? 0 : c1 ? 1 : c2 ? 2 : c3 ? 3 : -1;

        // ---- reward (:313-353) ----
        double rew = (P.dt * P.att_coef) * (1.0 - __ddiv_rn(att, P.max_attitude_error));
        rew += fuel;
        if (col_now) rew -= P.dt * P.collision_coef;
        if (rc_n < P.koz_radius && !collided) {
            double pos_err = sqrt(es.pos);
            if (pos_err < P.max_rd_error) {
                rew += P.dt * P.bonus_coef * (2.0 - __ddiv_rn(pos_err, P.max_rd_error));
                if (att < P.max_qd_error) rew += P.dt * P.bonus_coef * (2.0 - __ddiv_rn(att, P.max_qd_error));
            }
        }
        ep_ret += rew;

        // ---- outputs ----
        io.reward[i] = rew;
        io.done[i] = done ? 1 : 0;
        if (io.end_reason) io.end_reason[i] = (int8_t)reason;
        st.steps = 1; st.reward = rew; st.rk_acc = rk_acc; st.rk_rej = rk_rej; st.fail = fail;
        if (done) {
            st.episodes = 1; st.succeeded = success > 0; st.collided = collided; st.end[reason] = 1;
            st.ep_return = ep_ret; st.ep_length = (double)step; st.delta_v = tdv; st.delta_w = tdw;
            if (io.episode_record) {
                double *rec = io.episode_record + RDV_EP_NCOL * i;
                rec[RDV_EP_RETURN] = ep_ret; rec[RDV_EP_LENGTH] = (double)step; rec[RDV_EP_SUCCESS] = (double)success;
                rec[RDV_EP_COLLIDED] = (double)collided; rec[RDV_EP_DELTA_V] = tdv; rec[RDV_EP_DELTA_W] = tdw;
            }
            if (io.terminal_obs) {
                float *to = io.terminal_obs + RDV_OBS_DIM * i;
#pragma unroll
                for (int k = 0; k < RDV_OBS_DIM; ++k) to[k] = ov[k];
            }
            if (io.auto_reset) {
                int slot = atomicAdd(reset_list, 1);
                reset_list[2 + slot] = (int32_t)(i);       // local env index; the reset kernel rewrites state + obs
            }
        }
        store_env(S, i, e);
        S.f64[RDV_TDV * ld + i] = tdv; S.f64[RDV_TDW * ld + i] = tdw; S.f64[RDV_EPRET * ld + i] = ep_ret;
        S.i32[RDV_I_STEP * ld + i] = step; S.i32[RDV_I_SUCCESS * ld + i] = success;
        S.i32[RDV_I_COLLIDED * ld + i] = collided;
    }

    // ---- coalesced observation write-out: the CTA's rows are contiguous in obs[n][17] ----
    __syncthreads();
    {
        const int64_t rows = (n - base) < TPB ? (n - base) : TPB;
        const int total = (int)rows * RDV_OBS_DIM;
        float *dst = io.obs + base * RDV_OBS_DIM;              // base*17*4 B is a multiple of 16 (TPB = 64)
        const int nvec = total >> 2;
        const float4 *src4 = reinterpret_cast<const float4 *>(s_obs);
        float4 *dst4 = reinterpret_cast<float4 *>(dst);
        for (int k = threadIdx.x; k < nvec; k += TPB) dst4[k] = src4[k];
        for (int k = (nvec << 2) + threadIdx.x; k < total; k += TPB) dst[k] = s_obs[k];
    }
    if (io.stats) reduce_stats<TPB / 32>(st, io.stats, s_stats);
}

// ---------------------------------------------------------------------------------
// compacted auto-reset: one thread per finished env (list built by step_kernel).
// reset_list = {count, ticket, idx...}; the last CTA to finish clears count and ticket.
// ---------------------------------------------------------------------------------
RDV_DEV void write_reset_state(const RdvParams &P, const RdvState &S, int64_t i, const EnvRegs &e, int collided,
                               int success, int episode, float *obs_row)
{
    const int64_t ld = S.ld;
    store_env(S, i, e);
    S.f64[RDV_TDV * ld + i] = 0.0; S.f64[RDV_TDW * ld + i] = 0.0; S.f64[RDV_EPRET * ld + i] = 0.0;
    S.i32[RDV_I_STEP * ld + i] = 0; S.i32[RDV_I_SUCCESS * ld + i] = success;
    S.i32[RDV_I_COLLIDED * ld + i] = collided; S.i32[RDV_I_EPISODE * ld + i] = episode;
    if (obs_row) {
        float ov[RDV_OBS_DIM];
        make_obs(e, obs_scale(P), ov);
#pragma unroll
        for (int k = 0; k < RDV_OBS_DIM; ++k) obs_row[k] = ov[k];
    }
}

__global__ void __launch_bounds__(128) reset_list_kernel(const __grid_constant__ RdvParams P, const RdvState S,
                                                         float *obs, int32_t *reset_list, uint64_t seed,
                                                         int64_t env_offset)
{
    const int count = reset_list[0];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        const int64_t i = reset_list[2 + j];
        const int episode = S.i32[RDV_I_EPISODE * S.ld + i] + 1;
        double u[24];
        draw_uniforms(seed, env_offset + i, episode, u);
        EnvRegs e;
        int collided, success;
        reset_env(P, u, e, collided, success);
        write_reset_state(P, S, i, e, collided, success, episode, obs + RDV_OBS_DIM * i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int ticket = atomicAdd(reset_list + 1, 1);
        if (ticket == (int)gridDim.x - 1) { reset_list[0] = 0; reset_list[1] = 0; }
    }
}

// reset() for masked envs (rendezvous_env.py:223-270); uniforms from Philox or from the caller.
__global__ void __launch_bounds__(128) reset_kernel(const __grid_constant__ RdvParams P, const RdvState S,
                                                    const uint8_t *mask, const double *uniforms, float *obs,
                                                    int64_t n, uint64_t seed, int64_t env_offset, int bump)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    const int episode = S.i32[RDV_I_EPISODE * S.ld + i] + (bump ? 1 : 0);
    double u[24];
    if (uniforms) {
#pragma unroll
        for (int k = 0; k < 24; ++k) u[k] = uniforms[24 * i + k];
    } else {
        draw_uniforms(seed, env_offset + i, episode, u);
    }
    EnvRegs e;
    int collided, success;
    reset_env(P, u, e, collided, success);
    write_reset_state(P, S, i, e, collided, success, episode, obs ? obs + RDV_OBS_DIM * i : nullptr);
}

__global__ void __launch_bounds__(128) observe_kernel(const __grid_constant__ RdvParams P, const RdvState S,
                                                      float *obs, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    load_env(S, i, e);
    float ov[RDV_OBS_DIM];
    make_obs(e, obs_scale(P), ov);
#pragma unroll
    for (int k = 0; k < RDV_OBS_DIM; ++k) obs[RDV_OBS_DIM * i + k] = ov[k];
}

// get_errors / check_collision / check_success / dist_from_koz for evaluators
__global__ void __launch_bounds__(128) errors_kernel(const __grid_constant__ RdvParams P, const RdvState S,
                                                     double *errors, uint8_t *collision, uint8_t *success,
                                                     double *koz, int64_t n, int refresh_flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRegs e;
    load_env(S, i, e);
    const Rot Rc = rot_from_quat(e.qc), Rt = rot_from_quat(e.qt);
    const double rc_sq = dot3(e.rc, e.rc), rc_n = sqrt(rc_sq);
    const double att = attitude_error(P, e, Rc, rc_sq);
    const double th = corridor_angle(P, e, Rt, rc_sq);
    const bool col = rc_n < P.koz_radius && th > P.corridor_half_angle;
    ErrSq es = errors_sq(P, e, Rc, Rt);
    const double pe = sqrt(es.pos), ve = sqrt(es.vel), re = sqrt(es.rot);
    const bool within = pe <= P.max_rd_error && ve <= P.max_vd_error && att <= P.max_qd_error && re <= P.max_wd_error;
    if (refresh_flags) {
        S.i32[RDV_I_COLLIDED * S.ld + i] = col ? 1 : 0;
        S.i32[RDV_I_SUCCESS * S.ld + i] = (!col && within) ? 1 : 0;
        return;
    }
    const int sticky = S.i32[RDV_I_COLLIDED * S.ld + i];
    if (errors) { errors[4 * i] = pe; errors[4 * i + 1] = ve; errors[4 * i + 2] = att; errors[4 * i + 3] = re; }
    if (collision) collision[i] = col ? 1 : 0;
    if (success) success[i] = (!sticky && within) ? 1 : 0;
    if (koz) koz[i] = koz_distance(P, rc_n, th);
}

// chaser2lvlh / target2lvlh / lvlh2chaser / lvlh2target for evaluators (rendezvous_env.py:470-508)
__global__ void __launch_bounds__(128) frame_kernel(const double *q, const double *v, double *out, int64_t n,
                                                    int transpose)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qq[4] = {q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
    const double vv[3] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
    double o[3];
    const Rot R = rot_from_quat(qq);
    if (transpose) rot_apply_T(R, vv, o); else rot_apply(R, vv, o);
    out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}

// ---------------------------------------------------------------------------------
// fp32 MLP policy forward, 17 -> 64 -> 64 -> 6 with tanh (SB3 MlpPolicy, main.py:39-48);
// deterministic action = clip(mean, -1, 1) (monte_carlo.py:128-133).  One thread per env, fp32
// FFMA with sequential accumulation over the input index (the order torch's CPU kernel is
// compared against to ~1e-6); weights are staged in shared memory once per CTA.
// ---------------------------------------------------------------------------------
constexpr int PH = 64;
constexpr int PTPB = 128;
__global__ void __launch_bounds__(PTPB) policy_kernel(const RdvPolicy pi, const float *obs, float *actions, int64_t n)
{
    extern __shared__ __align__(16) float sm[];
    float *w0 = sm, *w1 = w0 + PH * 17, *w2 = w1 + PH * PH, *b0 = w2 + 6 * PH, *b1 = b0 + PH, *b2 = b1 + PH;
    float *hid = b2 + 8;                               // [PH][PTPB] hidden activations, column per thread
    for (int k = threadIdx.x; k < PH * 17; k += PTPB) w0[k] = pi.w0[k];
    for (int k = threadIdx.x; k < PH * PH; k += PTPB) w1[k] = pi.w1[k];
    for (int k = threadIdx.x; k < 6 * PH; k += PTPB) w2[k] = pi.w2[k];
    for (int k = threadIdx.x; k < PH; k += PTPB) { b0[k] = pi.b0[k]; b1[k] = pi.b1[k]; }
    if (threadIdx.x < 6) b2[threadIdx.x] = pi.b2[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * PTPB + threadIdx.x;
    if (i >= n) return;
    float x[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) x[k] = obs[17 * i + k];
    float h[PH];
#pragma unroll
    for (int j = 0; j < PH; ++j) {
        float acc = b0[j];
#pragma unroll
        for (int k = 0; k < 17; ++k) acc = fmaf(w0[j * 17 + k], x[k], acc);
        h[j] = tanhf(acc);
    }
#pragma unroll 1
    for (int j = 0; j < PH; ++j) {
        float acc = b1[j];
        const float4 *wr = reinterpret_cast<const float4 *>(w1 + j * PH);
#pragma unroll
        for (int k4 = 0; k4 < PH / 4; ++k4) {
            float4 w = wr[k4];
            acc = fmaf(w.x, h[4 * k4], acc); acc = fmaf(w.y, h[4 * k4 + 1], acc);
            acc = fmaf(w.z, h[4 * k4 + 2], acc); acc = fmaf(w.w, h[4 * k4 + 3], acc);
        }
        hid[j * PTPB + threadIdx.x] = tanhf(acc);
    }
#pragma unroll
    for (int k = 0; k < PH; ++k) h[k] = hid[k * PTPB + threadIdx.x];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float acc = b2[j];
#pragma unroll
        for (int k = 0; k < PH; ++k) acc = fmaf(w2[j * PH + k], h[k], acc);
        actions[6 * i + j] = fminf(1.0f, fmaxf(-1.0f, acc));
    }
}

// DFMA-saturating probe for the measured fp64 peak (16 independent accumulators per thread)
__global__ void fp64_peak_kernel(double *sink, int iters)
{
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double m = 1.0 + 1e-12 * threadIdx.x, c = 1e-15;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace rdv

// =====================================================================================
// C ABI
// =====================================================================================
using namespace rdv;

static int check_state(const RdvState *s, int64_t n)
{
    if (!s || !s->f64 || !s->i32) return RDV_ERR_NULL;
    if (n < 0 || s->ld < n) return RDV_ERR_SIZE;
    if (((uintptr_t)s->f64 & 7) || ((uintptr_t)s->i32 & 3)) return RDV_ERR_ALIGN;
    return RDV_OK;
}
static int launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RDV_OK : RDV_ERR_CUDA;
}

extern "C" {

int rdv_abi_version(void) { return RDV_ABI_VERSION; }
int rdv_sizeof_params(void) { return (int)sizeof(RdvParams); }

const char *rdv_strerror(int status)
{
    switch (status) {
        case RDV_OK: return "ok";
        case RDV_ERR_NULL: return "required pointer is NULL";
        case RDV_ERR_SIZE: return "bad size, leading dimension or enum value";
        case RDV_ERR_ALIGN: return "pointer not aligned for vector access";
        case RDV_ERR_PARAMS: return "invalid environment parameters";
        case RDV_ERR_CUDA: return "CUDA launch/runtime failure (is a B200 visible?)";
        case RDV_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

void rdv_params_default(RdvParams *p)
{
    memset(p, 0, sizeof(*p));
    const double rad = M_PI / 180.0;
    p->rc0[1] = -10.0; p->qc0[0] = 1.0; p->qt0[0] = 1.0;
    p->rc0_range = 1; p->vc0_range = 0.1; p->qc0_range = 1 * rad; p->wc0_range = 0.1 * rad;
    p->qt0_range = 45 * rad; p->wt0_range = 3 * rad;
    p->koz_radius = 5; p->corridor_half_angle = 30 * rad; p->h = 800e3; p->dt = 1; p->t_max = 120;
    p->collision_coef = 0.5; p->bonus_coef = 8; p->fuel_coef = 0.2; p->att_coef = 1;
    const double diag = 1.0 * 1 / 12 * 100 * 2;        // eye * 1/12 * m * (2*1**2), left to right
    for (int i = 0; i < 3; ++i) { p->inertia_c[4 * i] = diag; p->inertia_t[4 * i] = diag; }
    p->integrator = RDV_INTEGRATOR_RK45;
}

static int invert3(const double *m, double *o)
{
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (!(fabs(det) > 0.0)) return -1;
    bool diag = m[1] == 0 && m[2] == 0 && m[3] == 0 && m[5] == 0 && m[6] == 0 && m[7] == 0;
    if (diag) {
        for (int i = 0; i < 9; ++i) o[i] = 0.0;
        o[0] = 1.0 / m[0]; o[4] = 1.0 / m[4]; o[8] = 1.0 / m[8];
        return 0;
    }
    o[0] = c00 / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    o[3] = c01 / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    o[6] = c02 / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    return 0;
}
static int is_isotropic(const double *m, const double *tau)
{
    bool off = m[1] == 0 && m[2] == 0 && m[3] == 0 && m[5] == 0 && m[6] == 0 && m[7] == 0;
    bool eq = m[0] == m[4] && m[4] == m[8];
    bool t0 = !tau || (tau[0] == 0 && tau[1] == 0 && tau[2] == 0);
    return off && eq && t0;
}

int rdv_params_derive(RdvParams *p)
{
    if (!p) return RDV_ERR_NULL;
    const double rad = M_PI / 180.0;
    const double nominal_diag = 1.0 * 1 / 12 * 100 * 2;
    if (invert3(p->inertia_c, p->inv_inertia_c) || invert3(p->inertia_t, p->inv_inertia_t)) return RDV_ERR_PARAMS;
    p->max_delta_v = 10.0 / 100 * 0.5;                                   // rendezvous_env.py:81
    p->max_delta_w = 0.2 / nominal_diag * 0.5;                           // :82
    p->max_axial_distance = sqrt(p->rc0[0] * p->rc0[0] + p->rc0[1] * p->rc0[1] + p->rc0[2] * p->rc0[2]) + 10;  // :85
    p->max_axial_speed = 5; p->max_wc = 10 * rad; p->max_attitude_error = 30 * rad;   // :86-89
    p->max_rd_error = 0.5; p->max_vd_error = 0.1; p->max_qd_error = 5 * rad; p->max_wd_error = 1 * rad;  // :105-108
    p->rd[0] = 0; p->rd[1] = -2; p->rd[2] = 0;                           // :104
    p->capture_axis[0] = 0; p->capture_axis[1] = 1; p->capture_axis[2] = 0;        // :73
    p->corridor_axis[0] = 0; p->corridor_axis[1] = -1; p->corridor_axis[2] = 0;    // :95
    p->bubble0 = p->max_axial_distance;                                  // :113
    p->bubble_rate = 0.5 * p->dt;                                        // :114
    const double rd_n = 2.0;
    p->bubble_min = rd_n + 2 * p->max_rd_error;                          // :115
    const double ro = 6371e3 + p->h;
    p->n = sqrt(3.986004418e14 / (ro * ro * ro));                        // :122-126
    if (!(rd_n < p->koz_radius) || !(rd_n - p->max_rd_error > 0)) return RDV_ERR_PARAMS;   // :155-156
    if (!(p->dt > 0) || !(p->t_max > 0)) return RDV_ERR_PARAMS;
    // t = round(t + dt, 3) each step (:193) is reproduced as round(step * dt, 3), which is the same number only
    // when dt lies on the 1 ms grid the reference rounds to
    if (fabs(p->dt * 1000.0 - rint(p->dt * 1000.0)) > 1e-9 * p->dt * 1000.0) return RDV_ERR_UNSUPPORTED;
    // CW transition matrix, non-zero entries row by row (utils/dynamics.py:40-47)
    const double n = p->n, nt = n * p->dt, s = sin(nt), c = cos(nt);
    double *w = p->cw;
    w[0] = 4 - 3 * c;          w[1] = 1 / n * s;            w[2] = 2 / n * (1 - c);
    w[3] = 6 * (s - nt);       w[4] = 1;                    w[5] = -2 / n * (1 - c);  w[6] = 1 / n * (4 * s - 3 * nt);
    w[7] = c;                  w[8] = 1 / n * s;
    w[9] = 3 * n * s;          w[10] = c;                   w[11] = 2 * s;
    w[12] = -6 * n * (1 - c);  w[13] = -2 * s;              w[14] = 4 * c - 3;
    w[15] = -n * s;            w[16] = c;
    p->max_delta_v_f32 = (float)p->max_delta_v;
    p->fuel_num_f32 = (float)(p->dt * p->fuel_coef);
    p->fuel_den_f32 = (float)(3 * p->max_delta_v);
    p->iso_c = is_isotropic(p->inertia_c, p->torque_c);
    p->iso_t = is_isotropic(p->inertia_t, nullptr);
    if (p->integrator != RDV_INTEGRATOR_RK45 && p->integrator != RDV_INTEGRATOR_CLOSED_FORM) return RDV_ERR_SIZE;
    if (p->integrator == RDV_INTEGRATOR_CLOSED_FORM && !(p->iso_c && p->iso_t)) return RDV_ERR_UNSUPPORTED;
    return RDV_OK;
}

int rdv_step(const RdvParams *p, const RdvState *s, const RdvStepIO *io, int64_t n, uint64_t seed,
             int64_t env_offset, void *cuda_stream)
{
    if (!p || !io || !io->actions || !io->obs || !io->reward || !io->done) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    if (((uintptr_t)io->actions & (io->act_f64 ? 15 : 7)) || ((uintptr_t)io->obs & 15) || ((uintptr_t)io->reward & 7))
        return RDV_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const unsigned grid = (unsigned)((n + TPB - 1) / TPB);
    const bool iso = p->iso_c && p->iso_t;
    const bool closed = p->integrator == RDV_INTEGRATOR_CLOSED_FORM;
    int32_t *reset_list = nullptr;      // {count, ticket, env indices...}, self-clearing
    if (io->auto_reset < 0 || io->auto_reset > 2) return RDV_ERR_SIZE;
    if (io->auto_reset) {
        if (!io->reset_scratch) return RDV_ERR_NULL;
        reset_list = io->reset_scratch;
    }
#define RDV_LAUNCH(ISO_, F64_, CL_) \
    step_kernel<ISO_, F64_, CL_><<<grid, TPB, 0, st>>>(*p, *s, *io, n, reset_list)
    if (closed) { if (io->act_f64) RDV_LAUNCH(true, true, true); else RDV_LAUNCH(true, false, true); }
    else if (iso) { if (io->act_f64) RDV_LAUNCH(true, true, false); else RDV_LAUNCH(true, false, false); }
    else { if (io->act_f64) RDV_LAUNCH(false, true, false); else RDV_LAUNCH(false, false, false); }
#undef RDV_LAUNCH
    rc = launch_status();
    if (rc) return rc;
    if (io->auto_reset == 1) rc = rdv_auto_reset(p, s, io->obs, reset_list, n, seed, env_offset, cuda_stream);
    return rc;
}

int rdv_auto_reset(const RdvParams *p, const RdvState *s, float *obs, int32_t *reset_scratch, int64_t n,
                   uint64_t seed, int64_t env_offset, void *cuda_stream)
{
    if (!p || !obs || !reset_scratch) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    // grid sized for the common case (<= n/8 finished envs per step); grid-stride covers the rest
    unsigned rgrid = (unsigned)((n / 8 + 127) / 128);
    if (rgrid < 1) rgrid = 1;
    if (rgrid > 4096) rgrid = 4096;
    reset_list_kernel<<<rgrid, 128, 0, (cudaStream_t)cuda_stream>>>(*p, *s, obs, reset_scratch, seed, env_offset);
    return launch_status();
}

int rdv_reset(const RdvParams *p, const RdvState *s, const uint8_t *mask, const double *uniforms, float *obs,
              int64_t n, uint64_t seed, int64_t env_offset, int bump_episode, void *cuda_stream)
{
    if (!p) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    reset_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(*p, *s, mask, uniforms, obs, n,
                                                                                      seed, env_offset, bump_episode);
    return launch_status();
}

int rdv_observe(const RdvParams *p, const RdvState *s, float *obs, int64_t n, void *cuda_stream)
{
    if (!p || !obs) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    observe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(*p, *s, obs, n);
    return launch_status();
}

int rdv_errors(const RdvParams *p, const RdvState *s, double *errors, uint8_t *collision, uint8_t *success,
               double *koz, int64_t n, void *cuda_stream)
{
    if (!p) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    errors_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(*p, *s, errors, collision,
                                                                                       success, koz, n, 0);
    return launch_status();
}

int rdv_refresh_flags(const RdvParams *p, const RdvState *s, int64_t n, void *cuda_stream)
{
    if (!p) return RDV_ERR_NULL;
    int rc = check_state(s, n);
    if (rc) return rc;
    if (n == 0) return RDV_OK;
    errors_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(*p, *s, nullptr, nullptr,
                                                                                       nullptr, nullptr, n, 1);
    return launch_status();
}

int rdv_frame_transform(const double *q, const double *v, double *out, int64_t n, int transpose, void *cuda_stream)
{
    if (!q || !v || !out) return RDV_ERR_NULL;
    if (n < 0) return RDV_ERR_SIZE;
    if (n == 0) return RDV_OK;
    frame_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(q, v, out, n, transpose);
    return launch_status();
}

int rdv_policy_forward(const RdvPolicy *pi, const float *obs, float *actions, int64_t n, void *cuda_stream)
{
    if (!pi || !obs || !actions || !pi->w0 || !pi->b0 || !pi->w1 || !pi->b1 || !pi->w2 || !pi->b2) return RDV_ERR_NULL;
    if (pi->hidden != PH) return RDV_ERR_UNSUPPORTED;
    if (n < 0) return RDV_ERR_SIZE;
    if (n == 0) return RDV_OK;
    const size_t smem = sizeof(float) * (PH * 17 + PH * PH + 6 * PH + PH + PH + 8 + PH * PTPB);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    policy_kernel<<<(unsigned)((n + PTPB - 1) / PTPB), PTPB, smem, (cudaStream_t)cuda_stream>>>(*pi, obs, actions, n);
    return launch_status();
}

int rdv_fp64_peak_probe(double *sink, int blocks, int threads, int iters, void *cuda_stream)
{
    if (!sink) return RDV_ERR_NULL;
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return RDV_ERR_SIZE;
    fp64_peak_kernel<<<blocks, threads, 0, (cudaStream_t)cuda_stream>>>(sink, iters);
    return launch_status();
}

}  // extern "C"
