// rdv_step.cuh -- one RendezvousEnv.step (rendezvous_env.py:160-221) on register-resident state, as the
// building blocks shared by the thread-per-env step kernel and the fused multi-step rollout kernel:
//   ingest_action_*  :168-173, :201-202, :333   action -> impulses, fuel term, delta-v / delta-w totals
//   env_advance      :172-184                   CW translation + both attitude propagations
//   env_evaluate     :186-211                   latch, time, bubble, observation, done, reward
#pragma once
#include "rdv_env.cuh"

namespace rdv {

__constant__ double c_zero3[3] = {0.0, 0.0, 0.0};     // the target carries no torque (rendezvous_env.py:585)

// Per-env quantities that live next to the 20 state numbers.
struct EnvCounters {
    double tdv, tdw, ep_ret;          // total_delta_v, total_delta_w, running episode return
    int step, success, collided, episode;
};

struct ActionTerms { double dvb[3], dw[3], fuel; };

RDV_DEV void ingest_action_f64(const RdvParams &P, const double (&a)[6], EnvCounters &c, ActionTerms &t)
{
    t.dvb[0] = a[0] * P.max_delta_v; t.dvb[1] = a[1] * P.max_delta_v; t.dvb[2] = a[2] * P.max_delta_v;
    t.dw[0] = a[3] * P.max_delta_w; t.dw[1] = a[4] * P.max_delta_w; t.dw[2] = a[5] * P.max_delta_w;
    const double sv = fabs(a[0]) + fabs(a[1]) + fabs(a[2]);
    const double sw = fabs(a[3]) + fabs(a[4]) + fabs(a[5]);
    c.tdv += sv * P.max_delta_v;
    c.tdw += sw * P.max_delta_w;
    t.fuel = P.fuel_scale * sv;
}

// float32 actions follow NumPy-2 promotion: delta_v, total_delta_v and the fuel term are
// rounded in fp32; delta_w and total_delta_w are fp64.
RDV_DEV void ingest_action_f32(const RdvParams &P, const float (&a)[6], EnvCounters &c, ActionTerms &t)
{
    t.dvb[0] = (double)__fmul_rn(a[0], P.max_delta_v_f32);
    t.dvb[1] = (double)__fmul_rn(a[1], P.max_delta_v_f32);
    t.dvb[2] = (double)__fmul_rn(a[2], P.max_delta_v_f32);
    t.dw[0] = (double)a[3] * P.max_delta_w; t.dw[1] = (double)a[4] * P.max_delta_w; t.dw[2] = (double)a[5] * P.max_delta_w;
    const float sv = __fadd_rn(__fadd_rn(fabsf(a[0]), fabsf(a[1])), fabsf(a[2]));
    const float sw = __fadd_rn(__fadd_rn(fabsf(a[3]), fabsf(a[4])), fabsf(a[5]));
    c.tdv = (double)__fadd_rn((float)c.tdv, __fmul_rn(sv, P.max_delta_v_f32));
    c.tdw += (double)sw * P.max_delta_w;
    t.fuel = (double)__fdiv_rn(__fmul_rn(P.fuel_num_f32, sv), P.fuel_den_f32);
}

// Translation (impulse rotated by the OLD chaser attitude, then the CW transition; dynamics.py:24-55) and the
// two attitude propagations (:552-604).  LOCKSTEP: the two isotropic solves advance interleaved stage by stage
// (rk45_iso_plane_pair, 2x ILP) -- the right choice with <= 2 warps per SM sub-partition; otherwise one solve after
// the other through a single copy of the solver (rk45_iso_plane, 128 registers), which wins once 3-4 warps per
// sub-partition hide the latency instead.
template <bool ISO, bool CLOSED, bool LOCKSTEP>
RDV_DEV void env_advance(const RdvParams &P, EnvRegs &e, const ActionTerms &t, int &rk_acc, int &rk_rej, int &fail)
{
    {
        const Rot Rc_old = rot_from_quat(e.qc);
        double dv[3];
        rot_apply(Rc_old, t.dvb, dv);
        const double r0 = e.rc[0], r1 = e.rc[1], r2 = e.rc[2];
        const double v0 = e.vc[0] + dv[0], v1 = e.vc[1] + dv[1], v2 = e.vc[2] + dv[2];
        const double *c = P.cw;
        e.rc[0] = fma(c[2], v1, fma(c[1], v0, c[0] * r0));
        e.rc[1] = fma(c[6], v1, fma(c[5], v0, fma(c[3], r0, c[4] * r1)));
        e.rc[2] = fma(c[8], v2, c[7] * r2);
        e.vc[0] = fma(c[11], v1, fma(c[10], v0, c[9] * r0));
        e.vc[1] = fma(c[14], v1, fma(c[13], v0, c[12] * r0));
        e.vc[2] = fma(c[16], v2, c[15] * r2);
    }
    double y[7] = {e.qc[0], e.qc[1], e.qc[2], e.qc[3], e.wc[0] + t.dw[0], e.wc[1] + t.dw[1], e.wc[2] + t.dw[2]};
    double z[7] = {e.qt[0], e.qt[1], e.qt[2], e.qt[3], e.wt[0], e.wt[1], e.wt[2]};
    if (CLOSED) {
        closed_form_attitude(y, P.dt);
        closed_form_attitude(z, P.dt);
    } else {
        BodyConst bc, bt;
        bc.I = P.inertia_c; bc.Iinv = P.inv_inertia_c; bc.tau = P.torque_c;
        bt.I = P.inertia_t; bt.Iinv = P.inv_inertia_t; bt.tau = c_zero3;
        if (ISO && LOCKSTEP) {
            const int k = rk45_iso_plane_pair(y, z, P.dt, rk_rej);
            if (k < 0) fail = 1; else rk_acc += k;
        } else if (ISO) {
            // one solve after the other through a single copy of the solver code (bounded registers)
#pragma unroll 1
            for (int body = 0; body < 2; ++body) {
                const int k = rk45_iso_plane(y, P.dt, rk_rej);
                if (k < 0) fail = 1; else rk_acc += k;
#pragma unroll
                for (int j = 0; j < 7; ++j) { const double tmp = y[j]; y[j] = z[j]; z[j] = tmp; }
            }
        } else {
            const int kc = rk45_attitude<false>(y, P.dt, bc, rk_rej);
            const int kt = rk45_attitude<false>(z, P.dt, bt, rk_rej);
            if (kc < 0 || kt < 0) fail = 1; else rk_acc += kc + kt;
        }
    }
    const double ry = fast_rsqrt(dot4(y, y)), rz = fast_rsqrt(dot4(z, z));       // q / |q|  (:574-575, :601-602)
#pragma unroll
    for (int k = 0; k < 4; ++k) { e.qc[k] = y[k] * ry; e.qt[k] = z[k] * rz; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { e.wc[k] = y[4 + k]; e.wt[k] = z[4 + k]; }
}

struct StepResult { double rew; int done, reason; };

// Everything after the propagation: collision / success latch (:186-190), time and bubble (:193-198),
// observation (:205), done + end reason (:355-386), reward (:313-353).  Updates the counters.
RDV_DEV StepResult env_evaluate(const RdvParams &P, const EnvRegs &e, const double fuel, EnvCounters &c,
                                float (&ov)[RDV_OBS_DIM])
{
    const Rot Rc = rot_from_quat(e.qc);
    const double rc_sq = dot3(e.rc, e.rc);
    const double att = attitude_error(P, e, Rc, rc_sq);
    // The target-relative quantities (corridor angle, position / velocity / rate errors) only matter near the
    // target: a collision needs |rc| < koz, and success (:406-422) or the reward bonus (:348-351) need a position
    // error <= max_rd_error, impossible unless | |rc| - |rd| | <= max_rd_error (|R(qt) rd| = |rd|).  Far away
    // -- every step of a random-action episode -- the target rotation matrix is never formed.
    bool col_now = false;
    ErrSq es;
    es.pos = es.vel = es.rot = 1.0e300;
    if (rc_sq < P.near_sq) {
        const Rot Rt = rot_from_quat(e.qt);
        col_now = rc_sq < P.koz_radius_sq && corridor_angle(P, e, Rt, rc_sq) > P.corridor_half_angle;
        es = errors_sq(P, e, Rc, Rt);
    }
    if (!c.collided) {
        c.collided = col_now ? 1 : 0;
        if (!c.collided && es.pos <= P.max_rd_error_sq && es.vel <= P.max_vd_error_sq && att <= P.max_qd_error &&
            es.rot <= P.max_wd_error_sq)
            c.success += 1;
    }
    c.step += 1;
    const double bubble = fmax(fma(-(double)c.step, P.bubble_rate, P.bubble0), P.bubble_min);
    make_obs(e, obs_scale(P), ov);
    const bool c0 = !obs_in_box(ov), c1 = c.step >= P.done_steps, c2 = rc_sq > bubble * bubble,
               c3 = att > P.max_attitude_error;
    StepResult r;
    r.done = (c0 || c1 || c2 || c3) ? 1 : 0;
    r.reason = c0 ? 0 : c1 ? 1 : c2 ? 2 : c3 ? 3 : -1;
    double rew = P.att_scale * fma(-att, P.inv_max_attitude_error, 1.0);
    rew += fuel;
    if (col_now) rew -= P.collision_scale;
    if (rc_sq < P.koz_radius_sq && !c.collided && es.pos < P.max_rd_error_sq) {
        rew += P.bonus_scale * fma(-fast_sqrt(es.pos), P.inv_max_rd_error, 2.0);
        if (att < P.max_qd_error) rew += P.bonus_scale * fma(-att, P.inv_max_qd_error, 2.0);
    }
    r.rew = rew;
    c.ep_ret += rew;
    return r;
}

RDV_DEV void load_counters(const RdvState &S, int64_t i, EnvCounters &c)
{
    const int64_t ld = S.ld;
    c.tdv = S.f64[RDV_TDV * ld + i]; c.tdw = S.f64[RDV_TDW * ld + i]; c.ep_ret = S.f64[RDV_EPRET * ld + i];
    c.step = S.i32[RDV_I_STEP * ld + i]; c.success = S.i32[RDV_I_SUCCESS * ld + i];
    c.collided = S.i32[RDV_I_COLLIDED * ld + i]; c.episode = S.i32[RDV_I_EPISODE * ld + i];
}
RDV_DEV void store_counters(const RdvState &S, int64_t i, const EnvCounters &c)
{
    const int64_t ld = S.ld;
    S.f64[RDV_TDV * ld + i] = c.tdv; S.f64[RDV_TDW * ld + i] = c.tdw; S.f64[RDV_EPRET * ld + i] = c.ep_ret;
    S.i32[RDV_I_STEP * ld + i] = c.step; S.i32[RDV_I_SUCCESS * ld + i] = c.success;
    S.i32[RDV_I_COLLIDED * ld + i] = c.collided; S.i32[RDV_I_EPISODE * ld + i] = c.episode;
}

// U(-1,1) actions from the Philox stream (action_seed; env id, step index): blocks 0x40000000 | {0,1,2} of the
// counter space, disjoint from the reset() draws (blocks 0..11 keyed by the episode index).
RDV_DEV void philox_actions(uint64_t action_seed, int64_t env_id, int64_t step_index, double (&a)[6])
{
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        uint32_t c[4] = {(uint32_t)env_id, (uint32_t)((uint64_t)env_id >> 32), (uint32_t)step_index,
                         0x40000000u | blk | (((uint32_t)((uint64_t)step_index >> 32) & 0x00FFFFFFu) << 4)};
        philox4x32_10(c, (uint32_t)action_seed, (uint32_t)(action_seed >> 32));
        const double u0 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) * (1.0 / 9007199254740992.0);
        const double u1 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) * (1.0 / 9007199254740992.0);
        a[2 * blk] = fma(2.0, u0, -1.0);
        a[2 * blk + 1] = fma(2.0, u1, -1.0);
    }
}

}  // namespace rdv
