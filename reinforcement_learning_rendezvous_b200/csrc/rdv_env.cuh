// rdv_env.cuh -- per-environment logic of RendezvousEnv on top of rdv_math.cuh:
// observation, errors, collision / success checks, KOZ distance, and reset().
#pragma once
#include "rdv_math.cuh"

namespace rdv {

// Registers of one environment.
struct EnvRegs {
    double rc[3], vc[3], qc[4], wc[3], qt[4], wt[3];
};

RDV_DEV void load_env(const RdvState &S, int64_t i, EnvRegs &e)
{
    const double *f = S.f64 + i;
    const int64_t ld = S.ld;
#pragma unroll
    for (int k = 0; k < 3; ++k) e.rc[k] = f[(RDV_RCX + k) * ld];
#pragma unroll
    for (int k = 0; k < 3; ++k) e.vc[k] = f[(RDV_VCX + k) * ld];
#pragma unroll
    for (int k = 0; k < 4; ++k) e.qc[k] = f[(RDV_QCW + k) * ld];
#pragma unroll
    for (int k = 0; k < 3; ++k) e.wc[k] = f[(RDV_WCX + k) * ld];
#pragma unroll
    for (int k = 0; k < 4; ++k) e.qt[k] = f[(RDV_QTW + k) * ld];
#pragma unroll
    for (int k = 0; k < 3; ++k) e.wt[k] = f[(RDV_WTX + k) * ld];
}
RDV_DEV void store_env(const RdvState &S, int64_t i, const EnvRegs &e)
{
    double *f = S.f64 + i;
    const int64_t ld = S.ld;
#pragma unroll
    for (int k = 0; k < 3; ++k) f[(RDV_RCX + k) * ld] = e.rc[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) f[(RDV_VCX + k) * ld] = e.vc[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[(RDV_QCW + k) * ld] = e.qc[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) f[(RDV_WCX + k) * ld] = e.wc[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[(RDV_QTW + k) * ld] = e.qt[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) f[(RDV_WTX + k) * ld] = e.wt[k];
}

// get_observation (rendezvous_env.py:294-311): normalize_value maps [-hi, hi] -> [-1, 1] as
// 2 (v + hi) / (2 hi) - 1, cast to float32.  inv2hi = 1/(2 hi) precomputed per launch.
struct ObsScale { double hi_r, inv_r, hi_v, inv_v, hi_w, inv_w; };
RDV_DEV ObsScale obs_scale(const RdvParams &P)
{
    ObsScale s;
    s.hi_r = P.max_axial_distance; s.inv_r = P.obs_inv_r;
    s.hi_v = P.max_axial_speed;    s.inv_v = P.obs_inv_v;
    s.hi_w = P.max_wc;             s.inv_w = P.obs_inv_w;
    return s;
}
RDV_DEV void make_obs(const EnvRegs &e, const ObsScale &s, float *o /* stride 1 */)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = (float)fma(2.0 * (e.rc[k] + s.hi_r), s.inv_r, -1.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[3 + k] = (float)fma(2.0 * (e.vc[k] + s.hi_v), s.inv_v, -1.0);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[6 + k] = (float)e.qc[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[10 + k] = (float)fma(2.0 * (e.wc[k] + s.hi_w), s.inv_w, -1.0);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[13 + k] = (float)e.qt[k];
}
// gym 0.21 Box.contains on the float32 observation (rendezvous_env.py:367)
RDV_DEV bool obs_in_box(const float *o)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < RDV_OBS_DIM; ++k) ok = ok && (o[k] >= -1.0f) && (o[k] <= 1.0f);
    return ok;
}

// get_attitude_error (rendezvous_env.py:424-434): angle(-rc, R(qc) capture_axis)
RDV_DEV double attitude_error(const RdvParams &P, const EnvRegs &e, const Rot &Rc, double rc_sq)
{
    double cap[3];
    rot_apply(Rc, P.capture_axis, cap);
    return rounded_angle_from(-dot3(e.rc, cap), rc_sq, dot3(cap, cap));
}
// angle(rc, R(qt) corridor_axis), used by check_collision (:388-404) and dist_from_koz (:510-537)
RDV_DEV double corridor_angle(const RdvParams &P, const EnvRegs &e, const Rot &Rt, double rc_sq)
{
    double ax[3];
    rot_apply(Rt, P.corridor_axis, ax);
    return rounded_angle_from(dot3(e.rc, ax), rc_sq, dot3(ax, ax));
}
RDV_DEV bool collision_now(const RdvParams &P, const EnvRegs &e, const Rot &Rt, double rc_sq, double rc_norm)
{
    if (rc_norm < P.koz_radius) return corridor_angle(P, e, Rt, rc_sq) > P.corridor_half_angle;
    return false;
}
// get_errors (rendezvous_env.py:451-468); squares of the three vector errors (att is an angle)
struct ErrSq { double pos, vel, rot; };
RDV_DEV ErrSq errors_sq(const RdvParams &P, const EnvRegs &e, const Rot &Rc, const Rot &Rt)
{
    double wc_l[3], wt_l[3], rd_l[3], vd_l[3], d[3];
    rot_apply(Rc, e.wc, wc_l);
    rot_apply(Rt, e.wt, wt_l);
    rot_apply(Rt, P.rd, rd_l);
    cross3(wt_l, rd_l, vd_l);
    ErrSq s;
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = e.rc[k] - rd_l[k];
    s.pos = dot3(d, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = e.vc[k] - vd_l[k];
    s.vel = dot3(d, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = wc_l[k] - wt_l[k];
    s.rot = dot3(d, d);
    return s;
}

// dist_from_koz (rendezvous_env.py:510-537)
RDV_DEV double koz_distance(const RdvParams &P, double r, double th)
{
    const double rk = P.koz_radius, thc = P.corridor_half_angle;
    if (r < rk) {
        if (th >= thc) return -fmin(rk - r, r * sin(fmin(th - thc, 1.5707963267948966)));
        return r * sin(thc - th);
    }
    if (th >= thc) return r - rk;
    double d_rad = r - rk * cos(thc - th), d_tan = rk * sin(thc - th);
    return sqrt(fma(d_rad, d_rad, d_tan * d_tan));
}

// ---------------------------------------------------------------------------------
// reset() (rendezvous_env.py:223-270), computed by a TEAM of 8 adjacent lanes per environment.
//
// The reference consumes 24 uniform draws in a fixed order: rc dir(3)+mag, vc dir(3)+mag, theta_c,
// axis_c(3), wc dir(3)+mag, theta_t, axis_t(3), wt dir(3)+mag.  Lane `sub` (0..5) of a team owns
// quantity `sub` of (rc, vc, qc, wc, qt, wt) and the four draws u[4 sub .. 4 sub + 3], which are
// exactly Philox blocks 2 sub and 2 sub + 1 of the (seed; env id, episode) stream -- so the six
// quantities are generated concurrently, the two body rates wait only for their quaternion
// (lvlh2chaser / lvlh2target, :256, :258), and lane 0 evaluates the collided / success flags
// (:260-261) while lanes 1..7 store the state and the observation.  One serial reset (~5 k
// dependent instructions in the one-thread form) becomes ~0.7 k.
//
// `row` is a team-private scratch of RDV_TEAM_ROW doubles in shared memory.  Every lane of the warp
// must call the function (teams with valid == false compute on env i but store nothing).
// ---------------------------------------------------------------------------------
constexpr int RDV_TEAM = 8;
constexpr int RDV_TEAM_ROW = 24;

// Core: after the call row[0..19] holds the new state (rc vc qc wc qt wt) and, on lane 0's return,
// row[20] / row[21] the collided / success flags.  Ends with a __syncwarp().
RDV_DEV void team_reset_core(const RdvParams &P, uint64_t seed, int64_t env_id, int episode,
                             const double *uniforms /* nullable [24] of this env */, double *row)
{
    const int sub = threadIdx.x & (RDV_TEAM - 1);

    // ---- phase 1: own quantity from own four draws ----
    double val[4] = {0.0, 0.0, 0.0, 0.0};
    if (sub < 6) {
        double u0, u1, u2, u3;
        if (uniforms) {
            u0 = uniforms[4 * sub]; u1 = uniforms[4 * sub + 1]; u2 = uniforms[4 * sub + 2]; u3 = uniforms[4 * sub + 3];
        } else {
            philox_uniform_pair(seed, env_id, episode, 2 * sub, u0, u1);
            philox_uniform_pair(seed, env_id, episode, 2 * sub + 1, u2, u3);
        }
        const bool quat = (sub == 2) || (sub == 4);
        // random_unit_vector (utils/general.py:248-254): normalised U(-1,1)^3
        const double a = quat ? u1 : u0, b = quat ? u2 : u1, c = quat ? u3 : u2, m = quat ? u0 : u3;
        const double v[3] = {fma(2.0, a, -1.0), fma(2.0, b, -1.0), fma(2.0, c, -1.0)};
        const double rn = fast_rsqrt(dot3(v, v));
        const double dir[3] = {v[0] * rn, v[1] * rn, v[2] * rn};
        const double range = sub == 0 ? P.rc0_range : sub == 1 ? P.vc0_range : sub == 2 ? P.qc0_range
                           : sub == 3 ? P.wc0_range : sub == 4 ? P.qt0_range : P.wt0_range;
        const double mag = range * m;                          // np.random.uniform(0, range)
        if (!quat) {
            const double *nom = sub == 0 ? P.rc0 : sub == 1 ? P.vc0 : sub == 3 ? P.wc0 : P.wt0;
#pragma unroll
            for (int k = 0; k < 3; ++k) val[k] = fma(dir[k], mag, nom[k]);
        } else {
            // rot2quat (quaternions.py:11-27) then quat_product(dev, nominal) (:149-170); both normalise
            double sn, cs;
            sincos(0.5 * mag, &sn, &cs);
            const double ra = fast_rsqrt(dot3(dir, dir));
            double qa[4] = {cs, dir[0] * ra * sn, dir[1] * ra * sn, dir[2] * ra * sn};
            const double rq = fast_rsqrt(dot4(qa, qa));
            const double *qn = sub == 2 ? P.qc0 : P.qt0;
            const double rb = fast_rsqrt(dot4(qn, qn));
            const double qb[4] = {qn[0] * rb, qn[1] * rb, qn[2] * rb, qn[3] * rb};
#pragma unroll
            for (int k = 0; k < 4; ++k) qa[k] *= rq;
            val[0] = qa[0] * qb[0] - (qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3]);
            val[1] = qa[0] * qb[1] + qb[0] * qa[1] + (qa[2] * qb[3] - qa[3] * qb[2]);
            val[2] = qa[0] * qb[2] + qb[0] * qa[2] + (qa[3] * qb[1] - qa[1] * qb[3]);
            val[3] = qa[0] * qb[3] + qb[0] * qa[3] + (qa[1] * qb[2] - qa[2] * qb[1]);
        }
    }
    // ---- phase 2: the quaternions go to the scratch row; the rate lanes rotate into their body frame ----
    constexpr int OFF[6] = {RDV_RCX, RDV_VCX, RDV_QCW, RDV_WCX, RDV_QTW, RDV_WTX};
    if (sub == 2 || sub == 4) {
        const int o = sub == 2 ? RDV_QCW : RDV_QTW;
#pragma unroll
        for (int k = 0; k < 4; ++k) row[o + k] = val[k];
    }
    __syncwarp();
    if (sub == 3 || sub == 5) {
        const int o = sub == 3 ? RDV_QCW : RDV_QTW;
        const double q[4] = {row[o], row[o + 1], row[o + 2], row[o + 3]};
        const Rot R = rot_from_quat(q);
        double w[3];
        rot_apply_T(R, val, w);
        val[0] = w[0]; val[1] = w[1]; val[2] = w[2];
    }
    if (sub == 0 || sub == 1 || sub == 3 || sub == 5) {
        const int o = sub == 0 ? OFF[0] : sub == 1 ? OFF[1] : sub == 3 ? OFF[3] : OFF[5];
#pragma unroll
        for (int k = 0; k < 3; ++k) row[o + k] = val[k];
    }
    __syncwarp();

    // ---- phase 3: lane 0 -> collided / success flags (:260-261) ----
    if (sub == 0) {
        EnvRegs e;
#pragma unroll
        for (int k = 0; k < 3; ++k) { e.rc[k] = row[RDV_RCX + k]; e.vc[k] = row[RDV_VCX + k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { e.qc[k] = row[RDV_QCW + k]; e.qt[k] = row[RDV_QTW + k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { e.wc[k] = row[RDV_WCX + k]; e.wt[k] = row[RDV_WTX + k]; }
        // the nominal start is ~10 m out: beyond near_sq neither a collision nor a success is possible and
        // the rotation matrices, the corridor angle and the error vector are never formed
        const double rc_sq = dot3(e.rc, e.rc);
        int collided = 0, success = 0;
        if (rc_sq < P.near_sq) {
            const Rot Rc = rot_from_quat(e.qc), Rt = rot_from_quat(e.qt);
            collided = (rc_sq < P.koz_radius_sq && corridor_angle(P, e, Rt, rc_sq) > P.corridor_half_angle) ? 1 : 0;
            if (!collided) {
                const ErrSq es = errors_sq(P, e, Rc, Rt);
                if (es.pos <= P.max_rd_error_sq && es.vel <= P.max_vd_error_sq && es.rot <= P.max_wd_error_sq)
                    success = attitude_error(P, e, Rc, rc_sq) <= P.max_qd_error ? 1 : 0;
            }
        }
        row[20] = (double)collided;
        row[21] = (double)success;
    }
    __syncwarp();
}

// observation slot k (< 17) of a state value x: get_observation (:294-311) -- rc/20, vc/5, qc, wc/rad(10), qt
RDV_DEV float obs_of_state_row(const ObsScale &sc, int k, double x)
{
    if (k < RDV_QCW) {
        const bool pos = k < RDV_VCX;
        return (float)fma(2.0 * (x + (pos ? sc.hi_r : sc.hi_v)), pos ? sc.inv_r : sc.inv_v, -1.0);
    }
    if (k >= RDV_WCX && k < RDV_QTW) return (float)fma(2.0 * (x + sc.hi_w), sc.inv_w, -1.0);
    return (float)x;
}

// Reset of env i of the state arrays: team_reset_core, then the 8 lanes store the state rows, the counters
// and the observation.
RDV_DEV void team_reset(const RdvParams &P, const RdvState &S, uint64_t seed, int64_t env_id, int64_t i,
                        bool valid, int bump, const double *uniforms /* nullable [24] of this env */,
                        double *row, float *obs_row /* staging row (shared) or global row */)
{
    const int sub = threadIdx.x & (RDV_TEAM - 1);
    const int64_t ld = S.ld;
    const int episode = S.i32[RDV_I_EPISODE * ld + i] + bump;
    team_reset_core(P, seed, env_id, episode, uniforms, row);
    if (!valid) return;
    if (sub == 0) {
        S.f64[RDV_TDV * ld + i] = 0.0; S.f64[RDV_TDW * ld + i] = 0.0; S.f64[RDV_EPRET * ld + i] = 0.0;
        S.i32[RDV_I_STEP * ld + i] = 0; S.i32[RDV_I_SUCCESS * ld + i] = (int)row[21];
        S.i32[RDV_I_COLLIDED * ld + i] = (int)row[20]; S.i32[RDV_I_EPISODE * ld + i] = episode;
    } else {
        const ObsScale sc = obs_scale(P);
        for (int k = sub - 1; k < RDV_TDV; k += RDV_TEAM - 1) {
            const double x = row[k];
            S.f64[k * ld + i] = x;
            if (obs_row && k < RDV_OBS_DIM) obs_row[k] = obs_of_state_row(sc, k, x);
        }
    }
}

}  // namespace rdv
