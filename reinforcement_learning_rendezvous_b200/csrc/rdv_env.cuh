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
    s.hi_r = P.max_axial_distance; s.inv_r = 1.0 / (2.0 * P.max_axial_distance);
    s.hi_v = P.max_axial_speed;    s.inv_v = 1.0 / (2.0 * P.max_axial_speed);
    s.hi_w = P.max_wc;             s.inv_w = 1.0 / (2.0 * P.max_wc);
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
// reset() (rendezvous_env.py:223-270).  u[24] are the uniform draws in the reference's order:
// rc dir(3)+mag, vc dir(3)+mag, theta_c, axis_c(3), wc dir(3)+mag, theta_t, axis_t(3),
// wt dir(3)+mag.  Uses IEEE sqrt/div: this path is rare and follows the reference op by op.
// ---------------------------------------------------------------------------------
RDV_DEV void unit_from_cube(const double *u, double o[3])      // utils/general.py:248-254
{
    double v[3] = {fma(2.0, u[0], -1.0), fma(2.0, u[1], -1.0), fma(2.0, u[2], -1.0)};
    double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    o[0] = v[0] / nv; o[1] = v[1] / nv; o[2] = v[2] / nv;
}
RDV_DEV void quat_from_axis_angle(const double ax_in[3], double theta, double q[4])   // quaternions.py:11-27
{
    double na = sqrt(ax_in[0] * ax_in[0] + ax_in[1] * ax_in[1] + ax_in[2] * ax_in[2]);
    double s, c;
    sincos(theta / 2, &s, &c);
    q[0] = c; q[1] = ax_in[0] / na * s; q[2] = ax_in[1] / na * s; q[3] = ax_in[2] / na * s;
    double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] /= nq;
}
RDV_DEV void quat_mul(const double a_in[4], const double b_in[4], double o[4])        // quaternions.py:149-170
{
    double na = sqrt(a_in[0] * a_in[0] + a_in[1] * a_in[1] + a_in[2] * a_in[2] + a_in[3] * a_in[3]);
    double nb = sqrt(b_in[0] * b_in[0] + b_in[1] * b_in[1] + b_in[2] * b_in[2] + b_in[3] * b_in[3]);
    double a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = a_in[k] / na; b[k] = b_in[k] / nb; }
    o[0] = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
    o[1] = a[0] * b[1] + b[0] * a[1] + (a[2] * b[3] - a[3] * b[2]);
    o[2] = a[0] * b[2] + b[0] * a[2] + (a[3] * b[1] - a[1] * b[3]);
    o[3] = a[0] * b[3] + b[0] * a[3] + (a[1] * b[2] - a[2] * b[1]);
}

RDV_DEV void reset_env(const RdvParams &P, const double (&u)[24], EnvRegs &e, int &collided, int &success)
{
    double dir[3], qd[4], tmp[3];
    unit_from_cube(u + 0, dir);
#pragma unroll
    for (int k = 0; k < 3; ++k) e.rc[k] = P.rc0[k] + dir[k] * (P.rc0_range * u[3]);
    unit_from_cube(u + 4, dir);
#pragma unroll
    for (int k = 0; k < 3; ++k) e.vc[k] = P.vc0[k] + dir[k] * (P.vc0_range * u[7]);
    double theta_c = P.qc0_range * u[8];
    unit_from_cube(u + 9, dir);
    quat_from_axis_angle(dir, theta_c, qd);
    quat_mul(qd, P.qc0, e.qc);
    unit_from_cube(u + 12, dir);
#pragma unroll
    for (int k = 0; k < 3; ++k) tmp[k] = P.wc0[k] + dir[k] * (P.wc0_range * u[15]);
    Rot Rc = rot_from_quat(e.qc);
    rot_apply_T(Rc, tmp, e.wc);
    double theta_t = P.qt0_range * u[16];
    unit_from_cube(u + 17, dir);
    quat_from_axis_angle(dir, theta_t, qd);
    quat_mul(qd, P.qt0, e.qt);
    unit_from_cube(u + 20, dir);
#pragma unroll
    for (int k = 0; k < 3; ++k) tmp[k] = P.wt0[k] + dir[k] * (P.wt0_range * u[23]);
    Rot Rt = rot_from_quat(e.qt);
    rot_apply_T(Rt, tmp, e.wt);
    // collided = check_collision(); success = int(check_success())   (:260-261)
    double rc_sq = dot3(e.rc, e.rc), rc_n = sqrt(rc_sq);
    collided = collision_now(P, e, Rt, rc_sq, rc_n) ? 1 : 0;
    success = 0;
    if (!collided) {
        ErrSq s = errors_sq(P, e, Rc, Rt);
        double att = attitude_error(P, e, Rc, rc_sq);
        success = (sqrt(s.pos) <= P.max_rd_error && sqrt(s.vel) <= P.max_vd_error && att <= P.max_qd_error &&
                   sqrt(s.rot) <= P.max_wd_error) ? 1 : 0;
    }
}

RDV_DEV void draw_uniforms(uint64_t seed, int64_t env_id, int32_t episode, double (&u)[24])
{
#pragma unroll 1
    for (uint32_t blk = 0; blk < 12; ++blk) {
        double a, b;
        philox_uniform_pair(seed, env_id, episode, blk, a, b);
        // dynamic index into a register array would spill; select with predicated moves
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (k == (int)blk) { u[2 * k] = a; u[2 * k + 1] = b; }
    }
}

}  // namespace rdv
