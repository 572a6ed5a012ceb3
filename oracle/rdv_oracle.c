/* TEST INFRASTRUCTURE ONLY -- plain-C CPU oracle for the RendezvousEnv hot path.
 *
 * A scalar, one-environment-at-a-time restatement of the reference algorithm,
 * compiled with `gcc -O2 -ffp-contract=off` (see oracle/Makefile).  It is
 * the CHECKER the CUDA kernels are compared against at sizes the Python oracle
 * (oracle/rdv_oracle.py) is too slow for, and bench.py's `cpu_baseline` "port"
 * figure.  The product library never links, loads or calls it.
 *
 * Parity status: PINNED -- tests/test_c_oracle.py checks it against the golden
 * vectors in tests/golden/ that oracle/make_golden.py froze from the unmodified
 * reference (same states/actions -> same trajectories to <= 1e-12 relative;
 * rewards and all flags exact).
 *
 * Citations are reference file:line (relative to /root/reference) or, for the
 * un-vendored third-party integrator, SciPy 1.18.1 scipy/integrate/_ivp/ (rk.py, common.py, ivp.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_NSTATE 20   /* rc3 vc3 qc4 wc3 qt4 wt3 */
#define ORC_NAUX 4      /* total_delta_v, total_delta_w, t, bubble_radius */

typedef struct OrcParams {
    /* constructor arguments: rendezvous_env.py:17-70 */
    double rc0[3], vc0[3], qc0[4], wc0[3], qt0[4], wt0[3];
    double rc0_range, vc0_range, qc0_range, wc0_range, qt0_range, wt0_range;
    double koz_radius, corridor_half_angle, h, dt, t_max;
    /* reward_kwargs: rendezvous_env.py:313 */
    double collision_coef, bonus_coef, fuel_coef, att_coef;
    /* rigid bodies: rendezvous_env.py:75-101 (full matrices, row-major) */
    double inertia_c[9], inv_inertia_c[9], inertia_t[9], inv_inertia_t[9];
    double torque_c[3];      /* held chaser torque; the env always passes zeros (:181) */
    /* derived: rendezvous_env.py:81-126 */
    double max_delta_v, max_delta_w, max_axial_distance, max_axial_speed, max_wc;
    double max_attitude_error, max_rd_error, max_vd_error, max_qd_error, max_wd_error;
    double rd[3], capture_axis[3], corridor_axis[3];
    double bubble0, bubble_rate, bubble_min, n;
    int dt_is_integer;       /* python int dt keeps t an int; irrelevant numerically */
} OrcParams;

static double norm_n(const double *x, int n)
{   /* np.linalg.norm: sqrt(x.dot(x)) */
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

void orc_params_default(OrcParams *p)
{
    memset(p, 0, sizeof(*p));
    const double rad = M_PI / 180.0;
    p->rc0[1] = -10.0;
    p->qc0[0] = 1.0;
    p->qt0[0] = 1.0;
    p->rc0_range = 1; p->vc0_range = 0.1; p->qc0_range = 1 * rad; p->wc0_range = 0.1 * rad;
    p->qt0_range = 45 * rad; p->wt0_range = 3 * rad;
    p->koz_radius = 5; p->corridor_half_angle = 30 * rad; p->h = 800e3; p->dt = 1; p->t_max = 120;
    p->collision_coef = 0.5; p->bonus_coef = 8; p->fuel_coef = 0.2; p->att_coef = 1;
    /* np.array(eye) * 1/12 * m * (2*1**2), evaluated left to right: rendezvous_env.py:75-79 */
    double diag = 1.0 * 1 / 12 * 100 * 2;
    for (int i = 0; i < 3; ++i) {
        p->inertia_c[4 * i] = diag; p->inertia_t[4 * i] = diag;
        p->inv_inertia_c[4 * i] = 1.0 / diag; p->inv_inertia_t[4 * i] = 1.0 / diag;
    }
}

void orc_params_derive(OrcParams *p)
{
    const double rad = M_PI / 180.0;
    double nominal_diag = 1.0 * 1 / 12 * 100 * 2;
    p->max_delta_v = 10.0 / 100 * 0.5;                   /* :81 */
    p->max_delta_w = 0.2 / nominal_diag * 0.5;           /* :82 */
    p->max_axial_distance = norm_n(p->rc0, 3) + 10;      /* :85 */
    p->max_axial_speed = 5;
    p->max_wc = 10 * rad;
    p->max_attitude_error = 30 * rad;
    p->max_rd_error = 0.5; p->max_vd_error = 0.1; p->max_qd_error = 5 * rad; p->max_wd_error = 1 * rad;
    p->rd[0] = 0; p->rd[1] = -2; p->rd[2] = 0;
    p->capture_axis[0] = 0; p->capture_axis[1] = 1; p->capture_axis[2] = 0;
    p->corridor_axis[0] = 0; p->corridor_axis[1] = -1; p->corridor_axis[2] = 0;
    p->bubble0 = p->max_axial_distance;                  /* :113 */
    p->bubble_rate = 0.5 * p->dt;                        /* :114 */
    p->bubble_min = norm_n(p->rd, 3) + 2 * p->max_rd_error;
    double ro = 6371e3 + p->h;
    p->n = sqrt(3.986004418e14 / (ro * ro * ro));        /* :126 (ro**3 == ro*ro*ro in numpy/py) */
}

int orc_sizeof_params(void) { return (int)sizeof(OrcParams); }

/* ---- quaternion helpers ---------------------------------------------------- */
static void quat2mat(const double *q_in, double m[9])
{   /* utils/quaternions.py:48-68 */
    double nq = norm_n(q_in, 4);
    double w = q_in[0] / nq, x = q_in[1] / nq, y = q_in[2] / nq, z = q_in[3] / nq;
    m[0] = 2 * (w * w + x * x) - 1; m[1] = 2 * (x * y - w * z);     m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z);     m[4] = 2 * (w * w + y * y) - 1; m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y);     m[7] = 2 * (y * z + w * x);     m[8] = 2 * (w * w + z * z) - 1;
}
static void matvec(const double m[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = m[3 * i] * v[0] + m[3 * i + 1] * v[1] + m[3 * i + 2] * v[2];
}
static void matTvec(const double m[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = m[i] * v[0] + m[3 + i] * v[1] + m[6 + i] * v[2];
}
static void body2lvlh(const double *q, const double v[3], double o[3])
{   /* rendezvous_env.py:490-508 */
    double m[9]; quat2mat(q, m); matvec(m, v, o);
}
static void lvlh2body(const double *q, const double v[3], double o[3])
{   /* rendezvous_env.py:470-488 */
    double m[9]; quat2mat(q, m); matTvec(m, v, o);
}
static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double rounded_angle(const double a[3], const double b[3])
{   /* utils/general.py:163-181; round(x,5) on np.float64 == rint(x*1e5)/1e5 */
    double c = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (norm_n(a, 3) * norm_n(b, 3));
    return acos(rint(c * 1e5) / 1e5);
}

/* ---- attitude ODE: utils/dynamics.py:93-175 ---------------------------------- */
typedef struct { const double *I, *Iinv, *tau; } Body;

static void rhs(const double y[7], const Body *b, double f[7])
{
    double q[4];
    double nq = norm_n(y, 4);
    for (int i = 0; i < 4; ++i) q[i] = y[i] / nq;       /* :108 */
    nq = norm_n(q, 4);
    for (int i = 0; i < 4; ++i) q[i] = q[i] / nq;       /* :134 */
    double w1 = y[4], w2 = y[5], w3 = y[6];
    /* 0.5 * (skew @ q), rows as at :137-142 (zeros included in the dot products) */
    f[0] = 0.5 * (0 * q[0] + -w1 * q[1] + -w2 * q[2] + -w3 * q[3]);
    f[1] = 0.5 * (w1 * q[0] + 0 * q[1] + w3 * q[2] + -w2 * q[3]);
    f[2] = 0.5 * (w2 * q[0] + -w3 * q[1] + 0 * q[2] + w1 * q[3]);
    f[3] = 0.5 * (w3 * q[0] + w2 * q[1] + -w1 * q[2] + 0 * q[3]);
    double Lw[3], cr[3], rhs3[3];
    matvec(b->I, y + 4, Lw);                              /* :169 */
    cross3(y + 4, Lw, cr);                                /* :170 */
    for (int i = 0; i < 3; ++i) rhs3[i] = b->tau[i] - cr[i];
    matvec(b->Iinv, rhs3, f + 4);                         /* :171 */
}

/* RK_C (stage times) is not needed: the right-hand side is autonomous. */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525,
                               1.0 / 40};
static const double RK_P[7][4] = {
    {1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

static double rms7(const double *x) { return norm_n(x, 7) / sqrt(7.0); }   /* common.py:63-65 */

/* solve_ivp(RK45, (0, dt), y, t_eval=[dt], rtol=1e-7, atol=1e-6).y ; returns
 * number of accepted steps (>0) or -1 on TOO_SMALL_STEP / non-finite error. */
static int rk45(double y[7], double t_bound, const Body *b, int *n_reject)
{
    const double rtol = 1e-7, atol = 1e-6;                /* rendezvous_env.py:567-568 */
    double f[7], tmp[7], scale[7];
    double t = 0.0;
    rhs(y, b, f);
    /* select_initial_step: common.py:68-134 with order = 4 */
    for (int i = 0; i < 7; ++i) scale[i] = atol + fabs(y[i]) * rtol;
    for (int i = 0; i < 7; ++i) tmp[i] = y[i] / scale[i];
    double d0 = rms7(tmp);
    for (int i = 0; i < 7; ++i) tmp[i] = f[i] / scale[i];
    double d1 = rms7(tmp);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (t_bound < h0) h0 = t_bound;
    double y1[7], f1[7];
    for (int i = 0; i < 7; ++i) y1[i] = y[i] + h0 * 1.0 * f[i];
    rhs(y1, b, f1);
    for (int i = 0; i < 7; ++i) tmp[i] = (f1[i] - f[i]) / scale[i];
    double d2 = rms7(tmp) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5);
    double h_abs = fmin(fmin(100 * h0, h1), t_bound);

    double K[7][7];
    int accepted = 0;
    for (;;) {
        /* _step_impl: rk.py:111-179 */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int rejected = 0;
        double t_new, h, y_new[7];
        for (;;) {
            if (h_abs < min_step) return -1;
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            /* rk_step: rk.py:14-69 */
            for (int i = 0; i < 7; ++i) K[0][i] = f[i];
            for (int s = 1; s < 6; ++s) {
                double ys[7];
                for (int i = 0; i < 7; ++i) {
                    double acc = 0.0;
                    for (int j = 0; j < s; ++j) acc += K[j][i] * RK_A[s][j];
                    ys[i] = y[i] + acc * h;
                }
                rhs(ys, b, K[s]);
            }
            for (int i = 0; i < 7; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
                y_new[i] = y[i] + h * acc;
            }
            rhs(y_new, b, K[6]);
            double e[7];
            for (int i = 0; i < 7; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_E[j];
                double sc = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                e[i] = acc * h / sc;
            }
            double err = rms7(e);
            if (!(err == err) || isinf(err)) return -1;   /* reference would spin down to TOO_SMALL_STEP */
            if (err < 1) {
                double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
            rejected = 1;
            if (n_reject) ++*n_reject;
        }
        ++accepted;
        double t_old = t;
        t = t_new;
        if (t - t_bound >= 0) {
            /* dense output at t_eval=[t_bound]: ivp.py:710-728, rk.py:715-737 */
            double hh = t - t_old;
            double x = (t_bound - t_old) / hh;
            double pw[4] = {x, x * x, x * x * x, x * x * x * x};
            for (int i = 0; i < 7; ++i) {
                double acc = 0.0;
                for (int c = 0; c < 4; ++c) {
                    double qic = 0.0;
                    for (int s = 0; s < 7; ++s) qic += K[s][i] * RK_P[s][c];
                    acc += qic * pw[c];
                }
                tmp[i] = hh * acc + y[i];
            }
            for (int i = 0; i < 7; ++i) y[i] = tmp[i];
            return accepted;
        }
        for (int i = 0; i < 7; ++i) { y[i] = y_new[i]; f[i] = K[6][i]; }
    }
}

/* ---- environment logic ---------------------------------------------------------- */
typedef struct { double *rc, *vc, *qc, *wc, *qt, *wt; } View;
static View view(double *s) { View v = {s, s + 3, s + 6, s + 10, s + 13, s + 17}; return v; }

static double attitude_error(const OrcParams *p, const View *v)
{   /* rendezvous_env.py:424-434 */
    double cap[3], neg[3] = {-v->rc[0], -v->rc[1], -v->rc[2]};
    body2lvlh(v->qc, p->capture_axis, cap);
    return rounded_angle(neg, cap);
}
static int collision_now(const OrcParams *p, const View *v)
{   /* rendezvous_env.py:388-404 */
    if (norm_n(v->rc, 3) < p->koz_radius) {
        double ax[3];
        body2lvlh(v->qt, p->corridor_axis, ax);
        if (rounded_angle(v->rc, ax) > p->corridor_half_angle) return 1;
    }
    return 0;
}
static void errors4(const OrcParams *p, const View *v, double e[4])
{   /* rendezvous_env.py:451-468 */
    double wc_l[3], wt_l[3], rd_l[3], vd_l[3], d[3];
    body2lvlh(v->qc, v->wc, wc_l);
    body2lvlh(v->qt, v->wt, wt_l);
    body2lvlh(v->qt, p->rd, rd_l);
    cross3(wt_l, rd_l, vd_l);
    for (int i = 0; i < 3; ++i) d[i] = v->rc[i] - rd_l[i];
    e[0] = norm_n(d, 3);
    for (int i = 0; i < 3; ++i) d[i] = v->vc[i] - vd_l[i];
    e[1] = norm_n(d, 3);
    e[2] = attitude_error(p, v);
    for (int i = 0; i < 3; ++i) d[i] = wc_l[i] - wt_l[i];
    e[3] = norm_n(d, 3);
}
static int success_now(const OrcParams *p, const View *v, int collided)
{   /* rendezvous_env.py:406-422 */
    if (collided) return 0;
    double e[4];
    errors4(p, v, e);
    return e[0] <= p->max_rd_error && e[1] <= p->max_vd_error && e[2] <= p->max_qd_error &&
           e[3] <= p->max_wd_error;
}
static void observe(const OrcParams *p, const View *v, float o[17])
{   /* rendezvous_env.py:294-311 + utils/general.py:230-245 : (b-a)*(val-low)/(high-low)+a */
    int k = 0;
    const double hi[3] = {p->max_axial_distance, p->max_axial_speed, p->max_wc};
    for (int i = 0; i < 3; ++i) o[k++] = (float)(2 * (v->rc[i] - -hi[0]) / (hi[0] - -hi[0]) + -1);
    for (int i = 0; i < 3; ++i) o[k++] = (float)(2 * (v->vc[i] - -hi[1]) / (hi[1] - -hi[1]) + -1);
    for (int i = 0; i < 4; ++i) o[k++] = (float)v->qc[i];
    for (int i = 0; i < 3; ++i) o[k++] = (float)(2 * (v->wc[i] - -hi[2]) / (hi[2] - -hi[2]) + -1);
    for (int i = 0; i < 4; ++i) o[k++] = (float)v->qt[i];
}
static double koz_distance(const OrcParams *p, const View *v)
{   /* rendezvous_env.py:510-537 */
    double r = norm_n(v->rc, 3), rk = p->koz_radius, thc = p->corridor_half_angle, ax[3];
    body2lvlh(v->qt, p->corridor_axis, ax);
    double th = rounded_angle(v->rc, ax);
    if (r < rk) {
        if (th >= thc) {
            double d_rad = rk - r, d_tan = r * sin(fmin(th - thc, M_PI / 2));
            return -1 * fmin(d_rad, d_tan);
        }
        return r * sin(thc - th);
    }
    if (th >= thc) return r - rk;
    double d_rad = r - rk * cos(thc - th), d_tan = rk * sin(thc - th);
    return sqrt(d_rad * d_rad + d_tan * d_tan);
}

static void cw(const OrcParams *p, double r[3], double v[3])
{   /* utils/dynamics.py:24-55, full 6x6 product incl. the zero entries */
    double n = p->n, nt = n * p->dt, s = sin(nt), c = cos(nt);
    double M[6][6] = {
        {4 - 3 * c, 0, 0, 1 / n * s, 2 / n * (1 - c), 0},
        {6 * (s - nt), 1, 0, -2 / n * (1 - c), 1 / n * (4 * s - 3 * nt), 0},
        {0, 0, c, 0, 0, 1 / n * s},
        {3 * n * s, 0, 0, c, 2 * s, 0},
        {-6 * n * (1 - c), 0, 0, -2 * s, 4 * c - 3, 0},
        {0, 0, -n * s, 0, 0, c}};
    double x[6] = {r[0], r[1], r[2], v[0], v[1], v[2]}, o[6];
    for (int i = 0; i < 6; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += M[i][j] * x[j];
        o[i] = acc;
    }
    for (int i = 0; i < 3; ++i) { r[i] = o[i]; v[i] = o[3 + i]; }
}

/* One env.step(): rendezvous_env.py:160-221.  act_f32 selects the NumPy-2
 * promotion rules the reference follows when SB3 hands it float32 actions
 * (SURVEY.md 8a row a2): delta_v and total_delta_v in fp32, delta_w in fp64,
 * fuel reward term in fp32. */
static void step_one(const OrcParams *p, double *state, double *aux, int32_t *flags, const void *action,
                     int act_f32, float *obs, double *rew_out, uint8_t *done_out, int8_t *reason_out,
                     int32_t *rk_out)
{
    View v = view(state);
    double dv_b[3], dw[3], a64[6], sum_v, sum_w;
    float fuel32 = 0.f;
    if (act_f32) {
        const float *a = (const float *)action;
        for (int i = 0; i < 3; ++i) dv_b[i] = (double)(a[i] * (float)p->max_delta_v);
        for (int i = 0; i < 3; ++i) dw[i] = (double)a[3 + i] * p->max_delta_w;
        float sv = fabsf(a[0]); sv += fabsf(a[1]); sv += fabsf(a[2]);
        float sw = fabsf(a[3]); sw += fabsf(a[4]); sw += fabsf(a[5]);
        sum_v = sv; sum_w = sw;
        aux[0] = (double)((float)aux[0] + sv * (float)p->max_delta_v);          /* :201 in fp32 */
        aux[1] = aux[1] + (double)sw * p->max_delta_w;                          /* :202 in fp64 */
        fuel32 = ((float)(p->dt * p->fuel_coef) * sv) / (float)(3 * p->max_delta_v);   /* :333 */
    } else {
        const double *a = (const double *)action;
        for (int i = 0; i < 6; ++i) a64[i] = a[i];
        for (int i = 0; i < 3; ++i) dv_b[i] = a64[i] * p->max_delta_v;
        for (int i = 0; i < 3; ++i) dw[i] = a64[3 + i] * p->max_delta_w;
        sum_v = fabs(a64[0]); sum_v += fabs(a64[1]); sum_v += fabs(a64[2]);
        sum_w = fabs(a64[3]); sum_w += fabs(a64[4]); sum_w += fabs(a64[5]);
        aux[0] += sum_v * p->max_delta_v;
        aux[1] += sum_w * p->max_delta_w;
    }
    double dv[3];
    body2lvlh(v.qc, dv_b, dv);                                       /* :172 */
    for (int i = 0; i < 3; ++i) v.vc[i] += dv[i];                    /* :176 */
    cw(p, v.rc, v.vc);                                               /* :177 */
    for (int i = 0; i < 3; ++i) v.wc[i] += dw[i];                    /* :180 */

    int rk_c, rk_t, rej = 0;
    {   /* :552-577 */
        double y[7] = {v.qc[0], v.qc[1], v.qc[2], v.qc[3], v.wc[0], v.wc[1], v.wc[2]};
        Body b = {p->inertia_c, p->inv_inertia_c, p->torque_c};
        rk_c = rk45(y, p->dt, &b, &rej);
        double nq = norm_n(y, 4);
        for (int i = 0; i < 4; ++i) v.qc[i] = y[i] / nq;
        for (int i = 0; i < 3; ++i) v.wc[i] = y[4 + i];
    }
    {   /* :579-604 */
        double y[7] = {v.qt[0], v.qt[1], v.qt[2], v.qt[3], v.wt[0], v.wt[1], v.wt[2]};
        const double zero[3] = {0, 0, 0};
        Body b = {p->inertia_t, p->inv_inertia_t, zero};
        rk_t = rk45(y, p->dt, &b, &rej);
        double nq = norm_n(y, 4);
        for (int i = 0; i < 4; ++i) v.qt[i] = y[i] / nq;
        for (int i = 0; i < 3; ++i) v.wt[i] = y[4 + i];
    }
    if (rk_out) { rk_out[0] = rk_c; rk_out[1] = rk_t; rk_out[2] = rej; }

    if (!flags[0]) {                                                 /* :186-190 */
        flags[0] = collision_now(p, &v);
        if (success_now(p, &v, flags[0])) flags[1] += 1;
    }
    aux[2] = rint((aux[2] + p->dt) * 1000.0) / 1000.0;               /* :193 round(t+dt, 3) */
    aux[3] -= p->bubble_rate;                                        /* :196-198 */
    if (aux[3] < p->bubble_min) aux[3] = p->bubble_min;

    observe(p, &v, obs);                                             /* :205 */
    double att = attitude_error(p, &v);
    int inside = 1;                                                  /* gym 0.21 Box.contains on f32 obs */
    for (int i = 0; i < 17; ++i) if (!(obs[i] >= -1.0f) || !(obs[i] <= 1.0f)) inside = 0;
    int c0 = !inside, c1 = aux[2] >= p->t_max, c2 = norm_n(v.rc, 3) > aux[3], c3 = att > p->max_attitude_error;
    *done_out = (uint8_t)(c0 || c1 || c2 || c3);                     /* :355-386 */
    if (reason_out) *reason_out = (int8_t)(c0 ? 0 : c1 ? 1 : c2 ? 2 : c3 ? 3 : -1);

    /* reward: :313-353 */
    double rew = 0;
    rew += (p->dt * p->att_coef) * (1 - att / p->max_attitude_error);
    if (act_f32) rew += (double)fuel32;
    else rew += p->dt * p->fuel_coef * sum_v / (3 * p->max_delta_v);
    if (collision_now(p, &v)) rew -= p->dt * p->collision_coef;
    if (norm_n(v.rc, 3) < p->koz_radius && !flags[0]) {
        double e[4];
        errors4(p, &v, e);
        if (e[0] < p->max_rd_error) {
            rew += p->dt * p->bonus_coef * (2 - e[0] / p->max_rd_error);
            if (e[2] < p->max_qd_error) rew += p->dt * p->bonus_coef * (2 - e[2] / p->max_qd_error);
        }
    }
    *rew_out = rew;
}

/* reset(): rendezvous_env.py:223-270 with the 24 uniform draws supplied by the
 * caller in the reference's order; u in [0,1) is mapped like numpy's
 * uniform(low, high) = low + (high-low)*u. */
static void unit_vec(const double *u, double o[3])
{   /* utils/general.py:248-254 */
    double v[3] = {-1 + 2 * u[0], -1 + 2 * u[1], -1 + 2 * u[2]};
    double nv = norm_n(v, 3);
    for (int i = 0; i < 3; ++i) o[i] = v[i] / nv;
}
static void rot2quat(const double ax_in[3], double theta, double q[4])
{   /* utils/quaternions.py:11-27 */
    double na = norm_n(ax_in, 3), ax[3];
    for (int i = 0; i < 3; ++i) ax[i] = ax_in[i] / na;
    q[0] = cos(theta / 2);
    for (int i = 0; i < 3; ++i) q[1 + i] = ax[i] * sin(theta / 2);
    double nq = norm_n(q, 4);
    for (int i = 0; i < 4; ++i) q[i] /= nq;
}
static void quat_product(const double *a_in, const double *b_in, double o[4])
{   /* utils/quaternions.py:149-170 */
    double a[4], b[4], na = norm_n(a_in, 4), nb = norm_n(b_in, 4);
    for (int i = 0; i < 4; ++i) { a[i] = a_in[i] / na; b[i] = b_in[i] / nb; }
    double cr[3];
    cross3(a + 1, b + 1, cr);
    o[0] = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
    for (int i = 0; i < 3; ++i) o[1 + i] = a[0] * b[1 + i] + b[0] * a[1 + i] + cr[i];
}
static void reset_one(const OrcParams *p, const double *u, double *state, double *aux, int32_t *flags,
                      float *obs)
{
    View v = view(state);
    double dir[3], dev[3], q_dev[4], tmp[3];
    unit_vec(u + 0, dir);
    for (int i = 0; i < 3; ++i) v.rc[i] = p->rc0[i] + dir[i] * (0 + (p->rc0_range - 0) * u[3]);
    unit_vec(u + 4, dir);
    for (int i = 0; i < 3; ++i) v.vc[i] = p->vc0[i] + dir[i] * (0 + (p->vc0_range - 0) * u[7]);
    double theta_c = 0 + (p->qc0_range - 0) * u[8];
    unit_vec(u + 9, dir);
    rot2quat(dir, theta_c, q_dev);
    quat_product(q_dev, p->qc0, v.qc);
    unit_vec(u + 12, dir);
    for (int i = 0; i < 3; ++i) { dev[i] = dir[i] * (0 + (p->wc0_range - 0) * u[15]); tmp[i] = p->wc0[i] + dev[i]; }
    lvlh2body(v.qc, tmp, v.wc);
    double theta_t = 0 + (p->qt0_range - 0) * u[16];
    unit_vec(u + 17, dir);
    rot2quat(dir, theta_t, q_dev);
    quat_product(q_dev, p->qt0, v.qt);
    unit_vec(u + 20, dir);
    for (int i = 0; i < 3; ++i) { dev[i] = dir[i] * (0 + (p->wt0_range - 0) * u[23]); tmp[i] = p->wt0[i] + dev[i]; }
    lvlh2body(v.qt, tmp, v.wt);
    flags[0] = collision_now(p, &v);
    flags[1] = success_now(p, &v, flags[0]);
    aux[0] = 0; aux[1] = 0; aux[2] = 0; aux[3] = p->bubble0;
    if (obs) observe(p, &v, obs);
}

/* ---- batch entry points (arrays are [n, ...] row-major, one env per row) --------- */
void orc_step(const OrcParams *p, int64_t n, double *state, double *aux, int32_t *flags, const void *actions,
              int act_f32, float *obs, double *rew, uint8_t *done, int8_t *reason, int32_t *rk_steps,
              int reserved)
{   /* single-threaded; callers parallelise by giving disjoint row ranges to host threads
     * (ctypes releases the GIL) -- libgomp is not in this image */
    (void)reserved;
    size_t asz = act_f32 ? sizeof(float) : sizeof(double);
    for (int64_t i = 0; i < n; ++i)
        step_one(p, state + ORC_NSTATE * i, aux + ORC_NAUX * i, flags + 2 * i,
                 (const char *)actions + 6 * asz * i, act_f32, obs + 17 * i, rew + i, done + i,
                 reason ? reason + i : 0, rk_steps ? rk_steps + 3 * i : 0);
}

void orc_reset(const OrcParams *p, int64_t n, const double *uniforms, const uint8_t *mask, double *state,
               double *aux, int32_t *flags, float *obs)
{
    for (int64_t i = 0; i < n; ++i)
        if (!mask || mask[i])
            reset_one(p, uniforms + 24 * i, state + ORC_NSTATE * i, aux + ORC_NAUX * i, flags + 2 * i,
                      obs ? obs + 17 * i : 0);
}

void orc_observe(const OrcParams *p, int64_t n, double *state, float *obs)
{
    for (int64_t i = 0; i < n; ++i) { View v = view(state + ORC_NSTATE * i); observe(p, &v, obs + 17 * i); }
}

/* errors[4], collision_now, success_now (given sticky collided), dist_from_koz */
void orc_errors(const OrcParams *p, int64_t n, double *state, const int32_t *flags, double *errors,
                uint8_t *collision, uint8_t *success, double *koz)
{
    for (int64_t i = 0; i < n; ++i) {
        View v = view(state + ORC_NSTATE * i);
        errors4(p, &v, errors + 4 * i);
        collision[i] = (uint8_t)collision_now(p, &v);
        success[i] = (uint8_t)success_now(p, &v, flags ? flags[2 * i] : 0);
        koz[i] = koz_distance(p, &v);
    }
}

/* Philox4x32-10 (Salmon et al., SC'11; the Random123 reference constants), the
 * counter-based generator the CUDA reset kernel uses.  counter = (env_id lo, env_id hi,
 * episode_idx, block), key = (seed lo, seed hi).  Restated here so the oracle can
 * reproduce the device's uniform draws bit-for-bit. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* 24 uniforms in [0,1) with 53 random bits each ((a>>5)*2^26 + (b>>6)) / 2^53,
 * the same bit recipe numpy's random_sample uses on two 32-bit words. */
void orc_philox_uniforms(uint64_t seed, int64_t env_id, int32_t episode_idx, double *u24)
{
    for (uint32_t blk = 0; blk < 12; ++blk) {
        uint32_t c[4] = {(uint32_t)env_id, (uint32_t)((uint64_t)env_id >> 32), (uint32_t)episode_idx, blk};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        u24[2 * blk] = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0;
        u24[2 * blk + 1] = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) / 9007199254740992.0;
    }
}

/* U(-1,1) fp64 actions of the device's fused rollout (RDV_ACTIONS_PHILOX): counter = (env_id lo, env_id hi,
 * step index lo, 0x40000000 | block | step index hi << 4), blocks 0..2, key = action seed; a = 2u - 1. */
void orc_philox_actions(uint64_t seed, int64_t n, const int64_t *env_ids, int64_t step_index, double *actions)
{
    /* two blocks per env-step; six of the eight 32-bit words become a = (w + 0.5) 2^-31 - 1 (exact in double) */
    for (int64_t e = 0; e < n; ++e) {
        uint32_t w[8];
        for (uint32_t blk = 0; blk < 2; ++blk) {
            uint32_t c[4] = {(uint32_t)env_ids[e], (uint32_t)((uint64_t)env_ids[e] >> 32), (uint32_t)step_index,
                             0x40000000u | blk | (((uint32_t)((uint64_t)step_index >> 32) & 0x00FFFFFFu) << 4)};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            for (int j = 0; j < 4; ++j) w[4 * blk + j] = c[j];
        }
        for (int j = 0; j < 6; ++j) actions[6 * e + j] = (double)w[j] * 4.656612873077392578125e-10 + (-1.0 + 2.3283064365386962890625e-10);
    }
}

void orc_philox_uniforms_batch(uint64_t seed, int64_t n, const int64_t *env_ids, const int32_t *episode_idx, double *u)
{
    for (int64_t e = 0; e < n; ++e) orc_philox_uniforms(seed, env_ids[e], episode_idx[e], u + 24 * e);
}

/* raw block function, exported for the Random123 known-answer test */
void orc_philox_raw(uint32_t ctr[4], uint32_t k0, uint32_t k1) { philox4x32_10(ctr, k0, k1); }
