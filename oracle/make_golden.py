"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the UNMODIFIED
reference (run in the build container, where /root/reference exists):

    python oracle/make_golden.py            # everything (the MC sweep takes ~6 min)
    python oracle/make_golden.py --skip-mc  # all but the 1000-episode Monte-Carlo

The reference is imported through oracle/ref_stubs.py (gym/SB3 are stubbed; the
env, dynamics, quaternion and evaluator code runs verbatim).  Outputs are small
numpy archives that travel to the GPU box; the tests never read /root/reference
at run time.

Files written (T_max-padded with NaN, `length[c]` = number of valid steps):
  traj_f64.npz   CSV rows x seeded fp64 action streams (zero / fixed / random)
  traj_f32.npz   CSV rows x fp32 actions of the shipped MLP policy (deterministic)
  traj_cfg.npz   non-default constructor configs (sensitivity-style parameters)
  reset.npz      reset() states for given uniform draws (np.random.uniform patched)
  verify.npz     headless end states of the verification/*.py scenarios
  mc.npz         initial conditions (the reference CSV), the reference's
                 published workbook columns, and monte_carlo.evaluate re-run here
  policy.npz     fp32 weights of models/mlp_model_best.zip (policy.pth)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_stubs  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
STATE_KEYS = ("rc", "vc", "qc", "wc", "qt", "wt")
NO_RANGE = dict(rc0_range=0, vc0_range=0, qc0_range=0, wc0_range=0, qt0_range=0, wt0_range=0)


def state_of(env):
    return np.hstack([np.asarray(getattr(env, k), dtype=float) for k in STATE_KEYS])


def inject(env, row20):
    row20 = np.asarray(row20, dtype=float)
    env.rc, env.vc = row20[0:3].copy(), row20[3:6].copy()
    env.qc, env.wc = row20[6:10].copy(), row20[10:13].copy()
    env.qt, env.wt = row20[13:17].copy(), row20[17:20].copy()


def end_reason_of(env, obs):
    conds = [
        not env.observation_space.contains(obs),
        env.t >= env.t_max,
        np.linalg.norm(env.rc) > env.bubble_radius,
        env.get_attitude_error() > env.max_attitude_error,
    ]
    return conds.index(True) if any(conds) else -1


def record(env, ic, actions, stop_on_done=True, post_inject=None):
    """Roll the reference env; returns a dict of per-step arrays."""
    env.reset()
    if ic is not None:
        inject(env, ic)
    if post_inject:
        post_inject(env)
    T = len(actions)
    out = dict(
        state=np.full((T, 20), np.nan), obs=np.full((T, 17), np.nan, dtype=np.float32),
        rew=np.full(T, np.nan), done=np.zeros(T, dtype=np.int8), collided=np.zeros(T, dtype=np.int8),
        success=np.zeros(T, dtype=np.int32), tdv=np.full(T, np.nan), tdw=np.full(T, np.nan),
        reason=np.full(T, -1, dtype=np.int8), errors=np.full((T, 4), np.nan), koz=np.full(T, np.nan),
        collision_now=np.zeros(T, dtype=np.int8), t=np.full(T, np.nan), bubble=np.full(T, np.nan),
    )
    out["state0"] = state_of(env)
    out["obs0"] = env.get_observation()
    out["collided0"] = np.int8(bool(env.collided))
    out["success0"] = np.int32(env.success)
    n = 0
    for k in range(T):
        obs, rew, done, _ = env.step(actions[k])
        out["state"][k] = state_of(env)
        out["obs"][k] = obs
        out["rew"][k] = rew
        out["done"][k] = done
        out["collided"][k] = bool(env.collided)
        out["success"][k] = env.success
        out["tdv"][k] = env.total_delta_v
        out["tdw"][k] = env.total_delta_w
        out["reason"][k] = end_reason_of(env, obs)
        out["errors"][k] = env.get_errors()
        out["koz"][k] = env.dist_from_koz()
        out["collision_now"][k] = bool(env.check_collision())
        out["t"][k] = env.t
        out["bubble"][k] = env.bubble_radius
        n = k + 1
        if done and stop_on_done:
            break
    out["length"] = np.int32(n)
    return out


def stack(cases, actions_list, extra=None):
    keys = [k for k in cases[0] if k not in ("length",)]
    packed = {k: np.stack([c[k] for c in cases]) for k in keys}
    packed["length"] = np.array([c["length"] for c in cases], dtype=np.int32)
    packed["actions"] = np.stack(actions_list)
    if extra:
        packed.update(extra)
    return packed


def load_ics():
    import pandas as pd
    df = pd.read_csv(os.path.join(ref_stubs.REFERENCE_ROOT, "results",
                                  "data_monte_carlo_initial_conditions.csv"), index_col=0)
    cols = ['rcx', 'rcy', 'rcz', 'vcx', 'vcy', 'vcz', 'qcw', 'qcx', 'qcy', 'qcz',
            'wcx', 'wcy', 'wcz', 'qtw', 'qtx', 'qty', 'qtz', 'wtx', 'wty', 'wtz']
    ics = df[cols].to_numpy(dtype=float)
    # monte_carlo.py:66-67 normalises the quaternions before injecting them
    ics[:, 6:10] /= np.linalg.norm(ics[:, 6:10], axis=1, keepdims=True)
    ics[:, 13:17] /= np.linalg.norm(ics[:, 13:17], axis=1, keepdims=True)
    return df[cols].to_numpy(dtype=float), ics


def gen_traj_f64(env_mod, ics):
    """fp64 actions.  Streams: zero, the SURVEY 8c fixed action, U(-1,1), and
    0.25*U(-1,1) (keeps the attitude error small so episodes run long and reach
    the keep-out zone / time limit)."""
    rng = np.random.default_rng(20261018)
    cases, acts, meta = [], [], []
    T = 120
    for t_max in (60, 120):
        env = env_mod.RendezvousEnv(dt=1, t_max=t_max, quiet=True, **NO_RANGE)
        for row in range(12):
            for kind in ("zero", "fixed", "uniform", "gentle"):
                if kind == "zero":
                    a = np.zeros((T, 6))
                elif kind == "fixed":
                    a = np.tile(np.array([0.5, -0.25, 0.1, 0.2, -0.1, 0.05]), (T, 1))
                elif kind == "uniform":
                    a = rng.uniform(-1, 1, (T, 6))
                else:
                    a = 0.25 * rng.uniform(-1, 1, (T, 6))
                    a[:, 1] = np.abs(a[:, 1]) * 2      # push towards the target (+y)
                cases.append(record(env, ics[row], a))
                acts.append(a)
                meta.append((row, t_max, kind))
    return stack(cases, acts, dict(meta=np.array(json.dumps(meta)), ic=np.stack([ics[m[0]] for m in meta]),
                                   t_max=np.array([m[1] for m in meta], dtype=float)))


def gen_traj_f32(env_mod, ics, policy, rows=48):
    """fp32 actions of the shipped policy, deterministic (monte_carlo.py:126-133).
    The action fed at step k depends on the reference's own obs, so the stream is
    recorded and replayed open-loop by the tests."""
    env = env_mod.RendezvousEnv(dt=1, t_max=60, quiet=True, **NO_RANGE)
    cases, acts = [], []
    T = 60
    for row in range(rows):
        a_log = np.zeros((T, 6), dtype=np.float32)

        class Closed:
            """action list that queries the policy lazily from the env's last obs"""
            def __init__(self):
                self.obs = None

            def __len__(self):
                return T

            def __getitem__(self, k):
                obs = env.get_observation()
                a, _ = policy.predict(obs, deterministic=True)
                a_log[k] = a
                return a

        cases.append(record(env, ics[row], Closed()))
        acts.append(a_log)
    return stack(cases, acts, dict(ic=ics[:rows].copy(), t_max=np.full(rows, 60.0)))


def gen_traj_cfg(env_mod, ics):
    """Non-default constructor parameters (sensitivity_analysis.py:97-134 values)."""
    rng = np.random.default_rng(7)
    cfgs = [
        dict(h=400e3), dict(h=2000e3),
        dict(koz_radius=10), dict(koz_radius=3),
        dict(corridor_half_angle=float(np.radians(15))), dict(corridor_half_angle=float(np.radians(45))),
        dict(dt=0.5), dict(dt=0.25, t_max=20), dict(dt=2), dict(dt=4),
        dict(rc0=[0., -30., 0.]), dict(wt0=[0., 0., float(np.radians(2.5))]),
        dict(reward_kwargs=dict(collision_coef=1.0, bonus_coef=4, fuel_coef=0.1, att_coef=2.0)),
        dict(dt=0.1, t_max=5),
    ]
    T = 120
    cases, acts, ic_list, cfg_json = [], [], [], []
    for ci, cfg in enumerate(cfgs):
        kw = {k: (np.array(v) if isinstance(v, list) else v) for k, v in cfg.items()}
        env = env_mod.RendezvousEnv(quiet=True, **NO_RANGE, **kw)
        for rep in range(2):
            ic = ics[100 + 2 * ci + rep].copy()
            if "rc0" in cfg:
                ic[0:3] += np.array(cfg["rc0"]) - np.array([0., -10., 0.])
            a = (0.3 if rep else 1.0) * rng.uniform(-1, 1, (T, 6))
            if rep:
                a[:, 1] = np.abs(a[:, 1]) * 2
            cases.append(record(env, ic, a))
            acts.append(a)
            ic_list.append(ic)
            cfg_json.append(json.dumps(cfg))
    return stack(cases, acts, dict(ic=np.stack(ic_list), cfg=np.array(cfg_json)))


def gen_reset(env_mod):
    """reset() with np.random.uniform fed from a given stream of [0,1) numbers
    (low + (high-low)*u, numpy's own mapping) -- pins the draw order and the
    uniform->state map of rendezvous_env.py:223-270."""
    rng = np.random.default_rng(11)
    cfgs = [
        dict(),
        dict(rc0=[0., -30., 0.], wt0=[0., 0., float(np.radians(2.5))]),
        dict(qc0=[0.8, 0.2, -0.4, 0.4], qt0=[0.5, -0.5, 0.5, 0.5], wc0=[0.01, -0.02, 0.005],
             vc0=[0.01, 0.02, -0.03], wt0=[0.02, 0.01, -0.03], qt0_range=float(np.radians(90)),
             rc0_range=6.0),
    ]
    out_u, out_s, out_o, out_c, out_k, out_cfg = [], [], [], [], [], []
    real_uniform = np.random.uniform
    for cfg in cfgs:
        kw = {k: (np.array(v) if isinstance(v, list) else v) for k, v in cfg.items()}
        env = env_mod.RendezvousEnv(quiet=True, **kw)
        for _ in range(64):
            u = rng.random(24)
            it = iter(u.tolist())

            def fake(low=0.0, high=1.0, size=None):
                if size is None:
                    return low + (high - low) * next(it)
                return np.array([low + (high - low) * next(it) for _ in range(int(np.prod(size)))])
            np.random.uniform = fake
            try:
                obs = env.reset()
            finally:
                np.random.uniform = real_uniform
            out_u.append(u)
            out_s.append(state_of(env))
            out_o.append(obs)
            out_c.append(int(bool(env.collided)))
            out_k.append(int(env.success))
            out_cfg.append(json.dumps(cfg))
    return dict(uniforms=np.stack(out_u), state=np.stack(out_s), obs=np.stack(out_o),
                collided=np.array(out_c, dtype=np.int8), success=np.array(out_k, dtype=np.int32),
                cfg=np.array(out_cfg))


def gen_verify(env_mod, dyn_mod):
    """Headless end states of verification/verify_cw.py, verify_cw2.py,
    verify_attitude_torque.py, verify_attitude_racket.py (plain ctor kwargs
    replace the missing other.new_env.NewEnv; `done` is ignored like the scripts)."""
    out = {}
    # verify_cw.py:179-207 : P/8 of zero-action stepping vs one-shot analytic CW
    env = env_mod.RendezvousEnv(rc0=np.array([0., -10., 1.]), vc0=np.array([-0.01, 0.01, 0.]),
                                dt=1, quiet=True, **NO_RANGE)
    env.reset()
    r0, v0 = env.rc.copy(), env.vc.copy()
    steps = 755
    rs, vs = [], []
    for _ in range(steps):
        env.step(np.zeros(6))
        rs.append(env.rc.copy())
        vs.append(env.vc.copy())
    ana = [dyn_mod.clohessy_wiltshire_solution(r0, v0, env.n, (k + 1) * env.dt) for k in range(steps)]
    out["cw_r"] = np.array(rs)
    out["cw_v"] = np.array(vs)
    out["cw_r_analytic"] = np.array([a[0] for a in ana])
    out["cw_v_analytic"] = np.array([a[1] for a in ana])
    # verify_cw2.py:13-64 : constant radial thrust keeps vy ~ 2 m/s
    env = env_mod.RendezvousEnv(rc0=np.array([0., -120., 0.]), vc0=np.array([0., 2., 0.]),
                                dt=0.5, quiet=True, **NO_RANGE)
    env.reset()
    act = np.array([-0.041587163080124924, 0, 0, 0, 0, 0])
    st = []
    for _ in range(120):
        env.step(act)
        st.append(state_of(env))
    out["cw2_state"] = np.array(st)
    out["cw2_action"] = act
    # verify_attitude_torque.py:34-57 : constant body-z impulse train
    env = env_mod.RendezvousEnv(dt=0.5, quiet=True, **NO_RANGE)
    env.reset()
    act = np.array([0, 0, 0, 0, 0, 0.5])
    st = []
    for _ in range(65):
        env.step(act)
        st.append(state_of(env))
    out["torque_state"] = np.array(st)
    out["torque_action"] = act
    out["torque_max_delta_w"] = np.float64(env.max_delta_w)
    # verify_attitude_racket.py:34-76 : wc0 = [0, 5, 0.01] deg/s, 760 zero-action steps
    env = env_mod.RendezvousEnv(wc0=np.radians(np.array([0, 5, 0.01])), dt=1, quiet=True, **NO_RANGE)
    env.reset()
    st = []
    for _ in range(760):
        env.step(np.zeros(6))
        st.append(state_of(env))
    out["racket_state"] = np.array(st)
    return out


def read_xlsx_sheet(path, sheet_name):
    """Minimal xlsx reader (openpyxl is not installed): returns list of rows."""
    ns = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
          "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships"}
    with zipfile.ZipFile(path) as z:
        shared = []
        if "xl/sharedStrings.xml" in z.namelist():
            root = ET.fromstring(z.read("xl/sharedStrings.xml"))
            for si in root.findall("m:si", ns):
                shared.append("".join(t.text or "" for t in si.iter("{%s}t" % ns["m"])))
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
        relmap = {r.attrib["Id"]: r.attrib["Target"] for r in rels}
        target = None
        for s in wb.find("m:sheets", ns):
            if s.attrib["name"] == sheet_name:
                target = relmap[s.attrib["{%s}id" % ns["r"]]]
        if target is None:
            raise KeyError(sheet_name)
        target = target.lstrip("/")
        if not target.startswith("xl/"):
            target = "xl/" + target
        sheet = ET.fromstring(z.read(target))
        rows = []
        for row in sheet.find("m:sheetData", ns):
            vals = {}
            for c in row:
                ref = c.attrib["r"]
                col = 0
                for ch in ref:
                    if ch.isalpha():
                        col = col * 26 + (ord(ch.upper()) - 64)
                v = c.find("m:v", ns)
                if v is None:
                    continue
                if c.attrib.get("t") == "s":
                    vals[col] = shared[int(v.text)]
                else:
                    vals[col] = float(v.text)
            rows.append(vals)
        return rows


MC_COLS = ["ep_len", "num_collisions", "collided", "total_reward", "total_delta_v", "num_successes",
           "succeeded", "min_dist_from_koz", "pos_error", "vel_error", "att_error", "rot_error"]


def gen_mc(env_mod, mc_mod, eu_mod, raw_ics, policy, skip_mc, existing=None):
    out = dict(ic_raw=raw_ics)
    # the reference's own published results: results/data_monte_carlo_results_mlp.xlsx
    rows = read_xlsx_sheet(os.path.join(ref_stubs.REFERENCE_ROOT, "results",
                                        "data_monte_carlo_results_mlp.xlsx"), "results")
    header = rows[0]
    name_to_col = {str(v).strip(): k for k, v in header.items()}
    wb = {}
    for want in MC_COLS:
        match = [k for name, k in name_to_col.items() if name.split(" ")[0] == want]
        if not match:
            continue
        col = match[0]
        wb[want] = np.array([r.get(col, np.nan) for r in rows[1:1001]], dtype=float)
    for k, v in wb.items():
        out["workbook_" + k] = v
    if skip_mc:
        if existing is not None:
            for k in existing.files:
                if k.startswith("rerun_"):
                    out[k] = existing[k]
        return out
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        env = eu_mod.make_env(reward_kwargs=None, quiet=True, config=dict(dt=1, t_max=60), stochastic=False)
    res = {k: [] for k in MC_COLS}
    for i in range(len(raw_ics)):
        s = raw_ics[i]
        qc = s[6:10] / np.linalg.norm(s[6:10])
        qt = s[13:17] / np.linalg.norm(s[13:17])
        init = dict(rc=s[0:3].copy(), vc=s[3:6].copy(), qc=qc, wc=s[10:13].copy(), qt=qt, wt=s[17:20].copy())
        o = mc_mod.evaluate(policy, env, init)
        for k in MC_COLS:
            res[k].append(float(o[k]))
        if i % 100 == 0:
            print("  mc", i, flush=True)
    for k in MC_COLS:
        out["rerun_" + k] = np.array(res[k])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-mc", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    env_mod, mc_mod, eu_mod = ref_stubs.import_reference()
    import utils.dynamics as dyn_mod
    os.makedirs(GOLDEN, exist_ok=True)
    raw_ics, ics = load_ics()
    policy = ref_stubs.ReferencePolicy()

    def want(name):
        return args.only is None or args.only == name

    def save(name, d):
        path = os.path.join(GOLDEN, name)
        np.savez_compressed(path, **d)
        print(f"wrote {path}  {os.path.getsize(path) / 1024:.0f} KiB")

    if want("traj_f64"):
        save("traj_f64.npz", gen_traj_f64(env_mod, ics))
    if want("traj_f32"):
        save("traj_f32.npz", gen_traj_f32(env_mod, ics, policy))
    if want("traj_cfg"):
        save("traj_cfg.npz", gen_traj_cfg(env_mod, ics))
    if want("reset"):
        save("reset.npz", gen_reset(env_mod))
    if want("verify"):
        save("verify.npz", gen_verify(env_mod, dyn_mod))
    if want("policy"):
        save("policy.npz", {k.replace(".", "__"): v.numpy() for k, v in policy.sd.items()})
    if want("mc"):
        path = os.path.join(GOLDEN, "mc.npz")
        existing = np.load(path) if os.path.exists(path) else None
        save("mc.npz", gen_mc(env_mod, mc_mod, eu_mod, raw_ics, policy, args.skip_mc, existing))


if __name__ == "__main__":
    main()
