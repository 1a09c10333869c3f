"""Problem set-ups: the reference's three example programs restated against the mirror
classes of `lpopc_b200.problem`, plus the problems BASELINE.json names that the
reference does not ship (orbit raising, brachistochrone, quadrotor / cart-pole MPC,
synthetic 20-state stress dynamics).

Each builder returns an `OptimalProblem` whose phases already carry a mesh
(default: the reference's first mesh, one interval x 20 LGR nodes,
Lpopc/src/Core/LpMeshRefiner.cpp:30-31,50).
"""
import math

import numpy as np

from .problem import Linkage, OptimalProblem, Phase


def uniform_mesh(phase, intervals, nodes):
    phase.set_mesh(np.linspace(-1.0, 1.0, intervals + 1), [nodes] * intervals)


def hypersensitive(intervals=1, nodes=20, first_derive="finite-difference"):
    """Lpopc/example/hypersensitive/HyperSensitive.cpp:14-70."""
    t0, tf, x0, xf = 0.0, 5000.0, 1.5, 1.0
    xmin, xmax, umin, umax = -10, 10, -10, 10
    ph = Phase(1, 1, 1, 0, 0, 0)
    ph.SetTimeMin(t0, tf); ph.SetTimeMax(t0, tf)
    ph.SetStateMin(x0, xmin, xf); ph.SetStateMax(x0, xmax, xf)
    ph.SetcontrolMin(umin); ph.SetcontrolMax(umax)
    ph.SetTimeGuess(t0); ph.SetTimeGuess(tf)
    ph.SetStateGuess(1, x0); ph.SetStateGuess(1, xf)
    ph.SetControlGuess(1, -1); ph.SetControlGuess(1, 1)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "hypersensitive", [0.0], first_derive=first_derive)
    op.AddPhase(ph)
    return op


def bryson_denham(intervals=1, nodes=20):
    """Lpopc/example/bryson-denham/BrysonDenham.cpp:9-98."""
    ph = Phase(1, 3, 1, 0, 0, 5)
    ph.SetTimeMin(0.0, 0.0); ph.SetTimeMax(0, 50)
    ph.SetStateMin(0, 0, 0); ph.SetStateMax(1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0)
    ph.SetStateMin(-10, -10, -10); ph.SetStateMax(10, 10, 10)
    ph.SetStateMin(-10, -10, -10); ph.SetStateMax(10, 10, 10)
    ph.SetcontrolMin(-10); ph.SetcontrolMax(10)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(1.0)
    for v in (0, 1, 0, 0, -1):
        ph.SeteventMin(v); ph.SeteventMax(v)
    ph.SetStateGuess(1, 0); ph.SetStateGuess(1, 0)
    ph.SetStateGuess(2, 1.0); ph.SetStateGuess(2, -1.0)
    ph.SetStateGuess(3, 0.0); ph.SetStateGuess(3, 0.0)
    ph.SetControlGuess(1, 0.0); ph.SetControlGuess(1, 0.0)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "bryson_denham", [0.0])
    op.AddPhase(ph)
    return op


def _launch_oe2rv(oe, mu):
    """Launchoe2rv, Lpopc/example/launch/Launch.cpp:551-587 (host-side guess only)."""
    a, e, i, Om, om, nu = oe
    p = a * (1 - e * e)
    r = p / (1 + e * math.cos(nu))
    rv = np.array([r * math.cos(nu), r * math.sin(nu), 0.0])
    vv = np.array([-math.sin(nu), e + math.cos(nu), 0.0]) * math.sqrt(mu / p)
    cO, sO, co, so, ci, si = math.cos(Om), math.sin(Om), math.cos(om), math.sin(om), math.cos(i), math.sin(i)
    R = np.array([[cO * co - sO * so * ci, -cO * so - sO * co * ci, sO * si],
                  [sO * co + cO * so * ci, -sO * so + cO * co * ci, -cO * si],
                  [so * si, co * si, ci]])
    return R @ rv, R @ vv


def launch(intervals=1, nodes=20):
    """Lpopc/example/launch/Launch.cpp:13-545 (4 phases, 3 linkage pairs)."""
    PI = math.pi
    earthRadius, gravParam, initialMass = 6378145.0, 3.986012e14, 301454.0
    earthRotRate, seaLevelDensity, densityScaleHeight, g0 = 7.29211585e-5, 1.225, 7200.0, 9.80665
    s_length = earthRadius
    s_speed = math.sqrt(gravParam / s_length)
    s_time = s_length / s_speed
    s_acc = s_speed / s_time
    s_mass = initialMass
    s_force = s_mass * s_acc
    s_area = s_length * s_length
    s_volume = s_area * s_length
    s_density = s_mass / s_volume
    s_gravparam = s_acc * s_length * s_length
    omega = earthRotRate * s_time
    mu = gravParam / s_gravparam
    cd, sa = 0.5, 4 * PI / s_area
    rho0, H, Re, g0s = seaLevelDensity / s_density, densityScaleHeight / s_length, earthRadius / s_length, g0 / s_acc
    lat0 = 28.5 * PI / 180
    r0 = np.array([Re * math.cos(lat0), 0.0, Re * math.sin(lat0)])
    omat = np.array([[0, -omega, 0], [omega, 0, 0], [0, 0, 0]])
    v0 = omat @ r0
    bt_srb, bt_first, bt_second = 75.2 / s_time, 261.0 / s_time, 700.0 / s_time
    t0, t1, t2, t3, t4 = 0.0, 75.2 / s_time, 150.4 / s_time, 261 / s_time, 961 / s_time
    m_tot_srb, m_prop_srb = 19290 / s_mass, 17010 / s_mass
    m_dry_srb = m_tot_srb - m_prop_srb
    m_tot_first, m_prop_first = 104380 / s_mass, 95550 / s_mass
    m_dry_first = m_tot_first - m_prop_first
    m_tot_second, m_prop_second = 19300 / s_mass, 16820 / s_mass
    m_payload = 4164 / s_mass
    thrust_srb, thrust_first, thrust_second = 628500 / s_force, 1083100 / s_force, 110094 / s_force
    mdot_srb = m_prop_srb / bt_srb
    ISP_srb = thrust_srb / (g0s * mdot_srb)
    mdot_first = m_prop_first / bt_first
    ISP_first = thrust_first / (g0s * mdot_first)
    mdot_second = m_prop_second / bt_second
    ISP_second = thrust_second / (g0s * mdot_second)
    af, ef, incf, Omf, omf = 24361140 / s_length, 0.7308, 28.5 * PI / 180, 269.8 * PI / 180, 130.5 * PI / 180
    rout, vout = _launch_oe2rv([af, ef, incf, Omf, omf, 0.0], mu)
    m10 = m_payload + m_tot_second + m_tot_first + 9 * m_tot_srb
    m1f = m10 - (6 * mdot_srb + mdot_first) * t1
    m20 = m1f - 6 * m_dry_srb
    m2f = m20 - (3 * mdot_srb + mdot_first) * (t2 - t1)
    m30 = m2f - 3 * m_dry_srb
    m3f = m30 - mdot_first * (t3 - t2)
    m40 = m3f - m_dry_first
    m4f = m_payload
    rmin, rmax = -2 * Re, 2 * Re
    vmin, vmax = -10000 / s_speed, 10000 / s_speed
    consts = [omega, mu, cd, sa, rho0, H, Re, g0s, thrust_srb, thrust_first, thrust_second, ISP_srb, ISP_first, ISP_second]
    op = OptimalProblem(4, 3, "launch", consts)
    times = [(t0, t1), (t1, t2), (t2, t3), (t3, t4)]
    masses = [(m10, m1f), (m20, m2f), (m30, m3f), (m40, m4f)]
    for ip in range(4):
        ph = Phase(ip + 1, 7, 3, 0, 1, 5 if ip == 3 else 0)
        ta, tb = times[ip]
        if ip < 3:
            ph.SetTimeMin(ta, tb); ph.SetTimeMax(ta, tb)
        else:
            ph.SetTimeMin(t3, t3); ph.SetTimeMax(t3, t4)
        for j in range(3):
            if ip == 0:
                ph.SetStateMin(r0[j], rmin, rmin); ph.SetStateMax(r0[j], rmax, rmax)
            else:
                ph.SetStateMin(rmin, rmin, rmin); ph.SetStateMax(rmax, rmax, rmax)
        for j in range(3):
            if ip == 0:
                ph.SetStateMin(v0[j], vmin, vmin); ph.SetStateMax(v0[j], vmax, vmax)
            else:
                ph.SetStateMin(vmin, vmin, vmin); ph.SetStateMax(vmax, vmax, vmax)
        ma, mb = masses[ip]
        if ip == 0:
            ph.SetStateMin(m10, m1f, m1f); ph.SetStateMax(m10, m10, m10)
        else:
            ph.SetStateMin(mb, mb, mb); ph.SetStateMax(ma, ma, ma)
        for _ in range(3):
            ph.SetcontrolMin(-1); ph.SetcontrolMax(1)
        ph.SetpathMin(1); ph.SetpathMax(1)
        if ip == 3:
            for v in (af, ef, incf, Omf, omf):
                ph.SeteventMin(v); ph.SeteventMax(v)
        ph.SetTimeGuess(ta); ph.SetTimeGuess(tb)
        rg, vg = (r0, v0) if ip < 2 else (rout, vout)
        for j in range(3):
            ph.SetStateGuess(1 + j, rg[j]); ph.SetStateGuess(1 + j, rg[j])
        for j in range(3):
            ph.SetStateGuess(4 + j, vg[j]); ph.SetStateGuess(4 + j, vg[j])
        ph.SetStateGuess(7, ma); ph.SetStateGuess(7, mb)
        for j, g in enumerate((0, 1, 0)):
            ph.SetControlGuess(1 + j, g); ph.SetControlGuess(1 + j, g)
        uniform_mesh(ph, intervals, nodes)
        op.AddPhase(ph)
    for ip, drop in enumerate((-6 * m_dry_srb, -3 * m_dry_srb, -m_dry_first)):
        lk = Linkage(ip + 1, ip + 1, ip + 2)
        for j in range(7):
            v = drop if j == 6 else 0.0
            lk.SetLinkMin(v); lk.SetLinkMax(v)
        op.AddLinkage(lk)
    return op


def orbit_raising(intervals=200, nodes=10):
    """BASELINE config 2 (single phase, ns=5, nc=2, np=1)."""
    T, mu, mdot, tf = 0.1405, 1.0, 0.0749, 3.32
    ph = Phase(1, 5, 2, 0, 1, 2)
    ph.SetTimeMin(0.0, tf); ph.SetTimeMax(0.0, tf)
    lo = [0.5, -10.0, -5.0, -5.0, 0.1]
    hi = [5.0, 10.0, 5.0, 5.0, 1.5]
    x0 = [1.0, 0.0, 0.0, 1.0, 1.0]
    for j in range(5):
        ph.SetStateMin(x0[j], lo[j], lo[j]); ph.SetStateMax(x0[j], hi[j], hi[j])
    for _ in range(2):
        ph.SetcontrolMin(-1.0); ph.SetcontrolMax(1.0)
    ph.SetpathMin(1.0); ph.SetpathMax(1.0)
    for _ in range(2):
        ph.SeteventMin(0.0); ph.SeteventMax(0.0)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(tf)
    xf = [1.5, 2.4, 0.0, 0.8, 1.0 - mdot * tf]
    for j in range(5):
        ph.SetStateGuess(1 + j, x0[j]); ph.SetStateGuess(1 + j, xf[j])
    ph.SetControlGuess(1, 0.0); ph.SetControlGuess(1, 0.0)
    ph.SetControlGuess(2, 1.0); ph.SetControlGuess(2, 1.0)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "orbit_raising", [T, mu, mdot])
    op.AddPhase(ph)
    return op


def brachistochrone(intervals=10, nodes=8):
    ph = Phase(1, 3, 1, 0, 0, 0)
    ph.SetTimeMin(0.0, 0.1); ph.SetTimeMax(0.0, 10.0)
    ph.SetStateMin(0, 0, 2); ph.SetStateMax(0, 10, 2)
    ph.SetStateMin(0, -10, -2); ph.SetStateMax(0, 10, -2)
    ph.SetStateMin(0, -20, -20); ph.SetStateMax(0, 20, 20)
    ph.SetcontrolMin(-math.pi); ph.SetcontrolMax(math.pi)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(1.0)
    ph.SetStateGuess(1, 0); ph.SetStateGuess(1, 2)
    ph.SetStateGuess(2, 0); ph.SetStateGuess(2, -2)
    ph.SetStateGuess(3, 0); ph.SetStateGuess(3, 5)
    ph.SetControlGuess(1, 0.5); ph.SetControlGuess(1, 1.5)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "brachistochrone", [9.81])
    op.AddPhase(ph)
    return op


QUADROTOR_CONSTS = [0.5, 9.81, 2.3e-3, 2.3e-3, 4.0e-3, 0.17, 0.016,  # mass g Ixx Iyy Izz arm kM
                    10.0, 1.0, 2.0, 0.1, 0.5,                        # qp qv qa qw ru
                    1.0, 1.0, 1.0]                                   # pref


def quadrotor(intervals=8, nodes=8, x0=None, horizon=2.0):
    """BASELINE config 4 (ns=12, nc=4; shared mesh 8x8; MPC initial state through state0 bounds)."""
    if x0 is None:
        x0 = np.zeros(12)
    ph = Phase(1, 12, 4, 0, 0, 0)
    ph.SetTimeMin(0.0, horizon); ph.SetTimeMax(0.0, horizon)
    lo = [-10] * 3 + [-10] * 3 + [-1.2] * 3 + [-10] * 3
    hi = [10] * 3 + [10] * 3 + [1.2] * 3 + [10] * 3
    for j in range(12):
        ph.SetStateMin(x0[j], lo[j], lo[j]); ph.SetStateMax(x0[j], hi[j], hi[j])
    for _ in range(4):
        ph.SetcontrolMin(0.0); ph.SetcontrolMax(4.0)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(horizon)
    target = [1.0, 1.0, 1.0] + [0.0] * 9
    for j in range(12):
        ph.SetStateGuess(1 + j, x0[j]); ph.SetStateGuess(1 + j, target[j])
    hover = QUADROTOR_CONSTS[0] * QUADROTOR_CONSTS[1] / 4
    for j in range(4):
        ph.SetControlGuess(1 + j, hover); ph.SetControlGuess(1 + j, hover)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "quadrotor", QUADROTOR_CONSTS)
    op.AddPhase(ph)
    return op


CARTPOLE_CONSTS = [1.0, 0.3, 0.5, 9.81, 1.0, 5.0, 0.1, 0.05]  # mc mp l g qx qth qv ru


def cartpole(intervals=8, nodes=8, x0=None, horizon=2.0):
    """BASELINE config 4, second functor (ns=4, nc=1)."""
    if x0 is None:
        x0 = np.array([0.0, 0.3, 0.0, 0.0])
    ph = Phase(1, 4, 1, 0, 0, 0)
    ph.SetTimeMin(0.0, horizon); ph.SetTimeMax(0.0, horizon)
    lo, hi = [-5, -4, -20, -20], [5, 4, 20, 20]
    for j in range(4):
        ph.SetStateMin(x0[j], lo[j], lo[j]); ph.SetStateMax(x0[j], hi[j], hi[j])
    ph.SetcontrolMin(-30.0); ph.SetcontrolMax(30.0)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(horizon)
    for j in range(4):
        ph.SetStateGuess(1 + j, x0[j]); ph.SetStateGuess(1 + j, 0.0)
    ph.SetControlGuess(1, 0.0); ph.SetControlGuess(1, 0.0)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "cartpole", CARTPOLE_CONSTS)
    op.AddPhase(ph)
    return op


def synthetic20_consts(seed=6):
    """A (20x20), B (20x6) ~ U(-1,1)/sqrt(20), PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = rng.uniform(-1, 1, (20, 20)) / math.sqrt(20)
    B = rng.uniform(-1, 1, (20, 6)) / math.sqrt(20)
    return np.concatenate([A.ravel(), B.ravel()])


def synthetic20(intervals=10000, nodes=10, seed=6):
    """BASELINE config 5: ns=20, nc=6, 100k LGR nodes in 10 000 x 10; tf - t0 = 10."""
    ph = Phase(1, 20, 6, 0, 0, 0)
    ph.SetTimeMin(0.0, 10.0); ph.SetTimeMax(0.0, 10.0)
    for _ in range(20):
        ph.SetStateMin(-2, -2, -2); ph.SetStateMax(2, 2, 2)
    for _ in range(6):
        ph.SetcontrolMin(-2); ph.SetcontrolMax(2)
    ph.SetTimeGuess(0.0); ph.SetTimeGuess(10.0)
    for j in range(20):
        ph.SetStateGuess(1 + j, 0.5); ph.SetStateGuess(1 + j, -0.5)
    for j in range(6):
        ph.SetControlGuess(1 + j, 0.1); ph.SetControlGuess(1 + j, -0.1)
    uniform_mesh(ph, intervals, nodes)
    op = OptimalProblem(1, 0, "synthetic20", synthetic20_consts(seed))
    op.AddPhase(ph)
    return op


BUILDERS = {
    "hypersensitive": hypersensitive, "bryson_denham": bryson_denham, "launch": launch,
    "orbit_raising": orbit_raising, "brachistochrone": brachistochrone, "quadrotor": quadrotor,
    "cartpole": cartpole, "synthetic20": synthetic20,
}


def two_stage(intervals=2, nodes=6, first_derive="analytic"):
    """Two-stage transfer with an impulsive stage change (include/problems/two_stage.h): 2 phases, ns=2, nc=1, events in
    both phases, one link pair with 2 links, user-supplied first derivatives of everything (first-derive = analytic
    forwards DerivEvent / DerivLink / DerivMayer to the user, LpAnalyticDerive.hpp:18-47)."""
    op = OptimalProblem(2, 1, "two_stage", [1.3, 0.05, 0.12, 4.0, 0.9, 0.05], first_derive=first_derive)
    for ip in (1, 2):
        ph = Phase(ip, 2, 1, 0, 0, 2 if ip == 1 else 1)
        t0, t1 = (0.0, 1.0) if ip == 1 else (1.0, 2.5)
        ph.SetTimeMin(t0, t1 - 0.2); ph.SetTimeMax(t0, t1 + 0.2)
        ph.SetStateMin(-5, -5, -5); ph.SetStateMax(5, 5, 5)
        ph.SetStateMin(-5, -5, -5); ph.SetStateMax(5, 5, 5)
        ph.SetcontrolMin(-3); ph.SetcontrolMax(3)
        if ip == 1:
            for v in (0.2, 0.0):
                ph.SeteventMin(v); ph.SeteventMax(v)
        else:
            ph.SeteventMin(0.9); ph.SeteventMax(1.1)
        ph.SetTimeGuess(t0); ph.SetTimeGuess(t1)
        ph.SetStateGuess(1, 0.2 if ip == 1 else 0.6); ph.SetStateGuess(1, 0.6 if ip == 1 else 1.0)
        ph.SetStateGuess(2, 0.0 if ip == 1 else 0.5); ph.SetStateGuess(2, 0.6 if ip == 1 else 0.1)
        ph.SetControlGuess(1, 0.5); ph.SetControlGuess(1, -0.2)
        uniform_mesh(ph, intervals, nodes)
        op.AddPhase(ph)
    lk = Linkage(1, 1, 2)
    for _ in range(2):
        lk.SetLinkMin(0.0); lk.SetLinkMax(0.0)
    op.AddLinkage(lk)
    return op
