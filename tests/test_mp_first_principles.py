"""First-principles anchors for the movement-primitive half of the oracle (oracle/mp.py).

mp_pytorch (the reference's third-party MP library, <= 0.1.3) is absent and the reference holds no numeric vectors for it, so
the MP half of the oracle is a restatement of the PUBLISHED algorithms.  These tests tie it to what those
algorithms are defined by — the DMP second-order system and its closed-form (ProDMP) solution, normalised RBFs, finite
differences — independently of any implementation:
    tau^2 y'' = alpha (beta (g - y) - tau y') + f(x),   f = x(z) sum_k Phi_k(x) w_k,   x = exp(-alpha_x z),   z = t / tau
"""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from oracle import mp as omp


def _gens(mp_type, K=5, tau=1.5, alpha_phase=3.0, mode="gold", **traj_kw):
    pg = omp.get_phase_generator("exp" if mp_type != "promp" else "linear", mode=mode, tau=tau, alpha_phase=alpha_phase)
    basis = {"promp": "zero_rbf", "dmp": "rbf", "prodmp": "prodmp"}[mp_type]
    kw = dict(num_basis=K)
    if mp_type == "promp":
        kw.update(num_basis_zero_start=1, basis_bandwidth_factor=3.0)
    if mp_type == "prodmp":
        kw.update(alpha=traj_kw.pop("alpha", 10), dt=0.01)
    bg = omp.get_basis_generator(basis, pg, **kw)
    return pg, bg, omp.get_trajectory_generator(mp_type, 2, bg, **traj_kw)


def _forcing_ode(pg, bg, w, g, alpha, y0, v0, tau, t_eval):
    """reference solution of the DMP system in scaled time s = t / tau:  y'' = alpha (beta (g - y) - y') + f(s)"""
    beta = alpha / 4.0
    cen, bw = omp._gold_centres(bg)

    def f(s):
        x = np.exp(-float(pg.alpha0) * np.clip(s, 0, 1))
        b = np.exp(-((x - cen) ** 2 * bw) / 2)
        b = b / b.sum()
        return x * (b @ w)

    def rhs(s, u):
        return [u[1], alpha * (beta * (g - u[0]) - u[1]) + f(s)]
    sol = solve_ivp(rhs, (0.0, t_eval[-1] / tau), [y0, v0 * tau], t_eval=t_eval / tau, rtol=1e-11, atol=1e-13, max_step=1e-3)
    return sol.y[0], sol.y[1] / tau


def test_rbf_bases_are_a_partition_of_unity_on_linspace_centres():
    pg, bg, _ = _gens("promp", K=5, tau=2.0)
    t = np.linspace(0.01, 2.0, 200)
    b = bg.basis(t)
    assert b.shape == (200, 6) and np.allclose(b.sum(axis=1), 1.0, atol=1e-12) and (b >= 0).all()
    cen, bw = omp._gold_centres(bg)
    assert np.allclose(cen, np.linspace(0, 1, 6)) and np.allclose(bw, 3.0 / 0.2 ** 2)        # h = f / spacing^2 (App. B.3)
    assert (np.argmax(b, axis=1)[[0, -1]] == [0, 5]).all()                                  # first / last RBF dominate the ends


def test_promp_is_linear_in_the_weights_and_velocity_is_the_forward_difference():
    pg, bg, tg = _gens("promp", K=5, tau=2.0, weights_scale=2.0)
    rng = np.random.default_rng(0)
    w1, w2 = rng.standard_normal((2, 10))
    out = []
    for w in (w1, w2, 0.3 * w1 - 1.7 * w2):
        tg.set_params(w); tg.set_initial_conditions(0.0, np.zeros(2), np.zeros(2)); tg.set_duration(2.0, 0.01)
        out.append((tg.get_traj_pos(), tg.get_traj_vel()))
    assert np.allclose(out[2][0], 0.3 * out[0][0] - 1.7 * out[1][0], atol=1e-12)
    pos, vel = out[0]
    assert np.allclose(vel[:-1], np.diff(pos, axis=0) / np.diff(tg.times)[:, None], atol=1e-12) and np.array_equal(vel[-1], vel[-2])
    # zero padding: the first (unweighted) RBF owns t ~ 0, so the trajectory starts near 0 whatever the weights are
    assert np.abs(pos[0]).max() < 0.2 * np.abs(pos).max()


def test_dmp_euler_integration_converges_to_the_ode_with_first_order():
    alpha = 25.0
    errs = []
    for dt in (0.01, 0.005, 0.0025):
        pg, bg, tg = _gens("dmp", K=5, tau=2.0, alpha_phase=2.0, weights_scale=50.0, goal_scale=1.0)
        rng = np.random.default_rng(1)
        p = rng.standard_normal(12)
        tg.set_params(p); tg.set_initial_conditions(0.0, np.array([0.3, -0.2]), np.array([0.5, 0.0])); tg.set_duration(2.0, dt)
        pos, vel = tg.get_traj_pos(), tg.get_traj_vel()
        # the library assigns the initial state to the FIRST grid point (t = dt): compare on the grid shifted by dt (App. B.6)
        t = tg.times - tg.times[0]
        ref_p, ref_v = _forcing_ode(pg, bg, 50.0 * p[0:5], p[5], alpha, 0.3, 0.5, 2.0, t)
        # forcing is sampled at the library's grid (phase of times[i]) while the shifted ODE sees phase(t - dt): O(dt) as well
        errs.append(np.abs(pos[:, 0] - ref_p).max())
    assert errs[0] < 0.05 * max(1.0, np.abs(ref_p).max())
    assert errs[1] < 0.62 * errs[0] and errs[2] < 0.62 * errs[1]          # halving dt (about) halves the error


def test_dmp_without_forcing_is_the_critically_damped_analytic_solution():
    pg, bg, tg = _gens("dmp", K=5, tau=1.0, alpha_phase=2.0)
    g, y0 = 1.3, -0.4
    p = np.zeros(12); p[5] = g; p[11] = g
    tg.set_params(p); tg.set_initial_conditions(0.0, np.array([y0, y0]), np.zeros(2)); tg.set_duration(1.0, 0.001)
    pos = tg.get_traj_pos()[:, 0]
    s = tg.times - tg.times[0]
    a = 25.0
    exact = g + (y0 - g) * (1 + a / 2 * s) * np.exp(-a / 2 * s)          # beta = alpha / 4: double root -alpha / 2
    assert np.abs(pos - exact).max() < 1e-2                              # semi-implicit Euler at dt = 1e-3: O(dt * alpha) error


@pytest.mark.parametrize("tau,alpha", [(1.5, 10.0), (2.0, 25.0)])
def test_prodmp_closed_form_solves_the_dmp_ode_and_meets_its_boundary_conditions(tau, alpha):
    pg, bg, tg = _gens("prodmp", K=5, tau=tau, alpha_phase=3.0, alpha=alpha, weights_scale=1.0, goal_scale=1.0)
    rng = np.random.default_rng(2)
    p = rng.standard_normal(12)
    y_b, v_b = np.array([0.7, -0.3]), np.array([0.4, 1.1])
    tg.set_params(p); tg.set_initial_conditions(0.0, y_b, v_b); tg.set_duration(2.0, 0.01)
    pos, vel = tg.get_traj_pos(), tg.get_traj_vel()
    t = np.concatenate([[0.0], tg.times])
    for d in range(2):
        ref_p, ref_v = _forcing_ode(pg, bg, p[6 * d:6 * d + 5], p[6 * d + 5], alpha, y_b[d], v_b[d], tau, t)
        # the pre-integrated bases are cumulative trapezoids on the dt / tau grid: ~1e-6 quadrature error (App. B.7)
        assert np.abs(pos[:, d] - ref_p[1:]).max() < 2e-5 * max(1.0, np.abs(ref_p).max())
        assert np.abs(vel[:, d] - ref_v[1:]).max() < 2e-4 * max(1.0, np.abs(ref_v).max())
    # boundary conditions: at the boundary time the blend reproduces (y_b, v_b) exactly whatever the weights are
    pos_H, vel_H, xi = tg.tables()
    bgv = bg.general_solution_values(np.array([0.0]))
    assert np.allclose([bgv[0][0], bgv[1][0]], [1.0, 0.0]) and np.allclose(bg.basis(np.array([0.0])), 0.0, atol=1e-15)


def test_mirror_and_shipped_modes_stay_within_float32_of_the_float64_definition():
    for mp_type in ("promp", "dmp", "prodmp"):
        outs = {}
        for mode in ("gold", "shipped", "mirror"):
            kw = dict(weights_scale=2.0) if mp_type == "promp" else {}
            pg, bg, tg = _gens(mp_type, K=5, tau=2.0, mode=mode, **kw)
            p = np.random.default_rng(3).standard_normal(tg.num_params).astype(np.float32) * 0.5
            tg.set_params(p); tg.set_initial_conditions(0.0, np.array([0.5, -0.5]), np.array([0.1, 0.0])); tg.set_duration(2.0, 0.01)
            outs[mode] = tg.get_traj_pos().astype(np.float64)
        scale = max(1.0, np.abs(outs["gold"]).max())
        assert np.abs(outs["shipped"] - outs["gold"]).max() < 1e-5 * scale, mp_type
        assert np.abs(outs["mirror"] - outs["gold"]).max() < 1e-5 * scale, mp_type
