"""Small known-answer systems shared by the oracle tests and the GPU parity
tests.  They restate the set-ups of the reference's own unit tests
(opty/tests/test_direct_collocation.py, cited per case); expected values are
closed-form NumPy, expected COO indices are the literal vectors the reference
tests assert."""

from collections import OrderedDict

import numpy as np
import sympy as sm


class Case(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def collocator_args(self):
        return (self.eom, self.states, self.N, self.h)

    def collocator_kwargs(self):
        kw = dict(known_parameter_map=self.par_map,
                  known_trajectory_map=self.traj_map,
                  instance_constraints=self.instance_constraints,
                  integration_method=self.method, time_symbol=self.t)
        return kw


def msd_unknown_trajectory(method='backward euler'):
    """Mass-spring-damper, known force f(t), unknown stiffness trajectory
    k(t), known mass, unknown damping c; N = 4.
    opty/tests/test_direct_collocation.py:1020-1066 (set-up), :1127-1161
    (values), :1163-1177 (literal indices), :1179-1282 (Jacobian)."""
    m, c, t = sm.symbols('m, c, t')
    x, v, f, k = [s(t) for s in sm.symbols('x, v, f, k', cls=sm.Function)]
    eom = sm.Matrix([x.diff() - v, m * v.diff() + c * v + k * x - f])
    xs = np.array([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]])
    fs = np.array([9.0, 10.0, 11.0, 12.0])
    ks = np.array([13.0, 14.0, 15.0, 16.0])
    mv, cv, h = 1.0, 2.0, 0.01
    free = np.hstack((xs[0], xs[1], ks, [cv]))

    if method == 'backward euler':
        kin = np.array([(xs[0, i] - xs[0, i - 1]) / h - xs[1, i]
                        for i in (1, 2, 3)])
        dyn = np.array([mv * (xs[1, i] - xs[1, i - 1]) / h + cv * xs[1, i] +
                        ks[i] * xs[0, i] - fs[i] for i in (1, 2, 3)])
        # partials wrt [xi, vi, xp, vp, ki, c] per node, eom-row major
        jac = []
        for i in (1, 2, 3):
            jac += [1 / h, -1.0, -1 / h, 0.0, 0.0, 0.0]
            jac += [ks[i], mv / h + cv, 0.0, -mv / h, xs[0, i], xs[1, i]]
        rows = [0, 0, 0, 0, 0, 0, 3, 3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 4, 4, 4,
                4, 4, 4, 2, 2, 2, 2, 2, 2, 5, 5, 5, 5, 5, 5]
        cols = [1, 5, 0, 4, 9, 12, 1, 5, 0, 4, 9, 12, 2, 6, 1, 5, 10, 12, 2,
                6, 1, 5, 10, 12, 3, 7, 2, 6, 11, 12, 3, 7, 2, 6, 11, 12]
    else:
        kin, dyn, jac = [], [], []
        for i in (0, 1, 2):
            xi, vi, xn, vn = xs[0, i], xs[1, i], xs[0, i + 1], xs[1, i + 1]
            ki, kn, fi, fn = ks[i], ks[i + 1], fs[i], fs[i + 1]
            kin.append((xn - xi) / h - (vi + vn) / 2)
            dyn.append(mv * (vn - vi) / h + cv * (vi + vn) / 2 +
                       (ki + kn) / 2 * (xi + xn) / 2 - (fi + fn) / 2)
            # wrt [xi, vi, xn, vn, ki, kn, c]
            jac += [-1 / h, -0.5, 1 / h, -0.5, 0.0, 0.0, 0.0]
            jac += [(ki + kn) / 4, -mv / h + cv / 2, (ki + kn) / 4,
                    mv / h + cv / 2, (xi + xn) / 4, (xi + xn) / 4,
                    (vi + vn) / 2]
        kin, dyn = np.array(kin), np.array(dyn)
        rows = cols = None
    return Case(name='msd_unknown_trajectory_' + method.split()[0],
                eom=eom, states=(x, v), N=4, h=h, method=method, t=t,
                par_map=OrderedDict([(m, mv)]),
                traj_map=OrderedDict([(f, fs)]),
                instance_constraints=None, free=free,
                expected_con=np.hstack((kin, dyn)),
                expected_jac=np.array(jac), expected_rows=rows,
                expected_cols=cols)


def pendulum_instance_constraints():
    """Pendulum with four instance constraints, backward Euler, N = 4.
    opty/tests/test_direct_collocation.py:1411-1451 (set-up), :1556-1588
    (values), :1590-1603 (literal indices), :1645-1710 (instance parts)."""
    m, g, d, t = sm.symbols('m, g, d, t')
    theta, omega, T = [s(t) for s in sm.symbols('theta, omega, T',
                                                cls=sm.Function)]
    eom = sm.Matrix([theta.diff() - omega,
                     m * d**2 * omega.diff() + m * g * d * sm.sin(theta) - T])
    th, om, Tf = sm.symbols('theta, omega, T', cls=sm.Function)
    instance = (1.0 * th(0.0), 3.0 * th(0.03) - sm.pi, 4.0 * om(0.0),
                5.0 * om(0.03))
    thv = np.array([1.0, 2.0, 3.0, 4.0])
    omv = np.array([5.0, 6.0, 7.0, 8.0])
    Tv = np.array([9.0, 10.0, 11.0, 12.0])
    mv, gv, dv, h = 1.0, 9.81, 1.0, 0.01
    free = np.hstack((thv, omv, Tv))
    kin = np.array([(thv[i] - thv[i - 1]) / h - omv[i] for i in (1, 2, 3)])
    dyn = np.array([mv * dv**2 * (omv[i] - omv[i - 1]) / h +
                    mv * gv * dv * np.sin(thv[i]) - Tv[i] for i in (1, 2, 3)])
    inst = np.array([1.0 * thv[0], 3.0 * thv[3] - np.pi, 4.0 * omv[0],
                     5.0 * omv[3]])
    jac = []
    for i in (1, 2, 3):
        # wrt [thetai, omegai, thetap, omegap, Ti]
        jac += [1 / h, -1.0, -1 / h, 0.0, 0.0]
        jac += [mv * gv * dv * np.cos(thv[i]), mv * dv**2 / h, 0.0,
                -mv * dv**2 / h, -1.0]
    jac += [1.0, 3.0, 4.0, 5.0]
    rows = [0, 0, 0, 0, 0, 3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 4, 4, 4, 4, 4, 2, 2,
            2, 2, 2, 5, 5, 5, 5, 5, 6, 7, 8, 9]
    cols = [1, 5, 0, 4, 9, 1, 5, 0, 4, 9, 2, 6, 1, 5, 10, 2, 6, 1, 5, 10, 3,
            7, 2, 6, 11, 3, 7, 2, 6, 11, 0, 3, 4, 7]
    return Case(name='pendulum_instance_constraints', eom=eom,
                states=(theta, omega), N=4, h=h, method='backward euler', t=t,
                par_map=OrderedDict([(m, mv), (g, gv), (d, dv)]),
                traj_map=OrderedDict(), instance_constraints=instance,
                free=free, expected_con=np.hstack((kin, dyn, inst)),
                expected_jac=np.array(jac), expected_rows=rows,
                expected_cols=cols)


def pendulum_variable_duration():
    """Pendulum with a free node time interval ``h`` and instance constraints
    at integer multiples of ``h``; backward Euler, N = 4.
    opty/tests/test_direct_collocation.py:1746-1790 (set-up), :1878-1910
    (values), :1912-1933 (literal indices), :1935-2039 (Jacobian incl. the
    d/dh column)."""
    m, g, d, t, h = sm.symbols('m, g, d, t, h')
    theta, omega, T = [s(t) for s in sm.symbols('theta, omega, T',
                                                cls=sm.Function)]
    eom = sm.Matrix([theta.diff() - omega,
                     m * d**2 * omega.diff() + m * g * d * sm.sin(theta) - T])
    th, om = sm.symbols('theta, omega', cls=sm.Function)
    instance = (1.0 * th(0 * h), 3.0 * th(3 * h) - sm.pi, 4.0 * om(0 * h),
                5.0 * om(3 * h))
    thv = np.array([1.0, 2.0, 3.0, 4.0])
    omv = np.array([5.0, 6.0, 7.0, 8.0])
    Tv = np.array([9.0, 10.0, 11.0, 12.0])
    mv, gv, dv, hv = 1.0, 9.81, 1.0, 0.01
    free = np.hstack((thv, omv, Tv, [hv]))
    kin = np.array([(thv[i] - thv[i - 1]) / hv - omv[i] for i in (1, 2, 3)])
    dyn = np.array([mv * dv**2 * (omv[i] - omv[i - 1]) / hv +
                    mv * gv * dv * np.sin(thv[i]) - Tv[i] for i in (1, 2, 3)])
    inst = np.array([1.0 * thv[0], 3.0 * thv[3] - np.pi, 4.0 * omv[0],
                     5.0 * omv[3]])
    jac = []
    for i in (1, 2, 3):
        # wrt [thetai, omegai, thetap, omegap, Ti, h]
        jac += [1 / hv, -1.0, -1 / hv, 0.0, 0.0,
                -(thv[i] - thv[i - 1]) / hv**2]
        jac += [mv * gv * dv * np.cos(thv[i]), mv * dv**2 / hv, 0.0,
                -mv * dv**2 / hv, -1.0,
                -mv * dv**2 * (omv[i] - omv[i - 1]) / hv**2]
    jac += [1.0, 3.0, 4.0, 5.0]
    rows = [0, 0, 0, 0, 0, 0, 3, 3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 4, 4, 4, 4,
            4, 4, 2, 2, 2, 2, 2, 2, 5, 5, 5, 5, 5, 5, 6, 7, 8, 9]
    cols = [1, 5, 0, 4, 9, 12, 1, 5, 0, 4, 9, 12, 2, 6, 1, 5, 10, 12, 2, 6,
            1, 5, 10, 12, 3, 7, 2, 6, 11, 12, 3, 7, 2, 6, 11, 12, 0, 3, 4, 7]
    return Case(name='pendulum_variable_duration', eom=eom,
                states=(theta, omega), N=4, h=h, method='backward euler', t=t,
                par_map=OrderedDict([(m, mv), (g, gv), (d, dv)]),
                traj_map=OrderedDict(), instance_constraints=instance,
                free=free, expected_con=np.hstack((kin, dyn, inst)),
                expected_jac=np.array(jac), expected_rows=rows,
                expected_cols=cols)


def single_eom():
    """One equation of motion, one state: shapes of
    opty/tests/test_direct_collocation.py:2337-2393 (M*P is odd here, which
    exercises the non-TMA store path of the kernel)."""
    t, a = sm.symbols('t, a')
    x = sm.Function('x')(t)
    u = sm.Function('u')(t)
    eom = sm.Matrix([x.diff() + a * x**3 - u])
    N, h = 37, 0.05
    rng = np.random.default_rng(11)
    free = rng.standard_normal(2 * N + 1)
    xs, us, av = free[:N], free[N:2 * N], free[-1]
    con = (xs[1:] - xs[:-1]) / h + av * xs[1:]**3 - us[1:]
    jac = np.stack([1 / h + 3 * av * xs[1:]**2, -np.ones(N - 1) / h,
                    -np.ones(N - 1), xs[1:]**3], axis=1).ravel()
    return Case(name='single_eom', eom=eom, states=(x,), N=N, h=h,
                method='backward euler', t=t, par_map=OrderedDict(),
                traj_map=OrderedDict(), instance_constraints=None, free=free,
                expected_con=con, expected_jac=jac, expected_rows=None,
                expected_cols=None)


def implicit_known_trajectory():
    """Known trajectories that are implicit functions of time, theta(x(t)) and
    omega(v(t)), given as callables of ``free`` together with their
    derivatives; free node time interval; backward Euler, N = 4.
    opty/tests/test_direct_collocation.py:18-212."""
    import sympy.physics.mechanics as mech
    m, g, r, h = sm.symbols('m, g, r, h', real=True)
    x, v, f, s = mech.dynamicsymbols('x, v, f, s', real=True)
    t = mech.dynamicsymbols._t
    theta_of_x = sm.Function('theta', real=True)(x)
    omega_of_v = sm.Function('omega', real=True)(v)
    eom = sm.Matrix([x.diff() - v - s + r * omega_of_v,
                     m * v.diff() - f + m * g * sm.sin(theta_of_x)])
    N = 4
    xs = np.linspace(2.0, 5.0, num=N)
    ths = np.linspace(0.0, 10.0, num=N)

    def calc_theta_x(free):
        return np.interp(free[0:N], xs, ths)

    def calc_dtheta_dx(free):
        return np.array([3.9, 1.2, -5.6, 12.3])

    def calc_omega_v(free):
        return np.array([-0.01, -0.98, 3.45, 27.45])

    def calc_domega_dv(free):
        return np.array([0.1, 8.9, -43.4, -2.5])

    traj_map = OrderedDict([
        (omega_of_v.diff(v), calc_domega_dv), (omega_of_v, calc_omega_v),
        (s, np.array([121., 122., 123., 124.])), (theta_of_x, calc_theta_x),
        (theta_of_x.diff(x), calc_dtheta_dx)])
    free = np.array([2., 3., 4., 5., 6., 7., 8., 9., 10., 11., 12., 13., 14.])
    thetas, dthetas = calc_theta_x(free), calc_dtheta_dx(free)
    omegas, domegas = calc_omega_v(free), calc_domega_dv(free)
    con = np.array([
        (3. - 2.) / 14. - 7. - 122. + 7.1 * omegas[1],
        (4. - 3.) / 14. - 8. - 123. + 7.1 * omegas[2],
        (5. - 4.) / 14. - 9. - 124. + 7.1 * omegas[3],
        3.3 * (7. - 6.) / 14. - 11. + 3.3 * 10.2 * np.sin(thetas[1]),
        3.3 * (8. - 7.) / 14. - 12. + 3.3 * 10.2 * np.sin(thetas[2]),
        3.3 * (9. - 8.) / 14. - 13. + 3.3 * 10.2 * np.sin(thetas[3])])
    jac = []
    for i in (1, 2, 3):
        xi, xp = free[i], free[i - 1]
        vi, vp = free[N + i], free[N + i - 1]
        jac += [1. / 14., -1. + 7.1 * domegas[i], -1. / 14., 0., 0.,
                -(xi - xp) / 14.**2]
        jac += [3.3 * 10.2 * np.cos(thetas[i]) * dthetas[i], 3.3 / 14., 0.,
                -3.3 / 14., -1., -3.3 * (vi - vp) / 14.**2]
    return Case(name='implicit_known_trajectory', eom=eom, states=(x, v), N=N,
                h=h, method='backward euler', t=t,
                par_map=OrderedDict([(r, 7.1), (m, 3.3), (g, 10.2)]),
                traj_map=traj_map, instance_constraints=None, free=free,
                expected_con=con, expected_jac=np.array(jac),
                expected_rows=None, expected_cols=None)


def product_cases():
    """Cases for the CUDA path; the last one uses a reference feature the
    oracle does not restate (implicit known trajectories), its expected values
    are the reference test's closed-form numbers."""
    return all_cases() + [implicit_known_trajectory()]


def all_cases():
    return [msd_unknown_trajectory('backward euler'),
            msd_unknown_trajectory('midpoint'),
            pendulum_instance_constraints(), pendulum_variable_duration(),
            single_eom()]


# ---------------------------------------------------------------------------
# objective functions (opty/utils.py:329-470): the forms of the reference's
# tests (opty/tests/test_utils.py:67-220) and a tracking objective of the
# size of BASELINE config 2
# ---------------------------------------------------------------------------
def objective_cases():
    def small(method, which):
        def make():
            t = sm.symbols('t')
            x, v, f1, f2 = [f(t) for f in sm.symbols('x, v, f1, f2',
                                                      cls=sm.Function)]
            m, c, k = sm.symbols('m, c, k')
            exprs = {
                'single_state': sm.Integral(x ** 2, t),
                'single_input': sm.Integral(f1 ** 2, t),
                'single_unknown': m ** 2,
                'all': (sm.Integral(x ** 2 + m ** 2, t) +
                        sm.Integral(c ** 2 * f2 ** 2, t) + sm.sin(k) ** 2),
                'mixed': (3 * sm.Integral(sm.sin(x) * f1 * k +
                                          sm.exp(v * c), t) -
                          2 * sm.Integral(f2 * x, t) + m * k),
            }
            N = 20
            free = np.random.default_rng(11).random(4 * N + 3)
            return dict(objective=exprs[which], states=[x, v],
                        inputs=[f2, f1], params=[m, c, k], N=N, h=0.3,
                        method=method, t=t, free=free)
        return make

    def tracking(method):
        def make():
            t = sm.symbols('t')
            n, N = 22, 10000
            xs = [sm.Function('x{}'.format(i))(t) for i in range(n)]
            F = sm.Function('F')(t)
            w = sm.symbols('w')
            expr = sm.Integral(F ** 2 + sum((1 + 0.1 * i) * (x - 0.25 * i) ** 2
                                            for i, x in enumerate(xs[:11])) +
                               w * sm.cos(xs[11]) ** 2, t) + (w - 2) ** 2
            free = np.random.default_rng(12).standard_normal((n + 1) * N + 1)
            return dict(objective=expr, states=xs, inputs=[F], params=[w],
                        N=N, h=0.001, method=method, t=t, free=free)
        return make

    out = {}
    for method in ('backward euler', 'midpoint'):
        tag = 'be' if method == 'backward euler' else 'mid'
        for which in ('single_state', 'single_input', 'single_unknown', 'all',
                      'mixed'):
            out['{}_{}'.format(which, tag)] = small(method, which)
        out['tracking22_{}'.format(tag)] = tracking(method)
    return out
