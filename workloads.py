"""The BASELINE.json workloads: symbolic equations of motion, known values and
seeded free vectors, shared by ``bench.py``, the tests, the oracle scripts and
``__graft_entry__``.  Nothing here depends on the reference or on CUDA.

Each builder returns a :class:`Workload` whose ``collocator_kwargs()`` can be
passed to ``opty_b200.ConstraintCollocator`` -- and, unchanged, to the
reference's ``opty.direct_collocation.ConstraintCollocator`` (that is how the
golden vectors under ``tests/golden/`` were produced, see
``tests/golden/make_golden.py``).
"""

from collections import OrderedDict

import numpy as np
import sympy as sm
import sympy.physics.mechanics as me


class Workload(object):

    def __init__(self, name, eom, states, num_nodes, interval, method,
                 known_parameter_map=None, known_trajectory_map=None,
                 instance_constraints=None, time_symbol=None, seed=0,
                 free=None):
        self.name = name
        self.eom = eom
        self.states = tuple(states)
        self.num_nodes = num_nodes
        self.interval = interval
        self.method = method
        self.known_parameter_map = known_parameter_map or OrderedDict()
        self.known_trajectory_map = known_trajectory_map or OrderedDict()
        self.instance_constraints = instance_constraints
        self.time_symbol = time_symbol
        self.seed = seed
        self._free = free

    def collocator_args(self):
        return (self.eom, self.states, self.num_nodes, self.interval)

    def collocator_kwargs(self):
        kw = dict(known_parameter_map=self.known_parameter_map,
                  known_trajectory_map=self.known_trajectory_map,
                  instance_constraints=self.instance_constraints,
                  integration_method=self.method)
        if self.time_symbol is not None:
            kw['time_symbol'] = self.time_symbol
        return kw

    def free(self, num_free):
        """Seeded free vector.  For workloads built by ``n_link_pendulum``
        the draw continues the generator that produced the constants, as
        SURVEY.md §8(d) specifies."""
        if self._free is not None:
            return self._free(num_free)
        return np.random.default_rng(self.seed).standard_normal(num_free)


def n_link_pendulum(links=10, num_nodes=10000, interval=0.001,
                    method='midpoint', seed=0, name=None):
    """BASELINE config 2 (links=10, N=10 000) and 5 (links=50, N=50 000):
    n-link pendulum on a cart, all constants known, cart force unknown.

    Pattern of opty/tests/test_direct_collocation.py:2044-2048 with
    ``sympy.physics.mechanics.models.n_link_pendulum_on_cart``."""
    from sympy.physics.mechanics.models import n_link_pendulum_on_cart
    me.dynamicsymbols._t = sm.Symbol('t')
    kane = n_link_pendulum_on_cart(n=links, cart_force=True,
                                   joint_torques=False)
    states = kane.q.col_join(kane.u)
    eom = kane.mass_matrix_full @ states.diff() - kane.forcing_full
    t = me.dynamicsymbols._t
    constants = sorted((s for s in eom.free_symbols if s != t),
                       key=lambda s: s.name)
    rng = np.random.default_rng(seed)
    par_map = OrderedDict()
    for c in constants:
        par_map[c] = 9.81 if c.name == 'g' else 0.5 + rng.random()
    return Workload(name or 'pendulum{}_N{}'.format(links, num_nodes),
                    eom, list(states), num_nodes, interval, method,
                    known_parameter_map=par_map, seed=seed,
                    free=lambda nf: rng.standard_normal(nf))


def pendulum_swing_up(num_nodes=51, seed=1):
    """BASELINE config 1: single pendulum swing-up, 2 states, backward Euler,
    4 instance constraints (opty/tests/test_direct_collocation.py:471-519)."""
    duration = 10.0
    interval = duration / (num_nodes - 1)
    I, m, g, h, t = sm.symbols('I, m, g, h, t', real=True)
    theta, omega, T = sm.symbols('theta, omega, T', cls=sm.Function)
    states = (theta(t), omega(t))
    eom = sm.Matrix([theta(t).diff() - omega(t),
                     I * omega(t).diff() + m * g * h * sm.sin(theta(t)) -
                     T(t)])
    par_map = OrderedDict([(I, 1.0), (m, 1.0), (g, 9.81), (h, 1.0)])
    instance = (theta(0.0), theta(duration) - np.pi, omega(0.0),
                omega(duration))
    return Workload('pendulum_swing_up_N{}'.format(num_nodes), eom, states,
                    num_nodes, interval, 'backward euler',
                    known_parameter_map=par_map,
                    instance_constraints=instance, time_symbol=t, seed=seed)


def vyasarayani2011(num_nodes=5000, seed=3):
    """BASELINE config 3: pendulum parameter identification, one unknown
    parameter, midpoint (examples/vyasarayani2011.py:44-53, 83-85)."""
    duration = 50.0
    interval = duration / (num_nodes - 1)
    p, t = sm.symbols('p, t')
    y1, y2 = [f(t) for f in sm.symbols('y1, y2', cls=sm.Function)]
    y = sm.Matrix([y1, y2])
    eom = y.diff(t) - sm.Matrix([y2, -p * sm.sin(y1)])
    return Workload('vyasarayani2011_N{}'.format(num_nodes), eom, (y1, y2),
                    num_nodes, interval, 'midpoint', time_symbol=t,
                    seed=seed)


def n_link_pendulum_torques(links=4, num_nodes=2000, method='backward euler',
                            seed=4, name=None):
    """Stand-in for BASELINE config 4 (the human-gait EOM needs ``pygait2d``,
    which is not installable offline): an n-link pendulum
    with unknown joint torques, one known input trajectory (the cart force),
    some unknown parameters, a free node time interval and instance
    constraints -- the same structural class (q > 1, k > 0, r > 0, s = 1,
    o > 0, backward Euler)."""
    from sympy.physics.mechanics.models import n_link_pendulum_on_cart
    me.dynamicsymbols._t = sm.Symbol('t')
    kane = n_link_pendulum_on_cart(n=links, cart_force=True,
                                   joint_torques=True)
    states = kane.q.col_join(kane.u)
    eom = kane.mass_matrix_full @ states.diff() - kane.forcing_full
    t = me.dynamicsymbols._t
    constants = sorted((s for s in eom.free_symbols if s != t),
                       key=lambda s: s.name)
    rng = np.random.default_rng(seed)
    par_map = OrderedDict()
    for c in constants:
        if c.name in ('m0', 'l0'):
            continue                      # left unknown: r = 2
        par_map[c] = 9.81 if c.name == 'g' else 0.5 + rng.random()
    force = [f for f in me.find_dynamicsymbols(eom)
             if getattr(f, 'name', None) == 'F'][0]
    time = np.linspace(0.0, 1.0, num_nodes)
    traj_map = OrderedDict([(force, np.sin(3.0 * time))])
    h = sm.Symbol('h', real=True)
    q = list(kane.q)
    u = list(kane.u)
    instance = (q[0].subs(t, 0 * h), q[1].subs(t, 0 * h) - 0.5,
                u[0].subs(t, 0 * h),
                q[1].subs(t, (num_nodes - 1) * h) - 1.0,
                u[1].subs(t, (num_nodes - 1) * h))

    def draw(nf):
        free = rng.standard_normal(nf)
        free[-1] = 0.01 + 0.01 * rng.random()     # positive time interval
        free[-3:-1] = 0.5 + rng.random(2)         # positive mass / length
        return free
    return Workload(name or 'pendulum{}_torques_N{}'.format(links, num_nodes),
                    eom, list(states), num_nodes, h, method,
                    known_parameter_map=par_map,
                    known_trajectory_map=traj_map,
                    instance_constraints=instance, seed=seed, free=draw)


def n_link_pendulum_periodic(links=4, num_nodes=200, seed=6):
    """The config-4 stand-in with the PERIODICITY instance constraints of the
    human-gait example (examples-gallery/advanced/plot_human_gait.py:163-184):
    free node time interval, constraints that tie a state at the first node
    to a (different) state at the last node -- two function atoms each, so
    their Jacobian entries follow the reference's iteration over
    ``con.atoms(sm.Function)`` (opty/direct_collocation.py:2244, 2264) -- next
    to single-atom ones and one with a product."""
    w = n_link_pendulum_torques(links, num_nodes, seed=seed,
                                name='pendulum{}_periodic_N{}'.format(
                                    links, num_nodes))
    h = w.interval
    t = me.dynamicsymbols._t
    duration = (num_nodes - 1) * h
    half = len(w.states) // 2
    q, u = w.states[:half], w.states[half:]
    speed = 1.3

    def at(f, when):
        return f.subs(t, when)
    instance = [at(q[0], 0 * h) - 0.0,
                at(q[1], 0 * h) - 0.0,
                at(q[1], duration) - speed * at(q[0], duration),
                at(q[2], 0 * h) - at(q[2], duration)]
    # left / right swap pattern: q_a(0) = q_b(T), q_b(0) = q_a(T)
    for a, b in ((3, 4),):
        instance += [at(q[a], 0 * h) - at(q[b], duration),
                     at(q[b], 0 * h) - at(q[a], duration)]
    instance += [at(u[0], 0 * h) - at(u[0], duration),
                 at(u[1], 0 * h) - at(u[2], duration),
                 at(u[2], 0 * h) - at(u[1], duration),
                 at(u[3], 0 * h) * at(u[4], duration) - 0.25]
    w.instance_constraints = tuple(instance)
    return w
