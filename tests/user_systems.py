"""User-defined System subclasses used by the tests of the codegen path.

Each factory takes the `System` base class and the array namespace, so that the SAME method bodies run
  * under the UNMODIFIED reference (`/root/reference/python/class_files/systems/system_base.py` + jax.numpy,
    via oracle/jaxshim in tests/golden/make_golden.py) to produce golden vectors, and
  * under this package (`class_files.systems.system_base.System` + `class_files.symbolic`) on the GPU.
"""
import math

CARTPOLE = dict(dt=0.01, x_target=[0.0, math.pi, 0.0, 0.0], Q=[1.0, 5.0, 0.1, 0.1], R=[0.1],
                Q_f=[50.0, 200.0, 10.0, 10.0], mc=1.0, mp=0.2, l=0.5, g=9.81, b=0.1, p_max=1.5, w_bar=0.5)
SPRING = dict(dt=0.01, x_target=[math.pi, 0.0], Q=[1.0, 0.1], R=[0.5], Q_f=[50.0, 5.0], a=0.8, ks=3.0)


def make_cartpole_class(System, jnp):
    class MyCartPole(System):
        """Cart-pole, x = [p, theta, p_dot, theta_dot] (theta = 0 hanging down), u = [force].
        Stage cost = quadratic + a smooth (exponential) barrier keeping the cart inside |p| < p_max, so
        l_xx depends on the state -- a cost the shipped quadratic device cost cannot express (the kind
        of term the reference leaves commented out at pendulum_sys.py:84-85)."""

        def __init__(self, dt, x_target, Q, R, Q_f, mc=1.0, mp=0.2, l=0.5, g=9.81, b=0.1, p_max=1.5, w_bar=0.5,
                     use_jit=True, integrator="rk4", **kw):
            self.mc, self.mp, self.l, self.g, self.b = mc, mp, l, g, b
            self.p_max, self.w_bar = p_max, w_bar
            self.x_target, self.Q, self.R, self.Q_f = x_target, Q, R, Q_f
            super().__init__(n_x=4, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _f_cont_fcn(self, x, u):
            p, th, pd, thd = x[0], x[1], x[2], x[3]
            s, c = jnp.sin(th), jnp.cos(th)
            den = self.mc + self.mp * s * s
            pdd = (u[0] + self.mp * s * (self.l * thd * thd + self.g * c) - self.b * pd) / den
            thdd = (-u[0] * c - self.mp * self.l * thd * thd * c * s - (self.mc + self.mp) * self.g * s
                    + self.b * pd * c) / (self.l * den)
            return jnp.array([pd, thd, pdd, thdd])

        def _l_fcn(self, x, u):
            dx = x - self.x_target
            quad = 0.5 * dx.T @ self.Q @ dx + 0.5 * u.T @ self.R @ u
            barrier = self.w_bar * jnp.exp(4.0 * (x[0] * x[0] - self.p_max * self.p_max))
            return (quad + barrier) * self.dt

        def _l_f_fcn(self, x):
            dx = x - self.x_target
            return 0.5 * dx.T @ self.Q_f @ dx

    return MyCartPole


def make_user_ua_class(System, jnp):
    class UserUADoublePendulum(System):
        """The under-actuated double pendulum written as a USER system (M(q) qdd = h solved in closed form),
        to check the generated device code against the shipped device model and the reference's goldens."""

        def __init__(self, dt, x_target, Q, R, Q_f, g=9.81, m1=1.0, m2=1.0, l1=1.0, l2=1.0, d1=0.01, d2=0.01,
                     theta1=0.0, theta2=0.0, use_jit=True, integrator="rk4", **kw):
            self.ph = (g, m1, m2, l1, l2, d1, d2, theta1, theta2)
            self.x_target, self.Q, self.R, self.Q_f = x_target, Q, R, Q_f
            super().__init__(n_x=4, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _f_cont_fcn(self, x, u):
            g, m1, m2, l1, l2, d1, d2, th1, th2 = self.ph
            q1, q2, q1d, q2d = x[0], x[1], x[2], x[3]
            c = m2 * l1 * l2
            m11 = m1 * l1 * l1 / 4 + m2 * l1 * l1 + m2 * l2 * l2 / 4 + th1 + th2 + c * jnp.cos(q2)
            m22 = m2 * l2 * l2 / 4 + th2
            m12 = m22 + 0.5 * c * jnp.cos(q2)
            s12 = jnp.sin(q1 + q2)
            h1 = (u[0] + 0.5 * c * jnp.sin(q2) * (2 * q1d * q2d + q2d * q2d) - m2 * g * l2 / 2 * s12
                  - (m2 * g * l1 + m1 * g * l1 / 2) * jnp.sin(q1) - d1 * q1d)
            h2 = -0.5 * c * jnp.sin(q2) * q1d * q1d - m2 * g * l2 / 2 * s12 - d2 * q2d
            det = m11 * m22 - m12 * m12
            return jnp.array([q1d, q2d, (m22 * h1 - m12 * h2) / det, (m11 * h2 - m12 * h1) / det])

        def _l_fcn(self, x, u):
            dx = x - self.x_target
            return (0.5 * dx.T @ self.Q @ dx + 0.5 * u.T @ self.R @ u) * self.dt

        def _l_f_fcn(self, x):
            dx = x - self.x_target
            return 0.5 * dx.T @ self.Q_f @ dx

    return UserUADoublePendulum


def make_saturated_pendulum_class(System, jnp):
    class SaturatedPendulum(System):
        """A pendulum with data-dependent selects in BOTH user methods: the torque saturates (jnp.clip) and the cost has
        a one-sided quadratic wall (jnp.where) -- what the reference would write with jnp.where / lax.cond."""

        def __init__(self, dt, x_target, Q, R, Q_f, u_max=2.0, wall=1.0, use_jit=True, integrator="rk4", **kw):
            self.x_target, self.Q, self.R, self.Q_f, self.u_max, self.wall = x_target, Q, R, Q_f, u_max, wall
            super().__init__(n_x=2, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _f_cont_fcn(self, x, u):
            tau = jnp.clip(u[0], -self.u_max, self.u_max)
            return jnp.array([x[1], tau - 0.05 * x[1] - 9.81 * jnp.sin(x[0])])

        def _l_fcn(self, x, u):
            dx = x - self.x_target
            over = jnp.where(x[1] > self.wall, (x[1] - self.wall) ** 2, 0.0)
            return (0.5 * dx.T @ self.Q @ dx + 0.5 * u.T @ self.R @ u + 10.0 * over) * self.dt

        def _l_f_fcn(self, x):
            dx = x - self.x_target
            return 0.5 * dx.T @ self.Q_f @ dx

    return SaturatedPendulum


def make_implicit_spring_class(System, jnp, lax):
    class ImplicitSpringPendulum(System):
        """A pendulum whose joint spring is given IMPLICITLY: the deflection y solves y + a y^3 = r(x) and is found by a
        Newton iteration inside _f_cont_fcn -- a lax.while_loop whose trip count depends on the state (what the
        reference's own integrator does at system_base.py:139).  lax.cond switches the damping law with the sign of the
        velocity."""

        def __init__(self, dt, x_target, Q, R, Q_f, a=0.8, ks=3.0, use_jit=True, integrator="rk4", **kw):
            self.x_target, self.Q, self.R, self.Q_f, self.a, self.ks = x_target, Q, R, Q_f, a, ks
            super().__init__(n_x=2, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _deflection(self, r):
            a = self.a

            def cond(c):
                y, k = c
                return (jnp.abs(y + a * y * y * y - r) > 1e-13) & (k < 30)

            def body(c):
                y, k = c
                return (y - (y + a * y * y * y - r) / (1.0 + 3.0 * a * y * y), k + 1)

            y, _ = lax.while_loop(cond, body, (r / (1.0 + a * r * r), 0))
            return y

        def _f_cont_fcn(self, x, u):
            y = self._deflection(jnp.sin(x[0]) + 0.5 * x[1])
            damp = lax.cond(x[1] > 0.0, lambda w: 0.05 * w, lambda w: 0.15 * w, x[1])
            return jnp.array([x[1], u[0] - damp - 9.81 * jnp.sin(x[0]) - self.ks * y])

        def _l_fcn(self, x, u):
            dx = x - self.x_target
            return (0.5 * dx.T @ self.Q @ dx + 0.5 * u.T @ self.R @ u) * self.dt

        def _l_f_fcn(self, x):
            dx = x - self.x_target
            return 0.5 * dx.T @ self.Q_f @ dx

    return ImplicitSpringPendulum


def make_rich_math_class(System, jnp):
    class RichMath(System):
        """n = 3, m = 1; exercises the wider jnp surface: cross, linalg.det, arctan, arcsin, sinh, cosh, sign, trace, mean.
        Written so that the same body runs with numpy as `jnp` (the tests' expected values)."""

        def __init__(self, dt=0.01, use_jit=True, integrator="rk4", **kw):
            super().__init__(n_x=3, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _f_cont_fcn(self, x, u):
            c = jnp.cross(jnp.array([x[0], x[1], x[2]]), jnp.array([0.0, 0.0, 1.0]))          # [x1, -x0, 0]
            M = jnp.array([[1.5 + jnp.cosh(0.2 * x[0]), 0.3 * x[1]], [0.3 * x[1], 2.0]])
            sat = jnp.arctan(u[0])                                                            # smooth actuator saturation
            fric = 0.1 * jnp.sign(x[2]) * jnp.sinh(0.3 * x[2]) / jnp.cosh(0.3 * x[2])
            return jnp.array([c[0] - 0.2 * x[0] + 0.05 * jnp.trace(M),
                              c[1] - 0.2 * x[1] + jnp.arcsin(0.5 * jnp.sin(x[2])),
                              (sat - fric) / jnp.linalg.det(M) - 0.1 * jnp.mean(jnp.array([x[0], x[1], x[2]]))])

        def _l_fcn(self, x, u):
            return (0.5 * (x[0] * x[0] + x[1] * x[1] + 0.1 * x[2] * x[2]) + 0.05 * u[0] * u[0]) * self.dt

        def _l_f_fcn(self, x):
            return 5.0 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2])

    return RichMath


def make_two_loop_class(System, jnp, lax):
    class TwoLoops(System):
        """n = 2, m = 1: two lax.while_loops in sequence -- the second starts from the first one's result and its body
        reads it again -- and a fori_loop with trace-time bounds; exercises the chain rule through an earlier loop's
        tangents."""

        def __init__(self, dt=0.01, use_jit=True, integrator="rk4", **kw):
            super().__init__(n_x=2, n_u=1, dt=dt, use_jit=use_jit, integrator=integrator, **kw)

        def _f_cont_fcn(self, x, u):
            r = 1.5 + jnp.sin(x[0]) + 0.3 * u[0]                      # in [0.2, 2.8] for |u| <= 1
            # square root of r by Heron's iteration
            s, _ = lax.while_loop(lambda c: (jnp.abs(c[0] * c[0] - r) > 1e-14) & (c[1] < 40),
                                  lambda c: (0.5 * (c[0] + r / c[0]), c[1] + 1), (0.5 * (1.0 + r), 0))
            # fixed point z = cos(s * z) * 0.5 + 0.1 * x[1], started at s / 4
            z, _ = lax.while_loop(lambda c: (jnp.abs(c[0] - (0.5 * jnp.cos(s * c[0]) + 0.1 * x[1])) > 1e-14) & (c[1] < 200),
                                  lambda c: (0.5 * jnp.cos(s * c[0]) + 0.1 * x[1], c[1] + 1), (0.25 * s, 0))
            p = lax.fori_loop(0, 3, lambda i, v: v * z + (i + 1.0), 0.0)     # Horner: z^2 + 2 z + 3
            return jnp.array([x[1], u[0] - s * x[0] - p])

        def _l_fcn(self, x, u):
            return (0.5 * (x[0] * x[0] + 0.1 * x[1] * x[1]) + 0.05 * u[0] * u[0]) * self.dt

        def _l_f_fcn(self, x):
            return 5.0 * (x[0] * x[0] + x[1] * x[1])

    return TwoLoops
