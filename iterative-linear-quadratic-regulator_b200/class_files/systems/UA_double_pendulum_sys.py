"""MyUADoublePendulum -- drop-in for the reference's class_files/systems/UA_double_pendulum_sys.py:9-207.

Under-actuated double pendulum: same M(q), h as the fully actuated one but a single torque on joint 1,
f_act = [tau[0], 0] (UA_double_pendulum_sys.py:204), so n_u = 1.  Runs as DoublePendulumSys<T,1>.
"""
from .double_pendulum_sys import MyDoublePendulum


class MyUADoublePendulum(MyDoublePendulum):
    _MODEL = "ua_double_pendulum"
    _N_U = 1
