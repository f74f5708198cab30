# SPDX-License-Identifier: Apache-2.0
"""ctypes wrapper of the CPU oracle (``oracle/upkie_oracle.cpp``).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module. It binds the struct layouts of
``upkie_b200._abi`` (the public header's mirror) and nothing else from the
product.
"""

import ctypes as C
import os
import subprocess

import numpy as np

from upkie_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libupkie_oracle.so")
_lib = None


REF_SPINE_PATH = os.path.join(_HERE, "_ref", "libupkie_ref_spine.so")


def build_ref(reference: str = "/root/reference") -> str:
    """Compile the reference's own spine observers / controllers in place into ``oracle/_ref/`` (``make ref``) when
    the reference tree is present; returns the library path or "" (GPU box, fresh clone without the reference)."""
    if os.path.isdir(os.path.join(reference, "upkie", "cpp", "observers")):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"REF={reference}"])
    return REF_SPINE_PATH if os.path.exists(REF_SPINE_PATH) else ""


class RefSpine:
    """ctypes wrapper of oracle/_ref/libupkie_ref_spine.so: the REFERENCE'S observer and controller pipelines
    (oracle/ref_spine_shim.cpp). Test infrastructure; raises FileNotFoundError when the library was not built."""

    def __init__(self, observer_config, balancer_config, spine_frequency: int):
        if not os.path.exists(REF_SPINE_PATH):
            raise FileNotFoundError(REF_SPINE_PATH)
        L = C.CDLL(REF_SPINE_PATH)
        L.ref_spine_create.restype = C.c_void_p
        L.ref_spine_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        L.ref_spine_destroy.argtypes = [C.c_void_p]
        L.ref_spine_reset.argtypes = [C.c_void_p]
        L.ref_spine_observers_step.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_spine_controllers_step.argtypes = [C.c_void_p, _dp, _dp, _dp]
        self._L = L
        self._h = L.ref_spine_create(C.byref(observer_config), C.byref(balancer_config) if balancer_config is not None else None,
                                     int(spine_frequency))

    def __del__(self):
        try:
            self._L.ref_spine_destroy(self._h)
        except Exception:
            pass

    def reset(self):
        self._L.ref_spine_reset(self._h)

    def observers_step(self, spine_row):
        s = np.ascontiguousarray(spine_row, dtype=np.float64).reshape(_abi.SPINE_DIM)
        out = np.zeros(_abi.OBSV_DIM)
        self._L.ref_spine_observers_step(self._h, _d(s), _d(out))
        return out

    def controllers_step(self, obs3, target2, action):
        o = np.ascontiguousarray(obs3, dtype=np.float64).reshape(3)
        t = None if target2 is None else np.ascontiguousarray(target2, dtype=np.float64).reshape(2)
        a = np.ascontiguousarray(action, dtype=np.float64).reshape(36).copy()
        self._L.ref_spine_controllers_step(self._h, _d(o), _d(t) if t is not None else None, _d(a))
        return a.reshape(6, 6)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (``oracle/Makefile``)."""
    src = os.path.join(_HERE, "upkie_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "upkie_b200.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH)
        < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libupkie_oracle.so"])
    return _LIB_PATH


_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [
            C.POINTER(_abi.UpkieModel),
            C.POINTER(_abi.UpkieSimConfig),
            C.c_int,
            C.c_int,
        ]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_randomization.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_set_external_forces.argtypes = [C.c_void_p, _dp, C.c_uint32]
        L.oracle_reset.argtypes = [C.c_void_p, _u8p, _dp]
        L.oracle_step_servos.argtypes = [C.c_void_p, _dp, _dp, _dp, _u8p, _u8p]
        L.oracle_step_gyropod.argtypes = [
            C.c_void_p, _dp, C.c_int, _dp, _dp, _u8p, _u8p,
        ]
        L.oracle_reset_obs.argtypes = [C.c_void_p, C.c_int, _dp]
        L.oracle_spine_obs.argtypes = [C.c_void_p, _dp]
        L.oracle_get_state.argtypes = [C.c_void_p, _dp]
        L.oracle_get_body_contacts.argtypes = [C.c_void_p, _dp]
        L.oracle_set_state.argtypes = [C.c_void_p, _dp]
        L.oracle_get_lag.argtypes = [C.c_void_p, _dp]
        L.oracle_set_lag.argtypes = [C.c_void_p, _dp]
        L.oracle_error_flags.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_substep.argtypes = [C.c_void_p, _dp, C.c_double]
        L.oracle_observe.argtypes = [C.c_void_p]
        L.oracle_energy.argtypes = [C.c_void_p, C.c_int, _dp]
        L.oracle_mass_com.argtypes = [C.c_void_p, C.c_int, _dp]
        L.oracle_compute_joint_torque.restype = C.c_double
        L.oracle_compute_joint_torque.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 8
        L.oracle_low_pass_filter.restype = C.c_double
        L.oracle_low_pass_filter.argtypes = [C.c_double] * 4
        L.oracle_clamp.restype = C.c_double
        L.oracle_clamp.argtypes = [C.c_double] * 3
        L.oracle_rotation_matrix_from_quaternion.argtypes = [_dp, _dp]
        L.oracle_quaternion_from_rotation_matrix.argtypes = [_dp, _dp]
        L.oracle_mpc_create.restype = C.c_void_p
        L.oracle_mpc_create.argtypes = [C.POINTER(_abi.UpkieMpcConfig)]
        L.oracle_mpc_destroy.argtypes = [C.c_void_p]
        L.oracle_mpc_matrices.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.oracle_mpc_cost_vector.argtypes = [C.c_void_p, _dp, C.c_double, _dp]
        L.oracle_mpc_solve.restype = C.c_int
        L.oracle_mpc_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_mpc_step.argtypes = [
            C.c_void_p, C.c_int, _dp, _dp, _u8p, C.c_double, _dp, _dp, _u8p, _dp, C.c_int,
        ]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _u8(a):
    return a.ctypes.data_as(_u8p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleSim:
    """N independent robots stepped by the CPU oracle (fp64 by default)."""

    def __init__(self, model, config=None, n_envs=1, use_float=False, threads=1):
        self.model = model
        self.config = config if config is not None else _abi.default_sim_config()
        self.n = int(n_envs)
        self._m = model.to_struct()
        self._h = lib().oracle_create(
            C.byref(self._m), C.byref(self.config), self.n, 1 if use_float else 0
        )
        lib().oracle_set_threads(self._h, int(threads))

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, threads):
        lib().oracle_set_threads(self._h, int(threads))

    def set_randomization(self, friction=None, inertia_eps=None):
        f = _f64(friction, (self.n,)) if friction is not None else None
        e = _f64(inertia_eps, (self.n, 6)) if inertia_eps is not None else None
        lib().oracle_set_randomization(
            self._h, _d(f) if f is not None else None, _d(e) if e is not None else None
        )

    def set_external_forces(self, force=None, local_mask=0):
        """``force[n, 7, 3]`` newtons at the bodies' centres of mass (None clears); bit i of
        ``local_mask``: the force on body i is expressed in the body frame."""
        f = _f64(force, (self.n, 7, 3)) if force is not None else None
        lib().oracle_set_external_forces(self._h, _d(f) if f is not None else None, int(local_mask))

    def reset(self, init_state, mask=None):
        init = _f64(init_state, (self.n, _abi.INIT_DIM))
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8).reshape(self.n)
        lib().oracle_reset(self._h, _u8(m) if m is not None else None, _d(init))

    def step_servos(self, action):
        a = _f64(action, (self.n, _abi.ACT_DIM))
        obs = np.empty((self.n, 6, 5))
        rew = np.empty(self.n)
        term = np.empty(self.n, dtype=np.uint8)
        trunc = np.empty(self.n, dtype=np.uint8)
        lib().oracle_step_servos(self._h, _d(a), _d(obs), _d(rew), _u8(term), _u8(trunc))
        return obs, rew, term, trunc

    def step_gyropod(self, action, act_dim):
        a = _f64(action, (self.n, act_dim))
        od = 4 if act_dim == 1 else 6
        obs = np.empty((self.n, od))
        rew = np.empty(self.n)
        term = np.empty(self.n, dtype=np.uint8)
        trunc = np.empty(self.n, dtype=np.uint8)
        lib().oracle_step_gyropod(
            self._h, _d(a), act_dim, _d(obs), _d(rew), _u8(term), _u8(trunc)
        )
        return obs, rew, term, trunc

    def reset_obs(self, obs_dim):
        shape = (self.n, 6, 5) if obs_dim == 30 else (self.n, obs_dim)
        obs = np.empty(shape)
        lib().oracle_reset_obs(self._h, obs_dim, _d(obs))
        return obs

    def spine_obs(self):
        out = np.empty((self.n, _abi.SPINE_DIM))
        lib().oracle_spine_obs(self._h, _d(out))
        return out

    def get_state(self):
        out = np.empty((self.n, _abi.STATE_DIM))
        lib().oracle_get_state(self._h, _d(out))
        return out

    def get_lag(self):
        """Spine mode: lag records ``[n, LAG_DIM]`` (the first 49 entries: replies of the last two cycles, last IMU
        reading; the oracle keeps the assembled observation in its spine row, so the rest reads 0)."""
        out = np.zeros((self.n, _abi.LAG_DIM))
        lib().oracle_get_lag(self._h, _d(out))
        return out

    def set_lag(self, lag):
        a = np.ascontiguousarray(lag, dtype=np.float64).reshape(self.n, _abi.LAG_DIM)
        lib().oracle_set_lag(self._h, _d(a))

    def set_state(self, state):
        s = _f64(state, (self.n, _abi.STATE_DIM))
        lib().oracle_set_state(self._h, _d(s))

    def error_flags(self):
        out = np.empty(self.n, dtype=np.uint32)
        lib().oracle_error_flags(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def substep(self, tau, h):
        t = _f64(tau, (self.n, 6))
        lib().oracle_substep(self._h, _d(t), float(h))

    def observe(self):
        lib().oracle_observe(self._h)

    def energy(self, env=0):
        out = np.empty(8)
        lib().oracle_energy(self._h, env, _d(out))
        return {"kinetic": out[0], "potential": out[1], "linear_momentum": out[2:5].copy(),
                "angular_momentum": out[5:8].copy()}

    def mass_com(self, env=0):
        out = np.empty(4)
        lib().oracle_mass_com(self._h, env, _d(out))
        return out[0], out[1:4].copy()

    def compute_joint_torque(self, joint, q, qd, feedforward_torque, target_position,
                             target_velocity, kp_scale, kd_scale, maximum_torque):
        return lib().oracle_compute_joint_torque(
            self._h, joint, q, qd, feedforward_torque, target_position,
            target_velocity, kp_scale, kd_scale, maximum_torque,
        )


def low_pass_filter(prev_output, cutoff_period, new_input, dt):
    return lib().oracle_low_pass_filter(prev_output, cutoff_period, new_input, dt)


def clamp(value, lower, upper):
    return lib().oracle_clamp(value, lower, upper)


def rotation_matrix_from_quaternion(quat):
    q = _f64(quat, (4,))
    R = np.empty(9)
    lib().oracle_rotation_matrix_from_quaternion(_d(q), _d(R))
    return R.reshape(3, 3)


def quaternion_from_rotation_matrix(R):
    r = _f64(R, (9,))
    q = np.empty(4)
    lib().oracle_quaternion_from_rotation_matrix(_d(r), _d(q))
    return q


def sample_init_state(config, np_random):
    """``RobotState.sample_state`` (``upkie/utils/robot_state.py:175-196``) with
    ``RobotStateRandomization.sample_*``
    (``upkie/utils/robot_state_randomization.py:135-189``), restated in NumPy.

    Draw order: angular velocity (3), linear velocity (3), ZYX euler (3),
    position (3), each one ``np_random.uniform(low, high, size=3)``. Returns an
    ``init_state[25]`` row.
    """
    c = config
    om = np_random.uniform(
        low=np.array([-c.rand_omega_x, -c.rand_omega_y, 0.0]),
        high=np.array([+c.rand_omega_x, +c.rand_omega_y, 0.0]),
        size=3,
    )
    lv = np.array(list(c.rand_linear_velocity))
    v = np_random.uniform(low=-lv, high=lv, size=3)
    ypr_b = np.array([0.0, c.rand_pitch, c.rand_roll])
    ypr = np_random.uniform(low=-ypr_b, high=+ypr_b, size=3)
    pos = np_random.uniform(
        low=np.array([-c.rand_x, 0.0, 0.0]), high=np.array([+c.rand_x, 0.0, c.rand_z]), size=3
    )
    # intrinsic ZYX euler -> quaternion, q = qz(yaw) * qy(pitch) * qx(roll)
    cy, sy = np.cos(ypr[0] / 2), np.sin(ypr[0] / 2)
    cp, sp = np.cos(ypr[1] / 2), np.sin(ypr[1] / 2)
    cr, sr = np.cos(ypr[2] / 2), np.sin(ypr[2] / 2)
    q_rand = np.array(
        [
            cy * cp * cr + sy * sp * sr,
            cy * cp * sr - sy * sp * cr,
            cy * sp * cr + sy * cp * sr,
            sy * cp * cr - cy * sp * sr,
        ]
    )
    q0 = np.array(list(c.init_quat))

    def qmul(a, b):
        return np.array(
            [
                a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
            ]
        )

    q = qmul(q0, q_rand)  # rotation_base_to_world * rotation_rand_to_base
    out = np.zeros(_abi.INIT_DIM)
    out[_abi.INIT_POS:_abi.INIT_POS + 3] = np.array(list(c.init_position)) + pos
    out[_abi.INIT_QUAT:_abi.INIT_QUAT + 4] = q
    # nominal + random part; the joint configuration is kept, joint velocities are not set by the reset
    out[_abi.INIT_LINVEL:_abi.INIT_LINVEL + 3] = np.array(list(c.init_linear_velocity)) + v
    out[_abi.INIT_ANGVEL:_abi.INIT_ANGVEL + 3] = np.array(list(c.init_angular_velocity)) + om
    out[_abi.INIT_Q:_abi.INIT_Q + 6] = np.array(list(c.init_joint_configuration))
    return out


class OracleMpc:
    """``MPCBalancer`` restated (condensed QP + exact active-set, fp64)."""

    def __init__(self, config=None):
        self.config = config if config is not None else _abi.default_mpc_config()
        self._h = lib().oracle_mpc_create(C.byref(self.config))
        self.N = int(self.config.nb_timesteps)

    def close(self):
        if self._h:
            lib().oracle_mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def matrices(self):
        P = np.empty((self.N, self.N))
        A = np.empty((4, 4))
        B = np.empty(4)
        lib().oracle_mpc_matrices(self._h, _d(P), _d(A), _d(B))
        return P, A, B

    def cost_vector(self, x0, v_target):
        x = _f64(x0, (4,))
        q = np.empty(self.N)
        lib().oracle_mpc_cost_vector(self._h, _d(x), float(v_target), _d(q))
        return q

    def solve(self, q):
        qq = _f64(q, (self.N,))
        U = np.empty(self.N)
        ok = lib().oracle_mpc_solve(self._h, _d(qq), _d(U))
        return U, bool(ok)

    def step(self, x0, v_target, floor_contact, dt, v_cmd, threads=1):
        x = _f64(x0)
        n = x.shape[0]
        vt = _f64(v_target, (n,))
        fc = np.ascontiguousarray(floor_contact, dtype=np.uint8).reshape(n)
        vc = _f64(v_cmd, (n,)).copy()
        first = np.empty(n)
        found = np.empty(n, dtype=np.uint8)
        plan = np.empty((n, self.N))
        lib().oracle_mpc_step(
            self._h, n, _d(x), _d(vt), _u8(fc), float(dt), _d(vc), _d(first), _u8(found),
            _d(plan), int(threads),
        )
        return vc, first, found, plan
