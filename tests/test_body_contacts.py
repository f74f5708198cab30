# SPDX-License-Identifier: Apache-2.0
"""Body-ground contacts (config.body_contacts): a fallen robot rests on the collision points of its torso instead of
passing through the floor, as every link with a <collision> does in Bullet (pybullet_backend.py:115,121,306).

CPU part: the fp64 oracle (generic row solver over the full kinematic tree) against the kernels' arithmetic compiled
for the host (gate + general_contact_solve of sim_pair.cuh), plus what can be pinned analytically: the resting height
is the box face, the contact impulses carry the weight, a robot that never comes near the ground is bit-identical
with the rows on and off. GPU part: the same runs through the C ABI on the device.
"""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from hostsim_wrap import HostSim
from upkie_b200 import _abi
from upkie_b200.model import Model, body_contact_points, body_points_in_world, contact_points_from_state


def _oracle_rec(oracle_lib, osim):
    out = np.zeros((osim.n, _abi.BODY_REC_DIM))
    oracle_lib.lib().oracle_get_body_contacts(osim._h, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _on():
    """Default configuration with the body-ground contact rows on (what B200Backend runs; batched handles opt in)."""
    cfg = _abi.default_sim_config()
    assert cfg.body_contacts == 0 and cfg.joint_limits == 3
    cfg.body_contacts = 1
    return cfg


def _falling_setup(model, n, pitch_lo=0.05, pitch_hi=0.4):
    init = np.zeros((n, _abi.INIT_DIM))
    init[:, 2] = 0.6
    pitch = np.linspace(pitch_lo, pitch_hi, n) * np.where(np.arange(n) % 2, 1.0, -1.0)
    init[:, 3], init[:, 5] = np.cos(pitch / 2), np.sin(pitch / 2)
    act = np.zeros((n, 6, 6), dtype=np.float32)
    act[:, :, 0] = np.nan  # no position target: velocity damping only, the robot falls
    act[:, :, 3] = act[:, :, 4] = 1.0
    act[:, :, 5] = np.asarray(model.tau_max, dtype=np.float32)
    return init, act


def test_model_carries_the_torso_box(model):
    assert len(model.collision_body) == 8 and set(model.collision_link) == {"torso"}
    s = model.to_struct()
    assert s.n_collision_points == 8 and list(s.collision_body[:8]) == [0] * 8
    # the box the base inertia was authored from, centred on the base lump's centre of mass
    ext = model.collision_point.max(axis=0) - model.collision_point.min(axis=0)
    np.testing.assert_allclose(ext, [0.14, 0.22, 0.22], atol=1e-12)
    np.testing.assert_allclose(model.collision_point.mean(axis=0), model.com[0], atol=1e-12)


def test_urdf_round_trip_of_collision_shapes():
    from upkie_b200.urdf import write_urdf

    m = Model.standard_upkie()
    m.add_collision_shape("left_lower_leg", 2, "sphere", [0.03], [0.0, 0.0, -0.02])
    m.add_collision_shape("right_upper_leg", 4, "capsule", [0.02, 0.1], [0.0, 0.01, -0.08])
    for split in (True, False):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "u.urdf")
            write_urdf(m, path, split)
            m2 = Model.from_urdf(path)
        np.testing.assert_allclose(m2.collision_point, m.collision_point, atol=1e-12)
        np.testing.assert_array_equal(m2.collision_body, m.collision_body)
        np.testing.assert_allclose(m2.collision_radius, m.collision_radius, atol=1e-15)
        assert m2.collision_link[8:] == ["left_lower_leg", "right_upper_leg", "right_upper_leg"]
    with pytest.raises(ValueError):
        m.add_collision_shape("torso", 0, "box", (0.1, 0.1, 0.1), [0, 0, 0])  # 11 + 8 > UPKIE_MAX_COLLISION_POINTS
    # a URDF with more shapes than UpkieModel carries still loads: the surplus is reported and does not collide
    m3 = Model.standard_upkie()
    m3.add_collision_shape("left_lower_leg", 2, "box", (0.02, 0.02, 0.1), [0.0, 0.0, -0.08])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "u.urdf")
        write_urdf(m3, path, True)
        text = open(path).read().replace(
            '<link name="right_lower_leg">',
            '<link name="right_lower_leg"><collision><origin xyz="0 0 0" rpy="0 0 0"/><geometry><box size="0.02 0.02 0.1"/>'
            "</geometry></collision>")
        open(path, "w").write(text)
        with pytest.warns(UserWarning, match="ignored"):
            m4 = Model.from_urdf(path)
    assert len(m4.collision_body) == 16


def test_fallen_robot_rests_on_its_torso(model, oracle_lib):
    n = 8
    cfg = _on()
    init, act = _falling_setup(model, n)
    osim = oracle_lib.OracleSim(model, cfg, n)
    osim.reset(init)
    hs = HostSim(model, cfg, n)
    hs.reset(init.astype(np.float32))
    h = cfg.dt / cfg.nb_substeps
    first_contact = np.full(n, -1)
    for t in range(500):
        osim.step_servos(act)
        _, rec = hs.step_servos_rec(act)
        orec = _oracle_rec(oracle_lib, osim)
        touching = (orec[:, 0] != 0) & (first_contact < 0)
        first_contact[touching] = t
        # impact ticks: the 1 / h sensitivity of a contact row to the penetration depth (see
        # test_kernel_arithmetic_cpu.py) lets the two differ by a substep at touchdown
        settled = (first_contact >= 0) & (t > first_contact + 20)
        np.testing.assert_array_equal(rec[settled, 0], orec[settled, 0])
    so, sh = osim.get_state(), hs.state.astype(np.float64)
    assert (first_contact >= 0).all()
    # lying on the back face (x = -0.0819) or the front face (x = +0.0581) of the box: four corners each
    xs = np.sort(np.unique(np.round(model.collision_point[:, 0], 9)))
    back = init[:, 5] < 0  # negative pitch: falls backwards
    np.testing.assert_array_equal(orec[back, 0], 15.0)
    np.testing.assert_array_equal(orec[~back, 0], 240.0)
    np.testing.assert_allclose(so[back, 2], -xs[0], atol=2e-3)
    np.testing.assert_allclose(so[~back, 2], xs[1], atol=2e-3)
    # at rest, and the kernel arithmetic agrees with the oracle
    assert np.abs(so[:, 7:13]).max() < 1e-3 and np.abs(sh[:, 7:13]).max() < 1e-3
    assert np.abs(so[:, 2] - sh[:, 2]).max() < 1e-4
    pitch_o = np.arcsin(np.clip(2 * (so[:, 3] * so[:, 5] - so[:, 6] * so[:, 4]), -1, 1))
    pitch_h = np.arcsin(np.clip(2 * (sh[:, 3] * sh[:, 5] - sh[:, 6] * sh[:, 4]), -1, 1))
    assert np.abs(pitch_o - pitch_h).max() < 2e-3
    assert np.abs(so[:, 13:19] - sh[:, 13:19]).max() < 5e-3
    # the contact impulses carry the weight: sum of the normal impulses (tires + torso) = M g h
    weight_impulse = model.total_mass() * cfg.gravity * h
    for state, r in ((so, orec), (sh, rec.astype(np.float64))):
        lam = state[:, _abi.ST_CONTACT_IMPULSE:_abi.ST_CONTACT_IMPULSE + 2].sum(axis=1) + r[:, 2::4].sum(axis=1)
        np.testing.assert_allclose(lam, weight_impulse, rtol=2e-2)
    # host-side contact report: four points on the torso, on the floor, pushing up
    pts = body_contact_points(model, sh[0], rec[0], h)
    assert len(pts) == 4 and all(name == "torso" for name, _, _ in pts)
    assert all(abs(p[2]) < 5e-3 and f[2] > 0 for _, p, f in pts)
    contacts = contact_points_from_state(model, sh[0], cfg, None, rec[0])
    fz = sum(c.force_in_world[2] for c in contacts)
    assert abs(fz - model.total_mass() * cfg.gravity) < 0.03 * model.total_mass() * cfg.gravity
    assert [c.link_name for c in contact_points_from_state(model, sh[0], cfg, "torso", rec[0])] == ["torso"] * 4
    assert contact_points_from_state(model, sh[0], cfg, "imu", rec[0]) == []


def test_without_the_rows_the_torso_tunnels(model, oracle_lib):
    n = 4
    init, act = _falling_setup(model, n, 0.3, 0.4)
    z = {}
    for on in (0, 1):
        cfg = _abi.default_sim_config()
        cfg.body_contacts = on
        hs = HostSim(model, cfg, n)
        hs.reset(init.astype(np.float32))
        for _ in range(400):
            hs.step_servos(act)
        pts = np.stack([body_points_in_world(model, hs.state[i]) for i in range(n)])
        z[on] = pts[:, :, 2].min(axis=1)
    assert (z[1] > -5e-3).all()   # resting on the floor (rigid contact: sub-millimetre penetration)
    assert (z[0] < -0.03).all()   # round 1's physics: the box is centimetres below the floor


def test_upright_robots_are_bit_identical_with_the_rows_on_and_off(model):
    """The gate leaves the packed solvers alone: a robot whose collision points stay away from the ground takes
    exactly the arithmetic it took before the rows existed."""
    n = 16
    rng = np.random.default_rng(3)
    init = np.zeros((n, _abi.INIT_DIM), dtype=np.float32)
    init[:, 2] = 0.58
    pitch = rng.uniform(-0.1, 0.1, n)
    init[:, 3], init[:, 5] = np.cos(pitch / 2), np.sin(pitch / 2)
    act = np.zeros((n, 6, 6), dtype=np.float32)
    act[:, [2, 5], 0] = np.nan
    act[:, :, 3] = act[:, :, 4] = 1.0
    act[:, :, 5] = 0.99 * model.tau_max
    act[:, [2, 5], 1] = rng.uniform(-3, 3, (n, 2))
    states = []
    for on in (0, 1):
        cfg = _abi.default_sim_config()
        cfg.body_contacts = on
        hs = HostSim(model, cfg, n)
        hs.reset(init)
        for _ in range(40):
            hs.step_servos(act)
        states.append(hs.state.copy())
    np.testing.assert_array_equal(states[0], states[1])


def test_collision_points_on_leg_bodies(oracle_lib):
    """Spheres on the knees (lower-leg bodies) and a capsule along an upper leg: the general solver walks the force up
    from the body it acts on; oracle and kernel arithmetic agree while a crouched robot topples onto them."""
    m = Model.standard_upkie()
    m.add_collision_shape("left_lower_leg", 2, "sphere", [0.04], [0.0, 0.0, 0.0])
    m.add_collision_shape("right_lower_leg", 5, "sphere", [0.04], [0.0, 0.0, 0.0])
    m.add_collision_shape("left_upper_leg", 1, "capsule", [0.03, 0.12], [0.0, 0.0, -0.085])
    n = 6
    cfg = _on()
    init, act = _falling_setup(m, n, 0.2, 0.5)
    init[:, _abi.INIT_Q + 0] = init[:, _abi.INIT_Q + 3] = 0.9    # crouched: knees forward
    init[:, _abi.INIT_Q + 1] = init[:, _abi.INIT_Q + 4] = -1.8
    init[:, 2] = 0.45
    act[:, [0, 1, 3, 4], 0] = init[:, [13, 14, 16, 17]].astype(np.float32)  # hold the crouch
    osim = oracle_lib.OracleSim(m, cfg, n)
    osim.reset(init)
    hs = HostSim(m, cfg, n)
    hs.reset(init.astype(np.float32))
    seen = np.zeros(n, dtype=int)
    for t in range(400):
        osim.step_servos(act)
        _, rec = hs.step_servos_rec(act)
        seen |= _oracle_rec(oracle_lib, osim)[:, 0].astype(int)
    so, sh = osim.get_state(), hs.state.astype(np.float64)
    assert (seen >> 8).any()  # some leg point held rows at some time (bits 8.. are the points added above)
    assert np.abs(so[:, 2] - sh[:, 2]).max() < 2e-3
    assert np.abs(so[:, 13:19] - sh[:, 13:19]).max() < 2e-2
    assert np.abs(so[:, 7:13]).max() < 5e-2  # came to rest


def test_body_contacts_need_the_limit_kernels(model):
    """The rows live in the "extras + limits" kernels: with joint_limits = 0 they are off, as the header says."""
    n = 2
    init, act = _falling_setup(model, n, 0.3, 0.4)
    cfg = _on()
    cfg.joint_limits = 0
    hs = HostSim(model, cfg, n)
    hs.reset(init.astype(np.float32))
    for _ in range(300):
        _, rec = hs.step_servos_rec(act)
    assert (rec == 0).all()


# ---- on the device ---------------------------------------------------------------------------------------------------


@pytest.mark.gpu
def test_gpu_fallen_robots_rest_on_the_floor(model, oracle_lib):
    import torch

    from upkie_b200.sim import UpkieSim

    n = 96  # three warps: one falls early, one late, one mixed with robots that stay up for a while
    cfg = _on()
    init, act = _falling_setup(model, n, 0.0, 0.45)
    init[64:, 3], init[64:, 5] = 1.0, 0.0
    init[64::3, 3], init[64::3, 5] = np.cos(0.2), np.sin(0.2)
    sim = UpkieSim(n, model=model, config=cfg, device=0)
    sim.reset(init_state=torch.from_numpy(init.astype(np.float32)).cuda())
    osim = oracle_lib.OracleSim(model, cfg, n, threads=os.cpu_count() or 1)
    osim.reset(init)
    a_dev = torch.from_numpy(act).cuda()
    for t in range(450):
        sim.step_servos(a_dev)
        osim.step_servos(act)
    sg = sim.get_state().cpu().numpy().astype(np.float64)
    so = osim.get_state()
    rec = sim.get_body_contacts().cpu().numpy()
    orec = _oracle_rec(oracle_lib, osim)
    down = orec[:, 0] != 0
    assert down.sum() >= 64
    # the same points hold rows; a robot still rocking on an edge may differ by a point at the breaking threshold
    assert (rec[down, 0] == orec[down, 0]).mean() >= 0.95, (rec[down, 0], orec[down, 0])
    dz = np.abs(sg[down, 2] - so[down, 2])
    assert np.quantile(dz, 0.9) < 1e-5 and dz.max() < 1e-2, (np.quantile(dz, 0.9), dz.max())  # a robot may still be rocking on an edge
    pts_z = np.stack([body_points_in_world(model, sg[i]) for i in range(n)])[:, :, 2]
    assert (pts_z.min(axis=1) > -5e-3).all()  # nobody tunnels
    h = cfg.dt / cfg.nb_substeps
    rest = down & (np.abs(so[:, 7:13]).max(axis=1) < 1e-3)  # at rest in the oracle: the impulses carry the weight
    assert rest.sum() >= 48
    lam = sg[rest, _abi.ST_CONTACT_IMPULSE:_abi.ST_CONTACT_IMPULSE + 2].sum(axis=1) + rec[rest, 2::4].sum(axis=1)
    np.testing.assert_allclose(lam, model.total_mass() * cfg.gravity * h, rtol=3e-2)


@pytest.mark.gpu
def test_gpu_backend_reports_torso_contacts(model):
    """``B200Backend.get_contact_points`` (``pybullet_backend.py:660-716``): a robot left without actions falls and
    then reports contacts on its torso link next to those of the tires."""
    from upkie_b200.backend import B200Backend
    from upkie_b200.robot_state import RobotState

    backend = B200Backend(dt=0.005)
    pitch = 0.4
    backend.reset(RobotState(orientation_base_in_world=np.array([np.cos(pitch / 2), 0.0, np.sin(pitch / 2), 0.0])))
    assert backend.get_contact_points("torso") == []
    for _ in range(500):
        backend.step({})
    torso = backend.get_contact_points("torso")
    assert 1 <= len(torso) <= 4 and all(c.link_name == "torso" and c.force_in_world[2] > 0 for c in torso)
    assert all(abs(c.position_contact_in_world[2]) < 5e-3 for c in torso)
    everything = backend.get_contact_points()
    fz = sum(c.force_in_world[2] for c in everything)
    assert abs(fz - model.total_mass() * 9.81) < 0.05 * model.total_mass() * 9.81
    backend.close()
