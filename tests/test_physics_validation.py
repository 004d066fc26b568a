"""Independent PHYSICAL checks of the two pieces of the path the reference has
no runnable kernel for (MRT collision, zero_gradient outlet): their oracle
definitions are our own, so besides oracle <-> CUDA parity
they are checked against analytic solutions that do not depend on either.

Plane Poiseuille flow between halfway bounce-back plates, driven by a body
force g through Guo's forcing: u(y) = g / (2 nu) * (y + 1/2) (H - 1/2 - y) for
node rows y = 0 .. H-1 (the walls sit half a node outside the outermost fluid
rows).  A lattice Boltzmann scheme reproduces the parabola up to a constant
slip that is O(1/H^2) relative to u_max: second-order convergence.
"""
import numpy as np
import pytest

import bench
import cases


def _steady_profile(model, forcing, H, nu, u_max, n_threads=2):
    g = 8.0 * nu * u_max / (H * H)
    sim = cases.poiseuille(4, H, forcing=forcing, g=g, kin_visc=nu, model=model,
                           perturb=0.0)
    sim.initial_fields_dict["default"]["fluid"]["density"] = {
        "type": "fixed", "value": 1.0}
    orc, _ = bench.make_oracle(sim, n_threads)
    steps = int(4.0 * H * H / nu)            # ~ 4 diffusion times: converged to 1e-13
    orc.step(steps)
    u = orc.velocity.reshape(6, H + 2, 2)[1:-1, 1:-1]
    assert np.abs(u[:, :, 1]).max() < 1e-12          # no cross flow
    assert np.abs(u[:, :, 0] - u[0, :, 0]).max() < 1e-14   # x-invariant
    y = np.arange(H, dtype=np.float64)
    exact = g / (2.0 * nu) * (y + 0.5) * (H - 0.5 - y)
    return u[0, :, 0], exact


@pytest.mark.parametrize("model,forcing", [
    ("MRT", "guo_second_order"), ("MRT", "guo_linear"), ("BGK", "guo_second_order")])
def test_poiseuille_profile_converges_with_second_order(model, forcing):
    nu, u_max = 0.1, 0.01
    errors = []
    for H in (8, 16, 32):
        got, exact = _steady_profile(model, forcing, H, nu, u_max)
        errors.append(float(np.sqrt(((got - exact) ** 2).sum() /
                                    (exact ** 2).sum())))
    # second order: the relative L2 error falls by 4 per doubling of H
    assert errors[0] < 2.0e-2 and errors[2] < 1.5e-3, errors
    for coarse, fine in zip(errors, errors[1:]):
        assert 3.5 < coarse / fine < 4.5, errors


@pytest.mark.parametrize("model,nu", [("MRT", 0.1), ("MRT", 0.02), ("BGK", 0.1)])
def test_poiseuille_slip_matches_two_relaxation_time_theory(model, nu):
    """The whole discrepancy is the wall slip of halfway bounce back, a
    constant offset (the curvature of the profile is exact), and its size is
    the textbook result for a scheme whose odd (energy-flux) moments relax at
    rate s_q: u_slip = g / (2 nu) * 4/3 * (Lambda - 3/16) with the magic
    parameter Lambda = (tau - 1/2) (1/s_q - 1/2).  The reference's MRT rates
    (base/collision_operator.py:159-163) put s_q = 1, BGK s_q = omega -- an
    independent check that the MRT operator relaxes the moments it claims to."""
    H, u_max = 16, 0.01
    got, exact = _steady_profile(model, "guo_second_order", H, nu, u_max)
    diff = got - exact
    assert np.ptp(diff) < 1e-3 * np.abs(diff).mean()
    tau = 3.0 * nu + 0.5
    s_q = 1.0 if model == "MRT" else 1.0 / tau
    magic = (tau - 0.5) * (1.0 / s_q - 0.5)
    g_over_2nu = 4.0 * u_max / (H * H)
    predicted = g_over_2nu * 4.0 / 3.0 * (magic - 3.0 / 16.0)
    assert abs(diff.mean() - predicted) < 2e-3 * abs(predicted)


@pytest.mark.gpu
@pytest.mark.parametrize("general", ["0", "1"])
def test_cuda_mrt_reaches_the_analytic_poiseuille_profile(general, monkeypatch):
    """The same check through the C ABI on the B200: the stress-moment form of
    the MRT collision (default) and the nine-rate 9 x 9 transform
    (PLB_MRT_GENERAL=1), through the several-steps-per-pass path (the lattice
    is large enough for it), must reach the parabola with the predicted slip."""
    from pylabolt_b200 import capi
    from test_gpu_parity import make_solver
    monkeypatch.setenv("PLB_MRT_GENERAL", general)
    monkeypatch.setenv("PLB_FUSE", "2")
    H, nx, nu, u_max = 32, 256, 0.1, 0.01
    g = 8.0 * nu * u_max / (H * H)
    sim = cases.poiseuille(nx, H, forcing="guo_second_order", g=g, kin_visc=nu,
                           model="MRT", perturb=0.0)
    sim.initial_fields_dict["default"]["fluid"]["density"] = {
        "type": "fixed", "value": 1.0}
    s = make_solver(sim, strict=False)
    try:
        s.advance(int(4.0 * H * H / nu), store_moments_last=True)
        u = s.plb.download(capi.VELOCITY).reshape(nx + 2, H + 2, 2)[1:-1, 1:-1]
    finally:
        s.close()
    y = np.arange(H, dtype=np.float64)
    exact = g / (2.0 * nu) * (y + 0.5) * (H - 0.5 - y)
    assert np.abs(u[:, :, 1]).max() < 1e-12
    diff = u[:, :, 0] - exact
    tau = 3.0 * nu + 0.5
    predicted = 4.0 * u_max / (H * H) * 4.0 / 3.0 * ((tau - 0.5) * 0.5 - 3.0 / 16.0)
    assert np.ptp(diff) < 2e-3 * abs(predicted)
    assert abs(diff.mean() - predicted) < 2e-3 * abs(predicted)


def cylinder_in_uniform_stream(diameter, velocity, reynolds, model, outlet,
                               height_d=12, length_d=26, upstream_d=8):
    """Circular cylinder in a uniform stream: fixed_velocity inlet, plates
    that move with the stream (no wall boundary layers; blockage 1/height_d),
    outlet either zero_gradient (our definition) or fixed_pressure (pinned to
    the reference); the body sits one node off the centre line so that vortex
    shedding starts by itself."""
    from types import SimpleNamespace
    nx, ny = length_d * diameter, height_d * diameter
    seg = cases._walls(nx, ny)
    stream = {"type": "fixed_velocity", "value": [velocity, 0.0]}
    out = ({"type": "zero_gradient"} if outlet == "zero_gradient" else
           {"type": "fixed_pressure", "value": 1.0})
    return SimpleNamespace(
        control_dict=cases._control(1),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": 1, "ny": 1},
        transport_dict={"kin_visc": velocity * diameter / reynolds},
        initial_fields_dict={"default": {"fluid": {
            "velocity": {"type": "fixed", "value": [velocity, 0.0]},
            "density": {"type": "fixed", "value": 1.0},
            "pressure": {"type": "fixed", "value": 0.0}}}},
        boundary_dict={
            "options": {},
            "plates": {"wall": False, "segments": [seg["bottom"], seg["top"]],
                       "fluid": stream},
            "inlet": {"wall": False, "segments": [seg["left"]], "fluid": stream},
            "outlet": {"wall": False, "segments": [seg["right"]], "fluid": out}},
        obstacle_dict={"options": {"compute_force_torque": True},
                       "cyl": {"type": "circle", "radius": diameter / 2,
                               "center": [upstream_d * diameter, ny // 2 + 1],
                               "density": 1.0, "static": True}},
        collision_dict={"fluid": {"model": model,
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={})


def drag_and_strouhal(history, diameter, velocity):
    """Mean drag coefficient and the Strouhal number of the lift signal over
    the last 40 % of `history` = rows (step, F_x, F_y).  The lift also carries
    the transverse acoustic mode of the box (period 2 H / c_s, St ~ 0.5), so
    the shedding frequency is the spectral peak inside 0.1 < St < 0.3."""
    h = np.asarray(history, dtype=np.float64)
    tail = h[h[:, 0] > 0.6 * h[-1, 0]]
    scale = 2.0 / (velocity * velocity * diameter)       # rho = 1
    cd = float(tail[:, 1].mean() * scale)
    lift = tail[:, 2] - tail[:, 2].mean()
    dt = tail[1, 0] - tail[0, 0]
    strouhal = np.fft.rfftfreq(len(lift), d=dt) * diameter / velocity
    power = np.abs(np.fft.rfft(lift * np.hanning(len(lift))))
    band = (strouhal > 0.1) & (strouhal < 0.3)
    peak = int(np.argmax(np.where(band, power, 0.0)))
    # parabolic interpolation of the peak
    a, b, c = power[peak - 1], power[peak], power[peak + 1]
    shift = 0.5 * (a - c) / (a - 2 * b + c)
    st = float(strouhal[peak] + shift * (strouhal[1] - strouhal[0]))
    return cd, st, float(np.abs(lift).max() * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("model,outlet", [("MRT", "zero_gradient"),
                                          ("BGK", "fixed_pressure")])
def test_cylinder_re100_drag_and_strouhal(model, outlet):
    """Flow past a cylinder at Re = 100 (BASELINE.json configs[2]): mean drag
    and shedding frequency from the momentum-exchange force history, against
    the literature band for this blockage (unbounded flow: C_d 1.32 - 1.42,
    St 0.164 - 0.170, e.g. Williamson 1996, Park et al. 1998; a blockage of
    1/12 with co-moving plates raises both by a few per cent).  The MRT +
    zero_gradient run (our definitions) must also agree with the BGK +
    fixed_pressure run, whose kernels are pinned to the reference."""
    from test_gpu_parity import make_solver
    d, u = 20, 0.05
    s = make_solver(cylinder_in_uniform_stream(d, u, 100.0, model, outlet),
                    strict=False)
    try:
        every = int(d / u / 40)               # 40 samples per convective time
        history = []
        for n in range(140 * 40):             # 140 convective times
            s.advance(every, record_links_last=True)
            _, body = s.compute_forces()
            history.append(((n + 1) * every, body[0, 0], body[0, 1]))
    finally:
        s.close()
    cd, st, cl = drag_and_strouhal(history, d, u)
    # the CPU oracle gives C_d = 1.515 / 1.512 and St = 0.1720 / 0.1725 for
    # the two configurations (MRT + zero_gradient / BGK + fixed_pressure)
    assert 1.42 < cd < 1.60, (cd, st, cl)
    assert 0.165 < st < 0.180, (cd, st, cl)
    assert cl > 0.2, (cd, st, cl)             # it does shed
