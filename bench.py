#!/usr/bin/env python
"""bench.py -- fluidLB time-step throughput (D2Q9, fp64) in GLUPS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload channel|cavity]
                    [--impl reference]

Workload (BASELINE.json configs[4], the weak-scaling case the metric's
multi-GPU target is quoted on): periodic channel, 8192 x 16384 lattice nodes
per GPU (Nx = 8192 * N), MRT + Guo second-order forcing, gravity [1e-6, 0],
bounce_back plates top / bottom, x-periodic across the slabs, rho = 1 and a
seeded velocity perturbation.  `--workload cavity` runs configs[3] instead
(16384 x 16384 lid-driven cavity, BGK, strong scaling) and, at N = 1, the
default run reports it too under "extra".

One "step" = one reference time step (Solver.single_time_step) of the whole
lattice.  value = lattice-node updates of all ranks / max-over-ranks device
time, populations resident in HBM.  e2e = the same through the public Solver
API with host buffers: upload of the initial rho / u from pinned host memory,
initialisation, K steps issued one by one from Python through the
execute_single_time_step slot, the residue read-back and the download of
rho / u into pinned host memory, all inside the timed region.

The lattice (19.3 GB of populations per GPU) is far larger than L2, so no L2
flush is needed between steps (config.l2 says so).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

ALGORITHMIC_BYTES_PER_NODE = 144          # 9 x 8 B read + 9 x 8 B write
HBM_FALLBACK_GBS = 6650.0                 # B200_PROFILING.md fallback


# --------------------------------------------------------------------------
# workloads (reference case-file schema)
# --------------------------------------------------------------------------
PERIOD = 32        # the initial perturbation repeats every PERIOD nodes (x and y)


def periodic_noise(amplitude, period=PERIOD):
    """Deterministic, rank-independent integer-hash noise in [-amplitude/2,
    amplitude/2) with period `period` along x and y.  The periodicity costs
    the kernels nothing (they never look at the values) and is what lets the
    bench check its own result at full size: a node after S steps depends only
    on data within Chebyshev distance S, so the full-size field is determined
    by a small oracle run on the same pattern (parity_check below)."""
    def func(i, j):
        i = np.asarray(i, dtype=np.int64) % period
        j = np.asarray(j, dtype=np.int64) % period
        h = (i * 73856093) ^ (j * 19349663)
        h = (h ^ (h >> 13)) * 1274126177
        a = ((h >> 8) & 0xFFFF).astype(np.float64) / 65536.0 - 0.5
        b = ((h >> 24) & 0xFFFF).astype(np.float64) / 65536.0 - 0.5
        return amplitude * a, amplitude * b
    func.vectorized = True
    return func


def _control(steps):
    return {"start_time": 0, "end_time": steps, "std_out_interval": steps,
            "save_interval": steps, "checkpoint_interval": None,
            "precision": "double"}


def _initial_fields():
    return {"default": {"fluid": {
        "velocity": {"type": "func", "func": periodic_noise(0.01)},
        "density": {"type": "fixed", "value": 1.0},
        "pressure": {"type": "fixed", "value": 0.0}}}}


def channel_case(nx, ny, n_ranks, steps):
    """BASELINE.json configs[4]."""
    from types import SimpleNamespace
    return SimpleNamespace(
        control_dict=_control(steps),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": n_ranks, "ny": 1},
        transport_dict={"kin_visc": 0.1},
        initial_fields_dict=_initial_fields(),
        boundary_dict={
            "options": {},
            "inout": {"wall": False,
                      "segments": [[[0, 0], [0, ny - 1]],
                                   [[nx - 1, 0], [nx - 1, ny - 1]]],
                      "fluid": {"type": "periodic"}},
            "plates": {"wall": True,
                       "segments": [[[0, 0], [nx - 1, 0]],
                                    [[0, ny - 1], [nx - 1, ny - 1]]],
                       "fluid": {"type": "bounce_back"}}},
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": "MRT",
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": "guo_second_order"}},
        forcing_dict={"gravity": [1.0e-6, 0.0]})


def cavity_case(nx, ny, n_ranks, steps):
    """BASELINE.json configs[3] (docs/Setup.rst:15-64 of the reference at
    16384 x 16384: three bounce_back walls, lid [0.1, 0], nu = 0.1, BGK)."""
    from types import SimpleNamespace
    left = [[0, 0], [0, ny - 1]]
    right = [[nx - 1, 0], [nx - 1, ny - 1]]
    bottom = [[0, 0], [nx - 1, 0]]
    top = [[0, ny - 1], [nx - 1, ny - 1]]
    return SimpleNamespace(
        control_dict=_control(steps),
        mesh_dict={"grid": [nx, ny]},
        lattice_dict={"lattice_type": "D2Q9"},
        decompose_dict={"nx": n_ranks, "ny": 1},
        transport_dict={"kin_visc": 0.1},
        initial_fields_dict=_initial_fields(),
        boundary_dict={
            "options": {},
            "walls": {"wall": True, "segments": [left, right, bottom],
                      "fluid": {"type": "bounce_back"}},
            "lid": {"wall": True, "segments": [top],
                    "fluid": {"type": "fixed_velocity", "value": [0.1, 0.0]}}},
        obstacle_dict={"options": {}},
        collision_dict={"fluid": {"model": "BGK",
                                  "equilibrium": "density_based_second_order",
                                  "forcing_model": None}},
        forcing_dict={})


WORKLOADS = {
    # name: (factory, per-GPU nx (weak) or total nx (strong), ny, scaling, text)
    "channel": (channel_case, 8192, 16384, "weak",
                "periodic channel {nx}x{ny} ({per}x{ny} per GPU), fp64 MRT + "
                "Guo second-order forcing, x-periodic slabs, bounce_back plates"),
    "cavity": (cavity_case, 16384, 16384, "strong",
               "lid-driven cavity {nx}x{ny}, fp64 BGK, halfway bounce-back + "
               "moving wall"),
}


def workload_sizes(name, n_ranks, scale):
    factory, nx, ny, scaling, text = WORKLOADS[name]
    nx = max(n_ranks, int(nx * scale))
    ny = max(8, int(ny * scale))
    if scale != 1.0:
        # debugging sizes keep the lattice a multiple of the pattern period
        nx = max(PERIOD, nx - nx % PERIOD)
        ny = max(PERIOD, ny - ny % PERIOD)
    total_nx = nx * n_ranks if scaling == "weak" else nx
    per = total_nx // n_ranks
    return factory, total_nx, ny, scaling, text.format(nx=total_nx, ny=ny,
                                                       per=per)


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: the
    sampler runs from before the warm-up, samples carry nvidia-smi's own
    timestamp and only those inside [mark_start, mark_stop] are used."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.out, stderr=subprocess.DEVNULL)
            deadline = time.time() + 5.0
            while time.time() < deadline and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    stamp = datetime.datetime.strptime(
                        parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((stamp, float(parts[1]), float(parts[2]),
                                 float(parts[3]),
                                 [n for n, flag in zip(names, parts[4:8])
                                  if flag.lower().startswith("active")]))
                except ValueError:
                    continue
        os.unlink(self.path)
        inside = [r for r in rows
                  if self.t0 is None or self.t0 - 0.03 <= r[0] <= self.t1 + 0.03]
        where = "timed region"
        if not inside and rows:
            # region shorter than the sampling period: nearest samples
            mid = 0.5 * (self.t0 + self.t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:2]
            where = "nearest samples (region shorter than sampling period)"
        if not inside:
            return None
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])),
                "sm_max_mhz": float(max(r[2] for r in inside)),
                "power_w_max": float(max(r[3] for r in inside)),
                "samples": len(inside), "window": where, "reasons": reasons}


def hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """Per-launch DRAM bytes of the bulk kernel from the committed ncu
    capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def make_oracle(sim, n_threads):
    """CPU oracle on the same case, built from our own host-side State."""
    from oracle.oracle import Oracle
    from pylabolt_b200.comm import SingleComm
    from pylabolt_b200.operators import CollisionOperator, FluidLB, ForceOperator
    from pylabolt_b200.state import State
    comm = SingleComm()
    st = State(sim, comm, 0, verbose=False)
    col = CollisionOperator(sim, FluidLB(), st, comm, verbose=False)
    frc = ForceOperator(sim, FluidLB(), st, comm, collision_operator=col,
                        verbose=False)
    elements = [{"type": el.type_fluid, "nodes": el.boundary_nodes,
                 "out": el.out_list, "inv": el.inv_list, "normal": el.normal,
                 "vector": el.vector_fluid, "scalar": float(el.scalar_fluid)}
                for el in st.boundary.boundary_elements]
    orc = Oracle(st.domain.shape, st.fields.solid, st.fields.ghost_node,
                 st.fields.density, st.fields.velocity, elements,
                 col.omega_fluid, gravity=frc.gravity, forcing=col.forcing_fluid,
                 collision=col.collision_fluid,
                 x_periodic=st.boundary.x_periodic,
                 y_periodic=st.boundary.y_periodic, mrt_rates=col.mrt_rates,
                 n_threads=n_threads)
    orc.initialize_pop()
    return orc, int(st.domain.inner_size)


def time_cpu_port(workload, steps, warmup, budget_s=12.0, sample=2048):
    """The oracle (CPU port of the reference's five-pass step, OpenMP where
    the reference has numba prange) on a bounded sample of the workload."""
    factory = WORKLOADS[workload][0]
    cores = os.cpu_count() or 1
    sim = factory(sample, sample, 1, 1)
    orc, nodes = make_oracle(sim, cores)
    for _ in range(max(1, warmup)):
        orc.step(1)
    t0 = time.perf_counter()
    orc.step(1)
    one = time.perf_counter() - t0
    per_step = []
    n = max(1, steps)
    deadline = time.perf_counter() + budget_s
    for _ in range(n):
        t0 = time.perf_counter()
        orc.step(1)
        per_step.append(time.perf_counter() - t0)
        if time.perf_counter() > deadline:
            break
    mean = float(np.mean(per_step)) if per_step else one
    return {"value": nodes / mean / 1e9, "unit": "GLUPS", "cores": cores,
            "kind": "port",
            "sample": f"{sample}x{sample} nodes of the same case, "
                      f"{len(per_step)} timed steps, {cores} OpenMP threads "
                      "(oracle/plb_oracle.c, five-pass AoS like the reference)",
            "ms_per_step": mean * 1e3, "steps": len(per_step)}


# --------------------------------------------------------------------------
# parity of the benchmarked run itself (every N, full size)
# --------------------------------------------------------------------------
PARITY_RTOL = 1.0e-12      # BASELINE.json north_star: rho / u within 1e-12 relative


def _fold(n_full, n_small, half, period):
    """Index map full -> small lattice of a wall-bounded direction: the first
    and last `half` nodes map to the node at the same distance from their own
    wall, everything between to the node of the small lattice's wall-unaware
    middle period with the same phase of the initial pattern."""
    i = np.arange(n_full)
    return np.where(i < half, i,
                    np.where(i >= n_full - half, i - (n_full - n_small),
                             half + (i - half) % period))


class ParityCheck:
    """Checks the rho / u this rank's slab holds after `steps` steps against
    the CPU oracle at FULL size.  The step is local (data travels one node per
    step) and the initial state is PERIOD-periodic, so
      * channel (x-periodic, plates in y): the field is the PERIOD x ny oracle
        strip tiled along x;
      * cavity (walls all round): a node within `half` > steps of a wall
        equals the node at the same wall distance of a small cavity of
        2 * half + PERIOD nodes per direction, every other node the node of
        the small cavity's middle period with the same phase.
    The oracle (oracle/plb_oracle.c, pinned to the reference's own runs) is
    the checker here, never the thing measured."""

    def __init__(self, workload, nx, ny, x_offset, nx_rank, n_threads):
        self.workload, self.nx, self.ny = workload, nx, ny
        self.x_offset, self.nx_rank = int(x_offset), int(nx_rank)
        self.n_threads = n_threads
        self.orc = None
        self.done = 0

    def _build(self, steps_max):
        factory = WORKLOADS[self.workload][0]
        if self.workload == "channel":
            if self.nx % PERIOD:
                raise ValueError("parity check: nx must be a multiple of %d" % PERIOD)
            self.small = (PERIOD, self.ny)
            self.map_x = (self.x_offset + np.arange(self.nx_rank)) % PERIOD
            self.map_y = np.arange(self.ny)
        else:
            half = PERIOD * (-(-(steps_max + 2) // PERIOD))
            n_small = 2 * half + PERIOD
            if min(self.nx, self.ny) < n_small or self.nx % PERIOD or self.ny % PERIOD:
                # small lattices: the oracle runs the whole case
                self.small = (self.nx, self.ny)
                self.map_x = self.x_offset + np.arange(self.nx_rank)
                self.map_y = np.arange(self.ny)
            else:
                self.small = (n_small, n_small)
                self.map_x = _fold(self.nx, n_small, half, PERIOD)[
                    self.x_offset:self.x_offset + self.nx_rank]
                self.map_y = _fold(self.ny, n_small, half, PERIOD)
        sim = factory(self.small[0], self.small[1], 1, 1)
        self.orc, _ = make_oracle(sim, self.n_threads)

    def expected(self, steps, steps_max=None):
        """Oracle rho (sx, sy) and u (sx, sy, 2) of the small lattice after
        `steps` steps (must be called with non-decreasing `steps`)."""
        if self.orc is None:
            self._build(steps_max or steps)
        assert steps >= self.done
        self.orc.step(steps - self.done)
        self.done = steps
        sx, sy = self.small
        rho = self.orc.density.reshape(sx + 2, sy + 2)[1:-1, 1:-1]
        u = self.orc.velocity.reshape(sx + 2, sy + 2, 2)[1:-1, 1:-1]
        return rho.copy(), u.copy()

    def compare(self, want, rho_inner, u_inner):
        """max-norm relative error per field (a component-wise relative error
        is meaningless where u_y ~ 1e-16) of this rank's inner rho / u."""
        want_rho, want_u = want
        got_rho = rho_inner.reshape(self.nx_rank, self.ny)
        got_u = u_inner.reshape(self.nx_rank, self.ny, 2)
        identity_y = (len(self.map_y) == want_rho.shape[1] and
                      np.array_equal(self.map_y, np.arange(len(self.map_y))))
        err = {"density": 0.0, "velocity": 0.0}
        for x0 in range(0, self.nx_rank, 512):           # bounded temporaries
            rows = self.map_x[x0:x0 + 512]
            for key, got, ref in (("density", got_rho, want_rho),
                                  ("velocity", got_u, want_u)):
                expect = ref[rows] if identity_y else ref[rows][:, self.map_y]
                err[key] = max(err[key],
                               float(np.abs(got[x0:x0 + 512] - expect).max()))
        scale = {"density": float(np.abs(want_rho).max()),
                 "velocity": float(np.abs(want_u).max())}
        return max(err[k] / scale[k] for k in err)


# --------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------
def run_reference(args, rank, world, text, scaling):
    if rank != 0:
        return
    base = time_cpu_port(args.workload, args.steps, args.warmup,
                         budget_s=90.0, sample=args.cpu_sample)
    line = {
        "impl": "reference", "metric": "fluidLB D2Q9 fp64 lattice updates",
        "value": base["value"], "unit": "GLUPS", "n_gpus": args.gpus,
        "steps": base["steps"], "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": text,
                   "note": "reference's CPU path (oracle port: the reference "
                           "is Python/numba and cannot travel to the GPU box)"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind",
                                              "sample")},
        "e2e": {"value": base["value"], "unit": "GLUPS",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def measure(workload, comm, rank, world, steps, warmup, scale, device,
            parity=True, min_region_ms=400.0, max_repeats=12):
    from pylabolt_b200 import capi
    from pylabolt_b200.solver import Solver

    factory, nx, ny, scaling, text = workload_sizes(workload, world, scale)
    sim = factory(nx, ny, world, steps)
    t_setup = time.perf_counter()
    solver = Solver(comm, "b200", simulation=sim, device=device, verbose=False)
    st = solver.state
    total_nodes = nx * ny
    t_backend = time.perf_counter()
    solver.set_backend()
    solver.compile()
    # wall clock of the setup a user waits for: host containers (State: flags,
    # link lists, initial fields), then device allocation + uploads + node
    # classification (set_backend)
    setup_s = {"host_state": t_backend - t_setup,
               "set_backend": time.perf_counter() - t_backend}
    plb = solver.plb
    info = plb.info()

    copy_gbs = plb.copy_bandwidth()       # context for the roofline fraction

    def allmax(x):
        t = np.array([x], dtype=np.float64)
        out = np.zeros_like(t)
        comm.Allreduce(t, out, op="max")
        return float(out[0])

    # ---- device-resident throughput -------------------------------------
    # The timed region = K steps between two events, barrier + synchronize on
    # both sides.  It is repeated back to back (the lattice just keeps
    # advancing) until >= min_region_ms have been covered, so that the
    # nvidia-smi sampler sees the clocks of the timed work several times even
    # when K steps last 30 ms; the reported time is the median repeat.
    sampler = ClockSampler(device)
    sampler.start()
    plb.initialize_pop()
    plb.step(warmup, False)
    plb.sync()
    comm.Barrier()
    plb.kernel_launches(reset=True)
    plb.profile_enable(True)
    groups_before = tuple(plb.fused_info()[k] for k in ("pairs", "triples", "quads"))
    repeat_ms = []
    launches = 0
    sampler.mark_start()
    repeats = 1
    while len(repeat_ms) < repeats:
        plb.sync()
        comm.Barrier()
        plb.event_record(0)
        plb.step(steps, False)
        plb.event_record(1)
        plb.sync()
        comm.Barrier()
        repeat_ms.append(allmax(plb.event_elapsed_ms(0, 1)))
        if len(repeat_ms) == 1:
            launches = plb.kernel_launches()
            repeats = int(min(max_repeats,
                              max(1, np.ceil(min_region_ms / repeat_ms[0]))))
    sampler.mark_stop()
    bulk_ms, bulk_n = plb.profile_read()
    plb.profile_enable(False)
    clocks = sampler.stop()
    ms = float(np.median(repeat_ms))
    value = total_nodes * steps / (ms * 1e-3) / 1e9
    steps_timed = steps * repeats

    # ---- parity of exactly this run: one more (moment-storing) step, then
    # rho / u of every rank's slab against the oracle --------------------
    checker = None
    parity_out = None
    if parity:
        cores = os.cpu_count() or 1
        checker = ParityCheck(workload, nx, ny, st.domain.offset[0], plb.nx,
                              max(1, cores // world))
        s_dev = warmup + steps_timed + 1
        want_e2e = checker.expected(steps, steps_max=max(steps, s_dev))
        want_dev = checker.expected(s_dev)
        plb.step(1, True)
        rho_dev = plb.download(capi.DENSITY_INNER)
        u_dev = plb.download(capi.VELOCITY_INNER)
        mass = np.array([float(rho_dev.sum())])
        mass_all = np.zeros_like(mass)
        comm.Allreduce(mass, mass_all, op="sum")
        err_dev = allmax(checker.compare(want_dev, rho_dev, u_dev))
        del rho_dev, u_dev
        parity_out = {
            "max_rel_err": err_dev, "tolerance": PARITY_RTOL,
            "ok": bool(err_dev <= PARITY_RTOL),
            "steps_compared": s_dev,
            "what": "rho, u of every rank's whole slab after the warm-up, the "
                    "timed steps and one moment-storing step, against the CPU "
                    "oracle (max-norm relative error, max over ranks): " +
                    ("%d x %d oracle strip tiled along x" % checker.small
                     if workload == "channel" else
                     "%d x %d oracle cavity folded out from walls / corners / "
                     "periodic middle" % checker.small),
            "global_mean_density": float(mass_all[0]) / total_nodes}

    # roofline of the dominant kernel (bulk collide-stream), this rank.
    # `achieved` = the bytes a launch has to move, 144 B (9 x 8 B read + 9 x
    # 8 B write) x the nodes the launch covers, / the launch
    # duration: the kernel's real HBM rate, <= peak.  The single-step kernel
    # advances those nodes by one step; k_bulk_fused advances them by TWO
    # (THREE) steps on the same 144 B, so per node and STEP it moves 72 (48) B
    # -- `step_equivalent_*` restates the launch in SURVEY's per-step figure
    # (144 B x node-steps advanced) and is how a memory-bound kernel exceeds
    # the per-step roofline; it is a speed-up factor, not a bandwidth.
    peak, peak_src = hbm_peak()
    finfo = plb.fused_info()
    pairs = finfo["pairs"] - groups_before[0]
    triples = finfo["triples"] - groups_before[1]
    quads = finfo["quads"] - groups_before[2]
    singles = steps_timed - 2 * pairs - 3 * triples - 4 * quads
    bulk_ms_per_step = bulk_ms / steps_timed
    node_steps = (2 * finfo["n_deep"] * pairs + 3 * finfo["n_deep3"] * triples +
                  4 * finfo["n_deep4"] * quads + info["n_bulk_timed"] * singles)
    dram_nodes = (finfo["n_deep"] * pairs + finfo["n_deep3"] * triples +
                  finfo["n_deep4"] * quads + info["n_bulk_timed"] * singles)
    achieved = ALGORITHMIC_BYTES_PER_NODE * dram_nodes / (bulk_ms * 1e-3) / 1e9
    step_equiv = ALGORITHMIC_BYTES_PER_NODE * node_steps / (bulk_ms * 1e-3) / 1e9
    fused = pairs + triples + quads > 0
    kernel = ("k_bulk_fused<depth 4>" if quads else
              "k_bulk_fused<depth 3>" if triples else
              "k_bulk_fused<depth 2>" if pairs else
              "k_bulk_vec2" if info["variant"] else "k_bulk_scalar")
    traffic_key = workload + ("_fused4" if quads else "_fused3" if triples else
                              "_fused" if pairs else "")
    alg_per_launch = ALGORITHMIC_BYTES_PER_NODE * dram_nodes / max(1, bulk_n)
    traffic = ncu_traffic(traffic_key)
    if traffic is not None and not 0.9 <= traffic / alg_per_launch <= 1.3:
        # the committed capture is of another slab size (a strong-scaling
        # slab at N > 1): it says nothing about this launch
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": kernel,
                "algorithmic_bytes_per_launch": alg_per_launch,
                "steps_per_launch": 4 if quads else 3 if triples else 2 if pairs else 1,
                "bytes_per_node_and_step":
                    ALGORITHMIC_BYTES_PER_NODE * dram_nodes / max(1, node_steps),
                "step_equivalent_gbs": step_equiv,
                "step_equivalent_frac": step_equiv / peak,
                "frac_is": "144 B x the nodes one launch covers / launch time / peak: "
                           "the bytes that must cross HBM.  The launch advances "
                           "steps_per_launch steps on them; step_equivalent_frac "
                           "(= frac x steps per launch) is the fraction of SURVEY's "
                           "144 B per node and STEP roofline",
                "fused": {k: finfo[k] for k in ("active", "n_deep", "n_deep3", "n_deep4",
                                                "n_list1", "rows", "strips")},
                "pairs": pairs, "triples": triples, "quads": quads,
                "single_steps": singles,
                "face_transport": ["none", "own ghost rows", "nccl",
                                   "p2p stores"][info["faces"]],
                "launches_per_step": bulk_n / steps_timed,
                "kernel_ms_per_step": bulk_ms_per_step,
                "kernel_share_of_step": bulk_ms / float(np.sum(repeat_ms)),
                "peak_source": peak_src,
                "d2d_copy_gbs_this_box": copy_gbs,
                "kernel_build": plb.build_info()}
    mem = plb.memory_info()
    roofline["device_memory_gb"] = {k: round(v / 1e9, 3) for k, v in mem.items()}

    # ---- end to end through the public Solver API, host buffers -----------
    size = plb.size
    rho_in = plb.pinned((size,))
    u_in = plb.pinned((size, 2))
    rho_out = plb.pinned((plb.nx * plb.ny,))
    u_out = plb.pinned((plb.nx * plb.ny, 2))
    rho_in.array[:] = st.fields.density
    u_in.array[:] = st.fields.velocity
    # output buffers are reused by a real run: fault their pages in now, the
    # first DMA into never-touched pinned pages runs at a third of PCIe speed
    rho_out.array[:] = 0.0
    u_out.array[:] = 0.0
    # ... and let the device write them once: on the VM boxes the first DMA
    # into fresh pinned pages has been seen at a tenth of the PCIe rate even
    # after the CPU has touched them (572 ms instead of 59 ms for 3.2 GB)
    plb.download(capi.DENSITY_INNER, rho_out.array)
    plb.download(capi.VELOCITY_INNER, u_out.array)
    d2h = (rho_out.array.nbytes + u_out.array.nbytes + 48) * world

    def run_e2e(k):
        plb.sync()
        comm.Barrier()
        t0 = time.perf_counter()
        plb.event_record(2)
        moved = solver.upload_initial_fields(density=rho_in.array,
                                             velocity=u_in.array)
        plb.event_record(4)
        plb.initialize_pop()
        for _ in range(k - 1):
            solver.execute_single_time_step()
        solver.single_time_step(store_moments=True)
        plb.event_record(5)
        solver.residue_operator.compute_residues(st, solver.backend, comm, k)
        plb.event_record(6)
        plb.download(capi.DENSITY_INNER, rho_out.array)
        plb.download(capi.VELOCITY_INNER, u_out.array)
        plb.event_record(3)
        plb.sync()
        comm.Barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        total_ms = allmax(max(plb.event_elapsed_ms(2, 3), 0.0, wall_ms))
        h2d = moved * world
        return {"value": total_nodes * k / (total_ms * 1e-3) / 1e9,
                "unit": "GLUPS", "h2d_bytes_per_step": h2d / k,
                "h2d_bytes": h2d,
                "d2h_bytes_per_step": d2h / k, "ms_total": total_ms,
                "steps": k,
                "breakdown_ms": {
                    "upload": plb.event_elapsed_ms(2, 4),
                    "init_and_steps": plb.event_elapsed_ms(4, 5),
                    "residues": plb.event_elapsed_ms(5, 6),
                    "download": plb.event_elapsed_ms(6, 3)}}

    # the same run with more steps between the two copies (a production run
    # saves fields every few hundred steps): the PCIe share shrinks with K
    long_steps = 10 * steps if steps < 200 else 0
    e2e_long = run_e2e(long_steps) if long_steps else None
    e2e = run_e2e(steps)
    e2e["what"] = ("Solver.upload_initial_fields from pinned host arrays (a field "
                   "the case file fixes to one value -- here rho = 1 -- is "
                   "filled on the device, the velocity field is uploaded) + "
                   "initialize_pop + K python-issued steps + residues + "
                   "download rho,u (pinned)")
    mass = np.array([float(rho_out.array.sum())])
    mass_all = np.zeros_like(mass)
    comm.Allreduce(mass, mass_all, op="sum")
    e2e["global_mean_density"] = float(mass_all[0]) / total_nodes
    pcie_ms = e2e["breakdown_ms"]["upload"] + e2e["breakdown_ms"]["download"]
    e2e["pcie_share"] = pcie_ms / e2e["ms_total"]
    e2e["pcie_gbs"] = {"h2d": e2e["h2d_bytes"] / world / 1e6 / e2e["breakdown_ms"]["upload"],
                       "d2h": d2h / world / 1e6 / e2e["breakdown_ms"]["download"]}
    if e2e_long:
        e2e["same_run_with_10x_steps"] = {k: e2e_long[k] for k in
                                          ("value", "steps", "ms_total",
                                           "breakdown_ms")}
    if checker is not None:
        err = allmax(checker.compare(want_e2e, rho_out.array, u_out.array))
        e2e["parity_max_rel_err"] = err
        parity_out["e2e_max_rel_err"] = err
        parity_out["e2e_steps_compared"] = steps
        parity_out["ok"] = bool(parity_out["ok"] and err <= PARITY_RTOL)
    for buf in (rho_in, u_in, rho_out, u_out):
        buf.free()
    solver.close()
    return {"value": value, "ms": ms, "launches": launches, "scaling": scaling,
            "text": text, "roofline": roofline, "e2e": e2e, "clocks": clocks,
            "nx": nx, "ny": ny, "info": info, "parity": parity_out,
            "repeats": repeats, "repeat_ms": repeat_ms, "setup_s": setup_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="channel", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0,
                    help="shrink the lattice (debugging only; a scaled run is "
                         "not a benchmark value)")
    ap.add_argument("--cpu-sample", type=int, default=4096,
                    help="edge of the square lattice the CPU arm is timed on "
                         "(smaller: contract tests only)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the oracle check of the benchmarked run")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # relaunch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", os.environ.get("MASTER_PORT", "29511"),
                   os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    args.warmup = max(args.warmup, 3)

    _, _, _, scaling, text = workload_sizes(args.workload, world, args.scale)
    if args.impl == "reference":
        run_reference(args, rank, world, text, scaling)
        return

    from pylabolt_b200.comm import SingleComm, TorchComm
    comm = TorchComm() if world > 1 else SingleComm()
    res = measure(args.workload, comm, rank, world, args.steps, args.warmup,
                  args.scale, local_rank, parity=not args.no_parity)
    extra = {}
    other = None
    if not args.no_extras:
        # the other headline configuration (configs[3] strong scaling /
        # configs[4] weak scaling), same K, at this N
        other_name = "cavity" if args.workload == "channel" else "channel"
        other = measure(other_name, comm, rank, world, args.steps,
                        args.warmup, args.scale, local_rank,
                        parity=not args.no_parity)
        extra[other_name] = {
            "value": other["value"], "unit": "GLUPS", "n_gpus": world,
            "scaling": other["scaling"],
            "ms_per_step": other["ms"] / args.steps,
            "workload": other["text"], "repeats": other["repeats"],
            "roofline": other["roofline"], "e2e": other["e2e"],
            "parity": other["parity"], "clocks": other["clocks"],
            "setup_s": other["setup_s"]}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_cpu_port(args.workload, 200, 2, budget_s=15.0,
                            sample=args.cpu_sample)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        line = {
            "metric": "fluidLB D2Q9 fp64 lattice updates",
            "value": res["value"], "unit": "GLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms"] / args.steps, "higher_is_better": True,
            "scaling": res["scaling"], "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": res["text"], "nx": res["nx"], "ny": res["ny"],
                       "parallelism": f"x-slabs x{world}",
                       "l2": "lattice (19 GB/GPU) >> L2, no flush needed",
                       "kernel_variant": res["roofline"]["kernel"],
                       "steps_per_pass": res["roofline"]["steps_per_launch"],
                       "initial_state": "rho = 1, velocity noise of amplitude "
                                        f"0.01 with period {PERIOD} (x, y)",
                       "timed_region": f"{args.steps} steps, repeated "
                                       f"{res['repeats']}x back to back, median",
                       "scale": args.scale},
            "roofline": res["roofline"],
            "cpu_baseline": cpu,
            "e2e": res["e2e"],
            "parity": res["parity"],
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "repeats": res["repeats"],
            "repeat_ms": res["repeat_ms"],
            "setup_s": res["setup_s"],
            "hbm_roofline_frac_whole_step":
                res["value"] * 1e9 * ALGORITHMIC_BYTES_PER_NODE / world /
                (res["roofline"]["peak"] * 1e9),
        }
        if other is not None:
            # the second configuration's headline numbers at the top level too
            name = "cavity_16384_bgk" if args.workload == "channel" else \
                "channel_mrt_guo"
            line[name] = {
                "value": other["value"], "unit": "GLUPS", "n_gpus": world,
                "scaling": other["scaling"],
                "roofline_frac": other["roofline"]["frac"],
                "step_equivalent_frac": other["roofline"]["step_equivalent_frac"],
                "e2e": other["e2e"]["value"],
                "parity_max_rel_err": (other["parity"] or {}).get("max_rel_err"),
                "workload": other["text"]}
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
