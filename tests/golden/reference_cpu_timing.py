#!/usr/bin/env python
"""Times the REFERENCE's own CPU path (numba kernels, Solver.single_time_step)
next to the oracle port that bench.py's cpu_baseline / --impl reference use,
on the same case and thread count -- in a container that holds the reference
checkout (the GPU box does not, which is why the bench times the port).  Test
infrastructure, like make_golden.py next to it: it checks the checker.

    python tests/golden/reference_cpu_timing.py [--n 1024] [--steps 10] [--threads 8]
        > profiles/<round>_cpu_reference_vs_port.json

The point: the port is a fair stand-in -- it is not slower than the code it
stands for (it is the same five passes over the same AoS arrays, compiled by
gcc -O2 with OpenMP instead of numba prange).  The reference's setup has
per-node Python loops (base/fields.py:166-179), so the lattice is kept small.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, HERE)


def build_reference(n, steps, threads):
    import cases
    import make_golden
    comm = make_golden.install_shims()
    sim = cases.cavity(n, n, end_time=steps)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                return make_golden.build_reference_solver(sim, comm, n_threads=threads)
        finally:
            os.chdir(cwd)


def build_port(n, steps, threads):
    import bench
    import cases
    return bench.make_oracle(cases.cavity(n, n, end_time=steps), threads)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    import numba
    reference = build_reference(args.n, args.steps, args.threads)
    port = build_port(args.n, args.steps, args.threads)
    nodes = args.n * args.n
    out = {"case": f"lid-driven cavity {args.n}x{args.n}, BGK, fp64 (BASELINE configs[0] scaled up)",
           "host_cores": os.cpu_count(), "steps_per_round": args.steps,
           "rounds": args.rounds, "how": "the two codes timed alternately, best round of each"}
    for threads in (args.threads, 1):
        numba.set_num_threads(threads)
        port.lib.oracle_set_threads(threads)
        steps = args.steps if threads > 1 else max(2, args.steps // 3)
        best = {"reference_numba": 0.0, "oracle_port": 0.0}
        for _ in range(args.rounds + 1):          # first round = warm-up
            t0 = time.perf_counter()
            for _ in range(steps):
                reference.single_time_step()
            t1 = time.perf_counter()
            port.step(steps)
            t2 = time.perf_counter()
            if _ > 0:
                best["reference_numba"] = max(best["reference_numba"],
                                              nodes * steps / (t1 - t0) / 1e6)
                best["oracle_port"] = max(best["oracle_port"],
                                          nodes * steps / (t2 - t1) / 1e6)
        out[f"{threads}_threads"] = {
            "reference_numba_mlups": round(best["reference_numba"], 1),
            "oracle_port_mlups": round(best["oracle_port"], 1),
            "port_over_reference": round(best["oracle_port"] / best["reference_numba"], 2)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
