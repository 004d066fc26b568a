"""``pylabolt`` command line: mirror of pylabolt/pylabolt.py:5-57.

Same flags as the reference; ``--backend`` gains (and defaults to) ``b200``,
``--toVTK`` (README spelling) is accepted for ``--to_vtk``.  Reconstruction and
VTK conversion (pylabolt_b200/postprocess.py) work on the unchanged .npz /
metadata.json output, like the reference's own tools.
"""
from argparse import ArgumentParser


def build_parser():
    parser = ArgumentParser(description="A Lattice Boltzmann Python solver "
                                        "(B200 back end)")
    parser.add_argument("-s", "--solver",
                        choices=["fluidLB", "phaseFieldLB", "cgLB"], type=str,
                        help="choice of solver to run")
    parser.add_argument("-b", "--backend", choices=["b200", "gpu", "cpu"],
                        default="b200", type=str,
                        help="choice of backend (only b200 is built here)")
    parser.add_argument("-nt", "--n_threads", type=int, default=1,
                        help="kept for compatibility; unused on b200")
    parser.add_argument("--reconstruct", choices=["all", "time", None],
                        default=None, help="Domain reconstruction")
    parser.add_argument("-t", "--time", type=int, default=0,
                        help="Specify time which is to be reconstructed")
    parser.add_argument("--to_vtk", "--toVTK", dest="to_vtk",
                        choices=["all", "time", None], default=None,
                        help="Convert output data to VTK format")
    parser.add_argument("--debug", action="store_true", default=False,
                        help="Run in debug mode")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.solver == "fluidLB":
        from . import solver
        solver.main(args.backend, args.n_threads, debug_mode=args.debug)
    elif args.solver is not None:
        raise SystemExit(f"solver {args.solver} is outside the b200 build "
                         "(fluidLB is the accelerated path)")
    if args.reconstruct is not None:
        from .postprocess import reconstruct_data
        reconstruct_data(args.reconstruct, time_step=args.time)
    if args.to_vtk is not None:
        from .postprocess import convert_to_vtk
        convert_to_vtk(args.to_vtk, args.time)


if __name__ == "__main__":
    main()
