"""``python -m salib_b200.scripts.salib`` -- the Sobol' subset of the ``salib`` command line.

Same sub-commands, options and file formats as the reference CLI for the methods on the hot path
(src/SALib/scripts/salib.py:33-63, sample/common_args.py, analyze/common_args.py,
sample/sobol.py:192-251, sample/saltelli.py:205-264, analyze/sobol.py:442-491)::

    salib sample saltelli -n 1024 -p params.txt -o X.txt [--max-order 2] [--skip-values M] [--precision 8]
    salib sample sobol    -n 1024 -p params.txt -o X.txt [--max-order 2] [--scramble 1] [-s SEED]
    salib analyze sobol   -p params.txt -Y Y.txt [-c 0] [--max-order 2] [-r 100] [-s SEED]
    salib sample finite_diff -n 1000 -p params.txt -o X.txt [-d 0.01]
    salib analyze dgsm    -p params.txt -X X.txt -Y Y.txt [-c 0] [-r 1000] [-s SEED]
    salib analyze discrepancy -p params.txt -X X.txt -Y Y.txt [-c 0]

Text I/O stays NumPy's (``np.savetxt`` with ``%.<precision>e``, ``np.loadtxt``) so files are byte-identical to
the reference's; the numbers come from the CUDA path."""
import argparse
import sys

import numpy as np

from ..util.textio import loadtxt, savetxt
from ..util.util_funcs import read_param_file

SAMPLERS = ("saltelli", "sobol", "finite_diff")
ANALYZERS = ("sobol", "dgsm", "discrepancy")


def _sample_parser(method):
    p = argparse.ArgumentParser(prog=f"salib sample {method}",
                                description="Create parameter samples for sensitivity analysis")
    p.add_argument("-n", "--samples", type=int, required=True, help="Number of Samples")
    p.add_argument("-p", "--paramfile", type=str, required=True, help="Parameter Range File")
    p.add_argument("-o", "--output", type=str, required=True, help="Output File")
    if method != "saltelli":  # the reference removes the seed option for saltelli
        p.add_argument("-s", "--seed", type=int, required=False, default=None, help="Random Seed")
    p.add_argument("--delimiter", type=str, required=False, default=" ", help="Column delimiter")
    p.add_argument("--precision", type=int, required=False, default=8, help="Output floating-point precision")
    if method == "finite_diff":  # sample/finite_diff.py:96-110
        p.add_argument("-d", "--delta", type=float, required=False, default=0.01,
                       help="Finite difference step size (percent)")
        return p
    p.add_argument("--max-order", type=int, required=False, default=2, choices=[1, 2],
                   help="Maximum order of sensitivity indices to calculate")
    if method == "sobol":
        p.add_argument("--scramble", type=int, required=False, default=True, help="Use scrambled sequence")
    p.add_argument("--skip-values", type=int, required=False, default=None,
                   help="Number of sample points to skip (default: next largest power of 2 from `samples`)")
    return p


def _analyze_parser(method):
    p = argparse.ArgumentParser(prog=f"salib analyze {method}",
                                description="Perform sensitivity analysis on model output")
    p.add_argument("-p", "--paramfile", type=str, required=True, help="Parameter range file")
    p.add_argument("-Y", "--model-output-file", type=str, required=True, help="Model output file")
    p.add_argument("-c", "--column", type=int, required=False, default=0, help="Column of output to analyze")
    p.add_argument("--delimiter", type=str, required=False, default=" ", help="Column delimiter in model output file")
    p.add_argument("-s", "--seed", type=int, required=False, default=None, help="Random Seed")
    if method == "discrepancy":  # analyze/discrepancy.py:124-146 (its cli_action passes only the seed on)
        p.add_argument("-X", "--model-input-file", type=str, required=True, help="Model input file")
        return p
    if method == "dgsm":  # analyze/dgsm.py:143-161
        p.add_argument("-X", "--model-input-file", type=str, required=True, default=None, help="Model input file")
        p.add_argument("-r", "--resamples", type=int, required=False, default=1000,
                       help="Number of bootstrap resamples for confidence intervals")
        return p
    p.add_argument("--max-order", type=int, required=False, default=2, choices=[1, 2],
                   help="Maximum order of sensitivity indices to calculate")
    p.add_argument("-r", "--resamples", type=int, required=False, default=100,
                   help="Number of bootstrap resamples for Sobol confidence intervals")
    p.add_argument("--parallel", action="store_true", dest="parallel", help="Accepted for compatibility (ignored)")
    p.add_argument("--processors", type=int, required=False, default=None, dest="n_processors",
                   help="Accepted for compatibility (ignored)")
    return p


def run_sample(method, opts):
    args = _sample_parser(method).parse_args(opts)
    problem = read_param_file(args.paramfile)
    if method == "finite_diff":
        from ..sample import finite_diff

        X = finite_diff.sample(problem, args.samples, args.delta, seed=args.seed)
        savetxt(args.output, X, delimiter=args.delimiter, precision=args.precision)
        return
    second = args.max_order == 2
    if method == "saltelli":
        from ..sample import saltelli

        X = saltelli.sample(problem, args.samples, calc_second_order=second, skip_values=args.skip_values)
    else:
        from ..sample import sobol

        # like the reference's cli_action (sample/sobol.py:240-245), --skip-values and --seed are parsed
        # but not forwarded
        X = sobol.sample(problem, args.samples, calc_second_order=second, scramble=args.scramble)
    savetxt(args.output, X, delimiter=args.delimiter, precision=args.precision)


def run_analyze(method, opts):
    args = _analyze_parser(method).parse_args(opts)
    from ..analyze import sobol

    problem = read_param_file(args.paramfile)
    Y = loadtxt(args.model_output_file, delimiter=args.delimiter, usecols=(args.column,))
    if method == "discrepancy":
        from ..analyze import discrepancy

        X = loadtxt(args.model_input_file, delimiter=args.delimiter)
        discrepancy.analyze(problem, X, Y, print_to_console=True, seed=args.seed)
        return
    if method == "dgsm":
        from ..analyze import dgsm

        X = loadtxt(args.model_input_file, delimiter=args.delimiter, ndmin=2)
        dgsm.analyze(problem, X, Y, num_resamples=args.resamples, print_to_console=True, seed=args.seed)
        return
    sobol.analyze(problem, Y, (args.max_order == 2), num_resamples=args.resamples, print_to_console=True,
                  parallel=args.parallel, n_processors=args.n_processors, seed=args.seed)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    top = argparse.ArgumentParser(prog="salib", description="salib_b200 - Sobol' sensitivity analysis on B200")
    sub = top.add_subparsers(help="Sample or Analysis method", dest="action")
    sub.add_parser("sample", add_help=False).add_argument("method", choices=SAMPLERS)
    sub.add_parser("analyze", add_help=False).add_argument("method", choices=ANALYZERS)
    known, opts = top.parse_known_args(argv)
    if not known.action:
        top.print_help()
        return 0
    if known.action == "sample":
        if not opts:
            _sample_parser(known.method).print_help()
            return 0
        run_sample(known.method, opts)
    else:
        if not opts:
            _analyze_parser(known.method).print_help()
            return 0
        run_analyze(known.method, opts)
    return 0


if __name__ == "__main__":
    sys.exit(main())
