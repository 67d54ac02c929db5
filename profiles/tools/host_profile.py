"""Host-side cost of one C5 Bermudan valuation WITHOUT a GPU: the ctypes binding is pointed at profiles/tools/null_abi.c (handles only,
no arithmetic), so the time measured is the Python mirror + binding alone.  Profiling aid; the product never loads the stub.

    gcc -O2 -shared -fPIC -I include profiles/tools/null_abi.c -o /tmp/libnull_abi.so
    python profiles/tools/host_profile.py [--profile] [--paths N]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--paths", type=int, default=1_000_000)
    ap.add_argument("--lib", default="/tmp/libnull_abi.so")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    pkg = graft.load_package()
    nv = pkg.native
    nv.LIB_PATH = args.lib                                  # the stub instead of libfinmath_b200.so (profiling only)
    nv.load()
    from common import lmm_setup, bermudan_spec
    s = lmm_setup(pkg)
    b = bermudan_spec(s)
    factory = pkg.RandomVariableCudaFactory()
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, s["cov"], None, {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
    product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])

    def valuation():
        bm = pkg.BrownianMotionCuda(s["sim"], s["F"], args.paths, 3141, factory)
        sim = pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(model, bm, 2))
        t0 = time.perf_counter()
        sim.getProcess().getProcessValue(s["T"], s["N"] - 1)
        t1 = time.perf_counter()
        price = product.getValue(sim)
        return t1 - t0, time.perf_counter() - t1, price

    valuation()
    l0 = nv.launch_count()
    runs = [valuation()[:2] for _ in range(args.reps)]
    best = (min(r[0] for r in runs), min(r[1] for r in runs))
    calls = (nv.launch_count() - l0) / args.reps
    print("host only (null ABI): simulate set-up %.2f ms, Bermudan induction %.2f ms, %.0f native launches per valuation" % (1e3 * best[0], 1e3 * best[1], calls))
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            valuation()
        pr.disable()
        pstats.Stats(pr).sort_stats(os.environ.get("FMB_PROFILE_SORT", "tottime")).print_stats(int(os.environ.get("FMB_PROFILE_ROWS", "28")))


if __name__ == "__main__":
    main()
