"""CPU: the C-ABI library loads and exports every symbol include/bgp.h declares (no compute without a
GPU), fails loudly without CUDA, and the host-side term setup agrees with the oracle's constructors."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from bayesgp_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(lib, nm), nm
        assert nm in _lib.SIGNATURES, "no ctypes prototype for %s" % nm
    assert lib.bgp_version() >= 100
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for nm in _lib.SIGNATURES:
        getattr(raw, nm)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (BGP_ERR_CUDA), never fall back."""
    import subprocess
    import sys
    code = (
        "import numpy as np, bayesgp_b200 as b\n"
        "try:\n"
        "    b.LaplaceObjective(y=np.ones(8), family='Poisson')\n"
        "    print('CONSTRUCTED')\n"
        "except b.BgpError as e:\n"
        "    print('ERR', e.code)\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120).stdout
    assert "ERR 2" in out, out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesgp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_term_setup_matches_oracle_constructors():
    from bayesgp_b200 import terms as pt
    from oracle import basis as ob
    from oracle.fit import Term as OTerm, build_term
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.3, 1.2, 400)
    t = pt.prepare_term(pt.Term("sGP", "x", x, a=2 * np.pi * 3, k=9, m=2, region=np.array([0.0, 1.5])))
    o = build_term(OTerm("sGP", "x", x, a=2 * np.pi * 3, k=9, m=2, region=np.array([0.0, 1.5])))
    # design and precision of an sGP term are built on the device (tests/test_gpu_edge.py compares them with the
    # oracle's constructors); the host only sizes the blocks
    assert t.n_basis == o.B.shape[1] == o.P.shape[0] and t.n_boundary == o.X.shape[1]
    ti = pt.prepare_term(pt.Term("IWP", "x", x, order=3, k=11))
    oi = build_term(OTerm("IWP", "x", x, order=3, k=11))
    assert np.allclose(ti.knots, oi.knots) and ti.initial_location == oi.initial_location
    assert ti.n_basis == oi.B.shape[1] == 10 and ti.n_boundary == oi.X.shape[1] == 2
    assert np.array_equal(ti.observed_x, oi.observed_x)
    # mixed-sign knots: the negative / positive split of local_poly_helper
    tn = pt.prepare_term(pt.Term("IWP", "x", x, order=2, knots=np.array([-0.3, -0.1, 0.0, 0.5, 1.2]), initial_location=0.0))
    on = build_term(OTerm("IWP", "x", x, order=2, knots=np.array([-0.3, -0.1, 0.0, 0.5, 1.2]), initial_location=0.0))
    assert tn.n_basis == on.B.shape[1] == 4
    tid = pt.prepare_term(pt.Term("IID", "g", np.array([3, 1, 3, 2, 1.0])))
    Bi, Pi = pt.iid_design(tid)
    assert Bi.shape == (5, 3) and Pi.tolist() == [1, 1, 1] and Bi.sum() == 5


def test_argument_validation_mirrors_reference_messages():
    from bayesgp_b200 import terms as pt
    with pytest.raises(ValueError, match="should be >= 3"):
        pt.prepare_term(pt.Term("IWP", "x", np.arange(5.0), order=2, k=2))
    with pytest.raises(ValueError, match="should be >= 1"):
        pt.prepare_term(pt.Term("IWP", "x", np.arange(5.0), order=0, k=4))
    with pytest.raises(ValueError, match="should be positive"):
        pt.prepare_term(pt.Term("sGP", "x", np.arange(5.0), a=-1.0))


def test_workload_generators_are_seeded():
    from bayesgp_b200.workloads import c3_data, gh_nodes, iwp_knots
    x1, y1 = c3_data(1000)
    x2, y2 = c3_data(1000)
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2)
    x0, kn = iwp_knots(x1, 30)
    assert len(kn) == 30 and kn[0] == 0.0 and x0 == x1.min()
    from oracle.aghq import gh_rule
    assert np.allclose(gh_nodes(15), gh_rule(15)[0], atol=1e-13)


def _header_arity():
    """name -> number of parameters, parsed from include/bgp.h."""
    src = open(os.path.join(ROOT, "include", "bgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(bgp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def _call_arity(text, name):
    """argument counts of every call of `name(` in C source text (balanced parentheses)."""
    counts = []
    for m in re.finditer(r"\b%s\s*\(" % re.escape(name), text):
        depth, i, commas, seen = 1, m.end(), 0, False
        while depth and i < len(text):
            ch = text[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                commas += 1
            elif not ch.isspace():
                seen = True
            i += 1
        counts.append(commas + 1 if seen else 0)
    return counts


def test_r_shim_binds_only_declared_entry_points_with_matching_arity():
    """integration/r/bgp_rcall.c cannot be compiled here (no R headers); at least every bgp_* it calls must exist in
    include/bgp.h with the same number of arguments, and its .Call table must match its own definitions."""
    shim = open(os.path.join(ROOT, "integration", "r", "bgp_rcall.c")).read()
    code = re.sub(r"/\*.*?\*/", "", shim, flags=re.S)
    arity = _header_arity()
    used = sorted(set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", code)))
    assert len(used) >= 12
    for nm in used:
        assert nm in arity, "%s is not declared in include/bgp.h" % nm
        for k in _call_arity(code, nm):
            assert k == arity[nm], (nm, k, arity[nm])
    table = dict((n, int(k)) for n, k in re.findall(r'\{"(bgpR_[a-z_]+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', code))
    defs = dict((m.group(1), m.group(2).count("SEXP")) for m in re.finditer(r"^SEXP (bgpR_[a-z_]+)\(([^)]*)\)", code, flags=re.M))
    assert table and table == defs, (table, defs)
    rglue = open(os.path.join(ROOT, "integration", "r", "bgp_shim.R")).read()
    for nm in set(re.findall(r'\.Call\("(bgpR_[a-z_]+)"', rglue)):
        assert nm in table, nm


def test_fmm_spline_and_integrate_xy_match_the_oracle_restatement():
    """model_fit_loop's normalisation (sfsmisc::integrate.xy = stats::spline(method = "fmm") on max(1024, 3 n) points +
    trapezoid, R/02_model_fit.R:774): the product's tridiagonal recurrences against the oracle's dense solve of the
    defining end conditions, plus two closed forms."""
    from bayesgp_b200.post_fit import fmm_spline, integrate_xy
    from oracle.fit import fmm_spline_eval, integrate_xy as o_integrate
    rng = np.random.default_rng(3)
    for m in (2, 3, 4, 7, 25):
        x = np.sort(rng.uniform(-1.0, 4.0, m))
        y = np.exp(-0.5 * (x - 1.5) ** 2) + 0.05 * rng.standard_normal(m)
        xo, yo = fmm_spline(x, y, 301)
        assert xo[0] == x[0] and abs(xo[-1] - x[-1]) < 1e-15
        assert np.max(np.abs(yo - fmm_spline_eval(x, y, xo))) < 1e-12
        assert abs(integrate_xy(x, y) - o_integrate(x, y)) < 1e-12
    # a cubic is reproduced exactly by the fmm end conditions (four-point cubics at both ends)
    x = np.linspace(0.0, 2.0, 9)
    cubic = lambda t: 1.0 + 2.0 * t - 0.5 * t ** 2 + 0.25 * t ** 3
    xo, yo = fmm_spline(x, cubic(x), 101)
    assert np.max(np.abs(yo - cubic(xo))) < 1e-12
    assert abs(integrate_xy(x, cubic(x)) - (2.0 + 4.0 - 0.5 * 8.0 / 3.0 + 0.25 * 16.0 / 4.0)) < 1e-5   # trapezoid on 1024 points
