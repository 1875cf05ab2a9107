"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import titancna_b200 as T
from titancna_b200 import _lib, build as B
from oracle import titan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    B.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "titancna_b200.h")).read()
    declared = set(re.findall(r"\b(titan_[a-z_A-Z0-9]+)\s*\(", hdr))
    declared -= {"titan_solver"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(L, name) is not None


def test_compute_fails_loudly_without_a_device():
    if T.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(T.TitanError) as ei:
        T.fwd_back_clonalCN(np.zeros(4), np.zeros((4, 3)), np.zeros(2), np.array([0.0, -1.0]), 2, np.arange(3.0), 1e15, 1e15)
    assert ei.value.code == _lib.ERR_NO_DEVICE
    p = T.loadDefaultParameters(5, 2)
    data = {"chr": np.ones(4), "posn": np.arange(4.0), "ref": np.full(4, 9.0), "tumDepth": np.full(4, 12.0), "logR": np.zeros(4)}
    with pytest.raises(T.TitanError):
        T.runEMclonalCN(data, p, verbose=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "titancna_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "titan_port" not in src and "libtitanref" not in src, f


@pytest.mark.parametrize("cn", [3, 4, 5, 6, 7, 8])
def test_load_default_parameters_matches_oracle_restatement(cn):
    rng = np.random.default_rng(cn)
    data = {"ref": rng.integers(10, 40, 200).astype(float), "tumDepth": np.full(200, 45.0), "logR": rng.normal(0, 0.2, 200)}
    for kw in (dict(), dict(skew=0.1), dict(hetBaselineSkew=0.0), dict(data=data)):
        a = T.loadDefaultParameters(copyNumber=cn, numberClonalClusters=3, **kw)
        b = O.load_default_parameters(cn, 3, **kw)
        for grp in a:
            for k, v in a[grp].items():
                if isinstance(v, np.ndarray):
                    assert np.array_equal(v, b[grp][k]), (grp, k)
                elif isinstance(v, float):
                    assert v == pytest.approx(b[grp][k], rel=1e-15), (grp, k)
    assert {6: 16, 3: 6, 4: 9, 5: 12, 7: 20, 8: 25}[cn] == a["genotypeParams"]["rt"].size


def test_load_default_parameters_errors():
    with pytest.raises(ValueError, match="Fewer than 3 or more than 8"):
        T.loadDefaultParameters(copyNumber=9)
    with pytest.raises(ValueError, match="alleleEmissionModel"):
        T.loadDefaultParameters(alleleEmissionModel="poisson")


def test_chromosome_offsets():
    assert T.chromosome_offsets(np.array([1, 1, 2, 2, 2, 5])).tolist() == [0, 2, 5, 6]
    assert T.chromosome_offsets(np.array(["1", "1", "X"])).tolist() == [0, 2, 3]
    with pytest.raises(ValueError):
        T.chromosome_offsets(np.array([1, 2, 1]))


def test_run_em_argument_checks_mirror_reference():
    p = T.loadDefaultParameters(5, 2)
    data = {"chr": np.ones(4), "posn": np.arange(4.0), "ref": np.full(4, 9.0), "tumDepth": np.full(4, 12.0)}
    bad = dict(p)
    del bad["normalParams"]
    with pytest.raises(ValueError, match="params must include genotypeParams"):
        T.runEMclonalCN(data, bad)
    with pytest.raises(T.TitanError, match="useOutlierState"):
        T.runEMclonalCN(data, p, useOutlierState=True)


def test_host_mstep_matches_oracle(golden):
    """titan_mstep is host code (O(U*Z) scalars, the reference runs it in R): compare the C++
    coordinate descent with the oracle's restatement on statistics from an oracle E-step."""
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    po = O.load_default_parameters(5, 2, skew=0.1)
    g, nP, sP, pP = po["genotypeParams"], po["normalParams"], po["cellPrevParams"], po["ploidyParams"]
    U, Z = 12, 2
    rng = np.random.default_rng(1)
    rho = rng.dirichlet(np.full(24, 0.05), size=3000).T.reshape((U, Z, 3000), order="F")
    sub = {k: v[:3000] for k, v in data.items()}
    es = {"rho": rho, "rhoG": rho.sum(1), "rhoZ": rho.sum(0)}
    want = O.m_step(sub, es, 0.5, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], np.array([0]), po, maxiterUpdate=1500)
    st = np.concatenate([want["stats"][k].reshape(-1, order="F") for k in "abcdeg"])
    keep = []
    from titancna_b200.hmm import _hyper
    h = _hyper(T.loadDefaultParameters(5, 2, skew=0.1), keep)
    out_s, out_var, out_piG, out_piZ = np.zeros(Z), np.zeros(U), np.zeros(U), np.zeros(Z)
    n, phi, prior = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    sweeps = ctypes.c_int()
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_lib.c_double_p)
    rG0 = np.ascontiguousarray(es["rhoG"][:, 0])
    rZ0 = np.ascontiguousarray(es["rhoZ"][:, 0])
    s0, v0, pg, pz = (np.ascontiguousarray(a, dtype=np.float64) for a in (sP["s_0"], g["var_0"], g["piG_0"], sP["piZ_0"]))
    rc = T.lib().titan_mstep(ctypes.byref(h), U, Z, 1, 1500, 1, 1, 1, P(st), P(rG0), P(rZ0), 0.5, P(s0), P(v0), 2.0, P(pg),
                             P(pz), ctypes.byref(n), P(out_s), P(out_var), ctypes.byref(phi), P(out_piG), P(out_piZ),
                             ctypes.byref(sweeps), ctypes.byref(prior))
    assert rc == 0, T.lib().titan_last_error()
    assert sweeps.value == want["sweeps"] == 10
    assert abs(n.value - want["n"]) < 1e-12 and abs(phi.value - want["phi"]) < 1e-12
    assert np.abs(out_s - want["s"]).max() < 1e-12 and np.abs(out_var - want["var"]).max() < 1e-14
    assert np.abs(out_piG - want["piG"]).max() < 1e-15 and np.abs(out_piZ - want["piZ"]).max() < 1e-15
    pr = O.prior_probs(want["n"], want["s"], want["phi"], want["var"], want["piG"], want["piZ"], g, nP, sP, pP)
    assert abs(prior.value - pr["prior"]) < 1e-9 * abs(pr["prior"])


def test_position_overlap_host_entry():
    posn = np.array([5.0, 50.0, 500.0])
    out = np.zeros(3)
    P = lambda a: a.ctypes.data_as(_lib.c_double_p)
    st, sp = np.array([1.0, 40.0]), np.array([10.0, 60.0])
    assert T.lib().titan_get_position_overlap(P(posn), 3, P(st), P(sp), 2, P(out)) == 0
    assert out.tolist() == [1.0, 2.0, 0.0]


def test_r_glue_registers_the_reference_call_table():
    """titancna_b200/csrc/r_glue.c compiled against the shim R.h: same routine names and arities as the
    reference's src/register.c:8-13, reference error() text on a mis-shaped obslik, and a loud failure
    (not a CPU fallback) when no device exists."""
    from oracle import refshim
    glue = refshim.ShimLibrary(B.GLUE)
    assert glue.registered() == {"getPositionOverlapC": 3, "fwd_backC_clonalCN": 9, "viterbiC_clonalCN": 9}
    if refshim.have_reference():
        assert glue.registered() == refshim.reference().registered()
    # Level 2 (r_glue_fused.c): the solver routines ride in the same table
    assert glue.registered(("titan_solver_create_R", "titan_em_run_R", "titan_viterbi_R")) == {
        "titan_solver_create_R": 9, "titan_em_run_R": 8, "titan_viterbi_R": 10}
    keep = []
    with pytest.raises(refshim.RError, match="not a live titan solver"):
        glue.call("titan_viterbi_R", *[glue._real([0.0], keep) for _ in range(10)])
    with pytest.raises(refshim.RError, match="equal length"):
        glue.call("titan_solver_create_R", glue._real([0, 3], keep), glue._real([1, 2, 3], keep), glue._real([1, 2], keep),
                  glue._real([1, 2, 3], keep), glue._real([1, 2, 3], keep), glue._real([1e15], keep), glue._real([1.0], keep),
                  glue._real([0, 1], keep), glue._real([1], keep))
    glue.lib.shim_free_all()
    with pytest.raises(refshim.RError, match="The obslik must be 4-by-5 dimension"):
        glue.fwd_back(np.zeros(4), np.zeros((3, 5)), np.zeros(2), np.arange(2.0), 2, np.arange(5.0), 1e15, 1e15)
    with pytest.raises(refshim.RError, match="The obslik must be 4-by-5 dimension"):
        glue.viterbi(np.zeros(4), np.zeros((3, 5)), np.zeros(2), np.arange(2.0), 2, np.arange(5.0), 1e15, 1e15)
    assert glue.position_overlap([5, 50, 500], [1, 40], [10, 60]).tolist() == [1.0, 2.0, 0.0]
    if T.device_count() == 0:
        with pytest.raises(refshim.RError, match="CUDA"):
            glue.fwd_back(np.zeros(4), np.zeros((4, 5)), np.zeros(2), np.array([0.0, -1.0]), 2, np.arange(5.0), 1e15, 1e15)


def test_data_moments_reuse_gives_identical_parameters():
    """loadDefaultParameters(moments=dataMoments(data)) -- what the sweep does once per data set -- equals the
    per-call computation (R/utils.R:32-40,80-95) bit for bit."""
    rng = np.random.default_rng(3)
    data = {"ref": rng.integers(10, 40, 500).astype(float), "tumDepth": np.full(500, 45.0), "logR": rng.normal(0, 0.2, 500)}
    m = T.dataMoments(data)
    for Z in (1, 4):
        a = T.loadDefaultParameters(8, Z, data=data)
        b = T.loadDefaultParameters(8, Z, data=data, moments=m)
        for grp in a:
            for k, v in a[grp].items():
                assert np.array_equal(np.asarray(v, dtype=object), np.asarray(b[grp][k], dtype=object)), (grp, k)


def test_extended_genotype_table_is_opt_in():
    """BASELINE configs[3] as named needs maxCN = 10: the reference's limit (R/utils.R:9-12) stays the default."""
    with pytest.raises(ValueError, match="more than 8 copies"):
        T.loadDefaultParameters(copyNumber=10, numberClonalClusters=5)
    g = T.loadDefaultParameters(copyNumber=10, numberClonalClusters=5, extendedTable=True)["genotypeParams"]
    assert g["rt"].size == 36 and g["ct"].tolist().count(9) == 5 and g["ct"].tolist().count(10) == 6
    assert np.allclose(g["rt"][25:30], [1, 8 / 9, 7 / 9, 6 / 9, 5 / 9]) and np.allclose(g["rt"][30:35], [1, .9, .8, .7, .6])
    assert g["rt"][35] == g["rt"][3] and g["kappaGHyper"][35] == 500 and np.count_nonzero(g["ZS"] == -1) == 1
    # the first 25 states are the reference's table
    g8 = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=5)["genotypeParams"]
    assert np.array_equal(g["rt"][:25], g8["rt"]) and np.array_equal(g["ct"][:25], g8["ct"])


def test_solver_refuses_mismatched_column_lengths():
    """The C ABI takes bare pointers: the host mirror checks every length before the call (no device needed)."""
    p = T.loadDefaultParameters(5, 2)
    data = {"chr": np.ones(4), "posn": np.arange(4.0), "ref": np.full(3, 9.0), "tumDepth": np.full(4, 12.0), "logR": np.zeros(4)}
    with pytest.raises(ValueError, match=r"data\$ref has 3 entries"):
        T.Solver(data, p["genotypeParams"], 2, 1e15, 1.0)
    g = dict(p["genotypeParams"], ZS=p["genotypeParams"]["ZS"][:-1])
    data["ref"] = np.full(4, 9.0)
    with pytest.raises(ValueError, match="fewer entries"):
        T.Solver(data, g, 2, 1e15, 1.0)


def test_communicator_entry_points_fail_loudly_without_a_device_or_nccl():
    if T.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(T.TitanError):
        T.Comm(b"\0" * 128, 0, 1)
    with pytest.raises(ValueError, match="128 bytes"):
        T.Comm(b"short", 0, 1)


def test_reference_arm_sample_fits_its_time_budget():
    """bench.py --impl reference sizes its per-chromosome prefix sample from a calibration run so that steps + warm-up
    stay inside --ref-budget-s, and says which fraction of the workload that is."""
    import json
    import subprocess
    import sys
    import time
    t0 = time.time()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--snps", "60000", "--steps", "2",
                          "--warmup", "1", "--ref-budget-s", "6"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # stdout carries the one JSON line and nothing else
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["steps"] == 2 and rec["warmup"] == 1 and rec["gpu_launches"] == 0
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["e2e"]["h2d_bytes_per_step"] == 0
    assert rec["value"] > 0 and time.time() - t0 < 120


def test_interleaved_viterbi_table_builder_is_bit_identical(tmp_path):
    """Host code on the decode's critical path: the library builds the exact log-transition tables eight rows at a time
    (independent add chains, logs reused for equal sums); every table must equal the plain row-by-row restatement of the
    reference's sequential row sums bit for bit (1800 random (rhoG, rhoZ) pairs over six state layouts)."""
    import subprocess
    exe = tmp_path / "host_tables_check"
    subprocess.run(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "titancna_b200", "csrc"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "host_tables_check.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "mismatching tables: 0" in out.stdout, out.stdout + out.stderr


def test_current_device_query_fails_loudly_without_a_device():
    """titan_get_device is what run_sweep reads before it fans solutions out to worker threads (which start on CUDA
    device 0, not on their creator's device)."""
    if T.device_count() > 0:
        assert 0 <= _lib.current_device() < T.device_count()
        return
    with pytest.raises(T.TitanError) as ei:
        _lib.current_device()
    assert ei.value.code == _lib.ERR_NO_DEVICE
