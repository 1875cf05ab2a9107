"""GPU: one solution with its chromosomes sharded over ranks (SURVEY.md 8e, R/hmmClonal.R:203-234) -- the
in-library NCCL all-reduce of the statistics (titan_solver_set_comm) against the single-GPU EM on the same data.
Runs with 2 ranks when the box has 2+ GPUs; on a 1-GPU box the same code path runs with a 1-rank communicator."""
import ast
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import titancna_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_em_equals_single_gpu_em():
    world = 2 if T.device_count() >= 2 else 1
    env = dict(os.environ, SHARD_N="120000", SHARD_Z="3")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "run_sharded.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rec = [ast.literal_eval(l) for l in r.stdout.splitlines() if l.startswith("{'world'")][0]
    assert rec["world"] == world and rec["iters"] >= 2
    assert max(rec["max_abs_err"].values()) < 1e-7 and rec["loglik_rel_err"] < 1e-10
    assert rec["nccl_version"] >= 20000


def test_comm_layout_of_a_partial_shard():
    """A 1-rank communicator over a solver that holds only SOME chromosomes of a larger genome: the outputs come back
    in the whole-genome layout, zero where no rank holds the chromosome."""
    from tests import synth
    data, _ = synth.make_data(6000, Z=2, maxCN=5, n_chr=4, seed=5)
    p = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2, data=data)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    args = (0.4, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    full = T.Solver(data, g, 2, 1e15, 1.0)
    ref = full.estep(*args)
    full.close()
    off = T.chromosome_offsets(data["chr"])
    keep = np.array([1, 3])
    idx = np.concatenate([np.arange(off[c], off[c + 1]) for c in keep])
    sub = {k: np.asarray(v)[idx] for k, v in data.items()}
    comm = T.Comm(T.Comm.unique_id(), 0, 1)
    sol = T.Solver(sub, g, 2, 1e15, 1.0)
    sol.set_comm(comm, 4, keep)
    out = sol.estep(*args)
    alone = T.Solver(sub, g, 2, 1e15, 1.0)
    want = alone.estep(*args)
    alone.close()
    sol.close()
    comm.close()
    assert out["loglik_chr"].shape == (4,) and out["rhoG0"].shape[1] == 4
    assert np.all(out["loglik_chr"][[0, 2]] == 0) and np.all(out["rhoG0"][:, [0, 2]] == 0)
    assert np.array_equal(out["loglik_chr"][keep], want["loglik_chr"])
    assert np.allclose(out["loglik_chr"][keep], ref["loglik_chr"][keep], rtol=1e-12)
    assert np.array_equal(out["rhoG0"][:, keep], want["rhoG0"]) and np.array_equal(out["rhoZ0"][:, keep], want["rhoZ0"])
    assert np.array_equal(out["stats_flat"], want["stats_flat"])
