"""N>1 host logic on CPU: world_size 2, gloo backend, one or two radial tiles per rank.  The tiles'
B coefficients meet in an all-reduce of the patch-sized shared buffer (own block + 3-coefficient
halo written by each rank, zeros elsewhere), which reproduces the reference's SharedArray +
RemoteChannel overlap-add bit for bit because every shared coefficient has at most two addends."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import STATE_TOL, model_cases, pkg_model, run_oracle, slot_errs

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("case_name,ntiles,exchange", [("LinearAdvection1D", 2, "torch"), ("LinearAdvection1D", 2, "columns"),
                                                       ("LinearAdvectionRLZ", 2, "columns"), ("LinearAdvectionRZ", 4, "columns"),
                                                       ("LinearAdvectionRL", 2, "columns-p2p")])
def test_two_ranks_match_single_process_and_oracle(case_name, ntiles, exchange, emu_lib, tmp_path):
    out = tmp_path / "out"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 2000), OMP_NUM_THREADS="1")
    if case_name == "LinearAdvection1D":     # also: Model.cycle_host == set_state/cycle/get_state across two ranks
        env["SB_TEST_HOST_PIPELINE"] = "1"
    if case_name == "LinearAdvectionRLZ":    # also: checkpoint() on every rank with one path, restore(), exact continuation
        env["SB_TEST_CHECKPOINT"] = "1"
    if exchange == "columns-p2p":            # peer mapping fails on rank 1 only: all ranks must agree to fall back before enabling
        env.update(SB_P2P_FALLBACK="1", SB_TEST_P2P_FAIL_RANK="1")
    nsteps = min(model_cases()[case_name]["n"], 4)
    env["SB_TEST_STEPS"] = str(nsteps)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", env["MASTER_PORT"], str(ROOT / "tests" / "dist_worker.py"), case_name, str(out),
           str(emu_lib.path), str(ntiles), exchange]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    o0, o1 = np.load(f"{out}.rank0.npy"), np.load(f"{out}.rank1.npy")
    assert np.array_equal(o0, o1)                      # every rank holds the same patch coefficients
    case = dict(model_cases()[case_name], n=nsteps)
    m = pkg_model(case, ntiles, emu_lib)               # same tiles, one process
    m.initialize(case["ic"])
    m.run(case["n"])
    single = m.output()
    m.close()
    assert max(slot_errs(o0, single)) <= 1e-12
    oracle = run_oracle(case, ntiles).output_patch()
    assert max(slot_errs(o0[:, :, :2], oracle[:, :, :2])) <= STATE_TOL
