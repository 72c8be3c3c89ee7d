"""Worker for test_distributed_gloo.py: one rank = one radial tile; the shared spectral sum is a
torch.distributed all-reduce on a zero-copy view of the library's buffer (gloo on CPU with the
test-only emulation build; the same host code runs NCCL on the GPU box)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main(case_name, out_path, lib_path, ntiles, exchange="columns"):
    import scythe_jl_b200 as S  # noqa: F401
    from helpers import model_cases, pkg_model
    from oracle import grids as G
    from scythe_jl_b200 import _lib
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _lib.load(lib_path)
    case = dict(model_cases()[case_name])
    case["n"] = int(os.environ.get("SB_TEST_STEPS", case["n"]))
    m = pkg_model(case, int(ntiles), lib, distributed=True, exchange=exchange)
    assert m.tile_count == int(ntiles) // world and m.tile_first == rank * m.tile_count
    # each rank only ever sees its own slice of the initial state
    patch = G.createGrid(case["gp"])
    tp = G.calcTileSizes(patch, int(ntiles))
    pts = np.concatenate([[0], np.cumsum(tp[4]).astype(np.int64)])
    ics = [case["ic"][pts[t]:pts[t + 1]] for t in range(m.tile_first, m.tile_first + m.tile_count)]
    m_exchange = (m.exchange, m.p2p)
    m.initialize_tiles(ics)
    m.run(case["n"])
    out = m.output()
    np.save(f"{out_path}.rank{rank}.npy", out)
    dist.barrier()
    m.close()
    if os.environ.get("SB_TEST_P2P_FAIL_RANK") is not None:
        # asymmetric peer-mapping failure: every rank must have dropped to the message form, none enabled p2p
        assert m_exchange == ("columns", False), m_exchange
    if os.environ.get("SB_TEST_CHECKPOINT"):      # every rank checkpoints to the SAME path: per-rank files, exact restart
        a = pkg_model(case, int(ntiles), lib, distributed=True, exchange=exchange)
        a.initialize_tiles(ics)
        a.run(2)
        written = a.checkpoint(f"{out_path}.ck.npz")
        assert f".tiles{a.tile_first}-" in written
        a.run(2)
        b = pkg_model(case, int(ntiles), lib, distributed=True, exchange=exchange)
        b.restore(f"{out_path}.ck.npz")
        assert b.t == 2
        b.run(2)
        for i in range(len(a.tiles)):
            for k in ("var_np1", "expdot_nm1", "expdot_nm2"):
                assert np.array_equal(a.state(i, k), b.state(i, k)), (rank, i, k)
        dist.barrier()
        a.close(); b.close()
    if os.environ.get("SB_TEST_HOST_PIPELINE"):   # pipelined host-driven stepping with the exchange between ranks inside the cycle
        from helpers import check_host_pipeline
        check_host_pipeline(case, lib, ntiles=int(ntiles), nsteps=3, distributed=True, exchange=exchange)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(*sys.argv[1:6])
