"""Multi-GPU check (torchrun, NCCL): every exchange mode must give bit-identical tile states, and those states must
match the oracle run with the same number of tiles.
    torchrun --nproc-per-node 2 tests/dist_gpu_check.py
Run by tests/test_gpu_parity.py::test_multi_gpu_exchange_modes (skipped below 2 GPUs) and, in reduced form, by
bench.py's multi-rank preflight (`parity_nranks` in the bench line).  The same host logic is covered on CPU by
test_distributed_gloo.py."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    from helpers import STATE_TOL, model_cases, pkg_model, rel_err, run_oracle
    from oracle import grids as G
    from scythe_jl_b200 import _lib
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _lib.load()
    outs = {}
    for name in ("LinearAdvectionRLZ", "LinearAdvectionRLZ_z16_fused", "Oneway_ShallowWater_HeightResolvedBL"):
        case = dict(model_cases()[name])
        case["n"] = 3
        ntiles = world
        patch = G.createGrid(case["gp"])
        tp = G.calcTileSizes(patch, ntiles)
        pts = np.concatenate([[0], np.cumsum(tp[4]).astype(np.int64)])
        for ex in ("torch", "columns", "columns-native", "columns-p2p", "columns-p2p-native"):
            m = pkg_model(case, ntiles, lib, distributed=True, exchange=ex, device=local)
            ics = [case["ic"][pts[t]:pts[t + 1]] for t in range(m.tile_first, m.tile_first + m.tile_count)]
            m.initialize_tiles(ics)
            m.run(case["n"])
            outs[(name, ex)] = m.state(0, "var_np1")
            m.close()
            dist.barrier()
            if ex in ("columns", "columns-p2p"):
                # host-driven cycle (tendency -> exchange -> physics), the path bench.py's e2e times: same state as stepping
                m = pkg_model(case, ntiles, lib, distributed=True, exchange=ex, device=local)
                m.initialize_tiles(ics)
                for _ in range(case["n"]):
                    m.cycle()
                same = np.array_equal(m.state(0, "var_np1"), outs[(name, ex)])
                print(f"rank {rank} {name} {ex}: cycle() == step(): {same}", flush=True)
                assert same
                m.close()
                dist.barrier()
        # the default in-step tileTransform! (needed slots / fused K4) against the all-slots dataflow, tile by tile
        m = pkg_model(case, ntiles, lib, distributed=True, exchange="columns-p2p", device=local)
        m.set_k3_slots("all")
        m.initialize_tiles(ics)
        m.run(case["n"])
        same = np.array_equal(m.state(0, "var_np1"), outs[(name, "columns-p2p")])
        print(f"rank {rank} {name}: needed/fused slots == all slots: {same}", flush=True)
        assert same
        m.close()
        dist.barrier()
        # parity proper: this rank's tile against the oracle integrated with the same tiles
        omt = run_oracle(case, ntiles).mtiles[rank]
        for ex in ("torch", "columns-p2p"):
            err = rel_err(outs[(name, ex)], omt.var_np1)
            print(f"rank {rank} {name} {ex}: vs oracle rel_err={err:.2e} (tol {STATE_TOL:g})", flush=True)
            assert err <= STATE_TOL
        # asymmetric peer-mapping failure: rank 1 cannot map its peers -> every rank agrees to use messages, nobody enables p2p
        os.environ["SB_TEST_P2P_FAIL_RANK"] = "1"
        m = pkg_model(case, ntiles, lib, distributed=True, device=local)      # default exchange (columns-p2p under NCCL)
        del os.environ["SB_TEST_P2P_FAIL_RANK"]
        assert (m.exchange, m.p2p) == ("columns", False), (m.exchange, m.p2p)
        m.initialize_tiles(ics)
        m.run(case["n"])
        same = np.array_equal(m.state(0, "var_np1"), outs[(name, "columns")])
        print(f"rank {rank} {name}: injected IPC failure on rank 1 -> all ranks fell back to messages, state identical: {same}", flush=True)
        assert same
        m.close()
        dist.barrier()
        ref = outs[(name, "torch")]
        for ex in ("columns", "columns-native", "columns-p2p", "columns-p2p-native"):
            same = np.array_equal(outs[(name, ex)], ref)
            err = float(np.abs(outs[(name, ex)] - ref).max() / np.abs(ref).max())
            print(f"rank {rank} {name} {ex}: identical={same} rel_err={err:.2e}", flush=True)
            assert err <= 1e-12
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
