"""Launcher replacing ``run_Scythe.jl`` (/root/reference/run_Scythe.jl:1-46) for the B200 path.

    python -m scythe_jl_b200.run [-w WORKERS] [--gpus G] [--format csv|netcdf|both] \
        [--checkpoint FILE [--checkpoint-interval SECONDS]] [--restart FILE] model_file

* ``model_file`` is the reference's own argument: a Julia file whose body is ``model = ModelParameters(...)``
  (read by `modelfile.load_model_file`, not executed), or a ``.py`` file defining ``model``.
* ``-w / --workers`` keeps its meaning: the number of radial tiles (one per reference worker process,
  src/semiimplicit.jl:155-169).  ``--gpus G`` says how many GPUs of this node share them: with G = 1 all tiles run on one
  device in one launch sequence; with G > 1 the launcher re-executes itself under ``torch.distributed.run`` with one
  process per GPU and one tile per process (WORKERS defaults to G and must equal it), NCCL over NVLink for the
  exchange.  The reference's ``--sge`` / ``--email`` cluster options (ClusterManagers.addprocs_sge) have no equivalent
  on a single NVSwitch box and are rejected with a message rather than ignored.
* Under an external ``torchrun`` (RANK / WORLD_SIZE in the environment) the launcher joins that job instead of spawning.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time


def _args(argv=None):
    p = argparse.ArgumentParser(prog="python -m scythe_jl_b200.run", description=__doc__.split("\n\n")[0])
    p.add_argument("--workers", "-w", type=int, default=None, help="Number of worker processes (= radial tiles)")
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this node to spread the tiles over")
    p.add_argument("--sge", action="store_true", help="(reference option) not available: single-node launcher")
    p.add_argument("--email", default="none", help="(reference option) ignored unless --sge")
    p.add_argument("--format", choices=("csv", "netcdf", "both"), default=None, help="output format (default: model options or csv)")
    p.add_argument("--checkpoint", default=None, help="write exact restart files (AB3 history) to this .npz path")
    p.add_argument("--checkpoint-interval", type=float, default=0.0, help="model seconds between checkpoints (default: at the end)")
    p.add_argument("--restart", default=None, help="continue from a checkpoint written by --checkpoint")
    p.add_argument("--master-port", type=int, default=29531)
    p.add_argument("model", help="Name of model parameters file")
    return p.parse_args(argv)


def main(argv=None) -> int:
    a = _args(argv)
    if a.sge:
        print("--sge: Sun Grid Engine worker distribution is not available in the B200 launcher "
              "(one node, one process per GPU); use --gpus", file=sys.stderr)
        return 2
    under_torchrun = "RANK" in os.environ and "WORLD_SIZE" in os.environ
    world = int(os.environ.get("WORLD_SIZE", "1")) if under_torchrun else 1
    if a.gpus > 1 and not under_torchrun:
        workers = a.workers or a.gpus
        if workers != a.gpus:
            print(f"with --gpus {a.gpus} each GPU owns one tile: --workers must be {a.gpus} (got {workers})", file=sys.stderr)
            return 2
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(a.master_port), "-m", "scythe_jl_b200.run"]
        cmd += list(sys.argv[1:] if argv is None else argv)
        print(f"Initializing with {workers} workers on {a.gpus} GPUs")
        return subprocess.call(cmd)

    from . import api
    from .modelfile import load_model_file
    model = load_model_file(a.model)
    if a.format:
        model.options["output_format"] = a.format
    distributed = world > 1
    device = 0
    if distributed:
        import torch
        import torch.distributed as dist
        device = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    workers = a.workers or world
    if distributed and workers != world:
        print(f"one tile per rank: --workers must equal the world size {world}", file=sys.stderr)
        return 2
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        print(f"Initializing with {workers} workers" + (f" on {world} GPUs" if distributed else " on one GPU"))
        print(f"Model: {model.equation_set}, geometry {model.grid_params.geometry}, ts {model.ts}, "
              f"integration_time {model.integration_time}, output every {model.output_interval} -> {model.output_dir}")
    t0 = time.perf_counter()
    api.integrate_model(model, num_tiles=workers, write=True, device=device, distributed=distributed,
                        checkpoint=a.checkpoint, checkpoint_interval=a.checkpoint_interval, restart=a.restart)
    if rank == 0:
        print(f"Model complete!  ({time.perf_counter() - t0:.3f} s wall clock)")
    if distributed:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
