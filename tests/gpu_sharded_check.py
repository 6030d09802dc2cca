"""Run under torchrun on >= 2 GPUs: one diagonalisation sharded with an NCCL all-reduce on sigma.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/gpu_sharded_check.py [workload]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
group = fermion.ShardGroup()
for wl in (sys.argv[1:] or ["t", "c5"]):
    norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
    sa, sb = batches[0]
    ref = fermion.solve_sci((sa, sb), h, g, norb, nelec)             # unsharded, this GPU only
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = fermion.solve_sci((sa, sb), h, g, norb, nelec)
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    res = fermion.solve_sci_sharded((sa, sb), h, g, norb, nelec, group=group)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = fermion.solve_sci_sharded((sa, sb), h, g, norb, nelec, group=group)
    torch.cuda.synchronize()
    t_shard = time.perf_counter() - t0
    de = abs(res.energy - ref.energy)
    same = np.abs(np.abs(res.sci_state.amplitudes) - np.abs(ref.sci_state.amplitudes)).max()
    # every rank must hold the same answer
    e_all = [None] * world
    dist.all_gather_object(e_all, res.energy)
    if rank == 0:
        print(f"[{wl}] {len(sa)}x{len(sb)} dets: single-GPU {t_single * 1e3:.1f} ms, sharded x{world} "
              f"{t_shard * 1e3:.1f} ms, |dE| = {de:.2e}, max |d amp| = {same:.2e}, energies equal on all ranks: "
              f"{len(set(e_all)) == 1}", flush=True)
    assert de < 1e-9 and same < 1e-6 and len(set(e_all)) == 1
group.close()
dist.destroy_process_group()
