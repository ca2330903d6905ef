"""Lab driver of pcg_timeline.sh: the PCG leg of bench.py (100^3 small-strain elements per rank) on the timeline build."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from tahoe_b200 import capi, mesh as tmesh  # noqa: E402

capi.LIB_PATH = os.environ["TB2_LAB_LIB"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 100
    gx, gy, gz = tmesh.brick_grid(world)
    part = tmesh.partition_cube(n * gx, n * gy, n * gz, world, rank, jitter=0.1)
    X, conn, ns = part["coords"], part["conn"], part["nodesets"]
    m = capi.Mesh(X, conn, device=local)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    m.comm_init(rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"],
                all_gather=gather if os.environ.get("TB2_LAB_EXCHANGE", "peer") == "peer" else None)
    g = capi.Group(m, capi.SMALL_STRAIN, capi.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    eqs = capi.Equations(m, code)
    A = capi.Matrix(eqs)
    dev = torch.device("cuda", local)
    A.form_stiffness(g, torch.zeros(X.shape, dtype=torch.float64, device=dev))
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 1e-3
    b = torch.from_numpy(fext[eqs.eqnos() > 0]).to(dev)
    x = torch.zeros_like(b)
    for _ in range(2):
        x.zero_()
        m.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        it, rn = A.pcg(b, x, rtol=0.0, atol=0.0, max_iter=96)
        m.synchronize()
        if rank == 0:
            print("rank 0: %d iterations, %.1f us per iteration (host clock)" % (it, (time.perf_counter() - t0) / it * 1e6), file=sys.stderr)
    dist.barrier()
    A.close(); eqs.close(); g.close(); m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
