"""Worker of tests/test_plan_gloo.py: one process per rank, torch.distributed `gloo` backend.
Creates metadata-only plans through dtfft_b200.comm.TorchComm (the real allgather plumbing the
GPU ranks use with NCCL) and checks decomposition + exchange geometry against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import Config, Layout, Pencil, PlanC2C, PlanR2R, Reshape
    from oracle import layout as L
    from oracle import pipeline as P

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchComm()
    assert comm.size == world and comm.rank == rank

    # (1) default decomposition + neighbor_data of every transposition
    for dims in ([48, 21, 36], [37, 22], [128, 64, 96]):
        plan = PlanC2C(dims, comm=comm, config=Config(enable_z_slab=False), dry=True)
        comm_dims, _, _ = L.choose_grid(dims, world, z_slab=False)
        assert plan.grid_dims == comm_dims, (plan.grid_dims, comm_dims)
        gold = L.make_pencils(dims, comm_dims, rank)
        lay = [Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS]
        for d in range(len(dims)):
            got = plan.get_pencil(lay[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts)
        for t in ([1, -1] if len(dims) == 2 else [1, -1, 2, -2]):
            _, geos = L.plan_geometry(dims, comm_dims, t)
            d = plan.describe_exchange(t)
            g = geos[rank]
            assert d["members"] == g.members
            if g.comm_size > 1:
                assert np.array_equal(d["send_nd"], g.send_nd) and np.array_equal(d["recv_nd"], g.recv_nd)
                assert d["send_counts"].tolist() == g.send_counts
        plan.destroy()

    # (2) bricks -> pencils through the gathered user boxes; fused boxes of every rank gathered
    #     with torch.distributed and replayed against the global-array truth
    cuts = [[10, 6], [9, 11]] if world == 2 else [[10, 6], [9, 11], [12, 8]][: 2 if world < 8 else 3]
    nx = len(cuts[0])
    assert world % nx == 0
    if len(cuts) == 2 and world // nx != len(cuts[1]):
        cuts[1] = [20 // (world // nx)] * (world // nx)
    edges = [np.concatenate([[0], np.cumsum(c)]) for c in cuts]
    boxes = []
    for j in range(len(cuts[1])):
        for i in range(len(cuts[0])):
            boxes.append(([int(edges[0][i]), int(edges[1][j])], [int(cuts[0][i]), int(cuts[1][j])]))
    assert len(boxes) == world
    plan = PlanR2R(Pencil(*boxes[rank]), comm=comm, dry=True)
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    dims, comm_dims, coords, xs, xc, bgrid, _ = L.from_bricks(starts, counts)
    assert plan.dims == dims and plan.grid_dims == comm_dims
    xp = plan.get_pencil(Layout.X_PENCILS)
    assert (xp.starts, xp.counts) == (xs[rank], xc[rank])
    mine = plan.describe_exchange(Reshape.X_BRICKS_TO_PENCILS)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine["members"], mine["fused_boxes"].tolist()))
    G = P.global_array(dims, np.float64, kind="index")
    src = P.redistribute(G, [L.Pencil(1, starts[r], counts[r]) for r in range(world)])
    want = P.redistribute(G, [L.Pencil(1, xs[r], xc[r]) for r in range(world)])
    dsts = [np.full(w.size, -7.0) for w in want]
    for r in range(world):
        P.apply_boxes(src[r], dsts, gathered[r][1], gathered[r][0])
    for r in range(world):
        assert np.array_equal(dsts[r], want[r]), r
    plan.destroy()

    # (3) the copy-engine form of an exchanging transposition (geometry.h: DmaBlock) across real processes: every rank's
    #     (peer, slice) packs and strided 3-D copies, gathered and replayed against the datatype truth, whole and sliced
    from tests.test_plan_host import _dma_copy

    for sub in (None, "2048"):
        if sub is None:
            os.environ.pop("DTFFTB_DMA_SUB_BYTES", None)
        else:
            os.environ["DTFFTB_DMA_SUB_BYTES"] = sub
        dims = [96, 40, 64]  # x = 96: Y -> Z blocks can be cut into 3 slices along the source's slowest axis
        plan = PlanC2C(dims, comm=TorchComm(cart_dims=[1, 1, world]), config=Config(enable_z_slab=False), dry=True)
        comm_dims = plan.grid_dims
        G = P.global_array(dims, np.complex128, kind="index")
        nsliced = 0
        for t in (2, -2):
            d = plan.describe_dma(t)
            mine = (d["members"], d["me"], [(e["member"], e["nsub"], e["pack"].tolist(), e["fused"].tolist(), e["copy"]) for e in d["entries"]])
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            src = P.scatter_input(G, dims, comm_dims, t)
            want = P.transpose_datatype(G, dims, comm_dims, t)
            out = [np.full(w.size, np.complex128(-7 - 7j)) for w in want]
            hits = [np.zeros(w.size, np.int32) for w in want]
            for r in range(world):
                members, me, entries = gathered[r]
                remote = sum(int(np.prod(f[:3])) for (m, _, _, f, _) in entries if m != me and f[0] > 0)
                staging = np.zeros(max(remote, 1), np.complex128)
                for (m, nsub, pack, f, cp) in entries:
                    nsliced += nsub > 1
                    if f[0] <= 0:
                        continue
                    if m == me:
                        P.apply_boxes(src[r], out, [f], [members[m]])
                        P.apply_boxes(np.ones(src[r].size, np.int32), hits, [f], [members[m]])
                        continue
                    assert cp["ok"] == 1
                    P.apply_boxes(src[r], [staging], [pack], [0])
                    _dma_copy(staging, out[members[m]], cp, int(pack[4]))
                    _dma_copy(np.ones(staging.size, np.int32), hits[members[m]], cp, int(pack[4]))
            for r in range(world):
                assert np.array_equal(out[r], want[r]) and np.all(hits[r] == 1), (t, r, sub)
        assert (nsliced > 0) == (sub is not None), (nsliced, sub)
        plan.destroy()
    os.environ.pop("DTFFTB_DMA_SUB_BYTES", None)
    Config()._commit()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: gloo plan checks OK", flush=True)


if __name__ == "__main__":
    main()
