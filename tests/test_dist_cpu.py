"""World-size-2 `gloo` test of the row-block sharded loop (deepsir_b200/dist.py) on CPU: the exchange logic
(moment all-reduce, identical transforms on every rank, optional correspondence all-gather) with the ORACLE standing
in for the CUDA kernels.  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepsir_b200 import dist as D
from deepsir_b200 import synth
from oracle import deepsir_oracle as O


class OracleOps:
    match_argmin = staticmethod(lambda fs, fr: O.match_argmin(fs, fr))

    @staticmethod
    def moments(xyz_src, xyz_ref, idx, w):
        tgt = O.gather_neighbour_V3(xyz_ref, idx)
        return O.kabsch_moments_fp64(xyz_src.permute(0, 2, 1), tgt.permute(0, 2, 1), w)

    solve = staticmethod(O.kabsch_from_moments_fp64)
    transform = staticmethod(O.se3_transform_V2)
    compose = staticmethod(O.se3_concatenate)


def _make(J=900, B=2):
    b = synth.make_batch(B, J, 32, "kitti", config=4, first_pair=7)
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    return b["feat_src"], b["feat_ref"], xs, xr, b["weights"][:, :, 0].contiguous()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fs, fr, xs, xr, w = _make()
        J = fs.shape[2]
        lo, hi = D.row_block(J, world, rank)
        tr, pred, xyz, _ = D.align_rowblock(fs[:, :, lo:hi].contiguous(), fr, xs[:, :, lo:hi].contiguous(), xr,
                                            w[:, lo:hi].contiguous(), 3, ops=OracleOps, gather_pred_rows=J)
        q.put((rank, [t.numpy().copy() for t in tr], [p.numpy().copy() for p in pred], xyz.numpy().copy(), (lo, hi)))   # plain arrays: no fd passing
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_block_partition():
    for J, world in [(10, 3), (131072, 8), (5, 8), (16384, 2)]:
        blocks = [D.row_block(J, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == J
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1


def test_moments_are_additive_and_solve_matches_reference_form():
    fs, fr, xs, xr, w = _make()
    idx = O.match_argmin(fs, fr)
    tgt = O.gather_neighbour_V3(xr, idx)
    full = O.kabsch_moments_fp64(xs.permute(0, 2, 1), tgt.permute(0, 2, 1), w)
    parts = sum(O.kabsch_moments_fp64(xs[:, :, a:b].permute(0, 2, 1), tgt[:, :, a:b].permute(0, 2, 1), w[:, a:b])
                for a, b in [(0, 100), (100, 433), (433, xs.shape[2])])
    assert torch.allclose(full, parts, rtol=1e-12, atol=1e-12)
    T_m, _ = O.kabsch_from_moments_fp64(full)
    T_r, _ = O.compute_rigid_transform_2(xs.permute(0, 2, 1).contiguous(), tgt.permute(0, 2, 1).contiguous(), w[:, :, None])
    assert O.rotation_angle_deg(T_m[:, :, :3], T_r[:, :, :3]).max() < 1e-3
    assert (T_m[:, :, 3] - T_r[:, :, 3]).abs().max() < 1e-4


@pytest.mark.timeout(180)
def test_two_rank_gloo_loop_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda r: r[0])
    res = [(r[0], [torch.from_numpy(t) for t in r[1]], [torch.from_numpy(t) for t in r[2]], torch.from_numpy(r[3]), r[4]) for r in res]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    fs, fr, xs, xr, w = _make()
    tr_ref, pred_ref, xyz_ref = O.align_loop(fs, fr, xs, xr, w[:, :, None], 3)
    for it in range(3):
        assert torch.equal(res[0][1][it], res[1][1][it])                       # every rank holds the same transform
        assert O.rotation_angle_deg(res[0][1][it][:, :, :3], tr_ref[it][:, :, :3]).max() < 1e-3
        assert (res[0][1][it][:, :, 3] - tr_ref[it][:, :, 3]).abs().max() < 1e-4
        assert torch.equal(res[0][2][it], pred_ref[it]) and torch.equal(res[1][2][it], pred_ref[it])   # gathered [B,J]
    for rank, _, _, xyz, (lo, hi) in res:
        assert (xyz - xyz_ref[:, :, lo:hi]).abs().max() < 1e-3
