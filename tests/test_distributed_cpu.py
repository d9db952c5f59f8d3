"""CPU (-m "not gpu"): the N>1 host logic — tile partitioning and the final all-gather — with
world_size 2 and 3 over gloo. The per-tile 'compute' is a deterministic stand-in; the assembled result must
equal the single-rank result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vqengine_b200 import distributed as vd
import vqengine_b200 as vq


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_equal_tiles_cover():
    for n in (0, 1, 7, 384, 2160, 4320):
        for w in (1, 2, 3, 8):
            t = vd.equal_tiles(n, w)
            assert t[0][0] == 0 and t[-1][1] == n and all(t[i][1] == t[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in t]
            assert max(sizes) - min(sizes) <= 1
    assert vd.equal_tiles(4320, 8) == [(540 * r, 540 * (r + 1)) for r in range(8)]


def test_specular_tiles_balance_and_layout():
    res, mips = 512, 9
    rows, texels = vd.specular_tiles(res, mips, 8)
    total_rows = vq.cubemap_row_count(res, mips)
    assert rows[0][0] == 0 and rows[-1][1] == total_rows
    assert texels[0][0] == 0 and texels[-1][1] == vq.cubemap_texel_count(res, mips)
    costs = vd.specular_row_costs(res, mips)
    per = [sum(costs[a:b]) for a, b in rows]
    assert max(per) / (sum(per) / 8) < 1.15          # within 15 % of perfect balance
    # row -> texel offsets agree with the C-ABI layout helpers
    r0 = 0
    for m in range(mips):
        n = res >> m
        for f in range(6):
            assert vd.specular_row_to_texel(res, mips, r0 + f * n) == vq.cubemap_offset(res, m, f)
        r0 += 6 * n
    for w in (1, 2, 3, 5):
        rws, _ = vd.specular_tiles(64, 6, w)
        assert all(rws[i][1] == rws[i + 1][0] for i in range(w - 1)) and rws[-1][1] == vq.cubemap_row_count(64, 6)


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "rows":
            H, W = 37, 5
            tiles = vd.equal_tiles(H, world)
            buf = torch.zeros((H, W, 4))
            a, b = tiles[rank]
            buf[a:b] = torch.arange(a * W * 4, b * W * 4, dtype=torch.float32).reshape(b - a, W, 4) * 0.5 + 1.0
            vd.allgather_ranges(buf, tiles)
            want = torch.arange(0, H * W * 4, dtype=torch.float32).reshape(H, W, 4) * 0.5 + 1.0
        elif mode == "equal":
            H, W = 8 * world, 3
            tiles = vd.equal_tiles(H, world)
            buf = torch.zeros((H, W, 4))
            a, b = tiles[rank]
            buf[a:b] = float(rank + 1)
            vd.allgather_ranges(buf, tiles)
            want = torch.repeat_interleave(torch.arange(1, world + 1, dtype=torch.float32), 8)[:, None, None].expand(H, W, 4)
        elif mode == "interleaved":   # equal-cost + equal-size plan: pack -> one all_gather -> per-mip strided copy
            res, mips = 16, 4
            plan = vd.InterleavedSpecularPlan(res, mips, world)
            n = vq.cubemap_texel_count(res, mips)
            buf = torch.zeros((n, 4))
            want = torch.arange(0, n, dtype=torch.float32)[:, None] * 2.0 + torch.tensor([0.0, 0.25, 0.5, 1.0])
            covered = torch.zeros(n, dtype=torch.bool)
            for (ra, rb) in plan.row_ranges(rank):
                a, b = plan.texel_range(ra, rb)
                buf[a:b] = want[a:b]
            plan.gather(buf, rank)
        else:   # packed specular cubemap: unequal texel ranges
            res, mips = 16, 4
            rows, texels = vd.specular_tiles(res, mips, world, samples=64)
            n = vq.cubemap_texel_count(res, mips)
            buf = torch.zeros((n, 4))
            a, b = texels[rank]
            buf[a:b] = torch.arange(a, b, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.25, 0.5, 1.0])
            vd.allgather_ranges(buf, texels)
            want = torch.arange(0, n, dtype=torch.float32)[:, None] + torch.tensor([0.0, 0.25, 0.5, 1.0])
        q.put((rank, bool(torch.equal(buf, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["rows", "equal", "spec", "interleaved"])
def test_tiles_gather_gloo(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_interleaved_plan_partitions_every_row_once_or_replicates():
    for world in (1, 2, 3, 8):
        plan = vd.InterleavedSpecularPlan(512, 9, world)
        assert plan.total_rows == vq.cubemap_row_count(512, 9)
        cover = np.zeros(plan.total_rows, np.int32)
        sizes = []
        for r in range(world):
            rr = plan.row_ranges(r)
            sizes.append(sum(plan.texel_range(a, b)[1] - plan.texel_range(a, b)[0] for a, b in rr))
            for a, b in rr:
                cover[a:b] += 1
        rep = np.zeros(plan.total_rows, bool)
        for (_, r0, rows, _) in plan.replicated:
            rep[r0:r0 + rows] = True
        assert (cover[~rep] == 1).all() and (cover[rep] == world).all()
        assert len(set(sizes)) == 1                       # equal bytes per rank
    p8 = vd.InterleavedSpecularPlan(512, 9, 8)
    assert [m for (m, _, _, _) in p8.replicated] == [8]     # only the 2x2 mip does not divide by 8
