"""Multi-GPU plumbing: tile partitioning + the final gather of tiles.

The passes shard naturally — every output texel depends only on read-only inputs — so each rank runs
the SAME kernels on a contiguous range of rows and the only communication is one all-gather of the
finished tiles at the end (NCCL over NVLink on the GPU box; gloo in the CPU tests). One process per
GPU, torch.distributed for the plumbing. Nothing here computes pixels.

Row spaces:
  forward / post passes : image rows                      -> equal contiguous row blocks
  diffuse irradiance    : face*res + row                  -> equal contiguous row blocks
  specular prefilter    : flattened (mip, face, row)      -> contiguous blocks of equal COST
                          (a row of mip m has (res>>m) texels; mip 0 is roughness 0 = ~1 % of a full texel,
                          every other mip costs `samples` samples/texel)
Because both cubemap layouts are packed mip-major/face-minor/row-major, a contiguous row range is a
contiguous texel range of the packed buffer, so the gather needs no strided copies.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def equal_tiles(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """`world` contiguous [begin,end) row ranges whose sizes differ by at most one (some may be empty)."""
    return [((n_rows * r) // world, (n_rows * (r + 1)) // world) for r in range(world)]


def weighted_tiles(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """contiguous ranges over rows with per-row `costs`, cut where the cumulative cost crosses k/world."""
    total = float(sum(costs))
    cuts, acc, r = [0], 0.0, 1
    for i, c in enumerate(costs):
        acc += c
        while r < world and acc >= total * r / world - 1e-9:
            cuts.append(i + 1)
            r += 1
    while len(cuts) < world:
        cuts.append(len(costs))
    cuts.append(len(costs))
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def specular_row_costs(res: int, mips: int, samples: int = 512) -> List[float]:
    costs = []
    for m in range(mips):
        n = res >> m
        # mip 0 is roughness 0: one HDRI sample + a 4 x samples FADD replay, measured at ~1.2 % of a full texel on B200
        per_texel = 0.012 * float(samples) if m == 0 else float(samples)
        costs += [n * per_texel] * (6 * n)
    return costs


def specular_row_to_texel(res: int, mips: int, row: int) -> int:
    """texel offset (in the packed cubemap) of the first texel of flattened row `row` (row == total -> end)."""
    off = 0
    for m in range(mips):
        n = res >> m
        if row <= 6 * n:
            return off + row * n
        off += 6 * n * n
        row -= 6 * n
    return off


def specular_tiles(res: int, mips: int, world: int, samples: int = 512):
    rows = weighted_tiles(specular_row_costs(res, mips, samples), world)
    texels = [(specular_row_to_texel(res, mips, a), specular_row_to_texel(res, mips, b)) for a, b in rows]
    return rows, texels


def allgather_ranges(buf, ranges: Sequence[Tuple[int, int]], group=None):
    """In-place all-gather of UNEQUAL contiguous ranges of dim 0 of `buf` (same shape on every rank):
    on entry rank r has filled buf[ranges[r]]; on exit every rank has every range.
    Equal ranges covering the buffer use one all_gather_into_tensor straight into `buf`; unequal ranges use one
    broadcast per owner rank (no padding to the largest range)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(ranges) == world
    sizes = [b - a for a, b in ranges]
    if world == 1:
        return buf
    contiguous_cover = all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    if len(set(sizes)) == 1 and contiguous_cover and ranges[0][0] == 0 and ranges[-1][1] == buf.shape[0]:
        dist.all_gather_into_tensor(buf, buf[ranges[rank][0]:ranges[rank][1]].clone(), group=group)
        return buf
    # unequal ranges: one broadcast per owner, all issued asynchronously (NCCL runs them back to back on its stream).
    # Every rank receives exactly the bytes it is missing - no padding to the largest range (mip 0 of the specular
    # cubemap is 75 % of the texels but 1 % of the cost, so ranges differ by 50x in size).
    works = []
    for r, (a, b) in enumerate(ranges):
        if b > a:
            works.append(dist.broadcast(buf[a:b], src=dist.get_global_rank(group, r) if group is not None else r,
                                        group=group, async_op=True))
    for w in works:
        w.wait()
    return buf


class InterleavedSpecularPlan:
    """Equal-cost AND equal-size partition of the specular prefilter for one all-gather.

    Cost-balanced contiguous ranges differ 50x in size (mip 0 is 75 % of the texels and 1 % of the cost), which forces
    either padding or one broadcast per owner (8 collectives ~ 0.3 ms of launch latency at 8 GPUs, measured). Instead
    every mip whose 6*n rows divide by `world` is split into `world` equal row blocks (rank r takes block r of EVERY such
    mip: equal cost and equal bytes by construction); the few tiny mips that do not divide are simply computed by every
    rank (a few dozen texels). The gather is then: pack my blocks -> ONE all_gather_into_tensor -> one strided copy per
    split mip back into the packed cubemap.
    """

    def __init__(self, res: int, mips: int, world: int):
        self.res, self.mips, self.world = res, mips, world
        self.split, self.replicated = [], []      # (mip, row0 (flattened), rows_per_rank, n) / (mip, row0, rows, n)
        row0 = 0
        for m in range(mips):
            n = res >> m
            rows = 6 * n
            if rows % world == 0 and world > 1:
                self.split.append((m, row0, rows // world, n))
            else:
                self.replicated.append((m, row0, rows, n))
            row0 += rows
        self.total_rows = row0
        self.chunk_texels = sum(k * n for (_, _, k, n) in self.split)

    def row_ranges(self, rank: int):
        """flattened (mip, face, row) ranges this rank computes"""
        out = [(r0 + rank * k, r0 + (rank + 1) * k) for (_, r0, k, _) in self.split]
        out += [(r0, r0 + rows) for (_, r0, rows, _) in self.replicated]
        return sorted((a, b) for a, b in out if b > a)     # increasing and disjoint: what vq_specular_prefilter_ranges takes

    def texel_range(self, row_a: int, row_b: int):
        return specular_row_to_texel(self.res, self.mips, row_a), specular_row_to_texel(self.res, self.mips, row_b)

    def gather(self, cube_t, rank: int, group=None):
        """cube_t: [texels, 4] packed cubemap; on entry this rank's row_ranges are filled; on exit all are."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or not self.split:
            return cube_t
        parts = []
        for (_, r0, k, _) in self.split:
            a, b = self.texel_range(r0 + rank * k, r0 + (rank + 1) * k)
            parts.append(cube_t[a:b])
        send = torch.cat(parts, dim=0)
        recv = torch.empty((self.world * self.chunk_texels,) + tuple(cube_t.shape[1:]), dtype=cube_t.dtype, device=cube_t.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        recv = recv.view(self.world, self.chunk_texels, *cube_t.shape[1:])
        off = 0
        for (_, r0, k, n) in self.split:
            a, b = self.texel_range(r0, r0 + self.world * k)          # the whole mip: rank blocks are contiguous, in rank order
            cube_t[a:b].view(self.world, k * n, *cube_t.shape[1:]).copy_(recv[:, off:off + k * n])
            off += k * n
        return cube_t
