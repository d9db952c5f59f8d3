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
