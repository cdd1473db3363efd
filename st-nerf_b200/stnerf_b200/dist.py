"""Multi-GPU plumbing: one process per GPU, rays sharded by interleaved image rows, one all-gather of the
finished image planes.  Rays are independent, so there is no data-path collective inside the march;
the only exchange is the assembly of the image on every rank.

torch.distributed (NCCL over NVLink/NVSwitch; gloo in the CPU tests) is the transport; the kernels write each
rank's planes into a registered buffer that is the in-place input slice of the all-gather.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_rows(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows rendered by `rank`: row0, row_step, n_rows (interleaved for load balance against performer-dense rows)."""
    n = (height - rank + world - 1) // world if rank < height else 0
    return rank, world, n


def padded_rows(height: int, world: int) -> int:
    return (height + world - 1) // world


def assemble_image(gathered: torch.Tensor, height: int, width: int, world: int) -> torch.Tensor:
    """gathered (world, P, rows_pad*width*C) -> (P, height, width, C) undoing the row interleave.
    The trailing dim of each plane is C channels per pixel (C inferred)."""
    Wd, P = gathered.shape[0], gathered.shape[1]
    rp = padded_rows(height, world)
    c = gathered.shape[2] // (rp * width)
    g = gathered.view(Wd, P, rp, width, c).permute(1, 2, 0, 3, 4).reshape(P, rp * Wd, width, c)
    return g[:, :height]


def all_gather_planes(local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """local (P, n) contiguous -> (world, P, n) on every rank with a single collective."""
    if world == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)          # concatenation along dim 0 (NCCL and gloo agree)
    return out.view((world,) + tuple(local.shape))


class ShardedViewRenderer:
    """Renders HxW views with rows interleaved over the ranks of the default process group and assembles
    the fine images (mixed + per layer: rgb, depth, acc) on every rank."""

    def __init__(self, native, height: int, width: int, n1: int, n2: int, rank: int = 0, world: int = 1):
        self.nat, self.H, self.W, self.n1, self.n2 = native, height, width, n1, n2
        self.rank, self.world = rank, world
        self.row0, self.step, self.n_rows = shard_rows(height, rank, world)
        self.rp = padded_rows(height, world)
        self.l = native.l
        n_local = self.rp * width
        dev = torch.device("cuda", torch.cuda.current_device())
        self.out = torch.zeros((2, self.l + 1, 5 * n_local), dtype=torch.float32, device=dev)
        self.mask = torch.zeros((self.l, n_local), dtype=torch.uint8, device=dev)

    def rays_for(self, K, T, frame_ids):
        from . import ops
        rays = ops.generate_rays(K, T, self.H, self.W, frame_ids=frame_ids, row0=self.row0, row_step=self.step,
                                 n_rows=self.n_rows)
        if self.n_rows < self.rp:        # pad to the common shard size with copies of the last ray (discarded later)
            pad = rays[-1:].expand((self.rp - self.n_rows) * self.W, -1)
            rays = torch.cat([rays, pad], 0)
        return rays.contiguous()

    def render_local(self, rays: torch.Tensor, seed: int = 0):
        """This rank's rows of the fine images: (l+1, rows_pad*W*5), pixel-interleaved rgb(3), depth, acc."""
        # every pixel keeps the Philox stream it has in an unsharded render of the same seed
        self.nat.set_ray_ids(self.row0 * self.W, self.W, self.step * self.W)
        out, _ = self.nat.render(rays, self.n1, self.n2, seed=seed, out=self.out, ray_mask=self.mask)
        self.nat.set_ray_ids(0, 0, 0)
        n = self.rp * self.W
        fine = out[1]                                                     # (l+1, 5n): rgb(3n) | depth(n) | acc(n)
        local = torch.cat([fine[:, :3 * n].reshape(self.l + 1, n, 3), fine[:, 3 * n:4 * n].unsqueeze(-1),
                           fine[:, 4 * n:].unsqueeze(-1)], -1).reshape(self.l + 1, n * 5)
        return local

    def render(self, rays: torch.Tensor, seed: int = 0):
        """Returns the assembled fine images (l+1, H, W, 5) = rgb(3), depth, acc on every rank."""
        g = all_gather_planes(self.render_local(rays, seed), self.world)
        return assemble_image(g, self.H, self.W, self.world)
