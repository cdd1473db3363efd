"""Multi-GPU plumbing: one process per GPU, rays sharded by interleaved image rows, one in-place all-gather of the
finished image tiles.  Rays are independent, so there is no data-path collective inside the march; the
only exchange is the assembly of the image on every rank.

Layout.  Every rank owns ONE persistent buffer `gather` of shape (world, B, l+1, rows_pad*W, 5): slot [r] holds rank r's
rows (r, r+world, ...) of the fine images of a batch of B views, pixel-interleaved (rgb, depth, acc).  The fine compositing
kernel writes a rank's pixels straight into `gather[rank]` (`stnerf_render_views`, no staging, no repack) and
`all_gather_into_tensor(gather, gather[rank])` is the in-place form of the collective (send buffer = own slot of the receive
buffer: `ncclAllGather` in place over NVLink/NVSwitch; gloo in the CPU tests).  The assembled image is a strided VIEW of the
buffer -- `images()[b, p, k, r]` is row k*world + r -- so nothing is copied after the collective either; `assembled()` makes
the contiguous (l+1, H, W, 5) copy for consumers that want one.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows rendered by `rank`: row0, row_step, n_rows (interleaved for load balance against performer-dense rows)."""
    n = (height - rank + world - 1) // world if rank < height else 0
    return rank, world, n


def padded_rows(height: int, world: int) -> int:
    return (height + world - 1) // world


def all_gather_inplace(buf: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """buf (world, ...) contiguous, slot [rank] already filled by this rank -> every slot filled, with ONE collective whose
    send buffer is the rank's own slot of the receive buffer (no staging copy on either side)."""
    assert buf.is_contiguous() and buf.shape[0] == world
    if world > 1:
        dist.all_gather_into_tensor(buf.view(-1), buf[rank].view(-1), group=group)
    return buf


def rows_view(gathered: torch.Tensor, width: int) -> torch.Tensor:
    """(world, ..., rows_pad*W, C) -> (..., rows_pad, world, W, C): element [..., k, r] is image row k*world + r.  A view."""
    world = gathered.shape[0]
    c = gathered.shape[-1]
    rp = gathered.shape[-2] // width
    g = gathered.reshape(gathered.shape[:-2] + (rp, width, c))           # (world, ..., rp, W, C)
    nd = g.dim()
    return g.permute(*range(1, nd - 2), 0, nd - 2, nd - 1)               # (..., rp, world, W, C)


def assemble_image(gathered: torch.Tensor, height: int, width: int, world: int) -> torch.Tensor:
    """gathered (world, P, rows_pad*width*C) or (world, P, rows_pad*width, C) -> contiguous (P, height, width, C)."""
    if gathered.dim() == 3:
        rp = padded_rows(height, world)
        gathered = gathered.reshape(gathered.shape[0], gathered.shape[1], rp * width, gathered.shape[2] // (rp * width))
    v = rows_view(gathered, width)                                        # (P, rp, world, W, C)
    P, rp, Wd, W, C = v.shape
    return v.reshape(P, rp * Wd, W, C)[:, :height]


class ShardedViewRenderer:
    """Renders HxW views with rows interleaved over the ranks of the default process group and assembles the fine images
    (mixed + per layer: rgb, depth, acc) on every rank.  `views` are `stnerf_view` structs (NativeRenderer.make_view)."""

    def __init__(self, native, height: int, width: int, n1: int, n2: int, rank: int = 0, world: int = 1, group=None):
        self.nat, self.H, self.W, self.n1, self.n2 = native, int(height), int(width), int(n1), int(n2)
        self.rank, self.world, self.group = rank, world, group
        self.rp = padded_rows(height, world)          # every rank renders rows_pad rows (rows past H-1 are discarded)
        self.l = native.l
        self._gather = {}                             # batch size -> persistent (world, B, l+1, rp*W, 5) buffer
        self._coarse = {}                             # batch size -> scratch for the coarse images (with_coarse=True)
        self._timing = None                           # CUDA events around the last timed all-gather

    def gather_buffer(self, batch: int) -> torch.Tensor:
        buf = self._gather.get(batch)
        if buf is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            buf = torch.zeros((self.world, batch, self.l + 1, self.rp * self.W, 5), dtype=torch.float32, device=dev)
            self._gather[batch] = buf
        return buf

    def render_local(self, views: Sequence, with_coarse: bool = False) -> torch.Tensor:
        """This rank's rows of the fine images of every view, written in place into its slot of the gather buffer.
        `with_coarse`: also produce the coarse images (into a scratch buffer), i.e. all of forward()'s outputs."""
        buf = self.gather_buffer(len(views))
        coarse = None
        if with_coarse:
            coarse = self._coarse.get(len(views))
            if coarse is None:
                coarse = self._coarse[len(views)] = torch.empty_like(buf[self.rank])
        self.nat.render_views(views, self.H, self.W, self.n1, self.n2, row0=self.rank, row_step=self.world, n_rows=self.rp,
                              out=buf[self.rank], coarse_out=coarse)
        return buf

    def render(self, views: Sequence, time_collective: bool = False, with_coarse: bool = False) -> torch.Tensor:
        """(B, l+1, rows_pad, world, W, 5) view of the assembled images on every rank: [b, p, k, r] = row k*world + r."""
        buf = self.render_local(views, with_coarse)
        if time_collective and self.world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            all_gather_inplace(buf, self.rank, self.world, self.group)
            e1.record()
            self._timing = (e0, e1)
        else:
            all_gather_inplace(buf, self.rank, self.world, self.group)
        return rows_view(buf, self.W)

    def last_collective_ms(self) -> Optional[float]:
        t = self._timing
        if t is None:
            return None
        t[1].synchronize()
        return float(t[0].elapsed_time(t[1]))

    def assembled(self, rows: torch.Tensor) -> torch.Tensor:
        """Contiguous (B, l+1, H, W, 5) copy of what `render` returned."""
        B, P, rp, Wd, W, C = rows.shape
        return rows.reshape(B, P, rp * Wd, W, C)[:, :, :self.H]
