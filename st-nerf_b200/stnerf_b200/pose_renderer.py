"""Fast path of `LayeredNeuralRenderer.render_pose` / `render_path`.

The reference builds the H*W rays of a pose on the CPU (`data/datasets/ray_dataset.py:260-283` ->
`utils/render_helpers.py:42-126`), uploads 75 MB per 1080p frame (`render/layered_neural_renderer.py:372-375`), renders
chunk by chunk and copies every image back synchronously (`:451-454`).  Here the rays are generated on the device by
`stnerf_raygen`, one native call renders the whole frame, and the device->host copies of frame i overlap the rendering of
frame i+1 (double-buffered pinned staging).  With a process group initialised, each rank renders its interleaved rows
and one all-gather assembles the images (stnerf_b200.dist).

Returned values follow `render_pose` (`render/layered_neural_renderer.py:364-392`): `color (H,W,3)`, `depth (H,W,1)`
(negative depths zeroed, divided by `far`), `color_layer` and `depth_layer` lists over the l layers -- including the
reference's quirk that the per-layer depths are zeroed where the *mixed* depth is negative (`:386-388`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .dist import ShardedViewRenderer


class PoseRenderer:
    def __init__(self, model, height: int, width: int, far: float = 20.0, rank: int = 0, world: int = 1):
        self.model, self.H, self.W, self.far = model, int(height), int(width), float(far)
        self.rank, self.world = rank, world
        self.l = model.layer_num + 1
        self._svr: Optional[ShardedViewRenderer] = None
        self._pinned = [None, None]
        self._events = [None, None]
        self._slot = 0

    def _renderer(self) -> ShardedViewRenderer:
        dev = torch.device("cuda", torch.cuda.current_device())
        nat = self.model._ensure_native(dev)
        if self._svr is None:
            self._svr = ShardedViewRenderer(nat, self.H, self.W, self.model.coarse_ray_sample,
                                            self.model.fine_ray_sample, self.rank, self.world)
        return self._svr

    def frame_ids(self, layer_frame_pair: Sequence[Tuple[int, float]]) -> List[float]:
        """`frame_ids[:, layer_id] = frame_id` (data/datasets/ray_dataset.py:276-279)."""
        ids = [0.0] * self.l
        for layer_id, frame_id in layer_frame_pair:
            ids[int(layer_id)] = float(frame_id)
        return ids

    @torch.no_grad()
    def render_images(self, pose, K, layer_frame_pair, density_threshold=0, bkgd_density_threshold=0) -> torch.Tensor:
        """(l+1, H, W, 5) device tensor: image 0 = mixed, 1+i = layer i; channels rgb(3), raw depth, acc."""
        svr = self._renderer()
        ids = self.frame_ids(layer_frame_pair)
        # the prologue of forward(): boxes of ray 0's frame ids, edits, thresholds (layered_rfrender.py:190-242).
        # render_pose goes through layered_batchify_ray with N >= chunk size, so the thresholds are forwarded
        # (utils/batchify_rays.py:53-79).
        self.model.retiming = True
        svr.nat.set_scene(self.model._resolve_scene(torch.tensor(ids), density_threshold, bkgd_density_threshold))
        rays = svr.rays_for(torch.as_tensor(K, dtype=torch.float32), torch.as_tensor(pose, dtype=torch.float32), ids)
        self.model.seed += 1
        return svr.render(rays, seed=self.model.seed)

    def render_pose(self, pose, K, layer_frame_pair, density_threshold=0, bkgd_density_threshold=0):
        """Same return value as LayeredNeuralRenderer.render_pose (render/layered_neural_renderer.py:364-392)."""
        img = self.render_images(pose, K, layer_frame_pair, density_threshold, bkgd_density_threshold)
        return self._post(img)

    def _post(self, img: torch.Tensor):
        color = img[0, ..., :3]
        depth = img[0, ..., 3:4].clone()
        neg = depth < 0
        depth[neg] = 0                                               # :382
        depth = depth / self.far                                     # :383
        color_layer = [img[1 + i, ..., :3] for i in range(self.l)]   # :384
        depth_layer = []
        for i in range(self.l):
            d1 = img[1 + i, ..., 3:4].clone()
            d1[depth < 0] = 0        # quirk kept: tests the (already clamped, scaled) mixed depth -> never true (:387)
            depth_layer.append(d1 / self.far)
        return color, depth, color_layer, depth_layer

    def render_path(self, poses, Ks, layer_frame_pairs, density_threshold=0, bkgd_density_threshold=0,
                    per_frame_state=None):
        """Generator over the frames of a camera path (`render_path`, :401-488, without the file IO).  Yields CPU tensors
        `(color, depth, color_layer, depth_layer)`; the D2H copy of frame i runs while frame i+1 renders.
        `per_frame_state(idx, model)` may set model.shift / scale / alpha per frame (:435-440)."""
        copy_stream = torch.cuda.Stream()
        pending = None
        for idx in range(len(poses)):
            if per_frame_state is not None:
                per_frame_state(idx, self.model)
            img = self.render_images(poses[idx], Ks[idx], layer_frame_pairs[idx], density_threshold, bkgd_density_threshold)
            slot = self._slot
            self._slot ^= 1
            if self._pinned[slot] is None or self._pinned[slot].shape != img.shape:
                self._pinned[slot] = torch.empty(img.shape, dtype=img.dtype).pin_memory()
            done = torch.cuda.Event()
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                self._pinned[slot].copy_(img, non_blocking=True)      # img is a fresh tensor per frame; kept alive below
                done.record(copy_stream)
            if pending is not None:
                yield self._finish(*pending)
            pending = (slot, done, img)
        if pending is not None:
            yield self._finish(*pending)

    def _finish(self, slot, done, _keepalive):
        done.synchronize()
        host = self._pinned[slot]
        color = host[0, ..., :3].clone()
        depth = host[0, ..., 3:4].clone()
        depth[depth < 0] = 0
        depth = depth / self.far
        color_layer = [host[1 + i, ..., :3].clone() for i in range(self.l)]
        depth_layer = [host[1 + i, ..., 3:4].clone() / self.far for i in range(self.l)]
        return color, depth, color_layer, depth_layer
