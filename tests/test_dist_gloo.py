"""World-size-2 CPU (gloo) test of the multi-GPU plumbing: row-interleaved sharding + one all-gather + image assembly."""
import os
import socket
import subprocess
import sys
import textwrap

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, os.path.join(%r, "st-nerf_b200"))
    from stnerf_b200.dist import shard_rows, padded_rows, all_gather_inplace, assemble_image, rows_view
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, P, C = 13, 7, 3, 5                       # odd height: the last rank has one row less (padding path)
    row0, step, n_rows = shard_rows(H, rank, world)
    rp = padded_rows(H, world)
    rows = torch.arange(row0, row0 + step * n_rows, step)
    # fake renderer: pixel value encodes (plane, row, col, channel); each rank fills ITS slot of the gather buffer in place
    buf = torch.full((world, P, rp * W, C), -1.0)
    local = buf[rank].view(P, rp, W, C)
    for p in range(P):
        for i, r in enumerate(rows.tolist()):
            for c in range(W):
                local[p, i, c] = torch.arange(C) + 10 * c + 1000 * r + 100000 * p
    ptr = buf.data_ptr()
    g = all_gather_inplace(buf, rank, world)                        # one collective, send buffer = own slot of the receive buffer
    assert g.data_ptr() == ptr
    img = assemble_image(g, H, W, world)
    v = rows_view(g, W)                                              # the copy-free form: [p, k, r] = row k*world + r
    assert v.data_ptr() == ptr and tuple(v.shape) == (P, rp, world, W, C)
    want = torch.zeros(P, H, W, C)
    for p in range(P):
        for r in range(H):
            for c in range(W):
                want[p, r, c] = torch.arange(C) + 10 * c + 1000 * r + 100000 * p
    assert torch.equal(img, want), (rank, (img - want).abs().max())
    for y in range(H):
        assert torch.equal(v[:, y // world, y %% world], want[:, y])
    assert sum(shard_rows(H, k, world)[2] for k in range(world)) == H
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_sharding_and_assembly_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_shard_rows_cover_image():
    from stnerf_b200.dist import shard_rows
    for H in (1080, 2160, 13):
        for G in (1, 2, 4, 8):
            seen = []
            for r in range(G):
                r0, st, n = shard_rows(H, r, G)
                seen += list(range(r0, r0 + st * n, st))
            assert sorted(seen) == list(range(H))
