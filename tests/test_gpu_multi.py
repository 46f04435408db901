"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): sdfrt.dist.render_view over NCCL with world size 2
gives the same image and gradient as a single GPU (up to fp32 atomic ordering)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, W, SPP = 64, 96, 8


def _setup(dev):
    import sdfrt
    from sdfrt import synthetic as syn
    ctx = sdfrt.Context(dev)
    sdf = syn.make_sdf_device('blobs', R, torch.device('cuda', dev))
    ctx.set_sdf(sdf)
    cam = sdfrt.get_regular_cameras(12, resx=W, resy=W)[0]
    ctx.set_sensor(cam.to_world, W, W)
    ctx.set_integrator('sdf_direct_reparam')
    ctx.set_warp()
    return sdfrt, ctx, sdf


def _loss_grad(image):
    return torch.sign(image - 0.5) / image.numel()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    sdfrt, ctx, sdf = _setup(rank)
    g = torch.zeros_like(sdf)
    img = sdfrt.dist.render_view(ctx, 3, SPP * 4, 3 + 13, SPP, _loss_grad, g)
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out, image=img.cpu().numpy(), grad=g.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_match_one(tmp_path, built):
    import torch.multiprocessing as mp
    sdfrt, ctx, sdf = _setup(0)
    g1 = torch.zeros_like(sdf)
    img1 = sdfrt.dist.render_view(ctx, 3, SPP * 4, 3 + 13, SPP, _loss_grad, g1)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'r0.npz')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert np.abs(r['image'] - img1.cpu().numpy()).max() < 1e-4
    g1 = g1.cpu().numpy()
    assert np.abs(g1).max() > 0
    assert np.abs(r['grad'] - g1).max() < 1e-3 * np.abs(g1).max()
