"""world_size-2 gloo test of the multi-GPU host logic (dist.render_view): partition by film rows, film
all-reduces, gradient all-reduce.  The compute behind the backend interface is the host-compiled kernel
math (tests/hostmath.py), so this runs on CPU; on the GPU box the same function drives engine.Context."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, SPP, R = 20, 4, 24


class HostBackend:
    """engine.Context's rendering interface on top of hostmath.HostScene."""

    def __init__(self, prb=False):
        import hostmath as HM
        from oracle import sdf_oracle as O
        osc = O.make_scene('knot', res=R, width=W, height=W)
        self.prb = prb
        self.hs = HM.HostScene(osc.sdf.data.numpy(), osc.sensor.to_world, W, W, integrator=3 if prb else 2)
        self.hs.h.prb.max_depth = 3
        self.rank, self.world = 0, 1

    def set_partition(self, rank, world, rows=8):
        import hostmath as HM
        self.rank, self.world = rank, world
        HM.lib().hh_set_assign_rows(int(rows))

    def render_primal(self, seed, spp):
        if self.prb:
            return torch.from_numpy(self.hs.prb_forward(seed, spp, self.rank, self.world)[0])
        return torch.from_numpy(self.hs.forward(seed, spp, False, self.rank, self.world)[0])

    def develop(self, film):
        return torch.from_numpy(self.hs.develop(film.numpy()))

    def adjoint_accumulate(self, seed, spp):
        if self.prb:
            film, self.rec = self.hs.prb_forward(seed, spp, self.rank, self.world)
        else:
            film, self.rec, _ = self.hs.forward(seed, spp, True, self.rank, self.world)
        self.adj = (seed, spp)
        return torch.from_numpy(film)

    def adjoint_backprop(self, film, grad_image, grad_sdf, grad_albedo=None):
        fn = self.hs.prb_backward if self.prb else self.hs.backward
        g, _, _, _ = fn(self.adj[0], self.adj[1], film.numpy(), self.rec, grad_image.numpy(), self.rank, self.world)
        grad_sdf += torch.from_numpy(g)


def _loss_grad(image):
    return torch.sign(image - 0.5) / image.numel()      # d/dI of mean |I - 0.5|


def _worker(rank, world, port, out, prb=False):
    for p in (ROOT, os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sdfrt
    be = HostBackend(prb)
    g = torch.zeros(R, R, R)
    img = sdfrt.dist.render_view(be, 3, SPP * 4, 3 + 1 + 6, SPP, _loss_grad, g)
    if rank == 0:
        np.savez(out, image=img.numpy(), grad=g.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('prb', [False, True], ids=['sdf_direct_reparam', 'sdf_prb_reparam'])
def test_two_rank_view_matches_single_rank(tmp_path, built, prb):
    import sdfrt
    be = HostBackend(prb)
    g1 = torch.zeros(R, R, R)
    img1 = sdfrt.dist.render_view(be, 3, SPP * 4, 3 + 1 + 6, SPP, _loss_grad, g1)   # no process group: world 1
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'r0.npz')
    mp.spawn(_worker, args=(2, port, out, prb), nprocs=2, join=True)
    r = np.load(out)
    assert np.abs(r['image'] - img1.numpy()).max() < 1e-5
    assert np.abs(r['grad'] - g1.numpy()).max() < 1e-4 * np.abs(g1.numpy()).max()
    assert np.abs(g1.numpy()).max() > 0


def test_partition_covers_wavefront_exactly_once():
    import hostmath as HM
    from oracle import sdf_oracle as O
    osc = O.make_scene('sphere', res=8, width=13, height=7)
    hs = HM.HostScene(osc.sdf.data.numpy(), osc.sensor.to_world, 13, 7)
    for world in (1, 2, 3, 8):
        for spp in (1, 3, 16):
            n = [hs.local_lanes(spp, r, world) for r in range(world)]
            assert sum(n) == hs.wavefront(spp) and all(k % spp == 0 for k in n)   # whole pixels per rank
