#!/usr/bin/env python
"""bench.py -- primary+adjoint Mrays/s of the differentiable SDF rendering hot path.

Step = one optimisation-loop view as python/shape_opt.py:78-83 drives it: primal render at
spp*primal_spp_mult (256) + adjoint re-render/backprop at spp_grad (64), 256^3 SDF + 256^3x3 reflectance
grid, 512x512 film (BASELINE.json configs[2], the configuration the metric is quoted on).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the restated reference) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sdf kind, sdf res, albedo res, W, H, adjoint spp, primal mult, n_sensors)
    'c3_blobs256_tex_512x512_64spp': ('blobs', 256, 256, 512, 512, 64, 4, 12),
    'c2_blobs128_256x256_16spp': ('blobs', 128, None, 256, 256, 16, 4, 12),
    'c1_sphere64_128x128_4spp': ('sphere', 64, None, 128, 128, 4, 4, 6),
    'c4_blobs512_512x512_64spp': ('blobs', 512, None, 512, 512, 64, 4, 12),
    # BASELINE configs[4] (sdf_prb_reparam, 2-bounce GI): secondary line, `python bench.py --workload c5_...`
    'c5_prb_blobs256_512x512_128spp': ('blobs', 256, None, 512, 512, 128, 1, 12),
}
PRB_WORKLOADS = {'c5_prb_blobs256_512x512_128spp': dict(max_depth=3, rr_depth=5)}  # camera hit + 2 bounces
METRIC = "primary+adjoint Mrays/s"
# dram__bytes_read + dram__bytes_write of the k_march launches of ONE c3 step, from the committed ncu launch list
# (segment arrays, lane records and the L2 misses of the row-expanded grid copy; before that copy: 5.6 GB,
# profiles/r1_final_launches_c3.csv).  Only meaningful for the default workload.
NCU_DRAM_BYTES_PER_STEP = {'c3_blobs256_tex_512x512_64spp': 11.95e9}  # profiles/r1_launches_rows4_c3.csv: 24 k_march launches per step


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='sdfrt', choices=['sdfrt', 'reference'])
    ap.add_argument('--workload', default='c3_blobs256_tex_512x512_64spp', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith('active') for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the compiled all-cores CPU port of the restated reference (oracle/cpu_port.cpp) on a bounded sample
# ------------------------------------------------------------------------------------------------
_CPU_SCENE = {}


def cpu_scene(wl):
    if wl not in _CPU_SCENE:
        import sdfrt
        import sdfrt.synthetic as syn
        kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
        sdf = syn.make_sdf_device(kind, R, 'cpu').numpy()
        alb = None if AR is None else syn.make_albedo_grid(AR)
        cam = sdfrt.get_regular_cameras(nsens, resx=W, resy=H)[0]
        _CPU_SCENE[wl] = (sdf, alb, cam.to_world)
    return _CPU_SCENE[wl]


def cpu_sample(wl, n_primal, seed):
    """`n_primal` primal + n_primal/mult adjoint lanes (the workload's 4:1 ratio) from the centre of the wavefronts."""
    from oracle import cpu_port
    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
    sdf, alb, tw = cpu_scene(wl)
    r = cpu_port.time_sample(sdf, alb, tw, W, H, seed, spp * mult, spp, n_primal, max(spp, n_primal // mult),
                             prb=PRB_WORKLOADS.get(wl))
    r['seconds'] = r['seconds_primal'] + r['seconds_adjoint']
    return r


def cpu_calibrated_sample(wl, target_s, seed=100):
    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
    full = (W + 4) * (H + 4) * spp * mult
    probe = cpu_sample(wl, min(full, 1 << 19), seed)
    rate = probe['rays'] / probe['seconds']
    n = int(min(full, max(1 << 19, target_s * rate * mult / (mult + 1))))
    return n, cpu_sample(wl, n, seed)


def run_reference(args):
    """--impl reference: the reference's CPU implementation.  Mitsuba llvm_ad_rgb cannot be installed offline, so this is
    the compiled CPU port of the restated reference algorithm on all host cores (kind = "port")."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    wl = args.workload
    mult = WORKLOADS[wl][6]
    n, _ = cpu_calibrated_sample(wl, 8.0 / max(1, args.steps))   # also the warm-up
    for w in range(max(0, args.warmup - 1)):
        cpu_sample(wl, max(1 << 16, n // 8), w)
    rays = secs = 0.0
    threads = cores
    for s in range(args.steps):
        r = cpu_sample(wl, n, 100 + s)
        rays += r['rays']
        secs += r['seconds']
        threads = r['threads']
    val = rays / secs / 1e6
    sample = f"{n} primal + {max(WORKLOADS[wl][5], n // mult)} adjoint lanes per step from the wavefront centre"
    line = {"metric": METRIC, "value": val, "unit": "Mrays/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "note": "compiled CPU port of the restated reference (g++ -O3 -fopenmp); "
                                               "Mitsuba llvm_ad_rgb is not installable offline"},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import sdfrt
    from sdfrt import synthetic as syn

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[args.workload]
    ctx = sdfrt.Context(local)
    sdf = syn.make_sdf_device(kind, R, dev)
    alb = None if AR is None else torch.from_numpy(syn.make_albedo_grid(AR)).to(dev)
    sensor = sdfrt.get_regular_cameras(nsens, resx=W, resy=H)[0]
    ctx.set_sdf(sdf)
    ctx.set_albedo(alb)
    ctx.set_sensor(sensor.to_world, W, H)
    ctx.set_shading()
    ctx.set_integrator('sdf_direct_reparam')
    ctx.set_warp()
    ctx.set_partition(rank, world)
    Hb, Wb, b = ctx.film_dims()
    n_primal, n_adj = Hb * Wb * spp * mult, Hb * Wb * spp
    film_p, film_a = ctx.new_film(), ctx.new_film()
    image = torch.empty((H, W, 3), device=dev)
    grad_image = torch.ones((H, W, 3), device=dev)  # loss = dr.sum(img), figures/benchmark/benchmark.py:44
    grad_sdf = torch.zeros_like(sdf)
    grad_alb = None if alb is None else torch.zeros_like(alb)
    stream = torch.cuda.current_stream(dev)

    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)] for k in ('p0', 'p1', 'a0', 'a1', 'b1')}

    def step(i, seed, timed):
        # primal render (spp * mult) and adjoint sweep 1 (spp_grad) are independent: one pipelined call (sdfrt_render_view)
        if timed:
            ev['p0'][i].record(stream)
        ctx.render_view(seed, spp * mult, seed + 1 + nsens, spp, film_p, film_a)
        if world > 1:
            dist.all_reduce(film_p)
            dist.all_reduce(film_a)
        ctx.develop(film_p, image)
        if timed:
            ev['p1'][i].record(stream)
        grad_sdf.zero_()
        if grad_alb is not None:
            grad_alb.zero_()
        if timed:
            ev['a1'][i].record(stream)
        ctx.adjoint_backprop(film_a, grad_image, grad_sdf, grad_alb)
        if world > 1:
            dist.all_reduce(grad_sdf)
            if grad_alb is not None:
                dist.all_reduce(grad_alb)
        if timed:
            ev['b1'][i].record(stream)

    for w in range(args.warmup):
        step(0, w * (1 + nsens), False)
    torch.cuda.synchronize(dev)
    ctx.stats(reset=True)
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize(dev)
        t0.record(stream)
        for s in range(args.steps):
            step(s, 1000 + s * (1 + nsens), True)
        t1.record(stream)
        torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1)
    tt = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = tt.item()
    stats = ctx.stats(reset=True)
    rays_per_step = n_primal + n_adj
    value = rays_per_step * args.steps / (ms * 1e-3) / 1e6

    # per-kernel timings (this rank; CUDA events recorded by the library on the launching stream, inside the timed
    # region) and the roofline of the dominant kernel
    t_forward = sum(ev['p0'][i].elapsed_time(ev['p1'][i]) for i in range(args.steps)) / args.steps
    t_back = sum(ev['a1'][i].elapsed_time(ev['b1'][i]) for i in range(args.steps)) / args.steps
    # second timed loop, batches serialised on the launching stream so that the per-launch CUDA-event durations are additive
    ctx.set_overlap(False)
    ctx.set_profiling(True)
    ctx.timing()
    step(0, 7, False)
    torch.cuda.synchronize(dev)
    ctx.timing()
    ctx.stats(reset=True)
    for s_ in range(args.steps):
        step(s_, 1000 + s_ * (1 + nsens), True)
    torch.cuda.synchronize(dev)
    stats = ctx.stats(reset=True)
    kt = ctx.timing()
    ctx.set_profiling(False)
    ctx.set_overlap(True)
    kernel_ms = {k: v[0] / args.steps for k, v in kt.items()}
    kernel_ms['k_march (primary+shadow segments)'] = kernel_ms.pop('march_primary') + kernel_ms.pop('march_shadow')
    launches_per_step = sum(v[1] for v in kt.values()) / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'
    dom = max(kernel_ms, key=kernel_ms.get)
    # algorithmic bytes of k_march: one 4x4x4 fp32 neighbourhood = 256 B per B-spline evaluation it performs
    march_evals = (stats['n_evals_value'] + stats['n_evals_hess']) / args.steps
    march_bytes = 256.0 * march_evals
    march_ms = kernel_ms['k_march (primary+shadow segments)']
    achieved = march_bytes / (march_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_STEP.get(args.workload), "peak_source": peak_src, "kernel": "k_march (all launches of one step; batches serialised for this measurement)",
                "dominant_kernel": dom, "kernel_ms_per_step": kernel_ms, "pass_ms": {"primal+adjoint_sweep1": t_forward, "adjoint_backprop": t_back},
                "algorithmic_bytes_per_step": march_bytes, "evals_per_step": march_evals,
                "note": "fetch-level bytes: 256 B per B-spline evaluation; the SDF grid is L2-resident at <= 256^3, so the fetches are "
                        "served by L1/L2 and frac can exceed 1; `traffic` is the ncu dram__bytes of the same launches (profiles/)"}

    # e2e: HOST buffers in, host buffers out, copies inside the timed region.
    #  N = 1: the host-buffer C-ABI entry point sdfrt_render_view_host (H2D sdf; D2H image + grad_sdf every step)
    #  N > 1: the Python mirror's multi-GPU path (sdfrt.dist.render_view): every rank uploads the grid from pinned host memory,
    #         renders its film rows, NCCL all-reduces film and gradient, rank 0 downloads image + gradient.
    e2e = None
    e_steps = max(1, min(args.steps, 5))
    sdf_h = sdf.cpu().pin_memory()
    img_h = torch.empty((H, W, 3)).pin_memory()
    g_h = torch.empty((R, R, R)).pin_memory()
    if world == 1:
        for w in range(2):
            ctx.render_view_host(sdf_h, w, spp * mult, w + 1 + nsens, spp, None, img_h, g_h)
        torch.cuda.synchronize(dev)
        tw = time.perf_counter()
        for s in range(e_steps):
            ctx.render_view_host(sdf_h, 1000 + s, spp * mult, 1001 + s + nsens, spp, None, img_h, g_h)
        tw = time.perf_counter() - tw
        ctx.set_sdf(sdf)
    else:
        def one(seed):
            sdf.copy_(sdf_h, non_blocking=True)
            grad_sdf.zero_()
            img = sdfrt.dist.render_view(ctx, seed, spp * mult, seed + 1 + nsens, spp, lambda im: grad_image, grad_sdf)
            if rank == 0:
                img_h.copy_(img, non_blocking=True)
                g_h.copy_(grad_sdf, non_blocking=True)
            torch.cuda.synchronize(dev)
        for w in range(2):
            one(w)
        dist.barrier()
        tw = time.perf_counter()
        for s in range(e_steps):
            one(1000 + s)
        dist.barrier()
        tw = time.perf_counter() - tw
        twt = torch.tensor([tw], device=dev, dtype=torch.float64)
        dist.all_reduce(twt, op=dist.ReduceOp.MAX)
        tw = twt.item()
    e2e = {"value": rays_per_step * e_steps / tw / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": sdf_h.numel() * 4 * world,
           "d2h_bytes_per_step": img_h.numel() * 4 + g_h.numel() * 4, "steps": e_steps,
           "api": "sdfrt_render_view_host" if world == 1 else "sdfrt.dist.render_view"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        os.environ.setdefault('OMP_NUM_THREADS', str(cores))
        n, r = cpu_calibrated_sample(args.workload, 12.0)
        cpu_baseline = {"value": r['rays'] / r['seconds'] / 1e6, "unit": "Mrays/s", "cores": r['threads'], "kind": "port",
                        "sample": f"{n} primal + {max(spp, n // mult)} adjoint lanes from the wavefront centre, {r['seconds']:.1f} s; "
                                  "compiled CPU port (g++ -O3 -fopenmp) of the restated reference"}

    # second metric (protocol of figures/benchmark/benchmark.py:120-144): the per-iteration grid update on the
    # workload's SDF grid -- re-distancing (mean of 5, CUDA events) and the fused Adam step; the CPU side is the C oracle
    grid_update = None
    if rank == 0 and world == 1:
        gctx = sdfrt.redistancing.shared_context(local)
        phi = (sdf.reshape(R, R, R) * 1.7).contiguous()
        for _ in range(2):
            gctx.redistance(phi)
        gev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        gev[0].record()
        for _ in range(5):
            dist_t = gctx.redistance(phi)
        gev[1].record()
        torch.cuda.synchronize()
        rd_ms = gev[0].elapsed_time(gev[1]) / 5
        _, lists = gctx.redistance(phi, return_iters=True)
        pp = phi.clone().reshape(-1)
        gg, mm, vv = torch.randn_like(pp) * 0.01, torch.zeros_like(pp), torch.zeros_like(pp)
        gctx.adam_step(pp, gg, mm, vv, 4e-2, 1, grad_clamp=0.1)
        torch.cuda.synchronize()
        gev[0].record()
        for t in range(2, 7):
            gctx.adam_step(pp, gg, mm, vv, 4e-2, t, grad_clamp=0.1)
        gev[1].record()
        torch.cuda.synchronize()
        adam_ms = gev[0].elapsed_time(gev[1]) / 5
        grid_update = {"sdf_res": R, "redistance_ms": rd_ms, "redistance_work_lists": lists, "adam_step_ms": adam_ms,
                       "adam_step_GBps": 7 * 4 * pp.numel() / adam_ms / 1e6}
        if not args.no_cpu_baseline:
            from oracle import redistance_oracle as ro
            sub = phi[:: max(1, R // 128), :: max(1, R // 128), :: max(1, R // 128)].contiguous().cpu().numpy()
            t0 = time.time()
            want = ro.redistance(sub)
            grid_update["cpu_redistance_ms"] = (time.time() - t0) * 1e3
            grid_update["cpu_sample"] = f"C oracle (1 thread) on the grid subsampled to {sub.shape[0]}^3"
            got = gctx.redistance(torch.from_numpy(sub).cuda()).cpu().numpy()
            grid_update["max_abs_diff_vs_oracle"] = float(np.abs(got - want).max())

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "sdf_res": R, "albedo_res": AR, "film": [W, H], "spp_primal": spp * mult,
                           "spp_adjoint": spp, "rays_per_step": rays_per_step, "partition": f"interleaved 8-row film bands x{world}",
                           "l2": "inputs (sdf+albedo+grads+lane records > 1 GB) exceed L2; no explicit flush"},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(round(launches_per_step * args.steps)),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "stats_per_step": {k: v / args.steps for k, v in stats.items()},
                "grid_update": grid_update}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_gpu_prb(args):
    """Secondary line: the path replay integrator in its first, one-thread-per-sample form.  Same step
    structure and timing rules as the headline (primal + re-render in one call, then the replay/backprop), camera samples
    counted as rays; no roofline claim yet -- the wavefront form on k_march is the next step."""
    import torch
    import torch.distributed as dist
    import sdfrt
    from sdfrt import synthetic as syn
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[args.workload]
    ctx = sdfrt.Context(local)
    sdf = syn.make_sdf_device(kind, R, dev)
    sensor = sdfrt.get_regular_cameras(nsens, resx=W, resy=H)[0]
    ctx.set_sdf(sdf)
    ctx.set_albedo(None)
    ctx.set_sensor(sensor.to_world, W, H)
    ctx.set_shading()
    ctx.set_integrator('sdf_prb_reparam')
    ctx.set_prb(**PRB_WORKLOADS[args.workload])
    ctx.set_warp()
    ctx.set_partition(rank, world)
    Hb, Wb, b = ctx.film_dims()
    rays_per_step = Hb * Wb * spp * (mult + 1)
    film_p, film_a = ctx.new_film(), ctx.new_film()
    image = torch.empty((H, W, 3), device=dev)
    grad_image = torch.ones((H, W, 3), device=dev)
    grad_sdf = torch.zeros_like(sdf)
    stream = torch.cuda.current_stream(dev)

    def step(seed):
        ctx.render_view(seed, spp * mult, seed + 1 + nsens, spp, film_p, film_a)
        if world > 1:
            dist.all_reduce(film_p)
            dist.all_reduce(film_a)
        ctx.develop(film_p, image)
        grad_sdf.zero_()
        ctx.adjoint_backprop(film_a, grad_image, grad_sdf)
        if world > 1:
            dist.all_reduce(grad_sdf)

    for w in range(args.warmup):
        step(w * (1 + nsens))
    torch.cuda.synchronize(dev)
    ctx.stats(reset=True)
    if world > 1:
        dist.barrier()
    t0, t1, tm = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    with ClockSampler(local) as clk:
        torch.cuda.synchronize(dev)
        t0.record(stream)
        for s in range(args.steps):
            step(1000 + s * (1 + nsens))
        t1.record(stream)
        torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    tt = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = tt.item()
    stats = ctx.stats(reset=True)
    # e2e: host buffers in, host buffers out through the host entry point of the C ABI (N = 1 only)
    e2e = None
    if world == 1:
        sdf_h = sdf.cpu().pin_memory()
        img_h = torch.empty((H, W, 3)).pin_memory()
        g_h = torch.empty((R, R, R)).pin_memory()
        ctx.render_view_host(sdf_h, 0, spp * mult, 1 + nsens, spp, None, img_h, g_h)
        torch.cuda.synchronize(dev)
        e_steps = max(1, min(args.steps, 3))
        tw = time.perf_counter()
        for s in range(e_steps):
            ctx.render_view_host(sdf_h, 1000 + s, spp * mult, 1001 + s + nsens, spp, None, img_h, g_h)
        tw = time.perf_counter() - tw
        ctx.set_sdf(sdf)
        e2e = {"value": rays_per_step * e_steps / tw / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": sdf_h.numel() * 4,
               "d2h_bytes_per_step": img_h.numel() * 4 + g_h.numel() * 4, "steps": e_steps, "api": "sdfrt_render_view_host"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        os.environ.setdefault('OMP_NUM_THREADS', str(cores))
        n, r = cpu_calibrated_sample(args.workload, 12.0)
        cpu_baseline = {"value": r['rays'] / r['seconds'] / 1e6, "unit": "Mrays/s", "cores": r['threads'], "kind": "port",
                        "sample": f"{n} primal + {max(spp, n // mult)} adjoint lanes from the wavefront centre, {r['seconds']:.1f} s; "
                                  "compiled CPU port (g++ -O3 -fopenmp) of the restated sdf_prb_reparam"}
    if rank == 0:
        evals = (stats['n_evals_value'] + stats['n_evals_grad'] + stats['n_evals_hess'] + stats['n_evals_adjoint']) / args.steps
        line = {"metric": METRIC, "value": rays_per_step * args.steps / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "integrator": "sdf_prb_reparam", **PRB_WORKLOADS[args.workload], "sdf_res": R,
                           "film": [W, H], "spp_primal": spp * mult, "spp_adjoint": spp, "rays_per_step": rays_per_step,
                           "note": "camera samples counted as rays; every sample walks up to max_depth path segments plus "
                                   "their shadow rays (see evals_per_step)"},
                "clocks": clk.summary(), "e2e": e2e, "cpu_baseline": cpu_baseline, "gpu_launches": int(stats['n_launches']), "evals_per_step": evals,
                "fetch_GBps": 256.0 * evals / (ms / args.steps * 1e-3) / 1e9,
                "stats_per_step": {k: v / args.steps for k, v in stats.items()}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.workload in PRB_WORKLOADS:
        run_gpu_prb(a)
    else:
        run_gpu(a)
