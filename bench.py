#!/usr/bin/env python
"""bench.py -- primary+adjoint Mrays/s of the differentiable SDF rendering hot path.

Step = one optimisation-loop view as python/shape_opt.py:78-83 drives it: primal render at
spp*primal_spp_mult (256) + adjoint re-render/backprop at spp_grad (64), 256^3 SDF + 256^3x3 reflectance
grid, 512x512 film (BASELINE.json configs[2], the configuration the metric is quoted on).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port of the restated reference instead.

The line also carries (rank 0): `grad_check` = the metric's "grad MSE vs ref" on a fixed lane subsample against the fp64 oracle,
`secondary` = short runs of BASELINE configs[3] (512^3) and configs[4] (sdf_prb_reparam) at the same N.
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
    # BASELINE configs[4] (sdf_prb_reparam, 2-bounce GI)
    'c5_prb_blobs256_512x512_128spp': ('blobs', 256, None, 512, 512, 128, 1, 12),
}
PRB_WORKLOADS = {'c5_prb_blobs256_512x512_128spp': dict(max_depth=3, rr_depth=5)}  # camera hit + 2 bounces
SECONDARY = ('c4_blobs512_512x512_64spp', 'c5_prb_blobs256_512x512_128spp')
METRIC = "primary+adjoint Mrays/s"
# dram__bytes_read + dram__bytes_write of the k_march launches of ONE c3 step, from the committed ncu launch list
NCU_DRAM_BYTES_PER_STEP = {'c3_blobs256_tex_512x512_64spp': (16.80e9, 'profiles/r2_final_launches_c3.csv')}
# the same launch list, serialised by ncu (cold caches, no overlap): per step 44.0 ms of kernels, 34.1 ms of them k_march (22 launches)
NCU_SERIALISED = {'c3_blobs256_tex_512x512_64spp': {"step_kernel_ms": 44.04, "k_march_ms": 34.07, "k_march_share": 0.774, "k_march_launches": 22,
                                                    "source": "profiles/r2_final_launches_c3.csv"}}
# what binds each kernel, from the committed `ncu --set full` captures (profiles/): unit and its % of peak
BINDING = {   # round-2 captures of the shipped kernels (ncu --set full, one batch of a c3 step), summarised under profiles/r2_call*/
    'k_march<0,1,8> primary segments (value, rows8)': {"unit": "L1 data stage (l1tex__throughput) + issue", "pct": 66, "issue_active_pct": 58,
                                                        "top_stall": "long_scoreboard 1.8", "source": "profiles/r2_call3/r2_c3_value.md"},
    'k_march<0,1,8> shadow segments (value, rows8)': {"unit": "L1 data stage (l1tex__throughput), one wavefront per 32 B sector of the scattered row loads", "pct": 67,
                                                       "issue_active_pct": 44, "top_stall": "long_scoreboard 4.4", "source": "profiles/r2_call4/r2_c3_shadow_rows8.md"},
    'k_march<1,0,8> primary segments (differentiable)': {"unit": "issue (smsp__issue_active)", "pct": 72, "source": "profiles/r2_call4/r2_c3_diff.md"},
    'k_march<0,1,8,approach> shadow segments of adjoint sweep 1': {"unit": "L1 data stage + issue", "pct": 53, "issue_active_pct": 48, "source": "profiles/r2_call4/r2_c3_diff.md"},
    'k_march<1,1,8> deferred differentiable march': {"unit": "issue (smsp__issue_active)", "pct": 57, "source": "profiles/r2_call4/r2_c3_deferred.md"},
    'k_shade': {"unit": "L1 (l1tex__throughput)", "pct": 83, "issue_active_pct": 66, "source": "profiles/r2_call3/r2_c3_value.md"},
    'k_gen_primary': {"unit": "issue (smsp__issue_active)", "pct": 86, "source": "profiles/r2_call3/r2_c3_value.md"},
    'k_backward': {"unit": "latency: long_scoreboard 2.1 + no_instruction 1.8 stall cycles per issue (grouped transposed scatter)", "pct": None,
                   "issue_active_pct": 34, "source": "profiles/r2_call3/r2_c3_backward.md"},
    'k_prb_backward (c5, sites queued)': {"unit": "latency at 8 resident warps per SM (255 registers): long_scoreboard 3.1 + no_instruction 2.2 stall cycles per issue",
                                          "pct": None, "issue_active_pct": 20, "source": "profiles/r2_call15/r2_c5_prb_bwd_queue.md"},
    'k_scatter_sites (c5)': {"unit": "shared-memory transpose (short_scoreboard 3.0) + RED", "pct": None, "issue_active_pct": 50,
                             "source": "profiles/r2_call15/r2_c5_prb_bwd_queue.md"},
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='sdfrt', choices=['sdfrt', 'reference'])
    ap.add_argument('--workload', default='c3_blobs256_tex_512x512_64spp', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the short c4 / c5 records')
    ap.add_argument('--no-grad-check', action='store_true', help='skip the lane-subsample gradient check against the oracle')
    ap.add_argument('--no-e2e', action='store_true', help='profiling runs: skip the host-buffer leg')
    ap.add_argument('--views-in-flight', type=int, default=2, help='independent views (steps) enqueued concurrently on separate streams / view slots')
    ap.add_argument('--e2e-contexts', type=int, default=4, help='N = 1: contexts (host threads) of the host-buffer leg')
    ap.add_argument('--partition-rows', type=int, default=8, choices=[1, 2, 4, 8], help='film rows per unit of the interleaved row partition')
    ap.add_argument('--partition-rank', type=int, default=0, help='tuning: which rank of --partition-of to emulate')
    ap.add_argument('--partition-of', type=int, default=0, help='tuning: render rank 0 of a partition into this many ranks on one GPU (no collectives)')
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


def host_cores():
    """Usable host cores: the affinity mask, capped by the cgroup CPU quota (a container may see 128 CPUs and be allowed 16)."""
    n = len(os.sched_getaffinity(0))
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, p = f.read().split()[:2]
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except Exception:
        try:
            q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
            p = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if q > 0:
                n = max(1, min(n, int(q / p + 0.5)))
        except Exception:
            pass
    return n


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
    """`n_primal` primal + n_primal/mult adjoint lanes (the workload's 4:1 ratio), in chunks of 64 consecutive samples strided over
    the WHOLE wavefronts (bounding-box misses, empty space, grazing rays and hits in the workload's proportions), on all host
    cores of the affinity mask (set explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    from oracle import cpu_port
    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
    sdf, alb, tw = cpu_scene(wl)
    r = cpu_port.time_sample(sdf, alb, tw, W, H, seed, spp * mult, spp, n_primal, max(spp, n_primal // mult),
                             prb=PRB_WORKLOADS.get(wl), threads=_CPU_THREADS.get('n', host_cores()))
    r['seconds'] = r['seconds_primal'] + r['seconds_adjoint']
    return r


_CPU_THREADS = {}


def cpu_calibrated_sample(wl, target_s, seed=100):
    kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
    full = (W + 4) * (H + 4) * spp * mult
    # "all the host threads it can use": more threads are not always faster on a shared host (SMT siblings, CPU quota), so probe
    # the full affinity count and its half / quarter once and keep the fastest
    if 'n' not in _CPU_THREADS:
        best = None
        for t in sorted({host_cores(), max(1, host_cores() // 2), max(1, host_cores() // 4)}, reverse=True):
            _CPU_THREADS['n'] = t
            cpu_sample(wl, min(full, 1 << 17), seed)          # warm the thread pool
            p = cpu_sample(wl, min(full, 1 << 19), seed)
            r = p['rays'] / p['seconds']
            if best is None or r > best[0]:
                best = (r, t)
        _CPU_THREADS['n'] = best[1]
    probe = cpu_sample(wl, min(full, 1 << 19), seed)
    rate = probe['rays'] / probe['seconds']
    n = int(min(full, max(1 << 19, target_s * rate * mult / (mult + 1))))
    return n, cpu_sample(wl, n, seed)


def cpu_record(wl, n, r, kind_note):
    mult, spp = WORKLOADS[wl][6], WORKLOADS[wl][5]
    return {"value": r['rays'] / r['seconds'] / 1e6, "unit": "Mrays/s", "cores": r['threads'], "kind": "port",
            "sample": f"{n} primal + {max(spp, n // mult)} adjoint lanes, 64-sample chunks strided over the whole wavefronts, "
                      f"{r['seconds']:.1f} s; every ray marched (no empty-space cull); {kind_note} ({r['build']})",
            "evals_per_s": r['evals'] / r['seconds'], "hit_fraction": r['hits'] / r['rays']}


def run_reference(args):
    """--impl reference: the reference's CPU implementation.  Mitsuba llvm_ad_rgb cannot be installed offline, so this is
    the compiled CPU port of the restated reference algorithm on all host cores (kind = "port")."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = args.workload
    mult = WORKLOADS[wl][6]
    n, _ = cpu_calibrated_sample(wl, 8.0 / max(1, args.steps))   # also the warm-up
    for w in range(max(0, args.warmup - 1)):
        cpu_sample(wl, max(1 << 16, n // 8), w)
    rays = secs = evals = hits = 0.0
    threads, how = host_cores(), ''
    for s in range(args.steps):
        r = cpu_sample(wl, n, 100 + s)
        rays += r['rays']
        secs += r['seconds']
        evals += r['evals']
        hits += r['hits']
        threads, how = r['threads'], r['build']
    val = rays / secs / 1e6
    sample = (f"{n} primal + {max(WORKLOADS[wl][5], n // mult)} adjoint lanes per step, 64-sample chunks strided over the whole "
              f"wavefronts; every ray marched; g++ {how}")
    line = {"metric": METRIC, "value": val, "unit": "Mrays/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "note": "compiled CPU port of the restated reference; Mitsuba llvm_ad_rgb is not installable offline"},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                             "evals_per_s": evals / secs, "hit_fraction": hits / rays},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Job:
    """One workload bound to a context: buffers + the step function (device buffers resident; one optimisation-loop view)."""

    def __init__(self, wl, local, rank, world, inflight=2, partition_of=0, partition_rank=0, partition_rows=8):
        import torch
        import sdfrt
        from sdfrt import synthetic as syn
        self.torch, self.sdfrt = torch, sdfrt
        self.wl, self.rank, self.world = wl, rank, world
        kind, R, AR, W, H, spp, mult, nsens = WORKLOADS[wl]
        self.R, self.AR, self.W, self.H, self.spp, self.mult, self.nsens = R, AR, W, H, spp, mult, nsens
        self.dev = dev = torch.device('cuda', local)
        self.prb = PRB_WORKLOADS.get(wl)
        self.ctx = ctx = sdfrt.Context(local)
        self.sdf = syn.make_sdf_device(kind, R, dev)
        self.alb = None if AR is None else torch.from_numpy(syn.make_albedo_grid(AR)).to(dev)
        self.sensor = sdfrt.get_regular_cameras(nsens, resx=W, resy=H)[0]
        ctx.set_sdf(self.sdf)
        ctx.set_albedo(self.alb)
        ctx.set_sensor(self.sensor.to_world, W, H)
        ctx.set_shading()
        if self.prb:
            ctx.set_integrator('sdf_prb_reparam')
            ctx.set_prb(**self.prb)
        else:
            ctx.set_integrator('sdf_direct_reparam')
        ctx.set_warp()
        self.partition_rows = partition_rows
        self.e2e_contexts = 4
        if partition_of > 1:
            ctx.set_partition(partition_rank, partition_of, partition_rows)
        else:
            ctx.set_partition(rank, world, partition_rows)
        Hb, Wb, b = ctx.film_dims()
        self.n_primal, self.n_adj = Hb * Wb * spp * mult, Hb * Wb * spp
        self.rays_per_step = self.n_primal + self.n_adj
        self.grad_image = torch.ones((H, W, 3), device=dev)  # loss = dr.sum(img), figures/benchmark/benchmark.py:44
        self.stream = torch.cuda.current_stream(dev)
        # Views in flight: the views of an optimisation iteration are independent (python/shape_opt.py:77-83), so consecutive steps
        # are enqueued on alternating streams with their own view slot (sdfrt_set_view_slot), film, image and gradient buffers: the
        # persistent kernels of one view fill the latency tails of the other.  Every step still produces its own image + gradients.
        self.inflight = max(1, min(inflight, ctx.num_view_slots()))
        self.lanes = []
        for k in range(self.inflight):
            # the two film accumulators and the two gradient grids of a view live in one allocation each: one collective per pair
            films = torch.empty((2, Hb, Wb, 4), device=dev)
            nv, na = self.sdf.numel(), 0 if self.alb is None else self.alb.numel()
            grads = torch.zeros(nv + na, device=dev)
            self.lanes.append(dict(stream=self.stream if self.inflight == 1 else torch.cuda.Stream(dev), films=films, film_p=films[0], film_a=films[1],
                                   image=torch.empty((H, W, 3), device=dev), grads=grads, grad_sdf=grads[:nv].view(self.sdf.shape),
                                   grad_alb=None if self.alb is None else grads[nv:].view(self.alb.shape)))
        self.grad_sdf, self.grad_alb = self.lanes[0]['grad_sdf'], self.lanes[0]['grad_alb']
        self.n_step = 0

    def step(self, seed, ev=None):
        """primal render (spp * mult) and adjoint sweep 1 (spp_grad) in one pipelined call, film all-reduces, develop, adjoint
        sweep 2, gradient all-reduce (the one large collective per backward) -- on the stream / view slot of this step's lane."""
        import torch.distributed as dist
        torch, ctx, world = self.torch, self.ctx, self.world
        L = self.lanes[self.n_step % self.inflight]
        k = self.n_step % self.inflight
        self.n_step += 1
        with torch.cuda.stream(L['stream']):
            if ev:
                ev[0].record()
            ctx.set_view_slot(k)
            ctx.render_view(seed, self.spp * self.mult, seed + 1 + self.nsens, self.spp, L['film_p'], L['film_a'])
            if world > 1:
                dist.all_reduce(L['films'])     # primal film + film of adjoint sweep 1
            ctx.develop(L['film_p'], L['image'])
            if ev:
                ev[1].record()
            L['grads'].zero_()
            ctx.adjoint_backprop(L['film_a'], self.grad_image, L['grad_sdf'], L['grad_alb'])
            if world > 1:
                dist.all_reduce(L['grads'])     # dL/d(sdf) + dL/d(reflectance grid): the one large collective per backward
            if ev:
                ev[2].record()

    def fork(self):
        """the lanes' streams start after everything queued on the main stream ..."""
        if self.inflight > 1:
            e = self.torch.cuda.Event()
            e.record(self.stream)
            for L in self.lanes:
                L['stream'].wait_event(e)

    def join(self):
        """... and the main stream continues after all of them"""
        if self.inflight > 1:
            for L in self.lanes:
                e = self.torch.cuda.Event()
                e.record(L['stream'])
                self.stream.wait_event(e)

    def timed(self, steps, warmup, clocks=False):
        """W untimed steps, then exactly K steps bracketed by barrier + synchronize; device time, max over ranks."""
        torch = self.torch
        import torch.distributed as dist
        self.fork()
        for w in range(warmup):
            self.step(w * (1 + self.nsens))
        self.join()
        torch.cuda.synchronize(self.dev)
        self.ctx.stats(reset=True)
        if self.world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        clk = ClockSampler(self.dev.index) if clocks else None
        if clk:
            clk.__enter__()
        torch.cuda.synchronize(self.dev)
        t0.record(self.stream)
        self.fork()
        for s in range(steps):
            self.step(1000 + s * (1 + self.nsens), ev[s])
        self.join()
        t1.record(self.stream)
        torch.cuda.synchronize(self.dev)
        if clk:
            clk.__exit__()
        if self.world > 1:
            dist.barrier()
        ms = t0.elapsed_time(t1)
        tt = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
        stats = self.ctx.stats(reset=True)
        t_fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
        t_bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
        return ms, stats, {"primal+adjoint_sweep1": t_fwd, "adjoint_backprop+grad_allreduce": t_bwd}, (clk.summary() if clk else None)

    def kernel_split(self, steps):
        """Per-kernel CUDA-event timings (recorded by the library on the launching streams) over `steps` steps run exactly like the
        timed region (batches overlapping on the internal streams): summed launch durations and, per class, the union of the launch
        intervals -- the time that class was on the device."""
        torch, ctx = self.torch, self.ctx
        ctx.set_profiling(True)
        ctx.timing()
        self.fork()
        self.step(7)
        self.join()
        torch.cuda.synchronize(self.dev)
        ctx.timing()
        ctx.stats(reset=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(self.stream)
        self.fork()
        for s in range(steps):
            self.step(1000 + s * (1 + self.nsens))
        self.join()
        t1.record(self.stream)
        torch.cuda.synchronize(self.dev)
        stats = ctx.stats(reset=True)
        kt = ctx.timing(union=True)
        ctx.set_profiling(False)
        return kt, stats, t0.elapsed_time(t1) / steps

    def close(self):
        self.ctx = None
        for k in ('sdf', 'alb', 'grad_sdf', 'grad_alb', 'grad_image', 'lanes'):
            setattr(self, k, None)
        self.torch.cuda.empty_cache()


def march_roofline(job, kt, stats, steps, peak, peak_src, loop_ms):
    kernel_ms = {k: v[0] / steps for k, v in kt.items() if k != 'march_all'}
    union_ms = {k: v[2] / steps for k, v in kt.items()}
    march_evals = (stats['n_evals_value'] + stats['n_evals_hess']) / steps
    march_bytes = 256.0 * march_evals          # one 4x4x4 fp32 neighbourhood per B-spline evaluation
    march_ms = union_ms['march_all']
    achieved = march_bytes / (march_ms * 1e-3) / 1e9
    traffic = NCU_DRAM_BYTES_PER_STEP.get(job.wl)
    launches = sum(v[1] for k, v in kt.items() if k != 'march_all') / steps
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None, "peak_source": peak_src,
            "kernel": "k_march: all launches of one step", "ncu_serialised": NCU_SERIALISED.get(job.wl),
            "kernel_ms": march_ms, "kernel_ms_how": "union of the k_march launch intervals (CUDA events recorded by the library around every "
                                                    "launch), steps run as in the timed region with overlapping batches",
            "profiled_loop_ms_per_step": loop_ms,
            "launch_ms_sum_per_step": kernel_ms, "launch_ms_union_per_step": union_ms,
            "algorithmic_bytes_per_step": march_bytes, "evals_per_step": march_evals,
            "frac_vs_timed_step": None,
            "binding": BINDING,
            "note": "fetch-level bytes: 256 B per B-spline evaluation; the 256^3 grid is L2-resident, so the fetches are served by "
                    "L1/L2 (43 % by the register tap cache) and the figure is NOT an HBM utilisation: `traffic` is the ncu dram__bytes "
                    "of the same launches; `binding` names the unit that limits each kernel (ncu --set full, profiles/)"}, launches


def grad_check(job, n_lanes=16384):
    """Metric: "...; grad MSE vs ref".  A fixed window of wavefront lanes through the middle of the film (256 pixels x 64 spp at c3):
    sdfrt_adjoint_backprop restricted to those lanes (sdfrt_set_lane_window) under the completed film of the whole re-render, against
    the fp64 oracle evaluating the same lanes (oracle/window_check.py).  Whole wavefront on this rank (partition 0 of 1)."""
    import numpy as np
    torch = job.torch
    from oracle import sdf_oracle as O
    from oracle import window_check as WC
    ctx = job.ctx
    t0 = time.time()
    ctx.set_view_slot(0)
    ctx.set_partition(0, 1)
    sens = O.Sensor(to_world=job.sensor.to_world, width=job.W, height=job.H)
    osc = O.Scene(sdf=O.Grid3d(job.sdf.cpu().double()), sensor=sens, albedo_grid=None if job.alb is None else job.alb.cpu().double())
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    lo, hi = WC.centre_window(sens, job.spp, n_lanes)
    G = np.random.default_rng(4).uniform(-1, 1, size=(job.H, job.W, 3)).astype(np.float32)
    seed_grad = 13
    film = ctx.adjoint_accumulate(seed_grad, job.spp)
    ctx.set_lane_window(lo, hi)
    g = torch.zeros_like(job.sdf)
    ga = None if job.alb is None else torch.zeros_like(job.alb)
    ctx.adjoint_backprop(film, torch.from_numpy(G).to(job.dev), g, ga)
    ctx.set_lane_window(0, 0)
    ctx.set_partition(job.rank, job.world, job.partition_rows)
    _, g_o, ga_o = WC.window_gradient(osc, owf, 'sdf_direct_reparam', film.cpu().numpy(), G, seed_grad, job.spp, (lo, hi))
    out = {"vs": "fp64 oracle (oracle/sdf_oracle.py, torch autograd) on the same lanes", "lanes": [lo, hi],
           "grad_image": "uniform(-1,1), seed 4", "sdf": WC.gradient_errors(g.cpu().numpy(), g_o.numpy())}
    if ga is not None:
        out["reflectance"] = WC.gradient_errors(ga.cpu().numpy(), ga_o.numpy())
    out["grad_mse_vs_oracle"] = out["sdf"]["mse"]
    out["grad_max_rel"] = out["sdf"]["max_abs_over_ref_max"]
    out["seconds"] = time.time() - t0
    return out


def e2e_leg(job, steps):
    """Host buffers in, host buffers out, copies inside the timed region.
      N = 1: the host-buffer C-ABI entry point (sdfrt_render_view_host_ex): H2D sdf (+ reflectance grid), D2H image + both gradients
      N > 1: sdfrt.dist.render_view: every rank uploads the grids from pinned host memory, renders its film rows, NCCL all-reduces
             film and gradients, rank 0 downloads image + gradients."""
    torch, sdfrt = job.torch, job.sdfrt
    import torch.distributed as dist
    e_steps = max(1, min(steps, 5))
    sdf_h = job.sdf.cpu().pin_memory()
    alb_h = None if job.alb is None else job.alb.cpu().pin_memory()
    img_h = torch.empty((job.H, job.W, 3)).pin_memory()
    g_h = torch.empty(tuple(job.sdf.shape)).pin_memory()
    ga_h = None if job.alb is None else torch.empty(tuple(job.alb.shape)).pin_memory()
    nsens, spp, mult = job.nsens, job.spp, job.mult
    if job.world == 1:
        # Views in flight on the host API too: several contexts, one host thread each (the C call releases the GIL), every one with its own
        # pinned host buffers.  A context's view is upload -> kernels -> download in sequence, so with several contexts (default 4: 2406 / 2586 / 2644 Mrays/s with 2 / 3 / 4) the copies of one view
        # run while two others keep the SMs busy (PCIe is full duplex); with two, the kernels of one view ran alone a third of the time.
        import threading
        nctx = max(1, job.e2e_contexts)
        ctxs = [job.ctx] + [sdfrt.Context(job.dev.index) for _ in range(nctx - 1)]
        for c in ctxs[1:]:
            c.set_sensor(job.sensor.to_world, job.W, job.H)
            c.set_shading()
            c.set_integrator('sdf_prb_reparam' if job.prb else 'sdf_direct_reparam')
            if job.prb:
                c.set_prb(**job.prb)
            c.set_warp()
        bufs = [(img_h, g_h, ga_h)] + [(torch.empty_like(img_h).pin_memory(), torch.empty_like(g_h).pin_memory(),
                                        None if ga_h is None else torch.empty_like(ga_h).pin_memory()) for _ in range(nctx - 1)]

        def one(k, seed):
            im, g, ga = bufs[k]
            ctxs[k].render_view_host(sdf_h, seed, spp * mult, seed + 1 + nsens, spp, None, im, g, albedo_host=alb_h, grad_albedo_host=ga)

        def run(n_steps, seed0):
            def worker(k):
                for s in range(k, n_steps, nctx):
                    one(k, seed0 + s)
            th = [threading.Thread(target=worker, args=(k,)) for k in range(nctx)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        run(2 * nctx, 0)
        torch.cuda.synchronize(job.dev)
        e_steps = max(e_steps, 3 * nctx)     # enough views to amortise the fill and drain of the copy / kernel pipeline
        tw = time.perf_counter()
        run(e_steps, 1000)
        tw = time.perf_counter() - tw
        job.ctx.set_sdf(job.sdf)
        job.ctx.set_albedo(job.alb)
        del ctxs
        api = f"sdfrt_render_view_host_ex ({nctx} contexts / host threads, one view each in flight)" if nctx > 1 else "sdfrt_render_view_host_ex"
    else:
        # N > 1: sdfrt.dist.render_view per view (set_partition, film all-reduces, backprop, gradient all-reduce over NCCL).  The host
        # copies are double-buffered around it: the grids of view v+1 are uploaded on a copy stream into the other pair of device
        # buffers while view v renders, and rank 0's downloads of view v run on the copy stream behind the next view's kernels.
        copy_s = torch.cuda.Stream(job.dev)
        # ONE rank reads the host buffers: rank 0 uploads SDF + reflectance grid (one pinned allocation, one copy) and broadcasts them over
        # NVLink on a communicator of its own, so that the broadcast of view v+1 queues behind nothing of view v.  (Every rank uploading its
        # own replica was bound by the host side at 8 GPUs: 2.1 GB per step through shared PCIe uplinks, 12 ms per step against 5.9 of kernels.)
        bgroup = dist.new_group(backend='nccl')
        nv, na = job.sdf.numel(), 0 if job.alb is None else job.alb.numel()
        grid_h = torch.empty(nv + na).pin_memory()
        grid_h[:nv].copy_(sdf_h.reshape(-1))
        if na:
            grid_h[nv:].copy_(alb_h.reshape(-1))
        grid_d = [torch.empty(nv + na, device=job.dev) for _ in range(2)]
        sdf_d = [g[:nv].view(job.sdf.shape) for g in grid_d]
        alb_d = [None if job.alb is None else g[nv:].view(job.alb.shape) for g in grid_d]
        gs_d = [job.grad_sdf, torch.empty_like(job.grad_sdf)]
        ga_d = [job.grad_alb, None if job.grad_alb is None else torch.empty_like(job.grad_alb)]
        outs = [(img_h, g_h, ga_h), (torch.empty_like(img_h).pin_memory(), torch.empty_like(g_h).pin_memory(),
                                     None if ga_h is None else torch.empty_like(ga_h).pin_memory())]
        up_done = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]      # device buffers k may be overwritten (their view's kernels are done)
        main = torch.cuda.current_stream(job.dev)
        # Two views in flight here too: view v runs on context v % 2 and on that context's stream (one host thread, so every rank issues
        # its collectives in the same order).  Two CONTEXTS rather than two view slots of one, because each view binds its own upload of
        # the grid and the derived copies (rows8, brick minima) belong to the context.
        ctxs = [job.ctx, sdfrt.Context(job.dev.index)]
        ctxs[1].set_sensor(job.sensor.to_world, job.W, job.H)
        ctxs[1].set_shading()
        ctxs[1].set_integrator('sdf_prb_reparam' if job.prb else 'sdf_direct_reparam')
        if job.prb:
            ctxs[1].set_prb(**job.prb)
        ctxs[1].set_warp()
        lanes = [torch.cuda.Stream(job.dev), torch.cuda.Stream(job.dev)]
        job.ctx.set_view_slot(0)

        def upload(k):
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(free[k])
                if job.rank == 0:
                    grid_d[k].copy_(grid_h, non_blocking=True)
                dist.broadcast(grid_d[k], src=0, group=bgroup)
                up_done[k].record(copy_s)

        def view(k, seed):
            done = torch.cuda.Event()
            with torch.cuda.stream(lanes[k]):
                lanes[k].wait_event(up_done[k])
                ctxs[k].set_sdf(sdf_d[k])
                ctxs[k].set_albedo(alb_d[k])
                gs_d[k].zero_()
                if ga_d[k] is not None:
                    ga_d[k].zero_()
                img = sdfrt.dist.render_view(ctxs[k], seed, spp * mult, seed + 1 + nsens, spp, lambda im: job.grad_image, gs_d[k], ga_d[k], buffers_are_zero=True)
                done.record(lanes[k])
            if job.rank == 0:
                with torch.cuda.stream(copy_s):
                    copy_s.wait_event(done)
                    outs[k][0].copy_(img, non_blocking=True)
                    outs[k][1].copy_(gs_d[k], non_blocking=True)
                    if ga_d[k] is not None:
                        outs[k][2].copy_(ga_d[k], non_blocking=True)
                    img.record_stream(copy_s)
                    free[k].record(copy_s)
            else:
                free[k].record(lanes[k])

        def run(n_steps, seed0):
            for k in range(2):
                free[k].record(main)
            upload(0)
            for s_ in range(n_steps):
                if s_ + 1 < n_steps:
                    upload((s_ + 1) % 2)
                view(s_ % 2, seed0 + s_)
            torch.cuda.synchronize(job.dev)
        run(2, 0)
        dist.barrier()
        e_steps = max(e_steps, 12)           # the first upload and the last download are not overlapped: amortise them
        tw = time.perf_counter()
        run(e_steps, 1000)
        dist.barrier()
        tw = time.perf_counter() - tw
        twt = torch.tensor([tw], device=job.dev, dtype=torch.float64)
        dist.all_reduce(twt, op=dist.ReduceOp.MAX)
        tw = twt.item()
        job.ctx.set_sdf(job.sdf)
        job.ctx.set_albedo(job.alb)
        del ctxs
        api = "sdfrt.dist.render_view (rank 0 uploads, NVLink broadcast of the grids; two views in flight on two contexts, uploads / downloads on a copy stream)"
    h2d = (sdf_h.numel() + (alb_h.numel() if alb_h is not None else 0)) * 4 * (job.world if job.world == 1 else 1)
    d2h = (img_h.numel() + g_h.numel() + (ga_h.numel() if ga_h is not None else 0)) * 4
    return {"value": job.rays_per_step * e_steps / tw / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": e_steps, "api": api}


def secondary_record(wl, local, rank, world, peak, inflight=2):
    """A short run (2 warm-up + 4 timed steps, same timing rules) of another BASELINE config at the same N."""
    job = Job(wl, local, rank, world, inflight)
    steps = 4
    ms, stats, passes, _ = job.timed(steps, 2)
    evals = (stats['n_evals_value'] + stats['n_evals_grad'] + stats['n_evals_hess'] + stats['n_evals_adjoint']) / steps
    rec = {"value": job.rays_per_step * steps / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms / steps, "steps": steps, "warmup": 1,
           "n_gpus": world, "rays_per_step": job.rays_per_step, "evals_per_step_this_rank": evals,
           "fetch_GBps_this_rank": 256.0 * evals / (ms / steps * 1e-3) / 1e9, "pass_ms": passes,
           "marched_fraction": 1.0 - stats['n_skipped'] / max(1, stats['n_rays'] + stats['n_skipped']) if not job.prb else None,
           "config": {"sdf_res": job.R, "film": [job.W, job.H], "spp_primal": job.spp * job.mult, "spp_adjoint": job.spp,
                      **({"integrator": "sdf_prb_reparam", **job.prb} if job.prb else {"integrator": "sdf_direct_reparam"})}}
    if not job.prb:
        march_evals = (stats['n_evals_value'] + stats['n_evals_hess']) / steps
        rec["roofline"] = {"bound": "hbm", "achieved": 256.0 * march_evals / (ms / steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": 256.0 * march_evals / (ms / steps * 1e-3) / 1e9 / peak,
                           "note": "this rank's k_march fetch bytes over the whole step time (lower bound of the kernel's rate)"}
    job.close()
    return rec


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json, burst copy)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'

    job = Job(args.workload, local, rank, world, args.views_in_flight, args.partition_of, args.partition_rank, args.partition_rows)
    job.e2e_contexts = args.e2e_contexts
    ms, stats, passes, clocks = job.timed(args.steps, args.warmup, clocks=True)
    value = job.rays_per_step * args.steps / (ms * 1e-3) / 1e6
    n_cam = stats['n_rays'] + 0   # camera samples this rank shaded (marched or culled)
    skipped = stats['n_skipped']

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sdf_res": job.R, "albedo_res": job.AR, "film": [job.W, job.H],
                       "spp_primal": job.spp * job.mult, "spp_adjoint": job.spp, "rays_per_step": job.rays_per_step,
                       "partition": f"interleaved {args.partition_rows}-row film sub-bands x{world}, walked in 8-row tiles" + (f" (tuning: rank {args.partition_rank} of {args.partition_of} on one GPU, no collectives)" if args.partition_of > 1 else ""),
                       "views_in_flight": job.inflight,
                       **({"integrator": "sdf_prb_reparam", **job.prb} if job.prb else {}),
                       "l2": "inputs (sdf+albedo+grads+lane records > 1 GB) exceed L2; no explicit flush"},
            "clocks": clocks, "pass_ms": passes}
    # camera rays the empty-space test proved clear are counted in `value` (the reference's wavefront counts them too) but never
    # marched; the rate over marched rays only:
    if not job.prb:
        frac_marched = 1.0 - skipped / max(1, stats['n_rays'])
        line["value_marched_rays"] = value * frac_marched
        line["marched_fraction_this_rank"] = frac_marched
    ev_t = torch.tensor([float(stats['n_evals_value'] + stats['n_evals_grad'] + stats['n_evals_hess'] + stats['n_evals_adjoint'])],
                        device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ev_t)
    line["evals_per_s"] = ev_t.item() / (ms * 1e-3)     # B-spline evaluations of all ranks per second (the CPU arm prints the same)
    line["stats_per_step_this_rank"] = {k: v / args.steps for k, v in stats.items()}

    kt, kstats, loop_ms = job.kernel_split(max(2, min(args.steps, 10)))
    ksteps = max(2, min(args.steps, 10))
    if not job.prb:
        roof, launches = march_roofline(job, kt, kstats, ksteps, peak, peak_src, loop_ms)
        roof["frac_vs_timed_step"] = roof["algorithmic_bytes_per_step"] / (ms / args.steps * 1e-3) / 1e9 / peak
        line["roofline"] = roof
    else:
        launches = sum(v[1] for k, v in kt.items() if k != 'march_all') / ksteps
        evals = (kstats['n_evals_value'] + kstats['n_evals_grad'] + kstats['n_evals_hess'] + kstats['n_evals_adjoint']) / ksteps
        line["roofline"] = {"bound": "hbm", "achieved": 256.0 * evals / (ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": 256.0 * evals / (ms / args.steps * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                            "kernel": "all lookups of one step over the step time",
                            "launch_ms_sum_per_step": {k: v[0] / ksteps for k, v in kt.items()},
                            "launch_ms_union_per_step": {k: v[2] / ksteps for k, v in kt.items()}}
    line["gpu_launches"] = int(round(launches * args.steps))

    line["e2e"] = None if args.no_e2e else e2e_leg(job, args.steps)

    if rank == 0 and not args.no_grad_check and not job.prb:
        try:
            line["grad_check"] = grad_check(job)
        except Exception as e:  # the checker must never take the measurement down
            line["grad_check"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()

    line["cpu_baseline"] = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n, r = cpu_calibrated_sample(args.workload, 5.0)
        line["cpu_baseline"] = cpu_record(args.workload, n, r, "compiled CPU port of the restated reference")

    # second metric (protocol of figures/benchmark/benchmark.py:120-144): the per-iteration grid update on the
    # workload's SDF grid -- re-distancing (mean of 5, CUDA events) and the fused Adam step; the CPU side is the C oracle
    if rank == 0 and world == 1 and not job.prb:
        line["grid_update"] = grid_update_record(job, args)

    job.close()
    if not args.no_secondary and args.workload == 'c3_blobs256_tex_512x512_64spp':
        sec = {}
        for wl in SECONDARY:
            try:
                sec[wl] = secondary_record(wl, local, rank, world, peak, args.views_in_flight)
            except Exception as e:
                sec[wl] = {"error": repr(e)}
        line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def grid_update_record(job, args):
    import numpy as np
    torch, sdfrt = job.torch, job.sdfrt
    R = job.R
    gctx = sdfrt.redistancing.shared_context(job.dev.index)
    phi = (job.sdf.reshape(R, R, R) * 1.7).contiguous()
    for _ in range(2):
        gctx.redistance(phi)
    gev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    gev[0].record()
    for _ in range(5):
        gctx.redistance(phi)
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
    out = {"sdf_res": R, "redistance_ms": rd_ms, "redistance_work_lists": lists, "adam_step_ms": adam_ms,
           "adam_step_GBps": 7 * 4 * pp.numel() / adam_ms / 1e6}
    if not args.no_cpu_baseline:
        from oracle import redistance_oracle as ro
        sub = phi[:: max(1, R // 128), :: max(1, R // 128), :: max(1, R // 128)].contiguous().cpu().numpy()
        t0 = time.time()
        want = ro.redistance(sub)
        out["cpu_redistance_ms"] = (time.time() - t0) * 1e3
        out["cpu_sample"] = f"C oracle (1 thread) on the grid subsampled to {sub.shape[0]}^3"
        got = gctx.redistance(torch.from_numpy(sub).cuda()).cpu().numpy()
        out["max_abs_diff_vs_oracle"] = float(np.abs(got - want).max())
    return out


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)
