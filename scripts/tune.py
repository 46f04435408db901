#!/usr/bin/env python
"""A/B runs of the headline step under developer knobs (environment variables read at sdfrt_create) in ONE process.
usage: python scripts/tune.py [--workload W] [--steps K] cfg [cfg ...]     cfg = "name:ENV=val,ENV=val;inflight=2;partition_of=8"
Prints one line per configuration: ms_per_step (device time, CUDA events), Mrays/s, pass split."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3_blobs256_tex_512x512_64spp')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--split', action='store_true', help='also print the per-kernel union times')
    ap.add_argument('cfgs', nargs='+')
    a = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    for cfg in a.cfgs:
        name, _, rest = cfg.partition(':')
        env, inflight, part = {}, 2, 0
        for tok in filter(None, rest.replace(';', ',').split(',')):
            k, _, v = tok.partition('=')
            if k == 'inflight':
                inflight = int(v)
            elif k == 'partition_of':
                part = int(v)
            else:
                env[k] = v
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            job = bench.Job(a.workload, 0, 0, 1, inflight, part)
            ms, stats, passes, _ = job.timed(a.steps, a.warmup)
            out = {"cfg": name, "env": env, "inflight": inflight, "partition_of": part, "ms_per_step": round(ms / a.steps, 3),
                   "Mrays_s_full_frame": round(job.rays_per_step * a.steps / (ms * 1e-3) / 1e6, 1), "pass_ms": {k: round(v, 2) for k, v in passes.items()},
                   "evals": (stats['n_evals_value'] + stats['n_evals_hess']) // a.steps, "tap_loads": stats['n_tap_loads'] // a.steps}
            if a.split:
                kt, _, loop_ms = job.kernel_split(5)
                out["union_ms"] = {k: round(v[2] / 5, 2) for k, v in kt.items()}
                out["sum_ms"] = {k: round(v[0] / 5, 2) for k, v in kt.items()}
                out["profiled_loop_ms"] = round(loop_ms, 2)
            print(json.dumps(out), flush=True)
            job.close()
            del job
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


if __name__ == '__main__':
    main()
