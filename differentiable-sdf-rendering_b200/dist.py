"""Multi-GPU data path: one process per GPU, the SDF grid replicated, the wavefront split by
film rows (sdfrt_set_partition).  Exchange steps per view:
  1. all-reduce(sum) of the primal film accumulator   (small: (H+2b)(W+2b)*4 floats)
  2. all-reduce(sum) of the adjoint re-render's film  (the adjoint of I = V/Wt needs the global V, Wt)
  3. all-reduce(sum) of dL/d(sdf.data) (+ reflectance grid) -- the one large collective per backward.
The reference has no distributed path at all (single process, single device).
`backend` is anything with the engine.Context rendering interface; the CPU tests inject a host stand-in
under the gloo backend to exercise this logic without GPUs.
"""
import torch
import torch.distributed as dist

# granularity of the row partition (film rows per assignment unit, sdfrt_set_partition_rows).  Finer units (1, 2, 4 rows) level the
# row counts of the ranks but cost coherence; measured as each of 8 ranks on one GPU they do not lower the slowest rank's time
# (8 rows: 4.81-5.41 ms, 2 rows: 4.89-5.45 ms, 1 row: 4.98-5.39 ms per c3 step), so whole 8-row bands stay.
PARTITION_ROWS = 8


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def render_view(backend, seed, spp, seed_grad, spp_grad, grad_image_fn, grad_sdf, grad_albedo=None, group=None,
                async_grad_reduce=False, buffers_are_zero=False):
    """One optimisation-loop view (python/shape_opt.py:78-83) across the ranks of `group`.

    grad_image_fn(image) -> dLoss/dImage, evaluated redundantly on every rank from the all-reduced image.
    Gradients are ACCUMULATED into grad_sdf / grad_albedo (like dr.backward) and are complete (summed over
    ranks) on return unless async_grad_reduce=True, in which case the pending work handles are returned.
    buffers_are_zero=True: the caller has just zeroed grad_sdf / grad_albedo, so this view's contribution is written and all-reduced
    in place (no temporary grids).
    """
    rank, world = _world(group)
    backend.set_partition(rank, world, PARTITION_ROWS)
    if hasattr(backend, 'render_view'):
        # the primal render and adjoint sweep 1 are independent: one pipelined call, then both (small) film all-reduces
        film, film_a = backend.render_view(seed, spp, seed_grad, spp_grad)
        if world > 1:
            dist.all_reduce(film, group=group)
            dist.all_reduce(film_a, group=group)
        image = backend.develop(film)
        grad_image = grad_image_fn(image)
    else:
        film = backend.render_primal(seed, spp)
        if world > 1:
            dist.all_reduce(film, group=group)
        image = backend.develop(film)
        grad_image = grad_image_fn(image)
        film_a = backend.adjoint_accumulate(seed_grad, spp_grad)
        if world > 1:
            dist.all_reduce(film_a, group=group)
    local = grad_sdf if buffers_are_zero else torch.zeros_like(grad_sdf)
    local_alb = None if grad_albedo is None else (grad_albedo if buffers_are_zero else torch.zeros_like(grad_albedo))
    backend.adjoint_backprop(film_a, grad_image.contiguous(), local, local_alb)
    handles = []
    if world > 1:
        handles.append(dist.all_reduce(local, group=group, async_op=True))
        if local_alb is not None:
            handles.append(dist.all_reduce(local_alb, group=group, async_op=True))

    def finish():
        for h in handles:
            h.wait()
        if not buffers_are_zero:
            grad_sdf.add_(local)
            if local_alb is not None:
                grad_albedo.add_(local_alb)

    if async_grad_reduce:
        return image, finish
    finish()
    return image
