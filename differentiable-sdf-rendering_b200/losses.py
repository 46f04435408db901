"""Image losses of python/losses.py on torch tensors (H, W, C).  Host-side callers of the hot path: they turn the
rendered image into dLoss/dImage, the `grad_in` of render_backward."""
import torch


def l2(img, ref_img):
    return ((img - ref_img) ** 2).mean()


def l1(img, ref_img):
    return (img - ref_img).abs().mean()


def mape(img, ref_img):
    rel_error = (img - ref_img).abs() / (1e-2 + ref_img.mean(dim=-1, keepdim=True)).abs()
    return rel_error.mean()


def downsample(img):
    """python/losses.py:17-32: same-size image of 2x2 forward box averages with clamped indices."""
    h, w = img.shape[0], img.shape[1]
    ys = torch.arange(h, device=img.device)
    xs = torch.arange(w, device=img.device)
    y1 = (ys + 1).clamp_max(h - 1)
    x1 = (xs + 1).clamp_max(w - 1)
    return 0.25 * (img[ys][:, xs] + img[ys][:, x1] + img[y1][:, xs] + img[y1][:, x1])


def multiscale(img, ref_img, loss_fn=l1, levels=4):
    loss = loss_fn(img, ref_img)
    for _ in range(levels - 1):
        img = downsample(img)
        ref_img = downsample(ref_img)
        loss = loss + loss_fn(img, ref_img)
    return loss / levels


def multiscale_l1(img, ref_img, levels=4):
    return multiscale(img, ref_img, l1, levels)
