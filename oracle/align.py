"""Oracle restatement of reference ``cross.py:9-38`` + ``unet.py:6-31,119-189``
(the spatial-alignment network ``net_T``) and ``model.py:21-28`` (test infrastructure)."""
import torch
import torch.nn.functional as F

SLOPE = 0.01  # torch.nn.LeakyReLU default, unet.py:126,133,140 and cross.py:14


def _bn(sd, p, x, training):
    """nn.BatchNorm2d (affine, momentum 0.1, eps 1e-5) unet.py:125; updates the
    running buffers of ``sd`` in place in training mode like the module does."""
    if training and (p + "num_batches_tracked") in sd:
        sd[p + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"],
                        sd[p + "weight"], sd[p + "bias"], training, 0.1, 1e-5)


def conv_bn_act(sd, p, x, training):
    """``Conv2d()`` helper unet.py:119-126: conv3x3+bias -> BN -> LeakyReLU(0.01)."""
    x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
    return F.leaky_relu(_bn(sd, p + "1.", x, training), SLOPE)


def down(sd, p, x, training):
    """``Down()`` unet.py:135-140: avgpool2 -> conv1x1 -> BN -> LReLU."""
    x = F.avg_pool2d(x, 2, stride=2)
    x = F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"])
    return F.leaky_relu(_bn(sd, p + "2.", x, training), SLOPE)


def up(sd, p, x, training):
    """``Up()`` unet.py:128-133: nearest x2 -> conv1x1 -> BN -> LReLU."""
    x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    x = F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"])
    return F.leaky_relu(_bn(sd, p + "2.", x, training), SLOPE)


def res(sd, p, x, n, training):
    """ResSequential unet.py:15-24 of ``n`` Conv2d() blocks."""
    out = x
    for i in range(n):
        out = conv_bn_act(sd, f"{p}subnet.{i}.", out, training)
    return x + out


def _level(sd, p, x, depth, max_depth, training):
    """One CatSequential of UNet.__init__ unet.py:153-173; returns cat([module(x), x])."""
    y = down(sd, p + "0.", x, training)
    y = res(sd, p + "1.", y, 2, training)
    if depth < max_depth:
        y = _level(sd, p + "2.module.", y, depth + 1, max_depth, training)
        y = conv_bn_act(sd, p + "3.", y, training)
        y = res(sd, p + "4.", y, 1, training)
        y = up(sd, p + "5.", y, training)
    else:
        y = up(sd, p + "2.", y, training)
    return torch.cat([y, x], dim=1)             # module output first, unet.py:13


def unet(sd, p, x, num_levels, training):
    """UNet.forward unet.py:144-189 (``p`` ends with ``unet.``)."""
    x = conv_bn_act(sd, p + "0.", x, training)
    x = res(sd, p + "1.", x, 1, training)
    x = _level(sd, p + "2.module.", x, 1, num_levels, training)
    x = conv_bn_act(sd, p + "3.", x, training)
    x = res(sd, p + "4.", x, 1, training)
    return F.conv2d(x, sd[p + "5.weight"], sd[p + "5.bias"], padding=1)


def identity_grid(H, W, dtype=torch.float32, device=None):
    """affine_grid(identity, align_corners=False) cross.py:24-26: pixel centres
    x_j = (2j+1)/W - 1, y_i = (2i+1)/H - 1, last dim (x, y)."""
    xs = (2 * torch.arange(W, dtype=dtype, device=device) + 1) / W - 1
    ys = (2 * torch.arange(H, dtype=dtype, device=device) + 1) / H - 1
    return torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W)], dim=-1)[None]


def spatial_transformer(sd, p, moving, fixed, training=True, num_levels=4):
    """SpatialTransformer.forward cross.py:23-30 -> (offset [N,H,W,2], grid [N,H,W,2])."""
    x = unet(sd, p + "net.0.unet.", torch.cat([moving, fixed], 1), num_levels, training)
    x = F.leaky_relu(x, SLOPE)
    x = F.conv2d(x, sd[p + "net.2.weight"], sd[p + "net.2.bias"], padding=1)
    offset = x.permute(0, 2, 3, 1)
    grid = identity_grid(moving.shape[2], moving.shape[3], moving.dtype, moving.device) + offset
    return offset, grid


def warp(img, grid):
    """SpatialTransformer.warp cross.py:32-38 = grid_sample(bilinear, zeros,
    align_corners=False), restated explicitly (4-tap gather)."""
    N, C, H, W = img.shape
    gx, gy = grid[..., 0], grid[..., 1]
    ix = ((gx + 1) * W - 1) / 2
    iy = ((gy + 1) * H - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - x0, iy - y0
    out = torch.zeros(N, C, grid.shape[1], grid.shape[2], dtype=img.dtype, device=img.device)
    flat = img.reshape(N, C, H * W)
    for dy, wy in ((0, 1 - wy1), (1, wy1)):
        for dx, wx in ((0, 1 - wx1), (1, wx1)):
            xi, yi = (x0 + dx).long(), (y0 + dy).long()
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(N, 1, -1).expand(N, C, -1)
            v = torch.gather(flat, 2, idx).reshape(N, C, *grid.shape[1:3])
            out = out + v * (wx * wy * ok.to(img.dtype))[:, None]
    return out


def gradient_loss(s):
    """model.py:21-28."""
    assert s.shape[-1] == 2
    dx = torch.abs(s[:, :, 1:, :] - s[:, :, :-1, :])
    dy = torch.abs(s[:, 1:, :, :] - s[:, :-1, :, :])
    return (torch.mean(dx * dx) + torch.mean(dy * dy)) / 2.0
