"""Seeded synthetic EuRoC-shaped inputs.  No dataset is available offline, so
every test and benchmark uses these generators.  Pure numpy; no oracle, no GPU."""
import numpy as np


def _upsample_bilinear(small, H, W):
    h, w = small.shape
    ys = (np.arange(H) + 0.5) * h / H - 0.5
    xs = (np.arange(W) + 0.5) * w / W - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 1)
    y1 = np.clip(y0 + 1, 0, h - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    a = small[y0][:, x0] * (1 - fx) + small[y0][:, x1] * fx
    b = small[y1][:, x0] * (1 - fx) + small[y1][:, x1] * fx
    return a * (1 - fy) + b * fy


def scene_image(seed, W=752, H=480, n_shapes=None, noise=2.0):
    """Corner-rich 'machine hall' image: smooth background + random rectangles/triangles + noise."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(40, 216, (max(H // 8, 2), max(W // 8, 2))).astype(np.float64)
    img = _upsample_bilinear(bg, H, W)
    if n_shapes is None:
        n_shapes = int(rng.integers(300, 601) * (W * H) / (752 * 480))
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(n_shapes):
        g = float(rng.integers(0, 256))
        cx, cy = rng.integers(0, W), rng.integers(0, H)
        sw, sh = rng.integers(6, 60), rng.integers(6, 60)
        x0, x1 = max(cx - sw // 2, 0), min(cx + sw // 2 + 1, W)
        y0, y1 = max(cy - sh // 2, 0), min(cy + sh // 2 + 1, H)
        if x1 <= x0 or y1 <= y0:
            continue
        if rng.random() < 0.6:
            img[y0:y1, x0:x1] = g
        else:  # right triangle
            sub_y, sub_x = yy[y0:y1, x0:x1] - y0, xx[y0:y1, x0:x1] - x0
            m = sub_x * (y1 - y0) + sub_y * (x1 - x0) <= (x1 - x0) * (y1 - y0)
            if rng.random() < 0.5:
                m = m[:, ::-1]
            img[y0:y1, x0:x1][m] = g
    img = img + rng.normal(0, noise, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stereo_pair(seed, W=752, H=480, bf=47.9):
    """Left image + right image obtained by shifting horizontal bands by a per-band disparity bf/z."""
    rng = np.random.default_rng(seed + 7919)
    left = scene_image(seed, W + 64, H)
    right = np.empty((H, W), np.uint8)
    band = 0
    while band < H:
        bh = int(rng.integers(24, 96))
        z = rng.uniform(1.0, 15.0)
        d = int(round(bf / z))
        right[band:band + bh] = left[band:band + bh, d:d + W]
        band += bh
    return np.ascontiguousarray(left[:, :W]), right


def constant_image(value=128, W=752, H=480):
    return np.full((H, W), value, np.uint8)


def noise_image(seed, W=752, H=480):
    return np.random.default_rng(seed).integers(0, 256, (H, W)).astype(np.uint8)


def checkerboard(W=752, H=480, cell=16, lo=40, hi=200):
    yy, xx = np.mgrid[0:H, 0:W]
    return np.where(((yy // cell) + (xx // cell)) % 2 == 0, lo, hi).astype(np.uint8)
