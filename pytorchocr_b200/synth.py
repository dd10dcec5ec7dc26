"""Seeded synthetic inputs for the post-processing hot path.

These stand in for the conv-net heads' outputs (there is no network for datasets/checkpoints):
DB probability maps, PSENet kernel logits, PAN++ text/kernel/embedding maps and CRNN softmax
rows, with the pathologies the parity tests need (holes, low-score regions, specks, 1-px runs,
border-touching regions, merged text masks, sub-min-area kernels, ratio-flagged kernels).
Host-side numpy/cv2 only; nothing here is on the timed path.
"""
import math

import cv2
import numpy as np

BASE_SEED = 20221001


def _place_rects(rng, H, W, n, hh_rng=(5, 10), hw_rng=(15, 37), max_angle=20.0, margin=4, tries=40, occ=None):
    """Non-overlapping rotated rectangles (cx, cy, hw, hh, angle_deg); rejection-sampled on an
    occupancy grid (optionally shared between calls) so that neighbours stay >= margin px apart."""
    if occ is None:
        occ = np.zeros((H, W), np.uint8)
    rects = []
    for _ in range(n * tries):
        if len(rects) >= n:
            break
        hh = int(rng.integers(hh_rng[0], hh_rng[1] + 1))
        hw = int(rng.integers(hw_rng[0], hw_rng[1] + 1))
        ang = float(rng.uniform(-max_angle, max_angle))
        r = math.hypot(hw, hh) + margin + 2
        if W - r <= r or H - r <= r:
            continue
        cx = float(rng.uniform(r, W - r))
        cy = float(rng.uniform(r, H - r))
        box = cv2.boxPoints(((cx, cy), (2.0 * (hw + margin), 2.0 * (hh + margin)), ang))
        x0, y0 = np.floor(box.min(0)).astype(int)
        x1, y1 = np.ceil(box.max(0)).astype(int)
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, W - 1), min(y1, H - 1)
        m = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
        cv2.fillPoly(m, [np.round(box - [x0, y0]).astype(np.int32)], 1)
        if (occ[y0:y1 + 1, x0:x1 + 1] & m).any():
            continue
        occ[y0:y1 + 1, x0:x1 + 1] |= m
        rects.append((cx, cy, hw, hh, ang))
    return rects


def _fill_rect(img, rect, value, shrink=1.0):
    cx, cy, hw, hh, ang = rect
    box = cv2.boxPoints(((cx, cy), (2.0 * hw * shrink, 2.0 * hh * shrink), ang))
    cv2.fillPoly(img, [np.round(box).astype(np.int32)], value)


def db_map(seed, H=736, W=1280, n_regions=200, dtype=np.float32):
    """One DB/DB++ probability map [H,W] (SURVEY §8(d) Cfg 2 generator): background U(0,0.05);
    ~n_regions rotated rectangles at 0.9, blurred sigma=1, + U(-0.015,0.015) noise, clamp [0,1];
    5% regions with a 3x3..5x5 hole at 0.05; 5% low-score regions (core 0.4); ~20 specks;
    2 straight 1-px runs; 2 regions touching the image border."""
    rng = np.random.default_rng(seed)
    scale = (H * W) / float(736 * 1280)
    n = max(1, int(round(n_regions * min(1.0, scale * 1.0)))) if scale < 1 else n_regions
    img = np.zeros((H, W), np.float32)
    rects = _place_rects(rng, H, W, n)
    holes = []
    for r in rects:
        u = rng.random()
        if u < 0.05:
            _fill_rect(img, r, 0.9)
            holes.append(r)
        elif u < 0.10:
            _fill_rect(img, r, 0.4)
        else:
            _fill_rect(img, r, 0.9)
    # two regions touching the border (axis-aligned, clipped by the frame)
    if H >= 64 and W >= 128:
        img[0:9, W // 3:W // 3 + 50] = 0.9
        img[H // 2:H // 2 + 14, W - 30:W] = 0.9
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    # holes are punched after the blur so that they stay crisp (enclosed background)
    for (cx, cy, hw, hh, ang) in holes:
        k = int(rng.integers(3, 6))
        x0, y0 = int(cx) - k // 2, int(cy) - k // 2
        img[y0:y0 + k, x0:x0 + k] = 0.05
    bg = rng.uniform(0.0, 0.05, size=(H, W)).astype(np.float32)
    img = np.maximum(img, bg)
    img += rng.uniform(-0.015, 0.015, size=(H, W)).astype(np.float32)
    # specks (1-2 px) and straight 1-px runs: exercised by the "<= 2 contour points" rule
    for _ in range(int(round(20 * min(1.0, scale))) or 2):
        y, x = int(rng.integers(2, H - 2)), int(rng.integers(2, W - 3))
        if img[y - 2:y + 3, x - 2:x + 4].max() < 0.2:
            img[y, x] = 0.8
            if rng.random() < 0.5:
                img[y, x + 1] = 0.8
    for _ in range(2):
        y, x = int(rng.integers(2, H - 2)), int(rng.integers(2, max(3, W - 20)))
        ln = int(rng.integers(5, 12))
        if x + ln + 2 < W and img[y - 2:y + 3, x - 2:x + ln + 2].max() < 0.2:
            img[y, x:x + ln] = 0.8
    np.clip(img, 0.0, 1.0, out=img)
    return img.astype(dtype)


def db_batch(n, seed=BASE_SEED, H=736, W=1280, n_regions=200, dtype=np.float32):
    """[n,1,H,W]; image i uses seed+i."""
    out = np.empty((n, 1, H, W), dtype)
    for i in range(n):
        out[i, 0] = db_map(seed + i, H, W, n_regions, dtype)
    return out


def pse_scene(seed, H=736, W=1280, n_regions=200, K=7, hh_rng=(5, 10), hw_rng=(15, 37), n_abs=None):
    """Geometry of one PSENet scene: masks u8 [K,H,W] (channel k = rectangles shrunk to 1 - 0.1k),
    `weak` u8 [H,W] (regions whose text logit is only +1) and the rng to continue with."""
    rng = np.random.default_rng(seed)
    scale = (H * W) / float(736 * 1280)
    n = n_regions if scale >= 1 else max(2, int(round(n_regions * scale)))
    n = n_abs if n_abs is not None else n
    rects = _place_rects(rng, H, W, n, hh_rng, hw_rng, margin=6)
    masks = np.zeros((K, H, W), np.uint8)
    weak = np.zeros((H, W), np.uint8)
    for r in rects:
        cx, cy, hw, hh, ang = r
        u = rng.random()
        group = [r]
        if u < 0.20:
            # touching twin: same orientation, shifted along the long axis so the full-size
            # text rectangles overlap by ~2 px while the shrunk kernels stay apart
            a = math.radians(ang)
            d = 2 * hw - 2
            group.append((cx + d * math.cos(a), cy + d * math.sin(a), hw, hh, ang))
        tiny = 0.25 <= u < 0.30
        for g in group:
            for k in range(K):
                s = 1.0 - 0.1 * k
                if tiny and k == K - 1:
                    gx, gy = int(g[0]), int(g[1])
                    if 2 <= gy < H - 2 and 2 <= gx < W - 3:
                        masks[k, gy:gy + 2, gx:gx + 3] = 1  # 6 px < min_area 16
                    continue
                _fill_rect(masks[k], g, 1, s)
            if 0.20 <= u < 0.25:
                _fill_rect(weak, g, 1)
    return masks, weak, rng


def pse_maps(seed, H=736, W=1280, n_regions=200, K=7, hh_rng=(5, 10), hw_rng=(15, 37), n_abs=None):
    """PSENet logits [K,H,W] at processing resolution (SURVEY §8(d) Cfg 3 generator): channel k is
    the rectangle shrunk to (1 - 0.1k); +4 inside / -4 outside + N(0,0.5); 20% of regions placed
    as touching pairs (merged text masks => contested expansion); 5% with text logit +1
    (score 0.73 => rejected); 5% whose smallest kernel is < 16 px (seed dropped)."""
    masks, weak, rng = pse_scene(seed, H, W, n_regions, K, hh_rng, hw_rng, n_abs)
    logits = np.where(masks > 0, 4.0, -4.0).astype(np.float32)
    logits += rng.normal(0.0, 0.5, size=logits.shape).astype(np.float32)
    # low text score regions: logit +1 on the text channel only
    t = logits[0]
    t[(weak > 0) & (masks[0] > 0)] = 1.0 + 0.1 * rng.normal(size=int(((weak > 0) & (masks[0] > 0)).sum())).astype(np.float32)
    return logits


def pse_maps_torch(scenes, seed, device):
    """Same distribution as pse_maps for a batch of scenes [(masks, weak), ...], with the noise drawn
    on `device` (bench inputs: 26 MB of normals per image are too slow to draw on the host)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    masks = torch.from_numpy(np.stack([sc[0] for sc in scenes])).to(device)
    weak = torch.from_numpy(np.stack([sc[1] for sc in scenes])).to(device)
    logits = torch.where(masks > 0, 4.0, -4.0).to(torch.float32)
    logits += 0.5 * torch.randn(logits.shape, generator=g, device=device, dtype=torch.float32)
    sel = (weak > 0) & (masks[:, 0] > 0)
    t = logits[:, 0]
    t[sel] = (1.0 + 0.1 * torch.randn(t.shape, generator=g, device=device, dtype=torch.float32))[sel]
    return logits


def pan_scene(seed, H=736, W=1280, n_regions=200, hh_rng=(5, 10), hw_rng=(15, 37), n_abs=None):
    """Geometry of one PAN++ scene: text u8 [H,W], kernel u8 [H,W], instance id i32 [H,W], embedding
    centres f32 [n_inst+1,4] and the rng to continue with."""
    rng = np.random.default_rng(seed)
    scale = (H * W) / float(736 * 1280)
    n = n_regions if scale >= 1 else max(2, int(round(n_regions * scale)))
    n = n_abs if n_abs is not None else n
    occ = np.zeros((H, W), np.uint8)
    # ratio-flag scenes first (they are long): up to 4 per full-size image, non-overlapping with the rest
    longs = _place_rects(rng, H, W, 4, (14, 14), (90, 90), margin=6, occ=occ) if (H >= 200 and W >= 400) else []
    rects = _place_rects(rng, H, W, n, hh_rng, hw_rng, margin=6, occ=occ)
    text = np.zeros((H, W), np.uint8)
    kern = np.zeros((H, W), np.uint8)
    inst = np.zeros((H, W), np.int32)
    centres = [np.zeros(4, np.float32)]

    def centre(iid):   # embedding centres on a lattice of spacing 6 in 4-d => pairwise distance >= 6
        return np.array([(iid % 5), (iid // 5) % 5, (iid // 25) % 5, (iid // 125) % 5], np.float32) * 6.0

    for (bx, by, _, _, ang) in longs:
        # pa.pyx:42-54: one long text region holding a >= 3073-px kernel and a 3-px kernel blob
        # (3 >= min_kernel_area 2.6 survives; 3 * 1024 < big area => both flagged)
        iid = len(centres)
        centres.append(centre(iid))
        a = math.radians(ang)
        ca, sa = math.cos(a), math.sin(a)
        mb = np.zeros((H, W), np.uint8)
        _fill_rect(mb, (bx, by, 90, 14, ang), 1)
        text |= mb
        inst[mb > 0] = iid
        _fill_rect(kern, (bx - 10 * ca, by - 10 * sa, 75, 12, ang), 1, 0.95)
        qx, qy = int(round(bx + 80 * ca)), int(round(by + 80 * sa))
        kern[qy, qx - 1:qx + 2] = 1
    for r in rects:
        cx, cy, hw, hh, ang = r
        u = rng.random()
        group = [r]
        if u < 0.20:
            a = math.radians(ang)
            d = 2 * hw - 2
            group.append((cx + d * math.cos(a), cy + d * math.sin(a), hw, hh, ang))
        for g in group:
            iid = len(centres)
            centres.append(centre(iid))
            m = np.zeros((H, W), np.uint8)
            _fill_rect(m, g, 1)
            text |= m
            inst[m > 0] = iid
            _fill_rect(kern, g, 1, 0.5)
    kern &= text
    return text, kern, inst, np.stack(centres), rng


def pan_maps(seed, H=736, W=1280, n_regions=200, hh_rng=(5, 10), hw_rng=(15, 37), n_abs=None):
    """PAN++ maps [6,H,W] (SURVEY §8(d) Cfg 4 generator): ch0 text logit, ch1 kernel logit
    (half-sizes x0.5), ch2-5 embedding = per-instance centre (pairwise distance >= 6) + N(0,0.25);
    no pixel within 0.2 of the distance-3 gate; a few long text regions per image hold a >= 3073-px
    kernel next to a 3-px kernel blob so that the area-ratio flag and the embedding gate fire."""
    text, kern, inst, C, rng = pan_scene(seed, H, W, n_regions, hh_rng, hw_rng, n_abs)
    out = np.empty((6, H, W), np.float32)
    out[0] = np.where(text > 0, 4.0, -4.0) + rng.normal(0, 0.5, (H, W))
    out[1] = np.where(kern > 0, 4.0, -4.0) + rng.normal(0, 0.5, (H, W))
    emb = C[inst].transpose(2, 0, 1) + rng.normal(0, 0.25, (4, H, W)).astype(np.float32)
    # enforce the gate margin: push pixels whose distance to their own centre is within 0.2 of 3 back
    # onto the centre (cheap and sufficient: gates compare against kernel means ~ centres)
    flat = emb.reshape(4, -1).T
    d_own = np.linalg.norm(flat - C[inst.reshape(-1)], axis=1)
    bad = np.abs(d_own - 3.0) <= 0.2
    flat[bad] = C[inst.reshape(-1)][bad]
    out[2:] = flat.T.reshape(4, H, W)
    return out.astype(np.float32)


def pan_maps_torch(scenes, seed, device):
    """Same distribution as pan_maps for a batch of scenes [(text, kern, inst, centres), ...] with the
    noise drawn on `device` (bench inputs)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    N = len(scenes)
    H, W = scenes[0][0].shape
    out = torch.empty((N, 6, H, W), dtype=torch.float32, device=device)
    for i, (text, kern, inst, C) in enumerate(scenes):
        t = torch.from_numpy(text).to(device) > 0
        k = torch.from_numpy(kern).to(device) > 0
        out[i, 0] = torch.where(t, 4.0, -4.0) + 0.5 * torch.randn((H, W), generator=g, device=device)
        out[i, 1] = torch.where(k, 4.0, -4.0) + 0.5 * torch.randn((H, W), generator=g, device=device)
        cen = torch.from_numpy(C).to(device)[torch.from_numpy(inst).to(device).long()]      # [H,W,4]
        noise = 0.25 * torch.randn((H, W, 4), generator=g, device=device)
        bad = (noise.norm(dim=2) - 3.0).abs() <= 0.2
        noise[bad] = 0.0
        out[i, 2:] = (cen + noise).permute(2, 0, 1)
    return out


def char_dict(n=6623):
    """A synthetic dictionary of n distinct single characters (the reference's
    char_dict_6623.txt is data under /root/reference and is not copied)."""
    chars = []
    cp = 0x4E00
    while len(chars) < n:
        chars.append(chr(cp))
        cp += 1
    return chars


def write_char_dict(path, n=6623):
    with open(path, "wb") as f:
        f.write("\n".join(char_dict(n)).encode("utf-8"))
    return path


def ctc_probs_numpy(seed, T=80, B=64, C=6623):
    """Softmax rows [T,B,C] f32 (SURVEY §8(d) Cfg 5 generator): logits N(0,1) + 12 at a target class
    per step (blank w.p. 0.5, repeat-previous w.p. 0.3, else uniform in [1,C))."""
    rng = np.random.default_rng(seed)
    logits = rng.normal(0, 1, (T, B, C)).astype(np.float32)
    u = rng.random((T, B))
    tgt = rng.integers(1, C, (T, B))
    target = np.zeros((T, B), np.int64)
    for t in range(T):
        cur = np.where(u[t] < 0.5, 0, tgt[t])
        if t > 0:
            rep = (u[t] >= 0.5) & (u[t] < 0.8)
            cur = np.where(rep, target[t - 1], cur)
        target[t] = cur
    tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing="ij")
    logits[tt, bb, target] += 12.0
    logits -= logits.max(axis=2, keepdims=True)
    e = np.exp(logits)
    return (e / e.sum(axis=2, keepdims=True)).astype(np.float32), target


def ctc_probs_torch(seed, T, B, C, device, chunk=512):
    """Same distribution generated on `device` with torch, in chunks over B (for the 17 GB shard)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    out = torch.empty((T, B, C), dtype=torch.float32, device=device)
    prev = None
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        logits = torch.randn((T, nb, C), generator=g, device=device, dtype=torch.float32)
        u = torch.rand((T, nb), generator=g, device=device)
        tgt = torch.randint(1, C, (T, nb), generator=g, device=device)
        target = torch.zeros((T, nb), dtype=torch.long, device=device)
        for t in range(T):
            cur = torch.where(u[t] < 0.5, torch.zeros_like(tgt[t]), tgt[t])
            if t > 0:
                rep = (u[t] >= 0.5) & (u[t] < 0.8)
                cur = torch.where(rep, target[t - 1], cur)
            target[t] = cur
        logits.scatter_add_(2, target.unsqueeze(-1), torch.full((T, nb, 1), 12.0, device=device))
        out[:, b0:b1] = torch.softmax(logits, dim=2)
        del logits
    return out


# ------------------------------------------------------------------------------------------------
# pages + detected boxes for the text-line crop step
# ------------------------------------------------------------------------------------------------
def page_image(seed, H=736, W=1280, C=3):
    """uint8 page with structure at every scale (so that a wrong tap or weight changes the result)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 127 + 80 * np.sin(xx[..., None] * (0.05 + 0.03 * np.arange(C)) + yy[..., None] * 0.07)
    img = base + rng.integers(-47, 48, (H, W, C))
    return np.clip(img, 0, 255).astype(np.uint8)


def page_boxes(seed, n=200, H=736, W=1280, tall_frac=0.1, skew=2.0, scale=1.0):
    """int16 [n,4,2] text-line quads as a detector emits them (TL,TR,BR,BL, inside the page): rotated
    rectangles, a share of them tall (rot90 rule), corners jittered by `skew` px so that the transform is a
    true perspective map; clustered first-corner rows exercise the swap pass of sort_boxes."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4, 2), np.int16)
    rows = rng.integers(20, H - 20, max(1, n // 6))
    for i in range(n):
        hw, hh = rng.uniform(15, 60) * scale, rng.uniform(5, 14) * scale
        if rng.random() < tall_frac:
            hw, hh = hh, hw * 0.8
        a = rng.uniform(-0.35, 0.35) if rng.random() < 0.8 else 0.0
        cx = rng.uniform(hw + hh + 4, W - hw - hh - 4)
        cy = float(np.clip(rows[rng.integers(len(rows))] + rng.uniform(-6, 6), hw + hh + 4, H - hw - hh - 4))
        c, s_ = np.cos(a), np.sin(a)
        q = np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]])
        q = q @ np.array([[c, s_], [-s_, c]]) + [cx, cy] + rng.uniform(-skew, skew, (4, 2))
        out[i] = np.clip(np.round(q), 0, [W, H]).astype(np.int16)
    return out
