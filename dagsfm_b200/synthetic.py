"""Synthetic inputs of the hot path, shared by bench.py, __graft_entry__.smoke() and the tests.

`make_image_collection` is the C3 workload of SURVEY 8d / BASELINE.json configs[2]: a sequence of images of one long
3-D scene in which neighbouring images overlap, so that descriptor matching (sift.cc:76-198) of a candidate pair yields
correspondences that follow one relative pose and two-view verification (two_view_geometry.cc:292-489) has real work
to do on them -- both stages run on the SAME images, which the separate per-stage generators (random descriptors for
matching, pre-matched point pairs for verification) cannot give.

Layout of the scene: scene point k sits at x = k * dx with random height and depth (every seventh stretch of the scene
is a plane, so planar / panoramic configurations occur); image i sees the `shared` points of the window starting at
i * stride and `n_kp - shared` clutter keypoints with random descriptors.  Descriptors follow the reference's own test
recipe (sift_test.cc:243-253: squared uniform, L2-normalise, round(512 x), saturate) plus per-view noise.
"""
from __future__ import annotations

import numpy as np


def make_image_collection(n_img: int, n_kp: int, seed: int = 0, device="cuda", shared_frac: float = 0.5,
                          overlap_images: int = 50, width: int = 1000, kp_noise_px: float = 0.5,
                          desc_noise: float = 0.03 / 5.06, distractor_frac: float = 0.125):
    """-> dict(desc uint8 [n_img, n_kp, 128] (torch, on `device`), keypoints float64 [n_img, n_kp, 2] (numpy),
    cam_params (f, cx, cy, k) of the shared SIMPLE_RADIAL camera, prior bool [n_img], stride, shared).
    Images i and i + d share shared - d * stride scene points (none from d = overlap_images on).  `distractor_frac` of an
    image's keypoints carry descriptors of a second sliding pool (repeated texture) at random positions: they match
    across images like scene points do but follow no geometry -- the outlier matches verification exists to reject."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    shared = int(n_kp * shared_frac)
    stride = max(shared // overlap_images, 1)
    pool = stride * n_img + shared
    f, c = 1.2 * width, width / 2.0
    dx = 4.0 / shared                                   # a window spans 4 scene units at depth ~8: inside the field of view
    k = torch.arange(pool, device=device, dtype=torch.float64)
    X = torch.stack([k * dx,
                     (torch.rand(pool, generator=g, device=device, dtype=torch.float64) * 2 - 1) * 1.5,
                     6.0 + 4.0 * torch.rand(pool, generator=g, device=device, dtype=torch.float64)], 1)
    planar = ((k // (4 * shared)) % 7) == 6
    X[planar, 2] = 8.0 + 0.05 * (X[planar, 0] % 4.0)
    base = torch.rand((pool, 128), generator=g, device=device) ** 2
    base = base / base.norm(dim=1, keepdim=True)
    n_dis = int(n_kp * distractor_frac)
    dstride = max(n_dis // overlap_images, 1) if n_dis else 0
    dbase = torch.rand((dstride * n_img + n_dis, 128), generator=g, device=device) ** 2
    dbase = dbase / dbase.norm(dim=1, keepdim=True).clamp_min(1e-12)
    desc = torch.empty((n_img, n_kp, 128), dtype=torch.uint8, device=device)
    kps = torch.empty((n_img, n_kp, 2), dtype=torch.float64, device=device)
    ang = (torch.rand((n_img, 3), generator=g, device=device, dtype=torch.float64) * 2 - 1) * 0.08
    cen = torch.stack([(torch.arange(n_img, device=device, dtype=torch.float64) * stride + shared / 2.0) * dx,
                       (torch.rand(n_img, generator=g, device=device, dtype=torch.float64) * 2 - 1) * 0.3,
                       (torch.rand(n_img, generator=g, device=device, dtype=torch.float64) * 2 - 1) * 0.3], 1)
    cx, sx, cy, sy, cz, sz = ang[:, 0].cos(), ang[:, 0].sin(), ang[:, 1].cos(), ang[:, 1].sin(), ang[:, 2].cos(), ang[:, 2].sin()
    one, zero = torch.ones_like(cx), torch.zeros_like(cx)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).reshape(-1, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).reshape(-1, 3, 3)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).reshape(-1, 3, 3)
    R = Rz @ Ry @ Rx
    for i in range(n_img):
        w = slice(i * stride, i * stride + shared)
        Xc = (X[w] - cen[i]) @ R[i].T
        uv = f * Xc[:, :2] / Xc[:, 2:3] + c + kp_noise_px * torch.randn((shared, 2), generator=g, device=device, dtype=torch.float64)
        clutter = torch.rand((n_kp - shared, 2), generator=g, device=device, dtype=torch.float64) * width
        v = (base[w] + desc_noise * torch.randn((shared, 128), generator=g, device=device)).clamp_min(0)
        dv = (dbase[i * dstride:i * dstride + n_dis] + desc_noise * torch.randn((n_dis, 128), generator=g, device=device)).clamp_min(0)
        r = torch.rand((n_kp - shared - n_dis, 128), generator=g, device=device) ** 2
        d = torch.cat([v, dv, r])
        d = d / d.norm(dim=1, keepdim=True)
        d = torch.round(512.0 * d).clamp(0, 255).to(torch.uint8)
        perm = torch.randperm(n_kp, generator=g, device=device)
        desc[i] = d[perm]
        kps[i] = torch.cat([uv, clutter])[perm]
    prior = ((np.arange(n_img) // 3) % 2) == 0
    return {"desc": desc, "keypoints": kps.cpu().numpy(), "cam_params": (f, c, c, 0.0), "prior": prior,
            "stride": stride, "shared": shared, "width": width}


def candidate_pairs(n_img: int, per_image: int) -> np.ndarray:
    """The pair list a retrieval stage hands to the matcher: every image with its `per_image` successors in the
    sequence (the reference caps vocab-tree candidates per image the same way, graph/similarity_graph.h:52-54), sorted by
    the first image -- the locality order in which SiftFeatureMatcher::Match is fed (feature/matching.cc:1087-1105)."""
    i = np.repeat(np.arange(n_img), per_image)
    j = i + np.tile(np.arange(1, per_image + 1), n_img)
    ok = j < n_img
    return np.ascontiguousarray(np.stack([i[ok], j[ok]], 1).astype(np.uint32))


def make_vocabulary_device(desc, n_words: int, n_train: int = 1 << 20, seed: int = 0):
    """A synthetic vocabulary for a collection resident on the GPU (torch uint8 [..., 128]); same recipe as
    retrieval.make_vocabulary (words = sampled descriptors, projection = 64 rows of a random orthogonal matrix,
    thresholds = per-word medians of the projected training descriptors, embedding only for words with >= 5 of them;
    inverted_index.h:173-227, inverted_file.h:286-303) with the training assignment done by torch on the device.
    DATA GENERATOR for tests and benchmarks -- the reference loads pre-trained trees from disk."""
    import torch
    from .retrieval import Vocabulary
    d = desc.reshape(-1, 128)
    dev = d.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    words = d[torch.randperm(len(d), generator=g, device=dev)[:n_words]].contiguous()
    if len(words) < n_words:
        words = words[torch.randint(len(words), (n_words,), generator=g, device=dev)]
    q, _ = torch.linalg.qr(torch.randn((128, 128), generator=g, device=dev, dtype=torch.float64))
    proj = q.T[:64].to(torch.float32).contiguous()
    train = d[torch.randperm(len(d), generator=g, device=dev)[:n_train]]
    wf = words.to(torch.float32)
    wsq = (wf * wf).sum(1)
    assign = torch.empty(len(train), dtype=torch.int64, device=dev)
    for a in range(0, len(train), 8192):                      # integer-valued fp32 products < 2^24: exact
        x = train[a:a + 8192].to(torch.float32)
        assign[a:a + 8192] = torch.argmin(wsq[None, :] - 2.0 * (x @ wf.T), dim=1)
    pd = train.to(torch.float32) @ proj.T                     # [n_train, 64]
    cnt = torch.bincount(assign, minlength=n_words)
    start = torch.cumsum(cnt, 0) - cnt
    thr = torch.zeros((n_words, 64), dtype=torch.float32, device=dev)
    has = (cnt >= 5)
    lo = (start + (cnt - 1).clamp_min(0) // 2).clamp_max(max(len(train) - 1, 0))
    hi = (start + cnt // 2).clamp_max(max(len(train) - 1, 0))
    for j in range(64):
        o1 = torch.argsort(pd[:, j], stable=True)
        o2 = torch.argsort(assign[o1], stable=True)           # grouped by word, ascending value inside a word
        v = pd[o1[o2], j]
        thr[:, j] = torch.where(has, 0.5 * (v[lo] + v[hi]), torch.zeros((), device=dev))
    return Vocabulary(words.cpu().numpy(), proj.cpu().numpy(), thr.cpu().numpy(), has.to(torch.uint8).cpu().numpy())
