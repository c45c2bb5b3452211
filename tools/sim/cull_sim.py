#!/usr/bin/env python
"""CPU simulation (numpy) of warp-level culling for the BOUNCE rays of Algorithm B: how often does at least one of a
warp's 32 lanes need a given sphere pair / a given bounding sphere of a small group?  Decides the layout of the
two-level uniform-operand selection loop (rt_trace.cuh, brute_select_pkc2).  Development aid, not product code."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ray_tracer_v1_b200 as rtb
from ray_tracer_v1_b200 import scenes

def hit_all(O, D, C, R):
    # O,D [m,3]; C [n,3]; R [n] -> t [m,n] (nan = miss), reference semantics (tca<0 miss)
    L = C[None] - O[:, None]
    tca = (L * D[:, None]).sum(-1)
    d2 = (L * L).sum(-1) - tca * tca
    disc = R[None] ** 2 - d2
    ok = (tca >= 0) & (disc >= 0)
    t = tca - np.sqrt(np.where(ok, disc, 0))
    return np.where(ok, t, np.nan), tca, disc

def ball_pass(O, D, C, R):
    # forward half-line intersects ball: disc>=0 and (tca>=0 or origin inside)
    L = C[None] - O[:, None]
    tca = (L * D[:, None]).sum(-1)
    ll = (L * L).sum(-1)
    disc = R[None] ** 2 - (ll - tca * tca)
    return (disc >= 0) & ((tca >= 0) | (ll <= R[None] ** 2))

def main(scene="complex", nwarps=3000, seed=0):
    spec = scenes.build_complex() if scene == "complex" else scenes.build_chandelier()
    fs = rtb.flatten_scene(spec.spheres, background_colour=spec.background)
    C, R, mat = fs.centre, fs.radius, fs.material
    n = len(R)
    rs = np.random.RandomState(seed)
    W, H, fov = 1920, 1080, 60.0
    aspect = W / H; hh = np.tan(np.radians(fov) / 2); hw = hh * aspect
    # warps: 2x2 pixels x 8 samples
    bx = rs.randint(0, W // 2, nwarps) * 2; by = rs.randint(0, H // 2, nwarps) * 2
    px = (bx[:, None] + np.tile(np.repeat([0, 1], 8), 2)[None, :] * 0 + np.array([0, 1, 0, 1]).repeat(8)[None]).astype(float)
    py = (by[:, None] + np.array([0, 0, 1, 1]).repeat(8)[None]).astype(float)
    m = nwarps * 32
    px = px.reshape(m); py = py.reshape(m)
    jx, jy = rs.random_sample(m), rs.random_sample(m)
    sx = (2 * (px + jx) / W - 1) * aspect * hw; sy = (1 - 2 * (py + jy) / H) * hh
    D = np.stack([sx, sy, -np.ones(m)], 1); D /= np.linalg.norm(D, axis=1, keepdims=True)
    O = np.tile(np.array(spec.camera, float), (m, 1))
    alive = np.ones(m, bool)
    small = R < 50
    order = np.argsort(-R)            # big first
    records = []
    for depth in range(spec.max_bounces):
        t, tca, disc = hit_all(O, D, C, R)
        key = np.abs(t)
        key[~alive] = np.nan
        idx = np.where(np.all(np.isnan(key), 1), -1, np.nanargmin(np.where(np.isnan(key), np.inf, key), 1))
        if depth >= 1:
            records.append((O.copy(), D.copy(), alive.copy()))
        hit = alive & (idx >= 0)
        tt = t[np.arange(m), np.maximum(idx, 0)]
        P = O + D * tt[:, None]
        N = (P - C[np.maximum(idx, 0)]) / R[np.maximum(idx, 0)][:, None]
        emis = mat[np.maximum(idx, 0), 2] != 0
        alive = hit & ~emis
        mirror = mat[np.maximum(idx, 0), 0] > spec.mirror_threshold
        r1, r2 = rs.random_sample(m), rs.random_sample(m)
        ct, st = np.sqrt(r1), np.sqrt(1 - r1); ph = 2 * np.pi * r2
        deg = np.abs(N[:, 2]) > 0.9
        tg = np.where(deg[:, None], np.array([1.0, 0, 0])[None], np.stack([-N[:, 1], N[:, 0], np.zeros(m)], 1))
        tg /= np.linalg.norm(tg, axis=1, keepdims=True)
        bt = np.cross(N, tg); bt /= np.linalg.norm(bt, axis=1, keepdims=True)
        dd = (st * np.cos(ph))[:, None] * tg + (st * np.sin(ph))[:, None] * bt + ct[:, None] * N
        dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        dm = D - 2 * (D * N).sum(1)[:, None] * N
        D = np.where(mirror[:, None], dm, dd)
        O = P + 0.001 * N
        O[~alive] = 0; D[~alive] = np.array([0, 0, -1.0])
    return spec, fs, records

def greedy_groups(C, R, idxs, k):
    """greedy spatial grouping of sphere indices into groups of k minimising bounding radius (approx)."""
    left = list(idxs); groups = []
    while left:
        # start with the sphere farthest from the centroid
        cen = C[left].mean(0)
        s = max(left, key=lambda i: np.linalg.norm(C[i] - cen))
        g = [s]; left.remove(s)
        while len(g) < k and left:
            def rad(cand):
                pts = g + [cand]
                c = C[pts].mean(0)
                return max(np.linalg.norm(C[i] - c) + R[i] for i in pts)
            b = min(left, key=rad); g.append(b); left.remove(b)
        groups.append(g)
    return groups

def bound(C, R, g):
    c = C[g].mean(0)
    # a few Ritter-ish refinements
    for _ in range(20):
        d = np.array([np.linalg.norm(C[i] - c) + R[i] for i in g]); j = int(np.argmax(d))
        if len(g) == 1: break
        dirv = C[g[j]] - c; nv = np.linalg.norm(dirv)
        if nv < 1e-12: break
        d2 = np.sort(d)[-2] if len(d) > 1 else d[j]
        c = c + dirv / nv * (d[j] - d2) * 0.5
    r = max(np.linalg.norm(C[i] - c) + R[i] for i in g)
    return c, r

if __name__ == "__main__":
    scene = sys.argv[1] if len(sys.argv) > 1 else "complex"
    spec, fs, recs = main(scene)
    C, R = fs.centre, fs.radius
    n = len(R)
    O = np.concatenate([r[0] for r in recs]); D = np.concatenate([r[1] for r in recs]); A = np.concatenate([r[2] for r in recs])
    nw = len(O) // 32
    t, tca, disc = hit_all(O, D, C, R)
    hit = ~np.isnan(t) & A[:, None]                      # [m,n]
    hw = hit.reshape(nw, 32, n).any(1)                   # warp needs sphere
    act = A.reshape(nw, 32).any(1)
    print(f"{scene}: {n} spheres, {nw} warp-trips, {act.mean():.3f} active; lanes alive {A.reshape(nw,32)[act].sum(1).mean():.1f}")
    hw = hw[act]
    big = [i for i in range(n) if R[i] >= 0.9]
    sm = [i for i in range(n) if R[i] < 0.9]
    print("per-sphere P(warp needs): big", np.round(hw[:, big].mean(0), 2))
    print("per-sphere P(warp needs): small mean %.3f max %.3f" % (hw[:, sm].mean(), hw[:, sm].mean(0).max()))
    Oa, Da = O, D
    for k in (2, 4, 8):
        gs = greedy_groups(C, R, sm, k)
        pm, pb, rad = [], [], []
        for g in gs:
            c, r = bound(C, R, g)
            pas = ball_pass(Oa, Da, c[None], np.array([r * 1.01 + 1e-3]))[:, 0] & A
            pb.append(pas.reshape(nw, 32).any(1)[act].mean())
            pm.append(hw[:, g].any(1).mean())
            rad.append(r)
        print(f"groups of {k}: {len(gs)} groups; P(any lane hits a member) mean {np.mean(pm):.3f}; P(any lane passes bound) mean {np.mean(pb):.3f} "
              f"(min {np.min(pb):.2f} max {np.max(pb):.2f}); mean bound radius {np.mean(rad):.2f}")
    # scene-order pairs (current kernel layout)
    pm = [hw[:, [i, min(i + 1, n - 1)]].any(1).mean() for i in range(0, n, 2)]
    print("scene-order pairs P(any lane hits a member):", np.round(pm, 2))
