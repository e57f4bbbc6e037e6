"""Second, independent restatement of the reference's intersection and camera arithmetic in numpy float32
(vectorised; every numpy f32 operation is one IEEE rounding, no FMA).  TEST INFRASTRUCTURE: used only to
cross-check the C oracle, since the reference itself (Rust) cannot run here.

Follows /root/reference/src/renderer/cpu.rs:34-98 (hit_sphere, hit_cube), :199-202 and :234-251 (camera ray)."""
import numpy as np

f32 = np.float32


def dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def rust_min(a, b):
    # f32::min: NaN operand ignored; -0 < +0 (the oracle's and PTX's choice for the unspecified zero case)
    out = np.where(a < b, a, b)
    out = np.where(np.isnan(a), b, out)
    out = np.where(np.isnan(b), a, out)
    both_zero = (a == 0) & (b == 0)
    out = np.where(both_zero, np.where(np.signbit(a), a, b), out)
    return out.astype(f32)


def rust_max(a, b):
    out = np.where(a > b, a, b)
    out = np.where(np.isnan(a), b, out)
    out = np.where(np.isnan(b), a, out)
    both_zero = (a == 0) & (b == 0)
    out = np.where(both_zero, np.where(np.signbit(a), b, a), out)
    return out.astype(f32)


def hit_sphere(o, d, c, r):
    with np.errstate(all="ignore"):
        a = dot(d, d)
        k = dot(o, d) - dot(d, c)
        cc = dot(o, o) - f32(2.0) * dot(o, c) + dot(c, c) - r * r
        disc = k * k - a * cc
        sq = np.sqrt(disc)
        t1 = (-k - sq) / a
        t2 = (-k + sq) / a
        hit = ~(disc < 0) & ((t1 >= 0) | (t2 >= 0))
        t = np.where(t1 >= 0, t1, t2)
    return hit, np.where(hit, t, f32(0)).astype(f32)


def hit_cube(o, d, c, side):
    with np.errstate(all="ignore"):
        h = (side * f32(0.5))[..., None]
        mn = c - h
        mx = c + h
        t1 = (mn - o) / d
        t2 = (mx - o) / d
        tmin = rust_max(rust_max(rust_min(t1[..., 0], t2[..., 0]), rust_min(t1[..., 1], t2[..., 1])), rust_min(t1[..., 2], t2[..., 2]))
        tmax = rust_min(rust_min(rust_max(t1[..., 0], t2[..., 0]), rust_max(t1[..., 1], t2[..., 1])), rust_max(t1[..., 2], t2[..., 2]))
        hit = ~(tmax < 0) & ~(tmin > tmax)
        t = np.where(tmin < 0, tmax, tmin)
    return hit, np.where(hit, t, f32(0)).astype(f32)


def mat_vec4(m, v):
    # cgmath Matrix4 * Vector4: ((c0*v0 + c1*v1) + c2*v2) + c3*v3, m column-major flat[16]
    out = []
    for r in range(4):
        out.append(((m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3])
    return out


def camera_rays(width, height, inv_proj, inv_view):
    """All primary-ray directions, shape (H, W, 3)."""
    inv_proj = inv_proj.astype(f32); inv_view = inv_view.astype(f32)
    x = np.arange(width, dtype=f32)[None, :].repeat(height, 0)
    y = np.arange(height, dtype=f32)[:, None].repeat(width, 1)
    u = x / f32(width)
    v = f32(1.0) - y / f32(height)
    clip = [u * f32(2.0) - f32(1.0), v * f32(2.0) - f32(1.0), np.full_like(u, -1.0), np.full_like(u, -1.0)]
    cs = mat_vec4(inv_proj, clip)
    cs = [c / cs[3] for c in cs]
    ws = mat_vec4(inv_view, cs)
    w = np.stack(ws[:3], axis=-1).astype(f32)
    inv_len = f32(1.0) / np.sqrt(dot(w, w))
    return (-(w * inv_len[..., None])).astype(f32)


def first_hit(scene):
    """Brute-force nearest hit for every pixel; first minimum in object order wins (strict <)."""
    d = camera_rays(scene.width, scene.height, scene.inv_proj, scene.inv_view).reshape(-1, 3)
    o = np.broadcast_to(scene.cam_pos.astype(f32), d.shape)
    best_t = np.full(d.shape[0], np.inf, f32)
    best = np.full(d.shape[0], -1, np.int32)
    for i in range(scene.n_objects):
        c = np.broadcast_to(scene.geom[i, :3], d.shape)
        s = np.full(d.shape[0], scene.geom[i, 3], f32)
        hit, t = (hit_sphere if scene.kind[i] == 0 else hit_cube)(o, d, c, s)
        better = hit & ((best < 0) | (t < best_t))
        best = np.where(better, i, best)
        best_t = np.where(better, t, best_t)
    return best.reshape(scene.height, scene.width), np.where(best >= 0, best_t, 0).astype(f32).reshape(scene.height, scene.width)
