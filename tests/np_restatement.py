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


# ---- Camera::update_matrices (scene/camera.rs:210-231) restated from cgmath 0.18 in numpy f32 ------------------
# (look_at_lh, perspective, cofactor invert).  Reproduces the matrices stored in the .rscn fixtures bit for bit
# (tests/test_scene_io.py checks the C++ twin against the same fixtures).
def _dot3(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _norm3(a):
    inv = f32(f32(1) / np.sqrt(_dot3(a, a)))
    return np.array([f32(x * inv) for x in a], f32)


def _cross(a, b):
    return np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
                     f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], f32)


def look_at_lh(eye, center, up):
    eye = np.asarray(eye, f32); center = np.asarray(center, f32); up = np.asarray(up, f32)
    d = -np.array([f32(center[i] - eye[i]) for i in range(3)], f32)
    fw = _norm3(d); s = _norm3(_cross(fw, up)); u = _cross(s, fw)
    m = np.zeros(16, f32)
    m[0:4] = [s[0], u[0], -fw[0], 0]; m[4:8] = [s[1], u[1], -fw[1], 0]; m[8:12] = [s[2], u[2], -fw[2], 0]
    m[12:16] = [-_dot3(eye, s), -_dot3(eye, u), _dot3(eye, fw), 1]
    return m


def perspective(fov_deg, aspect, near, far):
    import math
    rad = f32(f32(fov_deg) * f32(math.pi / 180.0))
    fv = f32(f32(1) / f32(math.tan(float(f32(rad / f32(2))))))
    near = f32(near); far = f32(far); aspect = f32(aspect)
    m = np.zeros(16, f32)
    m[0] = f32(fv / aspect); m[5] = fv; m[10] = f32(f32(far + near) / f32(near - far)); m[11] = -1
    m[14] = f32(f32(f32(f32(2) * far) * near) / f32(near - far))
    return m


def _det3(m):
    return f32(f32(f32(m[0][0] * f32(f32(m[1][1] * m[2][2]) - f32(m[2][1] * m[1][2]))) -
                   f32(m[1][0] * f32(f32(m[0][1] * m[2][2]) - f32(m[2][1] * m[0][2])))) +
               f32(m[2][0] * f32(f32(m[0][1] * m[1][2]) - f32(m[1][1] * m[0][2]))))


def invert4(flat):
    M = [[f32(flat[c * 4 + r]) for r in range(4)] for c in range(4)]
    d = []
    for skip in range(4):
        cols = [c for c in range(4) if c != skip]
        d.append(_det3([[M[c][k + 1] for c in cols] for k in range(3)]))
    det = f32(f32(f32(f32(M[0][0] * d[0]) - f32(M[1][0] * d[1])) + f32(M[2][0] * d[2])) - f32(M[3][0] * d[3]))
    inv = f32(f32(1) / det)
    t = [[M[r][c] for r in range(4)] for c in range(4)]
    out = []
    for i in range(4):
        for j in range(4):
            mat = [[c[k] for k in range(4) if k != j] for ci, c in enumerate(t) if ci != i]
            sign = f32(-1) if (i + j) & 1 else f32(1)
            out.append(f32(f32(_det3(mat) * sign) * inv))
    return np.array(out, f32)
