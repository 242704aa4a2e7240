"""Procedural stand-ins for the scenes BASELINE.json names but the reference does not ship (only the Cornell
Box is in Assets/): a "Sponza-class" colonnaded atrium (config C4) and a "Subway-class" station tunnel (config C5).

They are assembled through the same SceneBuilder the glTF converter uses, so they arrive at the library as the reference's
flat buffers (Vertex / Material / RT::MeshInstance / RT::EmissiveTriangle): left-handed, +Y up, +Z forward,
clockwise front faces (cross(v1 - v0, v2 - v0) along the vertex normal), one instance per mesh primitive, repeated
geometry as several instances of one vertex / index range with quantised TRS transforms.

`detail` scales the tessellation (triangle count ~ detail^2): detail = 1.0 gives the sizes SURVEY 8d asks for
(atrium ~ 262 k triangles with >= 13 107 emissive ones so the host enables presampled sets and the light voxel grid;
tunnel ~ 1.2 M triangles with glass, glossy metal and emissive tubes), small values give scenes the brute-force oracle can
trace in seconds. Everything is deterministic (no RNG besides a fixed-seed generator for colours / placements)."""
import numpy as np

from .scene import SceneBuilder, make_material

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# mesh primitives: (positions, normals, uvs, indices) with cross(e1, e2) . n > 0
# ------------------------------------------------------------------------------------------------
def _orient(pos, nrm, idx):
    """Flip the winding of the whole mesh if its first non-degenerate triangle disagrees with the vertex normal."""
    t = idx.reshape(-1, 3)
    c = np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]])
    d = np.einsum("ij,ij->i", c, nrm[t[:, 0]])
    k = np.argmax(np.abs(d))
    if d[k] < 0:
        t = t[:, [0, 2, 1]]
    return np.ascontiguousarray(t).reshape(-1).astype(np.uint32)


def patch(f, nu, nv, flip=False, uv_scale=(1.0, 1.0), toward=None, away=None):
    """Parametric surface f(U, V) -> (..., 3) over [0,1]^2 with nu x nv quads; normals from central differences.
    toward / away: a point (or an array of points shaped like the surface) the front faces must look at / away from --
    single-sided materials are invisible from behind (GetMaterialData, RayQuery.hlsli:452-470), so every mesh states
    which way it faces instead of relying on the parametrisation's handedness."""
    nu = max(int(nu), 1); nv = max(int(nv), 1)
    u = np.linspace(0.0, 1.0, nu + 1); v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="ij")
    P = f(U, V)
    e = 1e-4
    du = f(U + e, V) - f(U - e, V)
    dv = f(U, V + e) - f(U, V - e)
    N = np.cross(du, dv)
    ln = np.linalg.norm(N, axis=-1, keepdims=True)
    ln[ln == 0] = 1.0
    N = N / ln
    if flip:
        N = -N
    if toward is not None or away is not None:
        ref = np.asarray(toward if toward is not None else away, dtype=np.float64)
        d = np.einsum("...k,...k->...", N, ref - P).mean()
        if (d < 0) == (toward is not None):
            N = -N
    i = (np.arange(nu)[:, None] * (nv + 1) + np.arange(nv)[None, :]).reshape(-1)
    quads = np.stack([i, i + nv + 1, i + nv + 2, i, i + nv + 2, i + 1], axis=1).reshape(-1)
    pos = P.reshape(-1, 3).astype(F32); nrm = N.reshape(-1, 3).astype(F32)
    uv = np.stack([U * uv_scale[0], V * uv_scale[1]], axis=-1).reshape(-1, 2).astype(F32)
    return pos, nrm, uv, _orient(pos.astype(np.float64), nrm.astype(np.float64), quads)


def box(size=(1, 1, 1), center=(0, 0, 0)):
    sx, sy, sz = [0.5 * s for s in size]
    pos, nrm, uv, idx = [], [], [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            n = np.zeros(3); n[axis] = sgn
            a = np.zeros(3); b = np.zeros(3)
            a[(axis + 1) % 3] = 1.0; b[(axis + 2) % 3] = 1.0
            h = np.array([sx, sy, sz])
            c = n * h
            quad = [c - a * h - b * h, c + a * h - b * h, c + a * h + b * h, c - a * h + b * h]
            base = len(pos)
            pos += quad; nrm += [n] * 4
            uv += [[0, 0], [1, 0], [1, 1], [0, 1]]
            idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    pos = np.array(pos) + np.asarray(center, dtype=np.float64)
    nrm = np.array(nrm)
    idx = np.array(idx, dtype=np.uint32)
    # orient each face on its own (faces alternate handedness)
    t = idx.reshape(-1, 3)
    c = np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]])
    bad = np.einsum("ij,ij->i", c, nrm[t[:, 0]]) < 0
    t[bad] = t[bad][:, [0, 2, 1]]
    return pos.astype(F32), nrm.astype(F32), np.array(uv, dtype=F32), t.reshape(-1).astype(np.uint32)


def cylinder(radius, height, nseg, nring=1, profile=None, flutes=0, flute_depth=0.0):
    """Open tube around +Y from y = 0 to height; profile(v) scales the radius along the axis, flutes ripple it around."""
    def f(U, V):
        ang = 2.0 * np.pi * U
        r = radius * (profile(V) if profile is not None else 1.0)
        if flutes:
            r = r * (1.0 - flute_depth * (0.5 + 0.5 * np.cos(flutes * ang)))
        return np.stack([r * np.cos(ang), V * height, r * np.sin(ang)], axis=-1)
    pos, nrm, uv, idx = patch(f, nseg, nring)
    # make the normals point away from the axis
    out = pos.copy(); out[:, 1] = 0
    if np.einsum("ij,ij->i", out, nrm).sum() < 0:
        nrm = -nrm
        idx = idx.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1).astype(np.uint32)
    return pos, nrm, uv, idx


def sphere(radius, nseg, nring):
    def f(U, V):
        th = 2.0 * np.pi * U
        ph = np.pi * (0.02 + 0.96 * V)          # leave tiny polar caps open: no degenerate triangles
        return np.stack([radius * np.sin(ph) * np.cos(th), radius * np.cos(ph), radius * np.sin(ph) * np.sin(th)], axis=-1)
    pos, nrm, uv, idx = patch(f, nseg, nring)
    if np.einsum("ij,ij->i", pos, nrm).sum() < 0:
        nrm = -nrm
        idx = idx.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1).astype(np.uint32)
    return pos, nrm, uv, idx


def _quat_y(angle):
    return (0.0, float(np.sin(0.5 * angle)), 0.0, float(np.cos(0.5 * angle)))


def _quat_axis(axis, angle):
    a = np.asarray(axis, dtype=np.float64); a = a / np.linalg.norm(a)
    s = np.sin(0.5 * angle)
    return (float(a[0] * s), float(a[1] * s), float(a[2] * s), float(np.cos(0.5 * angle)))


def _n(detail, base, lo=1):
    return max(int(round(base * detail)), lo)


# ------------------------------------------------------------------------------------------------
# "Sponza-class" atrium (config C4: ReSTIR GI + light voxel grid)
# ------------------------------------------------------------------------------------------------
ATRIUM_CAMERA = (0.0, 1.7, -13.0)


def atrium(detail=1.0, lamp_tris=None):
    """Two-storey colonnaded hall, x in [-9, 9], y in [0, 10], z in [-16, 16], closed (no sky), lit by strings of small
    emissive lanterns (>= 13 107 emissive triangles at detail 1) plus two emissive ceiling panels.
    lamp_tris: override the tessellation of one lantern (triangles), to force many lights in a small scene."""
    rng = np.random.default_rng(20240611)
    b = SceneBuilder()
    M = {}
    M["stone"] = b.add_material(make_material(base_color=(0.62, 0.58, 0.52), roughness=0.7))
    M["floor"] = b.add_material(make_material(base_color=(0.45, 0.43, 0.42), roughness=0.25, coat_weight=0.6, coat_roughness=0.08))
    M["brick"] = b.add_material(make_material(base_color=(0.55, 0.30, 0.24), roughness=0.85))
    M["plaster"] = b.add_material(make_material(base_color=(0.80, 0.78, 0.72), roughness=0.9))
    M["column"] = b.add_material(make_material(base_color=(0.74, 0.72, 0.66), roughness=0.5))
    M["marble"] = b.add_material(make_material(base_color=(0.85, 0.85, 0.88), roughness=0.15, coat_weight=1.0, coat_roughness=0.03))
    M["bronze"] = b.add_material(make_material(base_color=(0.71, 0.43, 0.18), metallic=1.0, roughness=0.22))
    M["gold"] = b.add_material(make_material(base_color=(1.0, 0.78, 0.34), metallic=1.0, roughness=0.12))
    M["iron"] = b.add_material(make_material(base_color=(0.35, 0.35, 0.37), metallic=1.0, roughness=0.45))
    M["glass"] = b.add_material(make_material(base_color=(0.96, 0.98, 0.97), roughness=0.02, transmission=1.0, ior=1.5, double_sided=True))
    M["wood"] = b.add_material(make_material(base_color=(0.40, 0.26, 0.13), roughness=0.55))
    drape_cols = [(0.70, 0.10, 0.10), (0.10, 0.25, 0.60), (0.12, 0.45, 0.18), (0.75, 0.60, 0.15), (0.50, 0.12, 0.45), (0.85, 0.85, 0.80)]
    for k, c in enumerate(drape_cols):
        M["drape%d" % k] = b.add_material(make_material(base_color=c, roughness=0.8, double_sided=True))
    lamp_cols = [(1.0, 0.78, 0.45), (1.0, 0.55, 0.30), (0.75, 0.85, 1.0), (1.0, 0.95, 0.85), (0.6, 1.0, 0.7), (1.0, 0.6, 0.8)]
    for k, c in enumerate(lamp_cols):
        M["lamp%d" % k] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=c, emissive_strength=6.0 + 3.0 * k))
    M["panel"] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=(1.0, 0.96, 0.9), emissive_strength=4.0))
    M["ceiling"] = b.add_material(make_material(base_color=(0.7, 0.7, 0.72), roughness=0.95))
    assert len(b.mats) == 25

    X, Y, Z = 9.0, 10.0, 16.0
    d = detail
    # floor: gently uneven flagstones
    b.add_mesh(*patch(lambda U, V: np.stack([(U * 2 - 1) * X, 0.012 * np.sin(37 * U) * np.sin(53 * V), (V * 2 - 1) * Z], -1),
                      _n(d, 90), _n(d, 160), uv_scale=(9, 16), toward=(0.0, 1000.0, 0.0)), M["floor"])
    # ceiling
    b.add_mesh(*patch(lambda U, V: np.stack([(U * 2 - 1) * X, Y + 0 * U, (V * 2 - 1) * Z], -1), _n(d, 24), _n(d, 40), toward=(0.0, -1000.0, 0.0)), M["ceiling"])
    # walls with a brick-like relief; normals face inwards
    def wall_x(sx):
        return lambda U, V: np.stack([sx * (X + 0.03 * np.cos(60 * np.pi * U) * np.cos(25 * np.pi * V)) , V * Y, (U * 2 - 1) * Z], -1)
    def wall_z(sz):
        return lambda U, V: np.stack([(U * 2 - 1) * X, V * Y, sz * (Z + 0.03 * np.cos(34 * np.pi * U) * np.cos(25 * np.pi * V))], -1)
    for sx in (-1.0, 1.0):
        p, n, uv, i = patch(wall_x(sx), _n(d, 150), _n(d, 56))
        if (n[:, 0] * sx).sum() > 0:
            n = -n; i = i.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
        b.add_mesh(p, n, uv, i, M["brick"])
    for sz in (-1.0, 1.0):
        p, n, uv, i = patch(wall_z(sz), _n(d, 90), _n(d, 56))
        if (n[:, 2] * sz).sum() > 0:
            n = -n; i = i.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
        b.add_mesh(p, n, uv, i, M["plaster"])

    # colonnade: fluted shafts (one mesh, many instances), bases and capitals, two storeys
    shaft = cylinder(0.33, 3.6, _n(d, 48, 6), _n(d, 24, 2), profile=lambda V: 1.0 - 0.12 * V, flutes=12, flute_depth=0.08)
    base_box = box((0.95, 0.3, 0.95), (0, 0.15, 0))
    cap_box = box((1.0, 0.28, 1.0), (0, 0.14, 0))
    g_shaft = g_base = g_cap = None
    ncol = 9
    for storey in range(2):
        y0 = 0.0 if storey == 0 else 4.7
        for side in (-1.0, 1.0):
            for k in range(ncol):
                z = -14.0 + 28.0 * k / (ncol - 1)
                x = side * 5.6
                rot = _quat_y(0.37 * k + storey)
                sc = (1.0, 1.0, 1.0) if storey == 0 else (0.8, 0.9, 0.8)
                mat = M["column"] if (k + storey) % 3 else M["marble"]
                if g_shaft is None:
                    g_shaft = b.add_mesh(*shaft, mat, (x, y0 + 0.3, z), rot, sc)
                    g_base = b.add_mesh(*base_box, M["stone"], (x, y0, z))
                    g_cap = b.add_mesh(*cap_box, M["stone"], (x, y0 + 0.3 + 3.6 * sc[1], z))
                else:
                    b.add_instance_of(g_shaft, mat, (x, y0 + 0.3, z), rot, sc)
                    b.add_instance_of(g_base, M["stone"], (x, y0, z))
                    b.add_instance_of(g_cap, M["stone"], (x, y0 + 0.3 + 3.6 * sc[1], z))
    # gallery floor slabs of the upper storey (left and right)
    for side in (-1.0, 1.0):
        b.add_mesh(*box((3.9, 0.35, 31.0), (side * 7.05, 4.5, 0.0)), M["stone"])
        # balustrade rail + iron posts
        b.add_mesh(*box((0.12, 0.1, 30.0), (side * 5.2, 5.65, 0.0)), M["wood"])
    post = cylinder(0.035, 0.95, _n(d, 10, 4), 1)
    g_post = None
    for side in (-1.0, 1.0):
        for k in range(_n(d, 60, 6)):
            z = -14.6 + 29.2 * k / max(_n(d, 60, 6) - 1, 1)
            if g_post is None:
                g_post = b.add_mesh(*post, M["iron"], (side * 5.2, 4.68, z))
            else:
                b.add_instance_of(g_post, M["iron"], (side * 5.2, 4.68, z))
    # arches between ground-floor columns: half tori
    def arch(U, V):
        a = np.pi * U
        r = 0.16
        c = 2.0 * np.pi * V
        R = 1.75
        return np.stack([r * np.cos(c), R * np.sin(a) + r * np.sin(c) * np.sin(a), -R * np.cos(a) - r * np.sin(c) * np.cos(a)], -1)
    def arch_axis(U, V):
        a = np.pi * U
        return np.stack([0 * a, 1.75 * np.sin(a), -1.75 * np.cos(a)], -1)
    _u, _v = np.meshgrid(np.linspace(0, 1, _n(d, 40, 4) + 1), np.linspace(0, 1, _n(d, 12, 3) + 1), indexing="ij")
    am = patch(arch, _n(d, 40, 4), _n(d, 12, 3), away=arch_axis(_u, _v))
    g_arch = None
    for side in (-1.0, 1.0):
        for k in range(ncol - 1):
            z = -14.0 + 28.0 * (k + 0.5) / (ncol - 1)
            if g_arch is None:
                g_arch = b.add_mesh(*am, M["stone"], (side * 5.6, 2.6, z))
            else:
                b.add_instance_of(g_arch, M["stone"], (side * 5.6, 2.6, z))
    # drapes: wavy cloth hanging from the gallery
    for k in range(6):
        side = -1.0 if k % 2 == 0 else 1.0
        z0 = -11.0 + 4.2 * k
        ph = 1.3 * k
        def cloth(U, V, side=side, z0=z0, ph=ph):
            return np.stack([side * (4.95 + 0.10 * np.sin(18 * U + ph) * (0.3 + V)), 4.4 - 3.2 * V + 0.05 * np.sin(9 * U + ph),
                             z0 + 2.6 * U + 0.04 * np.sin(14 * V)], -1)
        b.add_mesh(*patch(cloth, _n(d, 84), _n(d, 84)), M["drape%d" % k])
    # a row of vases / spheres along the centre line
    vase = cylinder(0.32, 0.9, _n(d, 40, 6), _n(d, 24, 3), profile=lambda V: 0.45 + 0.55 * np.sin(np.pi * (0.15 + 0.8 * V)))
    sph = sphere(0.38, _n(d, 48, 6), _n(d, 24, 4))
    g_vase = g_sph = None
    for k in range(7):
        z = -9.0 + 3.5 * k
        x = 1.6 * (-1) ** k
        b.add_mesh(*box((0.7, 0.8, 0.7), (x, 0.4, z)), M["wood"] if k % 2 else M["stone"])
        if k % 2 == 0:
            mat = [M["bronze"], M["gold"], M["iron"], M["marble"]][(k // 2) % 4]
            if g_vase is None:
                g_vase = b.add_mesh(*vase, mat, (x, 0.8, z))
            else:
                b.add_instance_of(g_vase, mat, (x, 0.8, z), _quat_y(0.9 * k), (1.0, 1.0 + 0.1 * k, 1.0))
        else:
            mat = [M["glass"], M["marble"], M["gold"]][(k // 2) % 3]
            if g_sph is None:
                g_sph = b.add_mesh(*sph, mat, (x, 1.18, z))
            else:
                b.add_instance_of(g_sph, mat, (x, 1.18, z), _quat_axis((1, 1, 0), 0.5 * k))
    # lanterns: four strings of small emissive spheres across the hall (the many-light workload)
    n_lamps = 64
    if lamp_tris is None:
        ls, lr = _n(d, 16, 4), _n(d, 8, 2)          # 2 * 16 * 8 = 256 triangles per lantern at detail 1
    else:
        lr = max(int(np.sqrt(lamp_tris / 4.0)), 1); ls = max(lamp_tris // (2 * lr), 3)
    lantern = sphere(0.09, ls, lr)
    g_l = {}
    for k in range(n_lamps):
        string = k % 4
        t = (k // 4 + 0.5) / (n_lamps // 4)
        z = -14.0 + 28.0 * t
        x = (-4.2 + 2.8 * string) + 0.5 * np.sin(7.0 * t + string)
        y = 4.1 - 0.9 * np.sin(np.pi * t) + 0.25 * string
        c = int(rng.integers(0, len(lamp_cols)))
        if c not in g_l:
            g_l[c] = b.add_mesh(*lantern, M["lamp%d" % c], (x, y, z))
        else:
            b.add_instance_of(g_l[c], M["lamp%d" % c], (x, y, z), _quat_y(0.4 * k))
    # two emissive ceiling panels (large area lights)
    for z in (-7.0, 7.0):
        p, n, uv, i = patch(lambda U, V, z=z: np.stack([(U * 2 - 1) * 2.0, Y - 0.05 + 0 * U, z + (V * 2 - 1) * 3.0], -1), _n(d, 6), _n(d, 8), toward=(0.0, -1000.0, z))
        b.add_mesh(p, n, uv, i, M["panel"])
    return b.finish()


# ------------------------------------------------------------------------------------------------
# "Subway-class" station tunnel (config C5: ReSTIR PT, 5 bounces, glass + glossy metal)
# ------------------------------------------------------------------------------------------------
TUNNEL_CAMERA = (-1.6, 1.7, -4.0)


def tunnel(detail=1.0):
    """Station tunnel along +Z, z in [-8, 112]: ribbed vault, tiled platform, track bed with rails and sleepers, steel pillars,
    glass platform screens, benches, emissive tube lights and signs. ~1.2 M triangles at detail 1."""
    b = SceneBuilder()
    M = {}
    M["concrete"] = b.add_material(make_material(base_color=(0.52, 0.52, 0.50), roughness=0.8))
    M["tile"] = b.add_material(make_material(base_color=(0.82, 0.84, 0.80), roughness=0.12, coat_weight=0.8, coat_roughness=0.05))
    M["platform"] = b.add_material(make_material(base_color=(0.38, 0.38, 0.40), roughness=0.35))
    M["yellow"] = b.add_material(make_material(base_color=(0.9, 0.75, 0.1), roughness=0.5))
    M["ballast"] = b.add_material(make_material(base_color=(0.22, 0.20, 0.19), roughness=0.95))
    M["rail"] = b.add_material(make_material(base_color=(0.56, 0.57, 0.58), metallic=1.0, roughness=0.18))
    M["sleeper"] = b.add_material(make_material(base_color=(0.30, 0.27, 0.24), roughness=0.9))
    M["steel"] = b.add_material(make_material(base_color=(0.62, 0.63, 0.65), metallic=1.0, roughness=0.28))
    M["brushed"] = b.add_material(make_material(base_color=(0.75, 0.76, 0.78), metallic=1.0, roughness=0.42))
    M["chrome"] = b.add_material(make_material(base_color=(0.95, 0.95, 0.95), metallic=1.0, roughness=0.04))
    M["glass"] = b.add_material(make_material(base_color=(0.97, 0.99, 0.98), roughness=0.01, transmission=1.0, ior=1.52, double_sided=True))
    M["frosted"] = b.add_material(make_material(base_color=(0.9, 0.95, 0.95), roughness=0.3, transmission=1.0, ior=1.5, double_sided=True))
    M["thin"] = b.add_material(make_material(base_color=(0.9, 0.9, 0.85), roughness=0.2, transmission=1.0, thin_walled=True, double_sided=True, subsurface=0.4))
    M["bench"] = b.add_material(make_material(base_color=(0.45, 0.28, 0.12), roughness=0.4, coat_weight=1.0, coat_roughness=0.1))
    M["red"] = b.add_material(make_material(base_color=(0.7, 0.08, 0.08), roughness=0.3, coat_weight=0.5))
    M["tube"] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=(0.95, 1.0, 1.0), emissive_strength=9.0))
    M["tube_warm"] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=(1.0, 0.85, 0.6), emissive_strength=7.0))
    M["sign"] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=(0.2, 0.5, 1.0), emissive_strength=3.0, double_sided=True))
    M["signal"] = b.add_material(make_material(base_color=(0, 0, 0), emissive_factor=(1.0, 0.1, 0.05), emissive_strength=12.0))

    d = detail
    Z0, Z1 = -8.0, 112.0
    L = Z1 - Z0
    W = 7.0          # half width of the vault
    H = 6.2          # crown height
    # vault: half-ellipse with ribs every 3 m and panel relief
    def vault(U, V):
        a = np.pi * U
        z = Z0 + L * V
        rib = 0.10 * np.maximum(0.0, np.cos(2 * np.pi * z / 3.0)) ** 8
        rel = 0.015 * np.cos(40 * a) * np.cos(2 * np.pi * z / 0.75)
        r = 1.0 - rib / W - rel / W
        return np.stack([-W * np.cos(a) * r, 0.6 + (H - 0.6) * np.sin(a) * r, z], -1)
    p, n, uv, i = patch(vault, _n(d, 220), _n(d, 1500), uv_scale=(14, 120))
    if n[:, 1].sum() > 0:           # must face down / inwards
        n = -n; i = i.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
    b.add_mesh(p, n, uv, i, M["tile"])
    # end walls
    for z, flip in ((Z0, False), (Z1, True)):
        p, n, uv, i = patch(lambda U, V, z=z: np.stack([(U * 2 - 1) * W, V * (H + 0.5), z + 0 * U], -1), _n(d, 16), _n(d, 12))
        want = 1.0 if not flip else -1.0
        if (n[:, 2] * want).sum() < 0:
            n = -n; i = i.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
        b.add_mesh(p, n, uv, i, M["concrete"])
    # platform (x < 0.4) and track bed (x > 0.4), both tessellated
    b.add_mesh(*patch(lambda U, V: np.stack([-W + (W + 0.4) * U, 1.05 + 0.004 * np.cos(2 * np.pi * 10 * U) * np.cos(2 * np.pi * (Z0 + L * V) / 0.6),
                                             Z0 + L * V], -1), _n(d, 64), _n(d, 900), toward=(0.0, 1000.0, 50.0)), M["platform"])
    b.add_mesh(*box((0.35, 0.02, L), (0.1, 1.065, 0.5 * (Z0 + Z1))), M["yellow"])
    b.add_mesh(*box((0.1, 1.05, L), (0.45, 0.525, 0.5 * (Z0 + Z1))), M["concrete"])
    def bed(U, V):
        z = Z0 + L * V
        h = 0.5 + 0.25 * np.sin(41.0 * U + 3.1 * z) * np.cos(17.3 * z) + 0.25 * np.sin(23.0 * U - 7.7 * z)
        return np.stack([0.5 + (W - 0.5) * U, 0.02 + 0.05 * h, z], -1)
    b.add_mesh(*patch(bed, _n(d, 70), _n(d, 1400), toward=(3.0, 1000.0, 50.0)), M["ballast"])
    # rails and sleepers
    for x in (2.3, 3.8):
        b.add_mesh(*box((0.08, 0.16, L), (x, 0.26, 0.5 * (Z0 + Z1))), M["rail"])
    sl = box((2.6, 0.14, 0.26), (0, 0, 0))
    g = None
    ns = _n(d, 200, 8)
    for k in range(ns):
        z = Z0 + 0.3 + (L - 0.6) * k / (ns - 1)
        if g is None:
            g = b.add_mesh(*sl, M["sleeper"], (3.05, 0.12, z))
        else:
            b.add_instance_of(g, M["sleeper"], (3.05, 0.12, z))
    # pillars along the platform edge: steel tubes with chrome collars
    pil = cylinder(0.16, 4.3, _n(d, 40, 6), _n(d, 16, 2))
    col = cylinder(0.19, 0.12, _n(d, 40, 6), 1)
    gp = gc = None
    npil = 20
    for k in range(npil):
        z = Z0 + 4.0 + (L - 8.0) * k / (npil - 1)
        x = -2.2
        if gp is None:
            gp = b.add_mesh(*pil, M["steel"], (x, 1.05, z))
            gc = b.add_mesh(*col, M["chrome"], (x, 2.0, z))
        else:
            b.add_instance_of(gp, M["steel"] if k % 2 else M["brushed"], (x, 1.05, z), _quat_y(0.3 * k))
            b.add_instance_of(gc, M["chrome"], (x, 2.0, z))
    # glass platform screens (solid slabs: two refractions) between pillars, with frosted and thin-walled variants
    for k in range(npil - 1):
        z = Z0 + 4.0 + (L - 8.0) * (k + 0.5) / (npil - 1)
        mat = [M["glass"], M["frosted"], M["glass"], M["thin"]][k % 4]
        if mat == M["thin"]:
            p, n, uv, i = patch(lambda U, V, z=z: np.stack([-0.2 + 0 * U, 1.1 + 2.0 * V, z - 2.2 + 4.4 * U], -1), _n(d, 8), _n(d, 4))
            b.add_mesh(p, n, uv, i, mat)
        else:
            b.add_mesh(*box((0.04, 2.0, 4.4), (-0.2, 2.1, z)), mat)
        b.add_mesh(*box((0.06, 0.06, 4.5), (-0.2, 3.13, z)), M["brushed"])
    # benches and bins
    seat = box((0.5, 0.06, 1.8), (0, 0, 0))
    leg = cylinder(0.025, 0.42, _n(d, 12, 4), 1)
    gs = gl = None
    for k in range(10):
        z = Z0 + 9.0 + 10.5 * k
        if gs is None:
            gs = b.add_mesh(*seat, M["bench"], (-5.6, 1.5, z))
        else:
            b.add_instance_of(gs, M["bench"] if k % 3 else M["red"], (-5.6, 1.5, z))
        for dz in (-0.75, 0.75):
            for dx in (-0.2, 0.2):
                if gl is None:
                    gl = b.add_mesh(*leg, M["chrome"], (-5.6 + dx, 1.05, z + dz))
                else:
                    b.add_instance_of(gl, M["chrome"], (-5.6 + dx, 1.05, z + dz))
    # emissive tube lights under the vault (two rows) -- finely tessellated so they are thousands of emissive triangles
    tube = cylinder(0.035, 1.8, _n(d, 16, 4), _n(d, 12, 2))
    gt = None
    ntube = 56
    for k in range(ntube):
        row = k % 2
        z = Z0 + 2.0 + (L - 4.0) * (k // 2) / (ntube // 2 - 1)
        x = -3.4 if row == 0 else 2.9
        y = 4.9 if row == 0 else 5.0
        rot = _quat_axis((1, 0, 0), 0.5 * np.pi)        # lay the tube along z
        mat = M["tube"] if (k // 2) % 4 else M["tube_warm"]
        if gt is None:
            gt = b.add_mesh(*tube, mat, (x, y, z - 0.9), rot)
        else:
            b.add_instance_of(gt, mat, (x, y, z - 0.9), rot)
    # signs (double-sided emissive quads) and two red signals at the far end
    for k in range(8):
        z = Z0 + 8.0 + 13.0 * k
        p, n, uv, i = patch(lambda U, V, z=z: np.stack([-4.6 + 1.6 * U, 3.3 + 0.4 * V, z + 0 * U], -1), _n(d, 4), _n(d, 2))
        b.add_mesh(p, n, uv, i, M["sign"])
    sig = sphere(0.08, _n(d, 16, 4), _n(d, 8, 2))
    gsig = b.add_mesh(*sig, M["signal"], (5.4, 2.4, Z1 - 6.0))
    b.add_instance_of(gsig, M["signal"], (5.4, 2.4, Z0 + 30.0))
    return b.finish()


SCENES = {"atrium": (atrium, ATRIUM_CAMERA), "tunnel": (tunnel, TUNNEL_CAMERA)}
