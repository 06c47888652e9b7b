"""A SECOND, independent restatement of the reference's `fragment` entry point — test infrastructure.

Written from the WGSL text alone (assets/shaders/raytrace.wgsl, random.wgsl, const.wgsl of the reference), without
looking at oracle/bvr_oracle.cpp, and in a different shape on purpose: where the oracle is scalar C++ (one pixel at a
time, nested loops), this one is data-parallel numpy — every statement of the shader runs for ALL pixels at once on
float32 / uint32 arrays, with index sets standing in for control flow.  numpy evaluates one IEEE-754 binary32 operation
per ufunc call, so nothing is contracted into an FMA and nothing is computed in a wider type.

The conventions WGSL leaves open are the ones DESIGN.md §3 fixes (they are the specification both restatements share):
  uv of a pixel = ((x + 0.5) / W, (y + 0.5) / H);  dot(a, b) = (a.x b.x + a.y b.y) + a.z b.z;
  normalize(v) = v / sqrt(dot(v, v));  min / max = IEEE minNum / maxNum;  pow(x, 5) = ((x x)(x x)) x;
  tan(fov / 2) evaluated in double and rounded once;  u32(f32) truncates;  `||` short-circuits.
tests/test_oracle_nversion.py compares the two, bit for bit, on every golden scene."""
import numpy as np

F = np.float32
U = np.uint32
INF = F(3.40282347e+38)          # const.wgsl:2
STACKSIZE = 32                   # raytrace.wgsl:310


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _normalize(v):
    return v / np.sqrt(_dot(v, v))[..., None]


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


class Shader:
    """One draw of the fullscreen triangle: bind the uniforms and storage buffers, then `fragment()` for every pixel."""

    def __init__(self, models, materials, nodes, camera, level, random_seed, width, height, raster_rgba=None, raster_depth=None):
        self.pos = np.ascontiguousarray(models["position"], F)
        self.radius = np.ascontiguousarray(models["radius"], F)
        self.material_id = np.ascontiguousarray(models["material_id"], U)
        self.base_color = np.ascontiguousarray(materials["base_color"], F)
        self.metallic = np.ascontiguousarray(materials["metallic"], F)
        self.roughness = np.ascontiguousarray(materials["roughness"], F)
        self.ior = np.ascontiguousarray(materials["ior"], F)
        self.transmission = np.ascontiguousarray(materials["specular_transmission"], F)
        self.bmin = np.ascontiguousarray(nodes["bounds_min"], F)
        self.bmax = np.ascontiguousarray(nodes["bounds_max"], F)
        self.index = np.ascontiguousarray(nodes["index"], U).astype(np.int64)
        self.count = np.ascontiguousarray(nodes["model_count"], U).astype(np.int64)
        self.sample_count, self.bounce_count = int(camera.sample_count), int(camera.bounce_count)
        self.near, self.far, self.fov, self.aspect = F(camera.near_plane), F(camera.far_plane), F(camera.fov), F(camera.aspect)
        self.cam_pos = np.array(list(camera.position), F)
        self.cam_dir = np.array(list(camera.direction), F)
        self.cam_up = np.array(list(camera.up), F)
        self.level = int(level)
        self.seed = F(random_seed)
        self.W, self.H = int(width), int(height)
        self.raster_rgba, self.raster_depth = raster_rgba, raster_depth
        self.rng = None

    # ---- random.wgsl -------------------------------------------------------------------------------------------
    def rng_next_int(self, idx):
        old = self.rng[idx] + U(747796405) + U(2891336453)
        word = ((old >> ((old >> U(28)) + U(4))) ^ old) * U(277803737)
        self.rng[idx] = (word >> U(22)) ^ word

    def rng_next_float(self, idx):
        self.rng_next_int(idx)
        return self.rng[idx].astype(F) / F(4294967296.0)        # f32(0xffffffffu) rounds to 2^32

    def random_unit_vec3(self, idx):
        """randomVec3InUnitSphere: rejection loop, every lane draws until ITS point lies in the ball."""
        out = np.zeros((len(idx), 3), F)
        todo = np.arange(len(idx))
        while len(todo):
            lanes = idx[todo]
            x = self.rng_next_float(lanes)
            y = self.rng_next_float(lanes)
            z = self.rng_next_float(lanes)
            p = F(2.0) * np.stack([x, y, z], axis=-1) - F(1.0)
            ok = _dot(p, p) <= F(1.0)
            out[todo[ok]] = p[ok]
            todo = todo[~ok]
        return out

    # ---- raytrace.wgsl -----------------------------------------------------------------------------------------
    def random_ray_from_uv(self, uv):
        every = np.arange(len(uv))
        rx = self.rng_next_float(every) - F(0.5)
        ry = self.rng_next_float(every) - F(0.5)
        height = F(self.H)
        width = F(self.H) * self.aspect
        delta_u = (F(1.0) / width) * rx
        delta_v = (F(1.0) / height) * ry
        ndc_x = (uv[:, 0] * F(2.0) - F(1.0)) + delta_u
        ndc_y = (F(1.0) - uv[:, 1] * F(2.0)) + delta_v
        right = _cross(self.cam_dir, self.cam_up)
        scale = F(np.tan(np.float64(self.fov * F(0.5))))
        d = (self.cam_dir[None, :] + (ndc_x * self.aspect * scale)[:, None] * right[None, :]) + (ndc_y * scale)[:, None] * self.cam_up[None, :]
        origin = np.broadcast_to(self.cam_pos, d.shape).copy()
        return origin, _normalize(d)

    def ray_bounding_dst(self, origin, direction, box_min, box_max):
        inv = F(1.0) / direction
        with np.errstate(invalid="ignore"):
            t_min = (box_min - origin) * inv
            t_max = (box_max - origin) * inv
        t1 = np.fmin(t_min, t_max)
        t2 = np.fmax(t_min, t_max)
        t_near = np.fmax(np.fmax(t1[:, 0], t1[:, 1]), t1[:, 2])
        t_far = np.fmin(np.fmin(t2[:, 0], t2[:, 1]), t2[:, 2])
        hit = (t_far >= t_near) & (t_far > F(0.0))
        return np.where(hit, np.where(t_near > F(0.0), t_near, F(0.0)), INF).astype(F)

    def hit_sphere(self, model, origin, direction):
        oc = self.pos[model] - origin
        a = _dot(direction, direction)
        h = _dot(direction, oc)
        c = _dot(oc, oc) - self.radius[model] * self.radius[model]
        disc = h * h - a * c
        with np.errstate(invalid="ignore"):
            root = (h - np.sqrt(disc)) / a
        return np.where(disc < F(0.0), F(-1.0), root).astype(F)

    def raycast(self, origin, direction):
        n = len(origin)
        dist = np.full(n, INF, F)
        position = np.zeros((n, 3), F)
        normal = np.zeros((n, 3), F)
        material = np.zeros(n, np.int64)
        front = np.ones(n, bool)
        model_hit = np.full(n, 0xFFFFFFFF, np.int64)          # not part of HitInfo: for the primary-id plane only
        stack = np.zeros((n, STACKSIZE + 1), np.int64)      # one spare slot: a push at index 32 ends the loop, it is never read
        top = np.ones(n, np.int64)
        with np.errstate(divide="ignore"):
            while True:
                live = np.nonzero((top > 0) & (top < STACKSIZE))[0]
                if not len(live):
                    break
                top[live] -= 1
                node = stack[live, top[live]]
                leaf = self.count[node] > 0
                # -- raycast_against_range --
                lr, ln = live[leaf], node[leaf]
                for k in range(int(self.count[ln].max()) if len(ln) else 0):
                    sel = self.count[ln] > k
                    r, m = lr[sel], self.index[ln[sel]] + k
                    t = self.hit_sphere(m, origin[r], direction[r])
                    better = (t != F(-1.0)) & (t > F(0.001)) & (t < dist[r])
                    r, m, t = r[better], m[better], t[better]
                    hp = origin[r] + t[:, None] * direction[r]
                    nn = _normalize(hp - self.pos[m])
                    dist[r], position[r], normal[r] = t, hp, nn
                    material[r] = self.material_id[m]
                    front[r] = _dot(direction[r], nn) < F(0.0)
                    model_hit[r] = m
                # -- inner node: first child, then second child --
                ir, inode = live[~leaf], node[~leaf]
                for child in (0, 1):
                    c = self.index[inode] + child
                    d = self.ray_bounding_dst(origin[ir], direction[ir], self.bmin[c], self.bmax[c])
                    push = (d != INF) & (d < dist[ir])
                    stack[ir[push], top[ir[push]]] = c[push]
                    top[ir[push]] += 1
        return dist, position, normal, material, front, model_hit

    def background_gradient(self, direction):
        unit = _normalize(direction)
        a = F(0.5) * (unit[:, 1] + F(1.0))
        white = np.array([1.0, 1.0, 1.0], F)
        blue = np.array([0.5, 0.7, 1.0], F)
        return (F(1.0) - a)[:, None] * white[None, :] + a[:, None] * blue[None, :]

    @staticmethod
    def reflect(v, n):
        return v - (F(2.0) * _dot(v, n))[:, None] * n

    @staticmethod
    def refract(v, n, ratio):
        cos_theta = np.fmin(_dot(-v, n), F(1.0))
        perp = ratio[:, None] * (v + cos_theta[:, None] * n)
        parallel = (-np.sqrt(np.abs(F(1.0) - _dot(perp, perp))))[:, None] * n
        return perp + parallel

    @staticmethod
    def reflectance(cosine, ri):
        r0 = (F(1.0) - ri) / (F(1.0) + ri)
        r0 = r0 * r0
        x = F(1.0) - cosine
        x2 = x * x
        return r0 + (F(1.0) - r0) * ((x2 * x2) * x)

    def scatter(self, lanes, origin, direction, hit_pos, hit_normal, hit_material, front):
        """Vectorised `scatter`: `lanes` are the pixels (RNG owners) of the rays.  Returns (origin, direction,
        attenuation, absorbed)."""
        n = len(lanes)
        mat = np.minimum(hit_material, len(self.metallic) - 1)        # out-of-range index: clamped buffer access
        new_dir = np.zeros((n, 3), F)
        att = np.zeros((n, 3), F)
        absorbed = np.zeros(n, bool)
        u1 = self.rng_next_float(lanes)
        metal = u1 < self.metallic[mat]
        rest = np.nonzero(~metal)[0]
        u2 = self.rng_next_float(lanes[rest])
        glass_sel = u2 < self.transmission[mat[rest]]
        glass, diffuse = rest[glass_sel], rest[~glass_sel]
        metal = np.nonzero(metal)[0]
        # metallic interaction
        if len(metal):
            refl = _normalize(self.reflect(direction[metal], hit_normal[metal])) + self.roughness[mat[metal]][:, None] * self.random_unit_vec3(lanes[metal])
            new_dir[metal] = refl
            att[metal] = self.base_color[mat[metal]]
            absorbed[metal] = _dot(refl, hit_normal[metal]) < F(0.0)
        # specular transmission
        if len(glass):
            ior = self.ior[mat[glass]]
            ri = np.where(front[glass], F(1.0) / ior, ior).astype(F)
            unit = _normalize(direction[glass])
            nrm = hit_normal[glass]
            cos_theta = np.fmin(_dot(-unit, nrm), F(1.0))
            sin_theta = np.sqrt(F(1.0) - cos_theta * cos_theta)
            cannot = ri * sin_theta > F(1.0)
            reflects = cannot.copy()
            ask = np.nonzero(~cannot)[0]                              # `||` short-circuits: only these draw
            reflects[ask] = self.reflectance(cos_theta[ask], ri[ask]) > self.rng_next_float(lanes[glass[ask]])
            out = np.where(reflects[:, None], self.reflect(unit, nrm), self.refract(unit, nrm, ri)).astype(F)
            new_dir[glass] = out
            att[glass] = F(1.0)
        # normal diffuse
        if len(diffuse):
            nrm = hit_normal[diffuse]
            b1 = self.random_unit_vec3(lanes[diffuse])
            b2 = self.random_unit_vec3(lanes[diffuse])
            sd = nrm + b1 + self.roughness[mat[diffuse]][:, None] * b2
            s = F(1e-8)
            near_zero = (np.abs(sd[:, 0]) < s) & (np.abs(sd[:, 1]) < s) & (np.abs(sd[:, 2]) < s)
            sd[near_zero] = nrm[near_zero]
            new_dir[diffuse] = sd
            att[diffuse] = self.base_color[mat[diffuse]]
            absorbed[diffuse] = _dot(sd, nrm) < F(0.0)
        return hit_pos.copy(), new_dir, att, absorbed

    def raytrace(self, origin, direction, sample_index, primary_id, primary_depth):
        P = len(origin)
        fallback_far = self.far + F(10.0) if self.level == 1 else self.far - F(1.0)
        first_depth = np.full(P, INF, F)
        ray_color = np.ones((P, 3), F)
        light = np.zeros((P, 3), F)
        running = np.arange(P)                       # pixels still inside the bounce loop
        for bounce in range(self.bounce_count + 1):
            if not len(running):
                break
            o, d = origin[running], direction[running]
            dist, hp, hn, hm, ff, model = self.raycast(o, d)
            if bounce == 0:
                first_depth[running] = dist
                if sample_index == 0:
                    primary_id[running] = np.where(dist == INF, 0xFFFFFFFF, model).astype(U)
                    primary_depth[running] = dist
            miss = dist == INF
            light[running[miss]] = self.background_gradient(d[miss])
            keep = ~miss
            lanes = running[keep]
            no, nd, att, absorbed = self.scatter(lanes, o[keep], d[keep], hp[keep], hn[keep], hm[keep], ff[keep])
            origin[lanes], direction[lanes] = no, nd
            cont = ~absorbed
            ray_color[lanes[cont]] = ray_color[lanes[cont]] * att[cont]
            running = lanes[cont]
        ray_color[running] = F(0.0)                  # the loop ran out: bounce_count == camera.bounce_count + 1
        first_depth[first_depth == INF] = fallback_far
        return np.sqrt(ray_color * light), first_depth

    def fragment(self):
        W, H = self.W, self.H
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        uv = np.stack([(xs.reshape(-1).astype(F) + F(0.5)) / F(W), (ys.reshape(-1).astype(F) + F(0.5)) / F(H)], axis=-1)
        P = W * H
        seed = ((self.seed * F(10000.0)) * (uv[:, 0] * F(402.0))) * (uv[:, 1] * F(31.5))
        self.rng = seed.astype(np.int64).astype(U)
        primary_id = np.full(P, 0xFFFFFFFF, U)
        primary_depth = np.full(P, INF, F)
        if self.level == 0:
            return {"rgba": self.raster_rgba.astype(F).copy(), "rt_depth": np.zeros((H, W), F),
                    "primary_id": primary_id.reshape(H, W), "primary_depth": primary_depth.reshape(H, W)}
        total_color = np.zeros((P, 3), F)
        total_depth = np.zeros(P, F)
        for s in range(self.sample_count):
            origin, direction = self.random_ray_from_uv(uv)
            color, depth = self.raytrace(origin, direction, s, primary_id, primary_depth)
            total_color = total_color + color
            total_depth = total_depth + depth
        with np.errstate(invalid="ignore", divide="ignore"):
            averaged_color = total_color / F(self.sample_count)
            averaged_depth = total_depth / F(self.sample_count)
        rgba = np.concatenate([averaged_color, np.ones((P, 1), F)], axis=1).reshape(H, W, 4)
        if self.level in (1, 2):
            depth = self.raster_depth.reshape(-1).astype(F)
            with np.errstate(divide="ignore"):
                rt = np.where(averaged_depth > self.far, F(-1.0), self.near / averaged_depth).astype(F)
            raster_wins = (depth > rt).reshape(H, W)
            rgba[raster_wins] = self.raster_rgba.astype(F)[raster_wins]
        return {"rgba": rgba, "rt_depth": averaged_depth.reshape(H, W), "primary_id": primary_id.reshape(H, W),
                "primary_depth": primary_depth.reshape(H, W)}


def render(models, materials, nodes, camera, level, random_seed, width, height, raster_rgba=None, raster_depth=None):
    return Shader(models, materials, nodes, camera, level, random_seed, width, height, raster_rgba, raster_depth).fragment()
