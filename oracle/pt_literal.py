"""ORACLE CROSS-CHECK (test infrastructure): PathTrace.render of integrator/PT_RGB.py:44-136 for one pixel and one frame,
transliterated a second time in plain Python with numpy float32 scalars (triangle scenes without environment light: the Cornell
box; emitters may be triangles, spheres, spot and laser lights).  It shares nothing with oracle/tiray_oracle.cpp's pt_rgb_pixel except the BVH queries (closet_hit / closet_hit_shadow
through orc_trace) and the Philox stream (orc_rng): hit attributes, light sampling, the Disney and glass BSDFs, NEE with the
power heuristic, the throughput recursion and the RNG block order are restated here from the reference.  numpy's float32
sin / cos / pow differ from glibc's by ULPs, so the comparison in tests/test_oracle.py uses a relative tolerance.
Only for small cases (pure-Python loops)."""
import numpy as np

from oracle.bdpt_literal import (f32, PI_REF, v3, dot, length, normalized, disney_evaluate_pdf, srgb_to_lrgb, offset_ray)

INF_VALUE = f32(1000000.0)
MAT_DISNEY, MAT_GLASS, MAT_LIGHT = 0, 1, 2
PRIMITIVE_TRI = 1
SHPAE_SPHERE, SHPAE_SPOT, SHPAE_LASER = 1, 3, 4
MAX_DEPTH = 15


def cross(a, b):
    return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def sign(x): return f32(int(x > 0) - int(x < 0))
def power_heuristic(a, b): t = a * a; return t / (b * b + t)                     # UtilsFunc.py:434-438
def reflect(i, n): return i - f32(2.0) * dot(n, i) * n                          # taichi_glsl


def cosine_sample_hemisphere(u1, u2):                                            # UtilsFunc.py:352-360
    r = np.sqrt(u1); phi = f32(2.0) * PI_REF * u2
    px, py = r * np.cos(phi, dtype=f32), r * np.sin(phi, dtype=f32)
    pz = np.sqrt(max(f32(0.0), f32(1.0) - px * px - py * py))
    return normalized(v3(px, py, pz))


def inverse_transform(d, N):                                                     # UtilsFunc.py:373-387
    n = normalized(N)
    b = v3(-n[1], n[0], 0.0) if abs(n[0]) > abs(n[2]) else v3(0.0, -n[2], n[1])
    b = normalized(b)
    t = normalized(cross(b, n))
    return (d[0] * t + d[1] * b) + d[2] * n


def disney_sample(d, N, metal, rough, prob, r1, r2):                             # brdf/Disney.py:17-40
    dr = f32(0.5) * (f32(1.0) - metal); alpha = max(f32(0.001), rough)
    if prob < dr:
        return inverse_transform(cosine_sample_hemisphere(r1, r2), N)
    phi = r1 * f32(2.0) * PI_REF
    cos_t = np.sqrt((f32(1.0) - r2) / (f32(1.0) + (alpha * alpha - f32(1.0)) * r2))
    sin_t = np.sqrt(f32(1.0) - (cos_t * cos_t))
    sin_p, cos_p = np.sin(phi, dtype=f32), np.cos(phi, dtype=f32)
    half = inverse_transform(v3(sin_t * cos_p, sin_t * sin_p, cos_t), N)
    return reflect(d, half)


def glass_sample(d, N, ior, prob):                                               # brdf/Glass.py:9-34, UtilsFunc.py:417-432
    cos_i = dot(d, N); eta = ior; f_or_b = f32(1.0); R = prob + f32(1.0)
    if cos_i > 0.0:
        N = -N
    else:
        cos_i = -cos_i; eta = f32(1.0) / ior
    ni = dot(N, d); k = f32(1.0) - eta * eta * (f32(1.0) - ni * ni)
    nxt = np.zeros(3, f32); suc = f32(-1.0)
    if k > 0.0:
        nxt = eta * d - (eta * ni + np.sqrt(k)) * N; suc = f32(1.0)
    if suc > 0.0:
        r0 = (f32(1.0) - ior) / (f32(1.0) + ior); r0 = r0 * r0
        R = r0 + (f32(1.0) - r0) * np.power(f32(1.0) - cos_i, f32(5.0), dtype=f32)
    if prob < R:
        nxt = reflect(d, N)
    else:
        f_or_b = f32(-1.0)
    return nxt, f_or_b


class Tracer:
    def __init__(self, scene, tables, cam, seed=0):
        self.s, self.t, self.seed = scene, tables, seed
        self.view_inv, self.eye = cam[1].astype(f32), cam[2].astype(f32)
        self.fx, self.fy, self.cx, self.cy = f32(cam[3]), f32(cam[4]), f32(cam[5]), f32(cam[6])
        self.nl = int(tables.light.size)

    def rng(self, i, j, frame, block):
        out = np.zeros(4, f32); self.s.lib.orc_rng(self.seed, i * 65536 + j, frame, block, out); return out

    def ray_direction(self, i, j, jx, jy):                                       # Camera.py:130-142
        x = (f32(i) + jx - self.cx) / self.fx; y = (f32(j) + jy - self.cy) / self.fy; z = f32(-1.0)
        m = self.view_inv
        return normalized(v3((m[0, 0] * x + m[0, 1] * y) + m[0, 2] * z, (m[1, 0] * x + m[1, 1] * y) + m[1, 2] * z, (m[2, 0] * x + m[2, 1] * y) + m[2, 2] * z))

    def closest(self, o, d):                                                     # Scene.closet_hit + intersect_prim (Scene.py:537-561,702-744)
        t, prim, uv = self.s.trace(o[None, :], d[None, :], shadow=False)
        t, prim, u, v = f32(t[0]), int(prim[0]), f32(uv[0, 0]), f32(uv[0, 1])
        if not t < INF_VALUE:
            return t, None, None, None, -1
        vi = int(self.t.primitive[prim, 1]); V = self.t.vertex
        v1, v2, v3_ = V[vi, 0:3], V[vi + 1, 0:3], V[vi + 2, 0:3]
        n1, n2, n3 = V[vi, 3:6], V[vi + 1, 3:6], V[vi + 2, 3:6]
        a, b, c = f32(1.0) - u - v, u, v
        gn = normalized(cross(v2 - v1, v3_ - v1))
        pos = (a * v1 + b * v2) + c * v3_
        nor = normalized((a * n1 + b * n2) + c * n3)
        return t, pos, gn, nor, prim

    def prim_area(self, p):                                                      # Scene.py:324-350
        if int(self.t.primitive[p, 0]) != PRIMITIVE_TRI:
            sh = self.t.shape[int(self.t.primitive[p, 1])]
            if int(sh[0]) in (SHPAE_SPHERE, SHPAE_SPOT, SHPAE_LASER):
                r = f32(sh[4]); return r * r * f32(3.1415926)
            return f32(0.0)
        vi = int(self.t.primitive[p, 1]); V = self.t.vertex
        a, b, c = length(V[vi, 0:3] - V[vi + 1, 0:3]), length(V[vi, 0:3] - V[vi + 2, 0:3]), length(V[vi + 2, 0:3] - V[vi + 1, 0:3])
        sm = ((a + b) + c) * f32(0.5)
        return np.sqrt(sm * (sm - a) * (sm - b) * (sm - c))

    def prim_random_point_normal(self, pi, a, b):                                # Scene.py:381-420
        pos, nor = np.zeros(3, f32), np.zeros(3, f32)
        if int(self.t.primitive[pi, 0]) == PRIMITIVE_TRI:
            vi = int(self.t.primitive[pi, 1]); V = self.t.vertex
            if a + b > 1.0: a = f32(1.0) - a; b = f32(1.0) - b
            pos = (V[vi, 0:3] + (V[vi + 2, 0:3] - V[vi, 0:3]) * a) + (V[vi + 1, 0:3] - V[vi, 0:3]) * b
            nor = normalized(((f32(1.0) - a - b) * V[vi, 3:6] + V[vi + 1, 3:6] * a) + V[vi + 2, 3:6] * b)
        else:
            sh = self.t.shape[int(self.t.primitive[pi, 1])].astype(f32); st = int(sh[0])
            if st == SHPAE_SPHERE:                                               # UniformSampleSphere, Scene.py:315-322
                z = f32(1.0) - f32(2.0) * a
                r = np.sqrt(min(max(f32(1.0) - z * z, f32(0.0)), f32(1.0))); phi = f32(2.0) * f32(3.1415926) * b
                nor = v3(r * np.cos(phi, dtype=f32), r * np.sin(phi, dtype=f32), z); pos = sh[1:4] + nor * sh[4]
            elif st in (SHPAE_SPOT, SHPAE_LASER):
                nor = sh[7:10].copy(); pos = sh[1:4].copy()
        return pos, normalized(nor)

    def sample_li(self, pos, u_idx, a, b):                                       # Scene.py:477-518
        index = int(u_idx * f32(self.nl))
        if index >= self.nl: index = self.nl - 1
        pi = int(self.t.light[index])
        lpos, nor = self.prim_random_point_normal(pi, a, b)
        emission = self.t.material[int(self.t.primitive[pi, 2]), 2:5].astype(f32)
        choice_pdf = f32(1.0) / (f32(self.nl) * self.prim_area(pi))
        nor = normalized(nor)
        d = pos - lpos; dist = length(d); d = d / dist
        ndl = abs(dot(d, nor))
        visable = f32(1.0)
        if int(self.t.primitive[pi, 0]) != PRIMITIVE_TRI:
            sh = self.t.shape[int(self.t.primitive[pi, 1])].astype(f32); st = int(sh[0])
            if st == SHPAE_SPOT:                                                 # :499-506
                x1, x2 = sh[4], sh[5]
                x = np.arccos(ndl, dtype=f32)
                if x > x2:
                    visable = f32(0.0)
                elif x > x1:
                    visable = visable * (f32(1.0) - (x - x1) / (x2 - x1))
            elif st == SHPAE_LASER:                                              # :508-515
                choice_pdf = f32(1.0) / f32(self.nl)
                proj = dot(d, nor) * dist
                r = np.sqrt(dist * dist - proj * proj)
                if r > sh[4]:
                    visable = f32(0.0)
        return lpos, nor, d, emission * visable, dist, choice_pdf

    def pixel(self, i, j, frame):
        """-> radiance (3,) f32, (closest-hit calls, shadow calls)"""
        jx = jy = f32(0.0)
        if frame != 0:
            r = self.rng(i, j, frame, 0); jx, jy = r[0] - f32(0.5), r[1] - f32(0.5)
        next_o, next_d = self.eye.copy(), self.ray_direction(i, j, jx, jy)
        depth = 0; brdf_pdf = f32(1.0); perfect_spec = 1
        T, L = np.ones(3, f32), np.zeros(3, f32)
        n_closest = n_shadow = 0
        while depth < MAX_DEPTH:
            o, d = next_o, next_d
            n_closest += 1
            t, pos, gn, nor, prim = self.closest(o, d)
            if not t < INF_VALUE:
                break                                                         # black environment
            fn = sign(dot(-d, gn)) * nor
            mid = int(self.t.primitive[prim, 2]); mrow = self.t.material[mid]
            mcol, mtype, p0, p1 = mrow[2:5].astype(f32), int(mrow[0]), f32(mrow[5]), f32(mrow[6])
            if mtype == MAT_LIGHT:
                f_cos = abs(dot(d, gn))
                if perfect_spec == 1:
                    L = L + T * mcol
                else:
                    area = self.prim_area(prim) * f32(self.nl)
                    light_pdf = (t * t) / (area * f_cos)
                    L = L + power_heuristic(brdf_pdf, light_pdf) * T * mcol
                break
            rc = srgb_to_lrgb(mcol)
            R0 = self.rng(i, j, frame, 1 + 2 * depth); R1 = self.rng(i, j, frame, 2 + 2 * depth)
            if mtype == MAT_GLASS:
                perfect_spec = 1
                next_d, f_or_b = glass_sample(d, nor, p0, R0[3])
                brdf, brdf_pdf = f32(1.0), f32(1.0)
            else:
                perfect_spec = 0
                lpos, lnor, ldir, emission, ldist, choice_pdf = self.sample_li(pos, R0[0], R0[1], R0[2])
                ndl_s, ndl_l = dot(fn, ldir), dot(lnor, ldir)
                if ndl_s < 0.0 and ndl_l > 0.0:
                    n_shadow += 1
                    st, sp, _ = self.s.trace(lpos[None, :], ldir[None, :], shadow=True)
                    if int(sp[0]) == prim:
                        b2, p2 = disney_evaluate_pdf(fn, -d, -ldir, p0, p1)
                        light_pdf = ldist * ldist * choice_pdf / ndl_l
                        if p2 > 0.0:
                            w = power_heuristic(light_pdf, p2) / max(f32(0.0001), light_pdf)
                            L = L + ((((w * emission) * T) * rc) * b2) * abs(ndl_s)
                f_or_b = f32(1.0)
                next_d = disney_sample(d, fn, p0, p1, R0[3], R1[0], R1[1])
                brdf, brdf_pdf = disney_evaluate_pdf(fn, -d, next_d, p0, p1)
                brdf = brdf * abs(dot(nor, next_d))
            next_o = offset_ray(pos, sign(f_or_b) * fn)
            if brdf_pdf > 0.0:
                if f_or_b < 0.0:
                    if R1[2] >= np.exp(-t / p1, dtype=f32):
                        break
                T = T * ((brdf / brdf_pdf) * rc)
                depth += 1
            else:
                break
        return L, (n_closest, n_shadow)
