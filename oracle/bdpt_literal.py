"""ORACLE CROSS-CHECK (test infrastructure): connect_path + mis_weight of integrator/BDPT_RGB.py:258-580 transliterated a second
time, in plain Python with numpy float32 scalars, as literally as the language allows -- dense per-pixel `Vertex` arrays with the
depth axis padded to 8 and indexed `& 7` (what Taichi 0.7.14's dense SNode does with the reference's -1 indices), field-by-field
`copy()` in the order of integrator/BDPT_Vertex.py:45-57, the save / overwrite / walk / restore sequence of mis_weight on those
arrays.  It consumes the sub-path vertices the C++ oracle dumps (orc_bdpt_pixel_dump) and re-derives every strategy's weighted
contribution; tests/test_bdpt_cpu.py compares the two.  The C++ restatement (oracle/bdpt_core.inc) keeps the same arrays but was
written independently of this file's control flow: aliasing or index slips in either show up as a mismatch.
Only for small cases (pure-Python loops).  Visibility queries go through the oracle's closet_hit_shadow (orc_trace)."""
import numpy as np

f32 = np.float32
PI_REF = f32(3.1415956)            # UtilsFunc.py:37 (sic)
EPS = f32(0.00001)
VERTEX_NONE, VERTEX_LIGHT, VERTEX_LENS, VERTEX_SURFACE = 0, 1, 2, 3
MAT_DISNEY = 0
MAX_DEPTH = 5
FIELDS = ("pos", "normal", "snormal", "beta", "wo", "fpdf", "rpdf", "type", "prim", "mat", "delta", "power")   # BDPT_Vertex.py:9-21


def v3(x, y, z): return np.array([x, y, z], f32)
def dot(a, b): return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
def length(a): return np.sqrt(dot(a, a))
def normalized(a): return (f32(1.0) / length(a)) * a
def clamp(x, lo, hi): return min(max(x, lo), hi)
def mix(a, b, t): return a * (f32(1.0) - t) + b * t


class Vertex:
    """integrator/BDPT_Vertex.py: one dense field per attribute, depth axis padded to 8"""

    def __init__(self):
        self.pos = np.zeros((8, 3), f32); self.normal = np.zeros((8, 3), f32); self.snormal = np.zeros((8, 3), f32)
        self.beta = np.zeros((8, 3), f32); self.wo = np.zeros((8, 3), f32)
        self.fpdf = np.zeros(8, f32); self.rpdf = np.zeros(8, f32); self.power = np.zeros(8, f32)
        self.type = np.zeros(8, np.int32); self.prim = np.zeros(8, np.int32); self.mat = np.zeros(8, np.int32); self.delta = np.zeros(8, np.int32)

    def copy(self, depth, temp, k):                      # BDPT_Vertex.py:44-57
        for name in FIELDS:
            getattr(self, name)[depth & 7] = getattr(temp, name)[k & 7]


def schlick_fresnel(u):
    m = clamp(f32(1.0) - u, f32(0.0), f32(1.0)); m2 = m * m
    return m2 * m2 * m


def gtr2(ndoth, a):
    a2 = a * a; t = f32(1.0) + (a2 - f32(1.0)) * ndoth * ndoth
    return a2 / (PI_REF * t * t)


def smithg(ndotv, alphag):
    a = alphag * alphag; b = ndotv * ndotv
    return f32(1.0) / (ndotv + np.sqrt(a + b - a * b))


def disney_pdf(N, V, L, metal, rough):                  # brdf/Disney.py:42-63
    pdf = f32(0.0)
    ndotl, ndotv = dot(N, L), dot(N, V)
    if ndotl > 0.0 and ndotv > 0.0:
        H = normalized(L + V)
        ndoth, ldoth = dot(H, N), dot(H, L)
        alpha = max(f32(0.001), rough)
        ds = gtr2(ndoth, alpha)
        dr = f32(0.5) * (f32(1.0) - metal); sr = f32(1.0) - dr
        pdf_spec = (ds * ndoth) / (f32(4.0) * abs(ldoth)); pdf_diff = f32(1.0) / PI_REF
        pdf = dr * pdf_diff + sr * pdf_spec
    return pdf


def disney_evaluate_pdf(N, V, L, metal, rough):         # brdf/Disney.py:65-108
    out, pdf = f32(0.0), f32(-1.0)
    ndotl, ndotv = dot(N, L), dot(N, V)
    if ndotl > 0.0 and ndotv > 0.0:
        H = normalized(L + V)
        ndoth, ldoth = dot(H, N), dot(H, L)
        cspec0 = mix(f32(0.04), f32(1.0), metal)
        fl, fv = schlick_fresnel(ndotl), schlick_fresnel(ndotv)
        fd90 = f32(0.5) + f32(2.0) * ldoth * ldoth * rough
        fd = mix(f32(1.0), fd90, fl) * mix(f32(1.0), fd90, fv)
        alpha = max(f32(0.001), rough)
        ds = gtr2(ndoth, alpha)
        fh = schlick_fresnel(ldoth)
        fs = mix(cspec0, f32(1.0), fh)
        rg = rough * f32(0.5) + f32(0.5); rg = rg * rg
        gs = smithg(ndotl, rg) * smithg(ndotv, rg)
        fsheen = fh * f32(0.5)
        out = (fsheen + f32(1.0) / PI_REF) * fd * (f32(1.0) - metal) + gs * fs * ds
        dr = f32(0.5) * (f32(1.0) - metal); sr = f32(1.0) - dr
        pdf = dr * (f32(1.0) / PI_REF) + sr * ((ds * ndoth) / (f32(4.0) * abs(ldoth)))
    return out, pdf


def cosine_hemisphere_pdf(c): return max(f32(0.01), c / PI_REF)      # UtilsFunc.py:348-350
def remap0(x): return f32(1.0) if x == 0.0 else x                    # BDPT_RGB.py:90-94


def srgb_to_lrgb(c):                                    # UtilsFunc.py:76-94
    out = np.zeros(3, f32)
    for k in range(3):
        out[k] = c[k] / f32(12.92) if c[k] < f32(0.04045) else np.power((c[k] + f32(0.055)) / f32(1.055), f32(2.4), dtype=f32)
    return out


def offset_ray(p, n):                                   # UtilsFunc.py:440-461
    out = np.zeros(3, f32)
    for k in range(3):
        i_of = np.int32(f32(256.0) * n[k])
        i_p = p[k:k + 1].copy().view(np.int32)[0]
        i_p = i_p - i_of if p[k] < 0.0 else i_p + i_of
        f_p = np.array([i_p], np.int32).view(f32)[0]
        out[k] = p[k] + f32(1.0 / 2048.0) * n[k] if abs(p[k]) < f32(1.0 / 256.0) else f_p
    return out


class Pixel:
    """the per-pixel state of BDPT.render for one (i, j): eye / light sub-paths from the C++ oracle's dump + the five scratch vertices"""

    def __init__(self, scene, tables, cam, W, H, verts, depths, i, j, frame, seed=0):
        self.s, self.t = scene, tables
        self.view, self.eye_pos, self.fx, self.fy, self.cx, self.cy = cam[0].astype(f32), cam[2].astype(f32), f32(cam[3]), f32(cam[4]), f32(cam[5]), f32(cam[6])
        self.W, self.H, self.i, self.j, self.frame, self.seed = W, H, i, j, frame, seed
        self.eye, self.light, self.sample = Vertex(), Vertex(), Vertex()
        self.ltemp, self.etemp, self.lminustemp, self.eminustemp = Vertex(), Vertex(), Vertex(), Vertex()
        self.eye_depth, self.light_depth = depths
        for k in range(7):
            self._load(self.eye, k, verts[k])
        for k in range(6):
            self._load(self.light, k, verts[7 + k])
        self.nl = int(tables.light.size)

    @staticmethod
    def _load(V, k, row):
        V.pos[k] = row[0:3]; V.normal[k] = row[3:6]; V.snormal[k] = row[6:9]; V.beta[k] = row[9:12]; V.wo[k] = row[12:15]
        V.fpdf[k] = row[15]; V.rpdf[k] = row[16]; fl = int(row[17]); V.type[k] = fl % 16; V.delta[k] = fl // 16
        V.prim[k] = int(row[18]); V.mat[k] = int(row[19])

    # ---- scene accessors
    def mat_rows(self, m): r = self.t.material[m]; return r[2:5].astype(f32), f32(r[5]), f32(r[6])

    def prim_area(self, p):                              # Scene.py:324-350 (triangles)
        vi = int(self.t.primitive[p, 1]); v = self.t.vertex
        v1, v2, v3_ = v[vi, 0:3], v[vi + 1, 0:3], v[vi + 2, 0:3]
        a, b, c = length(v1 - v2), length(v1 - v3_), length(v3_ - v2)
        sm = ((a + b) + c) * f32(0.5)
        return np.sqrt(sm * (sm - a) * (sm - b) * (sm - c))

    def sample_li(self, pos, u_idx, a, b):               # Scene.py:477-518 (triangle emitters)
        index = int(u_idx * f32(self.nl));
        if index >= self.nl: index = self.nl - 1
        pi = int(self.t.light[index]); vi = int(self.t.primitive[pi, 1]); v = self.t.vertex
        v1, v2, v3_ = v[vi, 0:3], v[vi + 1, 0:3], v[vi + 2, 0:3]
        n1, n2, n3 = v[vi, 3:6], v[vi + 1, 3:6], v[vi + 2, 3:6]
        if a + b > 1.0: a = f32(1.0) - a; b = f32(1.0) - b
        lpos = (v1 + (v3_ - v1) * a) + (v2 - v1) * b
        nor = normalized(((f32(1.0) - a - b) * n1 + n2 * a) + n3 * b)
        nor = normalized(nor); nor = normalized(nor)
        emission = self.t.material[int(self.t.primitive[pi, 2]), 2:5].astype(f32)
        choice_pdf = f32(1.0) / (f32(self.nl) * self.prim_area(pi))
        d = pos - lpos; dist = length(d); d = d / dist
        return lpos, nor, d, emission, dist, pi, choice_pdf

    def shadow(self, o, d):
        t, prim, _ = self.s.trace(o[None, :], d[None, :], shadow=True)
        return f32(t[0]), int(prim[0])

    def optical_axis(self): return self.view[2, 0:3]     # Camera.py:128-129

    def image_point(self, p):                            # Camera.py:144-158
        m = self.view
        px = ((m[0, 0] * p[0] + m[0, 1] * p[1]) + m[0, 2] * p[2]) + m[0, 3] * f32(1.0)
        py = ((m[1, 0] * p[0] + m[1, 1] * p[1]) + m[1, 2] * p[2]) + m[1, 3] * f32(1.0)
        pz = ((m[2, 0] * p[0] + m[2, 1] * p[1]) + m[2, 2] * p[2]) + m[2, 3] * f32(1.0)
        with np.errstate(all="ignore"):
            fu, fv = -px / pz * self.fx + self.cx, -py / pz * self.fy + self.cy
        u = int(fu) if np.isfinite(fu) and abs(fu) < 2e9 else -1
        v = int(fv) if np.isfinite(fv) and abs(fv) < 2e9 else -1
        wi = np.zeros(3, f32)
        if u < 0 or u >= self.W or v < 0 or v >= self.H or pz > 0.0:
            u = v = -1
        else:
            wi = p - self.eye_pos
        with np.errstate(all="ignore"):
            wi = normalized(wi)
        return u, v, wi

    # ---- BDPT_RGB.py:258-434
    def mis_weight(self, e, l):
        light, eye, sample = self.light, self.eye, self.sample
        weight_sum = f32(0.0)
        if l + e != 2:
            if l > 0: self.ltemp.copy(0, light, l - 1)
            if e > 0: self.etemp.copy(0, eye, e - 1)
            if l > 1: self.lminustemp.copy(0, light, l - 2)
            if e > 1: self.eminustemp.copy(0, eye, e - 2)
            if l == 1: light.copy(0, sample, 0)
            elif e == 1: eye.copy(0, sample, 0)
            if l > 0: light.delta[(l - 1) & 7] = 0
            if e > 0: eye.delta[(e - 1) & 7] = 0
            E, L = (lambda k: k & 7), (lambda k: k & 7)
            if e > 0:
                if l == 0:
                    eye.rpdf[E(e - 1)] = (f32(1.0) / self.prim_area(int(eye.prim[E(e - 1)]))) * (f32(1.0) / f32(self.nl))
                elif l == 1:
                    if eye.type[E(e - 1)] == VERTEX_SURFACE:
                        to = eye.pos[E(e - 1)] - light.pos[0]; dist = length(to); to = to / dist
                        pdf_dir = cosine_hemisphere_pdf(abs(dot(to, light.normal[0]))); ldotn = abs(dot(to, light.normal[0]))
                        eye.rpdf[E(e - 1)] = pdf_dir * ldotn / (dist * dist)
                    else:
                        eye.rpdf[E(e - 1)] = 1.0
                else:
                    wi = light.pos[L(l - 2)] - light.pos[L(l - 1)]; wo = eye.pos[E(e - 1)] - light.pos[L(l - 1)]
                    dist = length(wo); wi = normalized(wi); wo = normalized(wo)
                    pdf = f32(1.0); mat_id = int(light.mat[L(l - 1)])
                    if mat_id == MAT_DISNEY:
                        _, metal, rough = self.mat_rows(mat_id); pdf = disney_pdf(light.snormal[L(l - 1)], wi, wo, metal, rough)
                    eye.rpdf[E(e - 1)] = pdf * abs(dot(light.normal[L(l - 1)], wo)) / (dist * dist)
            if l > 0:
                if e > 1:
                    if eye.type[E(e - 1)] == VERTEX_SURFACE:
                        wi = eye.pos[E(e - 2)] - eye.pos[E(e - 1)]; wo = light.pos[L(l - 1)] - eye.pos[E(e - 1)]
                        dist = length(wo); wi = normalized(wi); wo = normalized(wo)
                        pdf = f32(1.0); mat_id = int(eye.mat[E(e - 1)])
                        if mat_id == MAT_DISNEY:
                            _, metal, rough = self.mat_rows(mat_id); pdf = disney_pdf(eye.snormal[E(e - 1)], wi, wo, metal, rough)
                        light.rpdf[L(l - 1)] = pdf * abs(dot(eye.normal[E(e - 1)], wo)) / (dist * dist)
                    else:
                        light.rpdf[L(l - 1)] = 1.0
                else:
                    to = eye.pos[0] - light.pos[L(l - 1)]; dist = length(to); to = to / dist
                    light.rpdf[L(l - 1)] = dot(to, self.optical_axis()) / (dist * dist)
            if e > 1:
                if l == 0:
                    to = eye.pos[E(e - 2)] - eye.pos[E(e - 1)]; dist = length(to); to = to / dist
                    pdf_dir = cosine_hemisphere_pdf(abs(dot(to, eye.normal[E(e - 1)]))); ldotn = dot(to, eye.normal[E(e - 1)])
                    eye.rpdf[E(e - 2)] = abs(pdf_dir * ldotn) / (dist * dist)
                else:
                    if eye.type[E(e - 1)] == VERTEX_SURFACE:
                        wi = light.pos[L(l - 1)] - eye.pos[E(e - 1)]; wo = eye.pos[E(e - 2)] - eye.pos[E(e - 1)]
                        dist = length(wo); wi = normalized(wi); wo = normalized(wo)
                        mat_id = int(eye.mat[E(e - 1)]); _, metal, rough = self.mat_rows(mat_id)
                        pdf = disney_pdf(eye.snormal[E(e - 1)], wi, wo, metal, rough)
                        eye.rpdf[E(e - 2)] = pdf / (dist * dist)
                        if eye.type[E(e - 2)] == VERTEX_SURFACE:
                            eye.rpdf[E(e - 2)] *= abs(dot(eye.normal[E(e - 1)], wo))
                    else:
                        eye.rpdf[E(e - 2)] = 1.0
            if l > 1:
                if eye.type[E(e - 1)] != VERTEX_LIGHT:
                    wi = eye.pos[E(e - 1)] - light.pos[L(l - 1)]; wo = light.pos[L(l - 2)] - light.pos[L(l - 1)]
                    dist = length(wo); wi = normalized(wi); wo = normalized(wo)
                    pdf = f32(1.0); mat_id = int(light.mat[L(l - 1)])
                    if mat_id == MAT_DISNEY:
                        _, metal, rough = self.mat_rows(mat_id); pdf = disney_pdf(light.normal[L(l - 1)], wi, wo, metal, rough)
                    light.rpdf[L(l - 2)] = pdf / (dist * dist)
                    if light.type[L(l - 2)] == VERTEX_SURFACE:
                        light.rpdf[L(l - 2)] *= abs(dot(light.normal[L(l - 1)], wo))
                else:
                    light.rpdf[L(l - 2)] = 1.0
            weight = f32(1.0); k = e - 1
            while k > 0:
                weight *= remap0(eye.rpdf[k]) / remap0(eye.fpdf[k])
                if eye.delta[k] == 0 and eye.delta[k - 1] == 0: weight_sum += weight
                k -= 1
            weight = f32(1.0); k = l - 1
            while k >= 0:
                weight *= remap0(light.rpdf[k]) / remap0(light.fpdf[k])
                if k == 0:
                    if light.delta[k] == 0: weight_sum += weight
                elif light.delta[k] == 0 and light.delta[k - 1] == 0: weight_sum += weight
                k -= 1
            light.copy(l - 1, self.ltemp, 0)
            eye.copy(e - 1, self.etemp, 0)
            if l > 0: light.copy(l - 2, self.lminustemp, 0)
            if e > 0: eye.copy(e - 2, self.eminustemp, 0)
        return f32(1.0) / (f32(1.0) + weight_sum)

    # ---- BDPT_RGB.py:436-580
    def connect_path(self, e, l, rng):
        eye, light, sample = self.eye, self.light, self.sample
        radiance = np.zeros(3, f32); new_pos = (self.i, self.j); misweight = f32(1.0)
        if l == 0:
            if eye.type[e - 1] == VERTEX_LIGHT: radiance = eye.beta[e - 1].copy()
        elif e == 1:
            prim = int(light.prim[l - 1]); surface = light.pos[l - 1]
            u, v, wi = self.image_point(surface); new_pos = (u, v)
            mat_id = int(light.mat[l - 1]); snormal = light.snormal[l - 1]
            with np.errstate(all="ignore"):
                ndotl = dot(wi, snormal)
            if u >= 0 and light.delta[l - 1] != 1 and ndotl < 0.0 and light.type[l - 1] == VERTEX_SURFACE:
                t, hit_prim = self.shadow(self.eye_pos, wi)
                if hit_prim == prim:
                    color, metal, rough = self.mat_rows(mat_id)
                    brdf, pdf = disney_evaluate_pdf(snormal, -light.wo[l - 1], -wi, metal, rough)
                    if pdf > 0.0:
                        G = abs(ndotl) / (t * t)
                        radiance = (((G * light.beta[l - 1]) * srgb_to_lrgb(color)) * brdf) / pdf
                        sample.pos[0] = self.eye_pos; sample.wo[0] = wi; sample.type[0] = VERTEX_LENS; sample.fpdf[0] = 1.0
        elif l == 1:
            surface = offset_ray(eye.pos[e - 1], eye.snormal[e - 1]); mat_id = int(eye.mat[e - 1])
            if eye.delta[e - 1] != 1:
                r0 = rng(1 + 2 * (e - 2))
                lpos, lnormal, wi, emission, ldist, lprim, choice_pdf = self.sample_li(surface, r0[0], r0[1], r0[2])
                ndotll, ndotle = dot(wi, lnormal), dot(wi, eye.snormal[e - 1])
                t, shadow_prim = self.shadow(surface, -wi)
                if shadow_prim == lprim and t > EPS:
                    light_pdf = choice_pdf; color, metal, rough = self.mat_rows(mat_id)
                    brdf, pdf = disney_evaluate_pdf(eye.snormal[e - 1], -eye.wo[e - 1], -wi, metal, rough)
                    if pdf > 0.0:
                        G = abs(ndotle * ndotll) / (t * t)
                        radiance = (((((G * eye.beta[e - 1]) * brdf) / pdf) * srgb_to_lrgb(color)) * emission) / light_pdf
                    sample.pos[0] = lpos; sample.wo[0] = wi; sample.type[0] = VERTEX_LIGHT; sample.fpdf[0] = light_pdf
                    sample.prim[0] = lprim; sample.normal[0] = lnormal; sample.snormal[0] = lnormal
        else:
            if light.delta[l - 1] != 1 and eye.delta[e - 1] != 1 and eye.type[e - 1] == VERTEX_SURFACE and light.type[l - 1] == VERTEX_SURFACE:
                prim_e = int(eye.prim[e - 1]); mat_e, mat_l = int(eye.mat[e - 1]), int(light.mat[l - 1])
                d = eye.pos[e - 1] - light.pos[l - 1]; dist = length(d); d = d / dist
                ndotll, ndotle = dot(d, light.snormal[l - 1]), dot(d, eye.snormal[e - 1])
                t, shadow_prim = self.shadow(light.pos[l - 1], d)
                if shadow_prim == prim_e and t > EPS:
                    col_l, metal_l, rough_l = self.mat_rows(mat_l); col_e, metal_e, rough_e = self.mat_rows(mat_e)
                    brdf_l, lpdf = disney_evaluate_pdf(light.snormal[l - 1], -light.wo[l - 1], d, metal_l, rough_l)
                    brdf_e, epdf = disney_evaluate_pdf(eye.snormal[e - 1], -eye.wo[e - 1], -d, metal_e, rough_e)
                    if brdf_l > 0.0 and brdf_e > 0.0:
                        G = abs(ndotle * ndotll) / (dist * dist)
                        radiance = (((((((G * eye.beta[e - 1]) * light.beta[l - 1]) * brdf_l) / lpdf) * brdf_e) / epdf) * srgb_to_lrgb(col_e)) * srgb_to_lrgb(col_l)
        if radiance[0] > 0.0 and radiance[1] > 0.0 and radiance[2] > 0.0:
            misweight = self.mis_weight(e, l)
        return radiance * misweight, new_pos

    def all_strategies(self, rng):
        """the loop of BDPT.render (:625-637) -> {(e, l): (rgb, (u, v))}"""
        out = {}
        for e in range(1, self.eye_depth + 1):
            for l in range(0, self.light_depth + 1):
                depth = l + e - 2
                if (l == 1 and e == 1) or depth < 0 or depth > MAX_DEPTH:
                    continue
                out[(e, l)] = self.connect_path(e, l, rng)
        return out
