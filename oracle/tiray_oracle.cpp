// tiray_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A literal CPU restatement, in f32 scalar C++, of the hot path of lyd405121/ti-raytrace
// (Taichi 0.7.14 kernels) that the CUDA library in ti-raytrace_b200/csrc replaces.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product path never does.
//
// Pinning status (see DESIGN.md "Oracle"):
//   * LBVH build (Morton, sort order, Karras topology, AABBs, flatten order): PINNED by the
//     reference's own golden vector nodelist.txt (tests/golden/nodelist.txt, 71/71 lines).
//   * Traversal / intersection / shading arithmetic: restated line by line from the
//     reference sources cited at every function; the reference's own runtime (Taichi,
//     taichi_glsl) is not installable here, and ti.random() is per-runtime-thread, so
//     per-pixel radiance is "parity unpinned" against Taichi itself.  It is pinned only
//     statistically against the reference's out.png (tests/golden/out.png).
//   * RNG: the reference uses ti.random() (xorshift per thread, unreproducible by design,
//     SURVEY F4); oracle and CUDA share Philox4x32-10 keyed by (seed; pixel, frame, block).
//
// Two builds of this file exist (oracle/Makefile):
//   liboracle.so       -O2 -ffp-contract=off -fno-fast-math   (parity checker, IEEE f32)
//   liboracle_fast.so  -O3 -march=native -ffast-math -fopenmp (CPU baseline timing;
//                       Taichi's default is fast_math=True)
//
// All citations are file:line in /root/reference (commit 70ccd57).

#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#include "../include/trmath.h"
#endif

namespace {

// ---------------------------------------------------------------- constants
// UtilsFunc.py:36-38
constexpr float REF_PIf     = 3.1415956f;   // (sic) reference typo, kept
constexpr float INF_VALUE = 1000000.0f;
// SceneData.py:33-55
constexpr int MAT_N = 10, VER_N = 9, PRI_N = 3, SHA_N = 10, NOD_N = 11, CPN_N = 9;
constexpr int SHAPE_SPHERE = 1, SHAPE_SPOT = 3, SHAPE_LASER = 4;
constexpr int PRIM_TRI = 1;
constexpr int MATT_DISNEY = 0, MATT_GLASS = 1, MATT_LIGHT = 2;

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// Taichi Vector.dot: left-to-right sum of products
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(V3 a) { return sqrtf(dot(a, a)); }
// Taichi 0.7.14 Vector.normalized(): invlen = 1/(norm()+eps), eps=0; return invlen*self
inline V3 normalized(V3 a) { float inv = 1.0f / length(a); return inv * a; }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // taichi_glsl mix
inline float signf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

// ---------------------------------------------------------------- RNG (shared spec with CUDA)
// Philox4x32-10.  counter = (pixel, frame, block, 0), key = (seed_lo, seed_hi).
struct U4 { uint32_t x, y, z, w; };
inline U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n; k0 += W0; k1 += W1;
    }
    return c;
}
inline float u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }
struct F4 { float x, y, z, w; };
inline F4 rng4(uint64_t seed, uint32_t pixel, uint32_t frame, uint32_t block) {
    U4 r = philox4x32_10(U4{pixel, frame, block, 0u}, (uint32_t)seed, (uint32_t)(seed >> 32));
    return F4{u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
}

// ---------------------------------------------------------------- scene container
struct Counters { uint64_t closest = 0, shadow = 0, node_visits = 0, leaf_tests = 0; };

// spectral tables of PT_Spec (integrator/PT_Spec.py:44-98)
struct Spectrum1 { std::vector<float> data; float lmin = 0, lmax = 0, lrange = 1; int size = 0; };
struct SpecData {
    std::vector<float> sensor; float s_lmin = 0, s_lmax = 0, s_lrange = 1; int s_size = 0;   // CIE 1931 2-deg CMFs, n x 3
    Spectrum1 sp[4];                                   // 0 d65, 1 white, 2 red, 3 green
    std::vector<float> rs_scale, rs_data; int rs_res = 0;   // Jakob-Hanika rgb2spec table
    float sky_cfg[11 * 9] = {0}, sky_rad[11] = {0}, sun_dir[3] = {0, 0, 1}; bool sky_on = false;
};

struct Scene {
    std::vector<float>   vertex;    // nv x 9
    std::vector<int32_t> prim;      // np x 3
    std::vector<float>   material;  // nm x 10
    std::vector<float>   shape;     // ns x 10
    std::vector<int32_t> light;     // nl
    int nv = 0, np = 0, nm = 0, ns = 0, nl = 0;
    float bmin[3], bmax[3];
    // build products (reference layouts)
    std::vector<int32_t> morton;    // np x 2  (code, prim)  == morton_code_s
    std::vector<float>   bvh_node;  // (2np-1) x 11
    std::vector<float>   compact;   // (2np-1) x 9
    int refit_sweeps = 0, refit_done = 0;
    // env texture (texture/Texture.py): buf[x][y], packed RGB
    std::vector<int32_t> env; int env_w = 0, env_h = 0; float env_power = 0.0f;
    // camera (Camera.py)
    float view_inv[16]; float eye[3]; float fx, fy, cx, cy;
    float view[16] = {0}; int wid = 0, hgt = 0;      // Camera.view / wid / hgt (BDPT: get_image_point, get_optical_axis)
    int stack_size = 64;
    SpecData spec;
    int max_stack_seen = 0; int overflow = 0;
};

// accessors, UtilsFunc.py:125-198
inline V3 vpos(const Scene& s, int i)    { const float* v = &s.vertex[(size_t)i * VER_N]; return {v[0], v[1], v[2]}; }
inline V3 vnor(const Scene& s, int i)    { const float* v = &s.vertex[(size_t)i * VER_N]; return {v[3], v[4], v[5]}; }
inline V3 vuv(const Scene& s, int i)     { const float* v = &s.vertex[(size_t)i * VER_N]; return {v[6], v[7], v[8]}; }
inline int prim_type(const Scene& s, int i)   { return s.prim[(size_t)i * 3 + 0]; }
inline int prim_vindex(const Scene& s, int i) { return s.prim[(size_t)i * 3 + 1]; }
inline int prim_mindex(const Scene& s, int i) { return s.prim[(size_t)i * 3 + 2]; }
inline int shape_type(const Scene& s, int i)  { return (int)s.shape[(size_t)i * SHA_N]; }
inline V3 shape_pos(const Scene& s, int i)    { const float* p = &s.shape[(size_t)i * SHA_N]; return {p[1], p[2], p[3]}; }
inline float shape_radius(const Scene& s, int i) { return s.shape[(size_t)i * SHA_N + 4]; }
inline int mat_type(const Scene& s, int m)    { return (int)s.material[(size_t)m * MAT_N]; }
inline V3 mat_color(const Scene& s, int m)    { const float* p = &s.material[(size_t)m * MAT_N]; return {p[2], p[3], p[4]}; }
inline float mat_p0(const Scene& s, int m)    { return s.material[(size_t)m * MAT_N + 5]; }  // metal | ior
inline float mat_p1(const Scene& s, int m)    { return s.material[(size_t)m * MAT_N + 6]; }  // rough | extinction

// ---------------------------------------------------------------- Morton (UtilsFunc.py:538-580)
inline int32_t expandBits(int32_t x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8))  & 0x0300F00F;
    x = (x | (x << 4))  & 0x030C30C3;
    x = (x | (x << 2))  & 0x09249249;
    return x;
}
inline int32_t morton3D(float x, float y, float z) {
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    int32_t xx = expandBits((int32_t)x), yy = expandBits((int32_t)y), zz = expandBits((int32_t)z);
    return xx | (yy << 1) | (zz << 2);
}
// UtilsFunc.py:555-566
inline int common_upper_bits(int32_t a, int32_t b) {
    int32_t x = a ^ b; int ret = 32;
    while (x > 0) { x >>= 1; ret -= 1; }
    return ret;
}

// accel/LBvh.py:318-336
void build_morton(Scene& s) {
    s.morton.assign((size_t)s.np * 2, 0);
    V3 mn = {s.bmin[0], s.bmin[1], s.bmin[2]}, mx = {s.bmax[0], s.bmax[1], s.bmax[2]};
    for (int i = 0; i < s.np; ++i) {
        int32_t code;
        if (prim_type(s, i) == PRIM_TRI) {
            int vid = prim_vindex(s, i);
            V3 v0 = vpos(s, vid), v1 = vpos(s, vid + 1), v2 = vpos(s, vid + 2);
            V3 c = ((v1 + v2) + v0) * (1.0f / 3.0f);
            V3 d = mx - mn, n = c - mn;
            code = morton3D(n.x / d.x, n.y / d.y, n.z / d.z);
        } else {
            // reads shape[id][0..2] = (type, pos.x, pos.y) un-normalised (quirk, LBvh.py:333-335)
            const float* sp = &s.shape[(size_t)prim_vindex(s, i) * SHA_N];
            code = morton3D(sp[0], sp[1], sp[2]);
        }
        s.morton[(size_t)i * 2] = code; s.morton[(size_t)i * 2 + 1] = i;
    }
}

// accel/LBvh.py:39-52,177-180
inline int pot_of(int n) { int m = 1; while (m < n) m <<= 1; return (m >> 1) << 1; }
inline int bit_of(int n) { int m = 1, c = 0; while (m < n) { m <<= 1; ++c; } return c; }

// Literal 1-bit x 30-pass LSD radix sort with Blelloch scan (accel/LBvh.py:55-72,339-386)
void radix_sort_literal(Scene& s) {
    int n = s.np, pot = pot_of(n), bit = bit_of(pot);
    if (n < 2) return;
    std::vector<int32_t> off((size_t)pot * 2), dst((size_t)n * 2);
    for (int b = 0; b < 30; ++b) {
        int32_t mask = 1 << b; int32_t zeros = 0;
        for (int i = 0; i < pot; ++i) {
            if (i < n) {
                int32_t one = (s.morton[(size_t)i * 2] & mask) >> b;
                off[(size_t)i * 2 + 1] = one; off[(size_t)i * 2] = 1 - one; zeros += 1 - one;
            } else { off[(size_t)i * 2] = 0; off[(size_t)i * 2 + 1] = 0; }
        }
        for (int l = 1; l <= bit; ++l) {            // up-sweep
            int mod = 1 << l;
            for (int i = 0; i < pot; ++i) if ((i + 1) % mod == 0) {
                int p = i - (mod >> 1);
                off[(size_t)i * 2] += off[(size_t)p * 2]; off[(size_t)i * 2 + 1] += off[(size_t)p * 2 + 1];
            }
        }
        for (int l = bit + 1; l > 0; --l) {         // down-sweep
            long mod = 1L << l;
            if (mod == (long)pot * 2) { off[(size_t)(pot - 1) * 2] = 0; off[(size_t)(pot - 1) * 2 + 1] = 0; continue; }
            for (int i = 0; i < pot; ++i) if ((i + 1) % mod == 0) {
                int p = i - (int)(mod >> 1);
                if (p >= 0) {
                    int32_t t0 = off[(size_t)p * 2], t1 = off[(size_t)p * 2 + 1];
                    off[(size_t)p * 2] = off[(size_t)i * 2]; off[(size_t)p * 2 + 1] = off[(size_t)i * 2 + 1];
                    off[(size_t)i * 2] += t0; off[(size_t)i * 2 + 1] += t1;
                }
            }
        }
        for (int i = 0; i < n; ++i) {               // fill
            int32_t one = (s.morton[(size_t)i * 2] & mask) >> b;
            int32_t o = one ? off[(size_t)i * 2 + 1] + zeros : off[(size_t)i * 2];
            dst[(size_t)o * 2] = s.morton[(size_t)i * 2]; dst[(size_t)o * 2 + 1] = s.morton[(size_t)i * 2 + 1];
        }
        std::copy(dst.begin(), dst.end(), s.morton.begin());
    }
}

// Same result, O(n log n): stable sort on the low 30 bits (ties keep original prim order).
void radix_sort_fast(Scene& s) {
    int n = s.np; std::vector<std::pair<int32_t, int32_t>> v(n);
    for (int i = 0; i < n; ++i) v[i] = {s.morton[(size_t)i * 2], s.morton[(size_t)i * 2 + 1]};
    std::stable_sort(v.begin(), v.end(), [](const auto& a, const auto& b) {
        return (a.first & 0x3FFFFFFF) < (b.first & 0x3FFFFFFF); });
    for (int i = 0; i < n; ++i) { s.morton[(size_t)i * 2] = v[i].first; s.morton[(size_t)i * 2 + 1] = v[i].second; }
}

// accel/LBvh.py:229-294
inline void determineRange(const Scene& s, int idx, int& r0, int& r1) {
    int n = s.np; r0 = 0; r1 = n - 1;
    auto code = [&](int i) { return s.morton[(size_t)i * 2]; };
    if (idx != 0) {
        int32_t self = code(idx), lc = code(idx - 1), rc = code(idx + 1);
        if (lc == self && rc == self) {
            r0 = idx;
            while (idx < n - 1) {
                idx += 1;
                if (idx >= n - 1) break;
                if (code(idx) != code(idx + 1)) break;
            }
            r1 = idx;
        } else {
            int Ld = common_upper_bits(self, lc), Rd = common_upper_bits(self, rc);
            int d = -1; if (Rd > Ld) d = 1;
            int dmin = std::min(Ld, Rd);
            int lmax = 2, delta = -1, it = idx + d * lmax;
            if (0 <= it && it < n) delta = common_upper_bits(self, code(it));
            while (delta > dmin) {
                lmax <<= 1; it = idx + d * lmax; delta = -1;
                if (0 <= it && it < n) delta = common_upper_bits(self, code(it));
            }
            int l = 0, t = lmax >> 1;
            while (t > 0) {
                it = idx + (l + t) * d; delta = -1;
                if (0 <= it && it < n) delta = common_upper_bits(self, code(it));
                if (delta > dmin) l += t;
                t >>= 1;
            }
            r0 = idx; r1 = idx + l * d;
            if (d < 0) std::swap(r0, r1);
        }
    }
}
// accel/LBvh.py:296-314
inline int findSplit(const Scene& s, int first, int last) {
    auto code = [&](int i) { return s.morton[(size_t)i * 2]; };
    int32_t fc = code(first), lc = code(last); int split = first;
    if (fc != lc) {
        int dn = common_upper_bits(fc, lc); int stride = last - first;
        while (true) {
            stride = (stride + 1) >> 1; int middle = split + stride;
            if (middle < last) { int d = common_upper_bits(fc, code(middle)); if (d > dn) split = middle; }
            if (stride <= 1) break;
        }
    }
    return split;
}

// accel/LBvh.py:389-450 + UtilsFunc.py:219-264 (flag-word arithmetic kept literally)
void build_lbvh(Scene& s) {
    int n = s.np, nn = 2 * n - 1;
    s.bvh_node.assign((size_t)nn * NOD_N, 0.0f);
    auto N = [&](int i) { return &s.bvh_node[(size_t)i * NOD_N]; };
    for (int i = 0; i < nn; ++i) {
        float* p = N(i);
        p[0] = p[1] = p[2] = p[3] = p[4] = -1.0f;
        p[5] = p[6] = p[7] = INF_VALUE; p[8] = p[9] = p[10] = -INF_VALUE;
    }
    for (int i = 0; i < nn; ++i) {
        float* p = N(i);
        if (i >= n - 1) {
            p[0] = (float)((int)p[0] & (0xfffe | 1));     // set_node_type(IS_LEAF)
            p[0] = (float)((int)p[0] & (0x0007 | 1));     // set_node_prim_size(1) -> 7.0
            int pi = s.morton[(size_t)(i - n + 1) * 2 + 1];
            p[4] = (float)pi;
            int vi = prim_vindex(s, pi);
            V3 mn = {0, 0, 0}, mx = {0, 0, 0};
            if (prim_type(s, pi) == PRIM_TRI) {
                V3 a = vpos(s, vi), b = vpos(s, vi + 1), c = vpos(s, vi + 2);
                mn = {fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z)};
                mx = {fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z)};
            } else if (shape_type(s, vi) == SHAPE_SPHERE) {
                V3 c = shape_pos(s, vi); float r = shape_radius(s, vi);
                mn = c + v3(-r, -r, -r); mx = c + v3(r, r, r);
            }
            p[5] = mn.x; p[6] = mn.y; p[7] = mn.z; p[8] = mx.x; p[9] = mx.y; p[10] = mx.z;
        } else {
            p[0] = (float)((int)p[0] & (0xfffe | 0));     // -> 65534.0
            int r0, r1; determineRange(s, i, r0, r1);
            int split = findSplit(s, r0, r1);
            int left = split, right = split + 1;
            if (std::min(r0, r1) == split) left += n - 1;
            if (std::max(r0, r1) == split + 1) right += n - 1;
            p[1] = (float)left; p[2] = (float)right;
            N(left)[3] = (float)i; N(right)[3] = (float)i;
        }
    }
}

// accel/LBvh.py:453-468 + host loop :206-218 (level-synchronous relaxation until fixed point).
// A sweep here reads the state left by the previous sweep (Jacobi); Taichi's in-place sweep may
// finish in fewer launches but the fixed point (min/max unions) is identical.
void gen_aabb(Scene& s) {
    int n = s.np, nn = 2 * n - 1;
    auto N = [&](int i) { return &s.bvh_node[(size_t)i * NOD_N]; };
    auto has_box = [&](int i) { const float* p = N(i); return p[5] <= p[8] && p[6] <= p[9] && p[7] <= p[10]; };
    int done = 0, prev = 0; s.refit_sweeps = 0;
    while (done < n - 1) {
        std::vector<int> ready;
        for (int i = 0; i < nn; ++i) if (!has_box(i)) {
            int l = (int)N(i)[1], r = (int)N(i)[2];
            if (l >= 0 && r >= 0 && has_box(l) && has_box(r)) ready.push_back(i);
        }
        for (int i : ready) {
            float* p = N(i); const float* a = N((int)p[1]); const float* b = N((int)p[2]);
            for (int k = 0; k < 3; ++k) { p[5 + k] = fminf(a[5 + k], b[5 + k]); p[8 + k] = fmaxf(a[8 + k], b[8 + k]); }
            ++done;
        }
        ++s.refit_sweeps;
        if (done == prev) break;
        prev = done;
    }
    s.refit_done = done;
}

// accel/LBvh.py:138-173 (recursion made iterative; same left-first pre-order)
void flatten(Scene& s) {
    int n = s.np, nn = 2 * n - 1;
    s.compact.assign((size_t)nn * CPN_N, 0.0f);
    struct Fr { int node; int slot; int stage; };
    std::vector<Fr> st; int offset = 0;
    st.push_back({0, -1, 0});
    // emulate: ret = offset++; copy; if internal: flatten(left); compact[ret][1] = flatten(right)
    std::vector<int> retslot;  // for right-child bookkeeping
    while (!st.empty()) {
        Fr f = st.back(); st.pop_back();
        if (f.stage == 1) {   // about to flatten the right child of compact slot f.slot
            s.compact[(size_t)f.slot * CPN_N + 1] = (float)offset;
            st.push_back({f.node, -1, 0});
            continue;
        }
        int ret = offset++;
        const float* b = &s.bvh_node[(size_t)f.node * NOD_N];
        float* c = &s.compact[(size_t)ret * CPN_N];
        c[0] = b[0];
        for (int i = 0; i < 6; ++i) c[2 + i] = b[5 + i];
        bool leaf = ((int)b[0]) & 1;
        if (!leaf) {
            st.push_back({(int)b[2], ret, 1});
            st.push_back({(int)b[1], -1, 0});
        } else c[1] = b[4];
    }
}

// ---------------------------------------------------------------- intersection
// UtilsFunc.py:494-523
inline int slabs(V3 o, V3 d, const float* mn, const float* mx) {
    int ret = 1; float tmin = 0.0f, tmax = INF_VALUE;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int i = 0; i < 3; ++i) {
        if (fabsf(dd[i]) < 0.000001f) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) ret = 0;
        } else {
            float ood = 1.0f / dd[i];
            float t1 = (mn[i] - oo[i]) * ood, t2 = (mx[i] - oo[i]) * ood;
            if (t1 > t2) std::swap(t1, t2);
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
            if (tmin > tmax) ret = 0;
        }
    }
    return ret;
}

// Scene.py:603-638
inline void intersect_tri(const Scene& s, V3 o, V3 d, int pid, float& t, float& u, float& v) {
    t = INF_VALUE; u = 0.0f; v = 0.0f;
    int vid = prim_vindex(s, pid);
    V3 v0 = vpos(s, vid), v1 = vpos(s, vid + 1), v2 = vpos(s, vid + 2);
    V3 E1 = v1 - v0, E2 = v2 - v0;
    V3 P = cross(d, E2);
    float det = dot(E1, P);
    V3 T;
    if (det > 0.0f) T = o - v0; else { T = v0 - o; det = -det; }
    if (det > 0.0f) {
        u = dot(T, P);
        if (u >= 0.0f && u <= det) {
            V3 Q = cross(T, E1);
            v = dot(d, Q);
            if (v >= 0.0f && u + v <= det) {
                t = dot(E2, Q);
                float inv = 1.0f / det;
                t *= inv; u *= inv; v *= inv;
            }
        }
    }
}

struct Hit { float t; V3 pos, gn, n, tex; int prim; float u, v; };

// sphere branch shared by intersect_prim / intersect_prim_any (Scene.py:565-596,653-665)
// returns true when the reference's `dis_cp < r` branch is taken; c_out = its scalar `c`
inline bool sphere_t(const Scene& s, V3 o, V3 d, int sid, float& t, float& c_out) {
    float r = shape_radius(s, sid); V3 ce = shape_pos(s, sid);
    V3 oc = ce - o; float oc2 = dot(oc, oc), op = dot(d, oc);
    float cp = sqrtf(oc2 - op * op);
    t = INF_VALUE; c_out = 0.0f;
    if (cp < r) {
        float a = dot(d, d), b = -2.0f * op, c = oc2 - r * r;
        t = (-b - sqrtf(b * b - 4.0f * a * c)) / 2.0f / a;
        c_out = c;
        return true;
    }
    return false;
}

// Scene.py:529-600
inline Hit intersect_prim(const Scene& s, V3 o, V3 d, int pid) {
    Hit h; h.t = INF_VALUE; h.pos = h.n = h.tex = h.gn = {0, 0, 0}; h.prim = pid; h.u = h.v = 0;
    if (prim_type(s, pid) == PRIM_TRI) {
        float t, u, v; intersect_tri(s, o, d, pid, t, u, v); h.t = t; h.u = u; h.v = v;
        if (t < INF_VALUE) {
            int vi = prim_vindex(s, pid);
            float a = 1.0f - u - v, b = u, c = v;
            V3 v1 = vpos(s, vi), v2 = vpos(s, vi + 1), v3_ = vpos(s, vi + 2);
            V3 n1 = vnor(s, vi), n2 = vnor(s, vi + 1), n3 = vnor(s, vi + 2);
            V3 t1 = vuv(s, vi), t2 = vuv(s, vi + 1), t3 = vuv(s, vi + 2);
            V3 v13 = v3_ - v1, v12 = v2 - v1;
            h.gn  = cross(v12, v13);
            h.pos = (a * v1 + b * v2) + c * v3_;
            h.tex = (a * t1 + b * t2) + c * t3;
            h.n   = (a * n1 + b * n2) + c * n3;
        }
    } else {
        int sid = prim_vindex(s, pid);
        if (shape_type(s, sid) == SHAPE_SPHERE) {
            float c, t;
            if (sphere_t(s, o, d, sid, t, c)) {
                h.t = t; h.pos = o + t * d;
                h.n = h.pos - v3(c, c, c);      // (sic) scalar c, Scene.py:595
                h.gn = h.n;
            }
        }
    }
    h.gn = normalized(h.gn); h.n = normalized(h.n);
    return h;
}
// Scene.py:642-669
inline float intersect_prim_any(const Scene& s, V3 o, V3 d, int pid) {
    if (prim_type(s, pid) == PRIM_TRI) { float t, u, v; intersect_tri(s, o, d, pid, t, u, v); return t; }
    int sid = prim_vindex(s, pid);
    if (shape_type(s, sid) == SHAPE_SPHERE) { float c, t; sphere_t(s, o, d, sid, t, c); return t; }
    return INF_VALUE;
}

// Scene.py:702-744 : explicit-stack DFS, unordered & unpruned, right child popped first
inline Hit closet_hit(Scene& s, V3 o, V3 d, std::vector<int>& stack, Counters* cnt) {
    Hit best; best.t = INF_VALUE; best.pos = best.n = best.gn = best.tex = {0, 0, 0}; best.prim = -1; best.u = best.v = 0;
    const int MAX = s.stack_size; stack[0] = 0; int sp = 0; int maxsp = 0;
    while (sp >= 0 && sp < MAX) {
        int ni = stack[sp]; sp -= 1;
        const float* c = &s.compact[(size_t)ni * CPN_N];
        if (((int)c[0]) & 1) {
            if (cnt) cnt->leaf_tests++;
            int pid = (int)c[1];
            Hit h = intersect_prim(s, o, d, pid);
            if (h.t < best.t && h.t > 0.0f) best = h;
        } else {
            if (cnt) cnt->node_visits++;
            if (slabs(o, d, c + 2, c + 5) == 1) {
                stack[++sp] = ni + 1; stack[++sp] = (int)c[1];
                if (sp > maxsp) maxsp = sp;
            }
        }
    }
    if (sp == MAX) s.overflow = 1;
    if (maxsp > s.max_stack_seen) s.max_stack_seen = maxsp;
    return best;
}
// Scene.py:671-699
inline void closet_hit_shadow(Scene& s, V3 o, V3 d, std::vector<int>& stack, float& ht, int& hp, Counters* cnt) {
    ht = INF_VALUE; hp = -1; const int MAX = s.stack_size; stack[0] = 0; int sp = 0;
    while (sp >= 0 && sp < MAX) {
        int ni = stack[sp]; sp -= 1;
        const float* c = &s.compact[(size_t)ni * CPN_N];
        if (((int)c[0]) & 1) {
            if (cnt) cnt->leaf_tests++;
            int pid = (int)c[1];
            float t = intersect_prim_any(s, o, d, pid);
            if (t < ht && t > 0.0f) { ht = t; hp = pid; }
        } else {
            if (cnt) cnt->node_visits++;
            if (slabs(o, d, c + 2, c + 5) == 1) { stack[++sp] = ni + 1; stack[++sp] = (int)c[1]; }
        }
    }
}

// ---------------------------------------------------------------- areas / lights
// Scene.py:324-350
inline float get_prim_area(const Scene& s, int idx) {
    float ret = 0.0f;
    if (prim_type(s, idx) == PRIM_TRI) {
        int vi = prim_vindex(s, idx);
        V3 v1 = vpos(s, vi), v2 = vpos(s, vi + 1), v3_ = vpos(s, vi + 2);
        float a = length(v1 - v2), b = length(v1 - v3_), c = length(v3_ - v2);
        float sum = ((a + b) + c) * 0.5f;
        ret = sqrtf(sum * (sum - a) * (sum - b) * (sum - c));
    } else {
        int sid = prim_vindex(s, idx); int st = shape_type(s, sid);
        if (st == SHAPE_SPHERE || st == SHAPE_SPOT || st == SHAPE_LASER) {
            float r = shape_radius(s, sid); ret = r * r * 3.1415926f;
        }
    }
    return ret;
}
// Scene.py:353-377
inline float get_prim_angle(const Scene& s, int idx, V3 v) {
    float ret = 0.0f;
    if (prim_type(s, idx) == PRIM_TRI) {
        int vi = prim_vindex(s, idx);
        V3 v1 = vpos(s, vi), v2 = vpos(s, vi + 1), v3_ = vpos(s, vi + 2);
        if (length(v1 - v) < 0.00001f)      ret = dot(normalized(v2 - v1), normalized(v3_ - v1));
        else if (length(v2 - v) < 0.00001f) ret = dot(normalized(v1 - v2), normalized(v3_ - v2));
        else                                ret = dot(normalized(v1 - v3_), normalized(v2 - v3_));
    }
    return tr_acosf(ret);
}
// Scene.py:315-322
inline V3 UniformSampleSphere(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float r = sqrtf(clampf(1.0f - z * z, 0.0f, 1.0f));
    float phi = 2.0f * 3.1415926f * u2;
    return {r * tr_cosf(phi), r * tr_sinf(phi), z};
}
// UtilsFunc.py:348-350
inline float CosineHemisphere_pdf(float c) { return fmaxf(0.01f, c / REF_PIf); }

struct LightSample { V3 pos, normal, dir, emission; float dist; int prim; float choice_pdf, dir_pdf; };
// Scene.py:477-518 with :423-428 and :381-420 inlined; randoms: u_idx, a, b
inline LightSample sample_li(const Scene& s, V3 p, float u_idx, float a, float b, bool li_terms = true) {
    LightSample L;
    int index = (int)(u_idx * (float)s.nl); if (index >= s.nl) index = s.nl - 1;
    int pi = s.light[index];
    V3 pos = {0, 0, 0}, nor = {0, 0, 0};
    if (prim_type(s, pi) == PRIM_TRI) {
        int vi = prim_vindex(s, pi);
        V3 v1 = vpos(s, vi), v2 = vpos(s, vi + 1), v3_ = vpos(s, vi + 2);
        V3 n1 = vnor(s, vi), n2 = vnor(s, vi + 1), n3 = vnor(s, vi + 2);
        if (a + b > 1.0f) { a = 1.0f - a; b = 1.0f - b; }
        pos = (v1 + (v3_ - v1) * a) + (v2 - v1) * b;
        nor = normalized(((1.0f - a - b) * n1 + n2 * a) + n3 * b);
    } else {
        int sid = prim_vindex(s, pi); int st = shape_type(s, sid);
        if (st == SHAPE_SPHERE) {
            float r = shape_radius(s, sid); V3 ce = shape_pos(s, sid);
            nor = UniformSampleSphere(a, b); pos = ce + nor * r;
        } else if (st == SHAPE_SPOT || st == SHAPE_LASER) {
            const float* sp = &s.shape[(size_t)sid * SHA_N]; nor = {sp[7], sp[8], sp[9]}; pos = shape_pos(s, sid);
        }
    }
    nor = normalized(nor);                 // get_prim_random_point_normal returns normal.normalized()
    int mid = prim_mindex(s, pi);
    L.emission = mat_color(s, mid);
    float area = get_prim_area(s, pi);
    L.choice_pdf = 1.0f / ((float)s.nl * area);
    nor = normalized(nor);                 // Scene.py:486
    V3 dir = p - pos; float dist = length(dir); dir = dir / dist;
    float NdotL = fabsf(dot(dir, nor));
    L.dir_pdf = CosineHemisphere_pdf(NdotL);
    L.pos = pos; L.normal = nor; L.dir = dir; L.dist = dist; L.prim = pi;
    if (li_terms && prim_type(s, pi) != PRIM_TRI) {     // Scene.py:493-516: spot falloff / laser radius cut-off scale the emission
        int sid = prim_vindex(s, pi); int st = shape_type(s, sid);
        const float* sp = &s.shape[(size_t)sid * SHA_N];
        float visable = 1.0f;
        if (st == SHAPE_SPOT) {
            L.dir_pdf = 1.0f;
            float x1 = sp[4], x2 = sp[5];   // UF.get_shape_xita
            float x = tr_acosf(NdotL);
            if (x > x2) visable = 0.0f;
            else if (x > x1) visable *= 1.0f - (x - x1) / (x2 - x1);
        } else if (st == SHAPE_LASER) {
            L.choice_pdf = 1.0f / (float)s.nl;
            float proj = dot(dir, nor) * dist;
            float r = sqrtf(dist * dist - proj * proj);
            float limit_r = shape_radius(s, sid);
            if (r > limit_r) visable = 0.0f;
            L.dir_pdf = 1.0f;
        }
        L.emission = L.emission * visable;
    }
    return L;
}

// ---------------------------------------------------------------- BRDF helpers
// UtilsFunc.py:352-360
inline V3 CosineSampleHemisphere(float u1, float u2) {
    float r = sqrtf(u1), phi = 2.0f * REF_PIf * u2;
    V3 p; p.x = r * tr_cosf(phi); p.y = r * tr_sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return normalized(p);
}
// UtilsFunc.py:373-387
inline V3 inverse_transform(V3 dir, V3 N) {
    V3 Nn = normalized(N), B;
    if (fabsf(Nn.x) > fabsf(Nn.z)) B = {-Nn.y, Nn.x, 0.0f}; else B = {0.0f, -Nn.z, Nn.y};
    B = normalized(B);
    V3 T = normalized(cross(B, Nn));
    return (dir.x * T + dir.y * B) + dir.z * Nn;
}
inline float SchlickFresnel(float u) { float m = clampf(1.0f - u, 0.0f, 1.0f); float m2 = m * m; return m2 * m2 * m; }
inline float GTR2(float NDotH, float a) { float a2 = a * a; float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH; return a2 / (REF_PIf * t * t); }
inline float smithG_GGX(float NDotv, float alphaG) { float a = alphaG * alphaG, b = NDotv * NDotv; return 1.0f / (NDotv + sqrtf(a + b - a * b)); }
inline V3 reflect(V3 I, V3 N) { return I - 2.0f * dot(N, I) * N; }   // taichi_glsl reflect
// UtilsFunc.py:417-432
inline V3 refract(V3 I, V3 N, float eta, float& suc) {
    suc = -1.0f; float NI = dot(N, I); float k = 1.0f - eta * eta * (1.0f - NI * NI);
    V3 R = {0, 0, 0};
    if (k > 0.0f) { R = eta * I - (eta * NI + sqrtf(k)) * N; suc = 1.0f; }
    return R;
}
inline float schlick(float cosine, float ior) { float r0 = (1.0f - ior) / (1.0f + ior); r0 = r0 * r0; return r0 + (1.0f - r0) * tr_powf(1.0f - cosine, 5.0f); }
inline float powerHeuristic(float a, float b) { float t = a * a; return t / (b * b + t); }
// UtilsFunc.py:440-461
inline V3 offset_ray(V3 p, V3 n) {
    const float int_scale = 256.0f, float_scale = 1.0f / 2048.0f, origin = 1.0f / 256.0f;
    float pp[3] = {p.x, p.y, p.z}, nn[3] = {n.x, n.y, n.z}, r[3];
    for (int k = 0; k < 3; ++k) {
        int32_t i_of = (int32_t)(int_scale * nn[k]);
        int32_t i_p; memcpy(&i_p, &pp[k], 4);
        if (pp[k] < 0.0f) i_p -= i_of; else i_p += i_of;
        float f_p; memcpy(&f_p, &i_p, 4);
        r[k] = (fabsf(pp[k]) < origin) ? pp[k] + float_scale * nn[k] : f_p;
    }
    return {r[0], r[1], r[2]};
}
// UtilsFunc.py:76-94,113-120
inline float srgb_to_lrgb1(float c) { return c < 0.04045f ? c / 12.92f : tr_powf((c + 0.055f) / 1.055f, 2.4f); }
inline V3 srgb_to_lrgb(V3 c) { return {srgb_to_lrgb1(c.x), srgb_to_lrgb1(c.y), srgb_to_lrgb1(c.z)}; }
inline float lrgb_to_srgb1(float c) { float r = c < 0.0031308f ? c * 12.92f : 1.055f * tr_powf(c, 1.0f / 2.4f) - 0.055f; return clampf(r, 0.0f, 1.0f); }
inline float tone_ACES1(float x) { const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f; return clampf((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f); }

// brdf/Disney.py:65-108
inline void disney_evaluate_pdf(V3 N, V3 V, V3 L, float metal, float rough, float& out, float& pdf) {
    out = 0.0f; pdf = -1.0f;
    float NDotL = dot(N, L), NDotV = dot(N, V);
    if (NDotL > 0.0f && NDotV > 0.0f) {
        V3 H = normalized(L + V);
        float NDotH = dot(H, N), LDotH = dot(H, L);
        float Cspec0 = mixf(0.04f, 1.0f, metal), Csheen = 0.5f;
        float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
        float Fd90 = 0.5f + 2.0f * LDotH * LDotH * rough;
        float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
        float alpha = fmaxf(0.001f, rough);
        float Ds = GTR2(NDotH, alpha);
        float FH = SchlickFresnel(LDotH);
        float Fs = mixf(Cspec0, 1.0f, FH);
        float rg = rough * 0.5f + 0.5f; rg = rg * rg;
        float Gs = smithG_GGX(NDotL, rg) * smithG_GGX(NDotV, rg);
        float Fsheen = FH * Csheen;
        out = (Fsheen + 1.0f / REF_PIf) * Fd * (1.0f - metal) + Gs * Fs * Ds;
        float dr = 0.5f * (1.0f - metal), sr = 1.0f - dr;
        float pdfGTR2 = Ds * NDotH, pdfSpec = pdfGTR2 / (4.0f * fabsf(LDotH)), pdfDiff = 1.0f / REF_PIf;
        pdf = dr * pdfDiff + sr * pdfSpec;
    }
}
// brdf/Disney.py:17-40  randoms: probability, r1, r2
inline V3 disney_sample(V3 dir, V3 N, float metal, float rough, float prob, float r1, float r2) {
    float dr = 0.5f * (1.0f - metal), alpha = fmaxf(0.001f, rough);
    V3 next;
    if (prob < dr) {
        next = inverse_transform(CosineSampleHemisphere(r1, r2), N);
    } else {
        float phi = r1 * 2.0f * REF_PIf;
        float cosT = sqrtf((1.0f - r2) / (1.0f + (alpha * alpha - 1.0f) * r2));
        float sinT = sqrtf(1.0f - (cosT * cosT));
        float sinP = tr_sinf(phi), cosP = tr_cosf(phi);
        V3 half = inverse_transform(v3(sinT * cosP, sinT * sinP, cosT), N);
        next = reflect(dir, half);
    }
    return next;
}
// brdf/Glass.py:9-34  random: probability
inline V3 glass_sample(V3 dir, V3 N, float ior, float prob, float& f_or_b) {
    float cos_i = dot(dir, N), eta = ior; f_or_b = 1.0f; float R = prob + 1.0f;
    if (cos_i > 0.0f) N = -N; else { cos_i = -cos_i; eta = 1.0f / ior; }
    float suc; V3 next = refract(dir, N, eta, suc);
    if (suc > 0.0f) R = schlick(cos_i, ior);
    if (prob < R) next = reflect(dir, N); else f_or_b = -1.0f;
    return next;
}

// texture/Texture.py:36-69
inline V3 tex_sample(const Scene& s, float fx, float fy) {
    int x = std::min(std::max((int)fx, 0), s.env_w - 1), y = std::min(std::max((int)fy, 0), s.env_h - 1);
    int32_t c = s.env[(size_t)x * s.env_h + y];
    return {(float)((c & 0x00FF0000) >> 16) / 255.0f, (float)((c & 0x0000FF00) >> 8) / 255.0f, (float)(c & 0xFF) / 255.0f};
}
inline V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
inline V3 texture2D(const Scene& s, float u, float v) {
    float x = clampf(u * (float)s.env_w, 0.0f, (float)s.env_w - 1.0f), y = clampf(v * (float)s.env_h, 0.0f, (float)s.env_h - 1.0f);
    float lx = floorf(x), ly = floorf(y);
    float wbt = y - floorf(y), wlr = x - floorf(x);
    return mix3(mix3(tex_sample(s, lx, ly), tex_sample(s, lx + 1, ly), wlr),
                mix3(tex_sample(s, lx, ly + 1), tex_sample(s, lx + 1, ly + 1), wlr), wbt);
}

// Camera.py:122-142
inline V3 ray_direction(const Scene& s, int i, int j, float jx, float jy) {
    float x = ((float)i + jx - s.cx) / s.fx, y = ((float)j + jy - s.cy) / s.fy, z = -1.0f;
    const float* m = s.view_inv;
    V3 w = {(m[0] * x + m[1] * y) + m[2] * z, (m[4] * x + m[5] * y) + m[6] * z, (m[8] * x + m[9] * y) + m[10] * z};
    return normalized(w);
}

// integrator/PT_RGB.py:44-132 for one pixel, one frame.  RNG blocks: 0 = jitter,
// 1+2*depth = (light index, a, b, lobe/Fresnel probability), 2+2*depth = (r1, r2, absorption, -)
V3 pt_rgb_pixel(Scene& s, int i, int j, int frame, int max_depth, uint64_t seed, std::vector<int>& stack, Counters& cnt) {
    uint32_t pixel_id = (uint32_t)i * 65536u + (uint32_t)j;   // RNG key: independent of image size and sharding
    float jx = 0.0f, jy = 0.0f;
    if (frame != 0) { F4 r = rng4(seed, pixel_id, (uint32_t)frame, 0); jx = r.x - 0.5f; jy = r.y - 0.5f; }
    V3 next_o = {s.eye[0], s.eye[1], s.eye[2]}, next_d = ray_direction(s, i, j, jx, jy);
    int depth = 0; float light_pdf = 1.0f, brdf_pdf = 1.0f; int perfect_spec = 1; float f_or_b = 1.0f, brdf = 1.0f;
    V3 T = {1, 1, 1}, L = {0, 0, 0};
    while (depth < max_depth) {
        V3 o = next_o, d = next_d;
        cnt.closest++;
        Hit h = closet_hit(s, o, d, stack, &cnt);
        if (h.t < INF_VALUE) {
            V3 fn = signf(dot(-d, h.gn)) * h.n;                       // UF.faceforward(normal, -direction, gnormal)
            int mid = prim_mindex(s, h.prim);
            V3 mcol = mat_color(s, mid); int mt = mat_type(s, mid);
            if (mt == MATT_LIGHT) {
                float fCos = fabsf(dot(d, h.gn));
                if (perfect_spec == 1) L = L + T * mcol;
                else {
                    float area = get_prim_area(s, h.prim) * (float)s.nl;
                    light_pdf = (h.t * h.t) / (area * fCos);
                    L = L + powerHeuristic(brdf_pdf, light_pdf) * T * mcol;
                }
                break;
            }
            V3 rc = srgb_to_lrgb(mcol);
            F4 R0 = rng4(seed, pixel_id, (uint32_t)frame, 1u + 2u * (uint32_t)depth);
            F4 R1 = rng4(seed, pixel_id, (uint32_t)frame, 2u + 2u * (uint32_t)depth);
            if (mt == MATT_GLASS) {
                perfect_spec = 1;
                next_d = glass_sample(d, h.n, mat_p0(s, mid), R0.w, f_or_b);
                brdf = 1.0f; brdf_pdf = 1.0f;
            } else {
                perfect_spec = 0;
                float metal = mat_p0(s, mid), rough = mat_p1(s, mid);
                if (s.nl > 0) {
                    LightSample ls = sample_li(s, h.pos, R0.x, R0.y, R0.z);
                    float NdotL_s = dot(fn, ls.dir), NdotL_l = dot(ls.normal, ls.dir);
                    if (NdotL_s < 0.0f && NdotL_l > 0.0f) {
                        float st; int sp; cnt.shadow++;
                        closet_hit_shadow(s, ls.pos, ls.dir, stack, st, sp, &cnt);
                        if (sp == h.prim) {
                            disney_evaluate_pdf(fn, -d, -ls.dir, metal, rough, brdf, brdf_pdf);
                            light_pdf = ls.dist * ls.dist * ls.choice_pdf / NdotL_l;
                            if (brdf_pdf > 0.0f) {
                                float w = powerHeuristic(light_pdf, brdf_pdf) / fmaxf(0.0001f, light_pdf);
                                L = L + ((((w * ls.emission) * T) * rc) * brdf) * fabsf(NdotL_s);
                            }
                        }
                    }
                }
                f_or_b = 1.0f;
                next_d = disney_sample(d, fn, metal, rough, R0.w, R1.x, R1.y);
                disney_evaluate_pdf(fn, -d, next_d, metal, rough, brdf, brdf_pdf);
                brdf *= fabsf(dot(h.n, next_d));
            }
            next_o = offset_ray(h.pos, signf(f_or_b) * fn);
            if (brdf_pdf > 0.0f) {
                if (f_or_b < 0.0f) {
                    float ext = mat_p1(s, mid); float R = tr_expf(-h.t / ext);
                    if (R1.z >= R) break;
                }
                T = T * ((brdf / brdf_pdf) * rc);
                depth += 1;
            } else break;
        } else {
            float dis = sqrtf(d.x * d.x + d.z * d.z);
            float tx = (tr_atan2f(d.z, d.x) + 3.1415926f) / 3.1415926f / 2.0f;
            float ty = tr_atan2f(d.y, dis) / 3.1415926f + 0.5f;
            if (s.env_w > 0) L = L + (srgb_to_lrgb(texture2D(s, tx, ty)) * T) * s.env_power;
            break;
        }
    }
    return L;
}

#include "spec_core.inc"
#include "bdpt_core.inc"

}  // namespace

// =============================================================================== C API
extern "C" {

void* orc_scene_create(const float* vertex, int nv, const int32_t* prim, int np, const float* material, int nm,
                       const float* shape, int ns, const int32_t* light, int nl, const float* bmin, const float* bmax) {
    Scene* s = new Scene();
    s->nv = nv; s->np = np; s->nm = nm; s->ns = ns; s->nl = nl;
    s->vertex.assign(vertex, vertex + (size_t)nv * VER_N);
    s->prim.assign(prim, prim + (size_t)np * PRI_N);
    s->material.assign(material, material + (size_t)nm * MAT_N);
    if (ns > 0) s->shape.assign(shape, shape + (size_t)ns * SHA_N); else s->shape.assign(SHA_N, 0.0f);
    if (nl > 0) s->light.assign(light, light + nl);
    for (int k = 0; k < 3; ++k) { s->bmin[k] = bmin[k]; s->bmax[k] = bmax[k]; }
    memset(s->view_inv, 0, sizeof(s->view_inv)); s->eye[0] = s->eye[1] = s->eye[2] = 0; s->fx = s->fy = 1; s->cx = s->cy = 0;
    return s;
}
void orc_scene_destroy(void* h) { delete (Scene*)h; }

// literal != 0 -> 30 x 1-bit Blelloch passes; else stable sort (identical result)
int orc_bvh_build(void* h, int literal_sort) {
    Scene& s = *(Scene*)h;
    build_morton(s);
    if (literal_sort) radix_sort_literal(s); else radix_sort_fast(s);
    build_lbvh(s); gen_aabb(s); flatten(s);
    return (s.refit_done == s.np - 1) ? 0 : -1;   // "aabb gen error", LBvh.py:215-216
}
void orc_morton_unsorted(void* h, int32_t* out) { Scene& s = *(Scene*)h; build_morton(s); memcpy(out, s.morton.data(), s.morton.size() * 4); }
void orc_bvh_get(void* h, int32_t* morton, float* bvh_node, float* compact) {
    Scene& s = *(Scene*)h;
    if (morton)   memcpy(morton, s.morton.data(), s.morton.size() * 4);
    if (bvh_node) memcpy(bvh_node, s.bvh_node.data(), s.bvh_node.size() * 4);
    if (compact)  memcpy(compact, s.compact.data(), s.compact.size() * 4);
}
int orc_refit_sweeps(void* h) { return ((Scene*)h)->refit_sweeps; }
void orc_vertex_get(void* h, float* out) { Scene& s = *(Scene*)h; memcpy(out, s.vertex.data(), s.vertex.size() * 4); }

void orc_camera_set(void* h, const float* view_inv, const float* eye, float fx, float fy, float cx, float cy) {
    Scene& s = *(Scene*)h; memcpy(s.view_inv, view_inv, 64); memcpy(s.eye, eye, 12); s.fx = fx; s.fy = fy; s.cx = cx; s.cy = cy;
}
void orc_env_set(void* h, const int32_t* buf, int w, int hgt, float power) {
    Scene& s = *(Scene*)h; s.env.assign(buf, buf + (size_t)w * hgt); s.env_w = w; s.env_h = hgt; s.env_power = power;
}
void orc_stack_size(void* h, int n) { ((Scene*)h)->stack_size = n; }
int orc_max_stack_seen(void* h) { return ((Scene*)h)->max_stack_seen; }
int orc_overflow(void* h) { return ((Scene*)h)->overflow; }

// Scene.py:747-750
float orc_total_area(void* h) { Scene& s = *(Scene*)h; float a = 0.0f; for (int i = 0; i < s.nl; ++i) a += get_prim_area(s, s.light[i]); return a; }

// Primary rays (Camera.py:122-142, frame 0: no jitter) -> dirs[W*H*3] indexed [i*H+j]
void orc_primary_rays(void* h, int W, int H, float* dirs) {
    Scene& s = *(Scene*)h;
    for (int i = 0; i < W; ++i) for (int j = 0; j < H; ++j) {
        V3 d = ray_direction(s, i, j, 0.0f, 0.0f); float* o = dirs + ((size_t)i * H + j) * 3; o[0] = d.x; o[1] = d.y; o[2] = d.z;
    }
}

// First hits of the frame-0 primary rays with the reference traversal (Scene.py:702-744).
// outputs indexed [i*H+j]: t, prim, uv(2), pos(3), gnormal(3), normal(3); stats[0..3] = rays, node visits, leaf tests, max stack
void orc_first_hit(void* h, int W, int H, float* t, int32_t* prim, float* uv, float* pos, float* gn, float* nrm, uint64_t* stats) {
    Scene& s = *(Scene*)h; Counters tot;
    #pragma omp parallel
    {
        std::vector<int> stack(s.stack_size + 2); Counters c;
        #pragma omp for schedule(dynamic, 64)
        for (int p = 0; p < W * H; ++p) {
            int i = p / H, j = p % H;
            V3 o = {s.eye[0], s.eye[1], s.eye[2]}, d = ray_direction(s, i, j, 0.0f, 0.0f);
            c.closest++;
            Hit hh = closet_hit(s, o, d, stack, &c);
            t[p] = hh.t; prim[p] = hh.prim;
            if (uv)  { uv[p * 2] = hh.u; uv[p * 2 + 1] = hh.v; }
            if (pos) { pos[p * 3] = hh.pos.x; pos[p * 3 + 1] = hh.pos.y; pos[p * 3 + 2] = hh.pos.z; }
            if (gn)  { gn[p * 3] = hh.gn.x; gn[p * 3 + 1] = hh.gn.y; gn[p * 3 + 2] = hh.gn.z; }
            if (nrm) { nrm[p * 3] = hh.n.x; nrm[p * 3 + 1] = hh.n.y; nrm[p * 3 + 2] = hh.n.z; }
        }
        #pragma omp critical
        { tot.closest += c.closest; tot.node_visits += c.node_visits; tot.leaf_tests += c.leaf_tests; }
    }
    if (stats) { stats[0] = tot.closest; stats[1] = tot.node_visits; stats[2] = tot.leaf_tests; stats[3] = (uint64_t)s.max_stack_seen; }
}

// Generic closest-hit / shadow queries for arbitrary rays (unit parity of the traversal kernels)
void orc_trace(void* h, int n, const float* o, const float* d, int shadow, float* t, int32_t* prim, float* uv) {
    Scene& s = *(Scene*)h;
    #pragma omp parallel
    {
        std::vector<int> stack(s.stack_size + 2);
        #pragma omp for schedule(dynamic, 64)
        for (int k = 0; k < n; ++k) {
            V3 oo = {o[k * 3], o[k * 3 + 1], o[k * 3 + 2]}, dd = {d[k * 3], d[k * 3 + 1], d[k * 3 + 2]};
            if (shadow) { float ht; int hp; closet_hit_shadow(s, oo, dd, stack, ht, hp, nullptr); t[k] = ht; prim[k] = hp; }
            else { Hit hh = closet_hit(s, oo, dd, stack, nullptr); t[k] = hh.t; prim[k] = hh.prim; if (uv) { uv[k * 2] = hh.u; uv[k * 2 + 1] = hh.v; } }
        }
    }
}

// integrator/Debug.py:44-66 -> hdr[(i*H+j)*3]
void orc_render_debug(void* h, int W, int H, float* hdr) {
    Scene& s = *(Scene*)h;
    #pragma omp parallel
    {
        std::vector<int> stack(s.stack_size + 2);
        #pragma omp for schedule(dynamic, 64)
        for (int p = 0; p < W * H; ++p) {
            int i = p / H, j = p % H;
            V3 o = {s.eye[0], s.eye[1], s.eye[2]}, d = ray_direction(s, i, j, 0.0f, 0.0f);
            Hit hh = closet_hit(s, o, d, stack, nullptr);
            V3 c = {0, 0, 0};
            if (hh.t < INF_VALUE) c = mat_color(s, prim_mindex(s, hh.prim));
            hdr[p * 3] = c.x; hdr[p * 3 + 1] = c.y; hdr[p * 3 + 2] = c.z;
        }
    }
}

// integrator/PT_RGB.py:44-136: frames [frame_begin, frame_begin+n_frames) accumulated into hdr (running mean).
// Only pixels with mask[p] != 0 are rendered when mask is given (multi-rank tile sharding tests).
// counters[0..3] += closest rays, shadow rays, node visits, leaf tests.
void orc_render_pt_rgb(void* h, int W, int H, int frame_begin, int n_frames, int max_depth, uint64_t seed,
                       float* hdr, const uint8_t* mask, uint64_t* counters) {
    Scene& s = *(Scene*)h; Counters tot;
    #pragma omp parallel
    {
        std::vector<int> stack(s.stack_size + 2); Counters c;
        #pragma omp for schedule(dynamic, 32)
        for (int p = 0; p < W * H; ++p) {
            if (mask && !mask[p]) continue;
            int i = p / H, j = p % H;
            for (int f = frame_begin; f < frame_begin + n_frames; ++f) {
                V3 L = pt_rgb_pixel(s, i, j, f, max_depth, seed, stack, c);
                float coff = 1.0f / ((float)f + 1.0f);           // PT_RGB.py:134-136
                float* o = hdr + (size_t)p * 3;
                o[0] = L.x * coff + o[0] * (1.0f - coff);
                o[1] = L.y * coff + o[1] * (1.0f - coff);
                o[2] = L.z * coff + o[2] * (1.0f - coff);
            }
        }
        #pragma omp critical
        { tot.closest += c.closest; tot.shadow += c.shadow; tot.node_visits += c.node_visits; tot.leaf_tests += c.leaf_tests; }
    }
    if (counters) { counters[0] += tot.closest; counters[1] += tot.shadow; counters[2] += tot.node_visits; counters[3] += tot.leaf_tests; }
}

// UtilsFunc.py:583-586
void orc_tonemap(int n, float exposure, const float* hdr, float* rgb) {
    for (int k = 0; k < n * 3; ++k) rgb[k] = lrgb_to_srgb1(tone_ACES1(hdr[k] * exposure));
}

// Scene.py:754-798 (uses Scene.stack[vertex_count, 32]; loop bound is part of the semantics)
void orc_process_normal(void* h) {
    Scene& s = *(Scene*)h; const int MAXS = 32;
    std::vector<float> smooth((size_t)s.nv * 3);
    std::vector<int32_t> vindex(s.nv);
    for (int p = 0; p < s.np; ++p) if (prim_type(s, p) == PRIM_TRI) { int vi = prim_vindex(s, p); vindex[vi] = vindex[vi + 1] = vindex[vi + 2] = p; }
    #pragma omp parallel
    {
        std::vector<int> stack(MAXS + 2);
        #pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < s.nv; ++i) {
            V3 v = vpos(s, i), n = normalized(vnor(s, i)); int f = vindex[i];
            V3 sm = (n * get_prim_angle(s, f, v)) * get_prim_area(s, f);
            stack[0] = 0; int sp = 0;
            while (sp >= 0 && sp < MAXS) {
                int ni = stack[sp]; sp -= 1;
                const float* c = &s.compact[(size_t)ni * CPN_N];
                if (((int)c[0]) & 1) {
                    int pi = (int)c[1];
                    if (prim_type(s, pi) == PRIM_TRI) {
                        int vi = prim_vindex(s, pi);
                        for (int j = 0; j < 3; ++j) {
                            int nb = j + vi;
                            if (i != nb) {
                                V3 nv = vpos(s, nb), nn = normalized(vnor(s, nb));
                                if (length(v - nv) < 0.000001f && dot(nn, n) > 0.5f) {
                                    float ang = get_prim_angle(s, pi, nv);
                                    sm = sm + (nn * ang) * get_prim_area(s, pi);
                                }
                            }
                        }
                    }
                } else {
                    const float* mn = c + 2; const float* mx = c + 5;
                    if (v.x >= mn[0] && v.y >= mn[1] && v.z >= mn[2] && v.x <= mx[0] && v.y <= mx[1] && v.z <= mx[2]) {
                        stack[++sp] = ni + 1; stack[++sp] = (int)c[1];
                    }
                }
            }
            smooth[(size_t)i * 3] = sm.x; smooth[(size_t)i * 3 + 1] = sm.y; smooth[(size_t)i * 3 + 2] = sm.z;
        }
    }
    for (int i = 0; i < s.nv; ++i) {
        V3 n = normalized(v3(smooth[(size_t)i * 3], smooth[(size_t)i * 3 + 1], smooth[(size_t)i * 3 + 2]));
        float* p = &s.vertex[(size_t)i * VER_N]; p[3] = n.x; p[4] = n.y; p[5] = n.z;
    }
}

// ---- unit hooks (same signatures as the CUDA test hooks in include/tiray.h)
void orc_disney_evaluate_pdf(int n, const float* N, const float* V, const float* L, float metal, float rough, float* out /*n x 2*/) {
    for (int k = 0; k < n; ++k) {
        float o, p; disney_evaluate_pdf(v3(N[k * 3], N[k * 3 + 1], N[k * 3 + 2]), v3(V[k * 3], V[k * 3 + 1], V[k * 3 + 2]),
                                        v3(L[k * 3], L[k * 3 + 1], L[k * 3 + 2]), metal, rough, o, p);
        out[k * 2] = o; out[k * 2 + 1] = p;
    }
}
void orc_disney_sample(int n, const float* dir, const float* N, float metal, float rough, const float* u /*n x 3*/, float* out /*n x 3*/) {
    for (int k = 0; k < n; ++k) {
        V3 r = disney_sample(v3(dir[k * 3], dir[k * 3 + 1], dir[k * 3 + 2]), v3(N[k * 3], N[k * 3 + 1], N[k * 3 + 2]), metal, rough, u[k * 3], u[k * 3 + 1], u[k * 3 + 2]);
        out[k * 3] = r.x; out[k * 3 + 1] = r.y; out[k * 3 + 2] = r.z;
    }
}
void orc_glass_sample(int n, const float* dir, const float* N, float ior, const float* u /*n*/, float* out /*n x 4: dir, f_or_b*/) {
    for (int k = 0; k < n; ++k) {
        float fb; V3 r = glass_sample(v3(dir[k * 3], dir[k * 3 + 1], dir[k * 3 + 2]), v3(N[k * 3], N[k * 3 + 1], N[k * 3 + 2]), ior, u[k], fb);
        out[k * 4] = r.x; out[k * 4 + 1] = r.y; out[k * 4 + 2] = r.z; out[k * 4 + 3] = fb;
    }
}
void orc_offset_ray(int n, const float* p, const float* nrm, float* out) {
    for (int k = 0; k < n; ++k) { V3 r = offset_ray(v3(p[k * 3], p[k * 3 + 1], p[k * 3 + 2]), v3(nrm[k * 3], nrm[k * 3 + 1], nrm[k * 3 + 2])); out[k * 3] = r.x; out[k * 3 + 1] = r.y; out[k * 3 + 2] = r.z; }
}
void orc_rng(uint64_t seed, uint32_t pixel, uint32_t frame, uint32_t block, float* out4) { F4 r = rng4(seed, pixel, frame, block); out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w; }
void orc_philox_raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) { U4 r = philox4x32_10(U4{c0, c1, c2, c3}, k0, k1); out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w; }
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// include/trmath.h on the host (fn 0 sin, 1 cos, 2 exp, 3 acos, 4 atan2(a, b), 5 pow(a, b)): the CUDA kernels compile the same header
void orc_math(int fn, int n, const float* a, const float* b, float* out) {
    for (int k = 0; k < n; ++k) {
        const float x = a[k], y = b ? b[k] : 0.0f;
        out[k] = fn == 0 ? tr_sinf(x) : fn == 1 ? tr_cosf(x) : fn == 2 ? tr_expf(x) : fn == 3 ? tr_acosf(x) : fn == 4 ? tr_atan2f(x, y) : tr_powf(x, y);
    }
}
// bench.py's reference arm: torchrun exports OMP_NUM_THREADS=1, the CPU baseline is defined on ALL host cores
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
int orc_slabs(const float* o, const float* d, const float* mn, const float* mx) { return slabs(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), mn, mx); }

#include "spec_api.inc"
#include "bdpt_api.inc"

}  // extern "C"
