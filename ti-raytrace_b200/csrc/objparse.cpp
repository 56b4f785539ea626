// objparse.cpp — native Wavefront OBJ / MTL ingest behind Scene.add_obj (host code, no CUDA).
//
// Replaces what the reference gets from PyWavefront 1.3.3 plus its own per-vertex Python loops (Scene.py:59-141: seconds
// for the 392 160 vertices of mc.obj + Teapot.obj): one pass over the file in memory, numbers through strtod (correctly
// rounded, i.e. the doubles Python's float() produces), output = per material the de-indexed vertex soup
// (pos3, normal3, tex3 as f64 rows) in the order PyWavefront builds it:
//   * materials in `newmtl` order of the MTL named by `mtllib`; a `usemtl` of an unknown name and faces before any
//     `usemtl` ("default<k>") create materials on the fly with PyWavefront's defaults (Kd .8, d 1, Ke 0, Ns 0, Ni 1)
//   * the vertex format of a material (has vt / has vn) is fixed by its first face; polygons are fanned as
//     (v1, v2, v3), (vj, v1, v(j-1)); negative indices are relative to the current element count
//   * `vt` keeps two components; `d` sets transparency, `Tr` sets 1 - Tr, `Ni` optical_density, `Ns` shininess
// The independent restatement this is checked against bit for bit is oracle/objload.py (tests/test_host.py).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_map>
#include "../../include/tiray.h"

namespace {

struct ObjMat {
    std::string name;
    double diffuse[3] = {0.8, 0.8, 0.8}, emissive[3] = {0.0, 0.0, 0.0};
    double transparency = 1.0, shininess = 0.0, optical_density = 1.0;
    int has_vt = -1, has_vn = -1;
    std::vector<int32_t> corners;        // (v, vt, vn) per corner, three corners per triangle
};

thread_local std::string g_obj_error;

struct Tok { const char* p; int n; };

// whitespace-split one line [b, e) like Python's str.split()
inline void split_line(const char* b, const char* e, std::vector<Tok>& out) {
    out.clear();
    const char* p = b;
    while (p < e) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
        if (p >= e) break;
        const char* q = p;
        while (q < e && !(*q == ' ' || *q == '\t' || *q == '\r' || *q == '\f' || *q == '\v')) ++q;
        out.push_back(Tok{p, (int)(q - p)});
        p = q;
    }
}
inline bool tok_is(const Tok& t, const char* s) { int n = (int)strlen(s); return t.n == n && memcmp(t.p, s, n) == 0; }
// Decimal -> double.  Fast path (Clinger): a mantissa below 2^53 and a power of ten up to 10^22 are both exact doubles, so one
// multiplication or division is correctly rounded -- the same value strtod / Python's float() return.  Everything else
// (long mantissas, big exponents, inf / nan, hex) goes to strtod.
inline bool to_double(const Tok& t, double& v) {
    static const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    {
        const char* p = t.p; const char* e = t.p + t.n;
        bool neg = false;
        if (p < e && (*p == '-' || *p == '+')) { neg = (*p == '-'); ++p; }
        uint64_t mant = 0; int digits = 0, exp10 = 0; bool any = false, ok = true;
        while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant || digits) ++digits; } else ok = false; ++p; any = true; }
        if (p < e && *p == '.') {
            ++p;
            while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant || digits) ++digits; --exp10; } else ok = false; ++p; any = true; }
        }
        if (any && p < e && (*p == 'e' || *p == 'E')) {
            ++p; bool eneg = false; int ev = 0, ed = 0;
            if (p < e && (*p == '-' || *p == '+')) { eneg = (*p == '-'); ++p; }
            while (p < e && *p >= '0' && *p <= '9') { if (ev < 10000) ev = ev * 10 + (*p - '0'); ++p; ++ed; }
            if (ed == 0) ok = false;
            exp10 += eneg ? -ev : ev;
        }
        if (any && ok && p == e && mant <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
            double d = (double)mant;
            d = exp10 < 0 ? d / p10[-exp10] : d * p10[exp10];
            v = neg ? -d : d;
            return true;
        }
    }
    char buf[64]; if (t.n <= 0 || t.n >= (int)sizeof(buf)) return false;
    memcpy(buf, t.p, t.n); buf[t.n] = 0;
    char* end = nullptr; v = strtod(buf, &end);
    return end == buf + t.n;
}
inline bool to_int(const char* p, int n, long& v) {
    char buf[32]; if (n <= 0 || n >= (int)sizeof(buf)) return false;
    memcpy(buf, p, n); buf[n] = 0;
    char* end = nullptr; v = strtol(buf, &end, 10);
    return end == buf + n;
}
bool read_file(const std::string& path, std::vector<char>& data) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    data.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(data.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == data.size();
}

}  // namespace

struct tr_obj {
    std::vector<double> P, N, T;                       // positions x3, normals x3, texcoords x2
    std::vector<ObjMat> mats;
    std::unordered_map<std::string, int> index;

    int find_or_add(const std::string& name) {
        auto it = index.find(name);
        if (it != index.end()) return it->second;
        ObjMat m; m.name = name; mats.push_back(std::move(m));
        index[name] = (int)mats.size() - 1;
        return (int)mats.size() - 1;
    }

    bool read_mtl(const std::string& path) {
        std::vector<char> data;
        if (!read_file(path, data)) { g_obj_error = "cannot open material library " + path; return false; }
        std::vector<Tok> tok; int cur = -1;
        const char* p = data.data(); const char* end = p + data.size();
        while (p < end) {
            const char* e = (const char*)memchr(p, '\n', end - p); if (!e) e = end;
            split_line(p, e, tok); p = e + 1;
            if (tok.empty() || tok[0].p[0] == '#') continue;
            if (tok_is(tok[0], "newmtl")) { if (tok.size() < 2) { g_obj_error = "newmtl without a name in " + path; return false; } cur = find_or_add(std::string(tok[1].p, tok[1].n)); continue; }
            if (cur < 0) continue;
            ObjMat& m = mats[cur];
            bool ok = true;
            if (tok_is(tok[0], "Kd") && tok.size() >= 4) { for (int k = 0; k < 3; ++k) ok &= to_double(tok[1 + k], m.diffuse[k]); }
            else if (tok_is(tok[0], "Ke") && tok.size() >= 4) { for (int k = 0; k < 3; ++k) ok &= to_double(tok[1 + k], m.emissive[k]); }
            else if (tok_is(tok[0], "d") && tok.size() >= 2) ok = to_double(tok[1], m.transparency);
            else if (tok_is(tok[0], "Tr") && tok.size() >= 2) { double v; ok = to_double(tok[1], v); m.transparency = 1.0 - v; }
            else if (tok_is(tok[0], "Ns") && tok.size() >= 2) ok = to_double(tok[1], m.shininess);
            else if (tok_is(tok[0], "Ni") && tok.size() >= 2) ok = to_double(tok[1], m.optical_density);
            if (!ok) { g_obj_error = "bad number in " + path + ": " + std::string(tok[0].p, tok[0].n); return false; }
        }
        return true;
    }

    bool parse(const std::string& path) {
        std::vector<char> data;
        if (!read_file(path, data)) { g_obj_error = "cannot open " + path; return false; }
        std::string dir; { size_t k = path.find_last_of('/'); if (k != std::string::npos) dir = path.substr(0, k + 1); }
        std::vector<Tok> tok; std::vector<int32_t> idx; int cur = -1;
        const char* p = data.data(); const char* end = p + data.size();
        long line_no = 0;
        auto fail = [&](const char* what) { char b[160]; snprintf(b, sizeof(b), "%s:%ld: %s", path.c_str(), line_no, what); g_obj_error = b; return false; };
        while (p < end) {
            const char* e = (const char*)memchr(p, '\n', end - p); if (!e) e = end;
            split_line(p, e, tok); p = e + 1; ++line_no;
            if (tok.empty()) continue;
            const Tok& key = tok[0];
            if (tok_is(key, "v") || tok_is(key, "vn")) {
                if (tok.size() < 4) return fail("vertex / normal needs three numbers");
                double v[3]; for (int k = 0; k < 3; ++k) if (!to_double(tok[1 + k], v[k])) return fail("bad number");
                std::vector<double>& dst = (key.n == 1) ? P : N; dst.insert(dst.end(), v, v + 3);
            } else if (tok_is(key, "vt")) {
                if (tok.size() < 3) return fail("texture coordinate needs two numbers");
                double v[2]; for (int k = 0; k < 2; ++k) if (!to_double(tok[1 + k], v[k])) return fail("bad number");
                T.insert(T.end(), v, v + 2);
            } else if (tok_is(key, "f")) {
                if (tok.size() < 4) return fail("face with fewer than three corners");
                if (cur < 0) { char nm[32]; snprintf(nm, sizeof(nm), "default%d", (int)mats.size()); cur = find_or_add(nm); }
                ObjMat& m = mats[cur];
                // vertex format of this face from its first corner: a | a/b | a//c | a/b/c
                const char* s1 = (const char*)memchr(tok[1].p, '/', tok[1].n);
                const char* s2 = s1 ? (const char*)memchr(s1 + 1, '/', tok[1].p + tok[1].n - s1 - 1) : nullptr;
                const int nparts = s1 ? (s2 ? 3 : 2) : 1;
                const int has_vt = (nparts == 2 || (nparts == 3 && s2 != s1 + 1)) ? 1 : 0, has_vn = (nparts == 3) ? 1 : 0;
                if (m.has_vt < 0) { m.has_vt = has_vt; m.has_vn = has_vn; }
                else if (m.has_vt != has_vt || m.has_vn != has_vn) return fail("material mixes vertex formats");
                const long np_ = (long)(P.size() / 3), nt_ = (long)(T.size() / 2), nn_ = (long)(N.size() / 3);
                idx.clear();
                for (size_t c = 1; c < tok.size(); ++c) {
                    const char* b0 = tok[c].p; const char* e0 = b0 + tok[c].n;
                    const char* a1 = (const char*)memchr(b0, '/', e0 - b0);
                    const char* a2 = a1 ? (const char*)memchr(a1 + 1, '/', e0 - a1 - 1) : nullptr;
                    long a = 0, bb = 0, cc = 0;
                    if (!to_int(b0, (int)((a1 ? a1 : e0) - b0), a)) return fail("bad vertex index");
                    a = a < 0 ? a + np_ : a - 1;
                    if (has_vt) { if (!a1 || !to_int(a1 + 1, (int)((a2 ? a2 : e0) - a1 - 1), bb)) return fail("bad texture index"); bb = bb < 0 ? bb + nt_ : bb - 1; }
                    if (has_vn) { if (!a2 || !to_int(a2 + 1, (int)(e0 - a2 - 1), cc)) return fail("bad normal index"); cc = cc < 0 ? cc + nn_ : cc - 1; }
                    if (a < 0 || a >= np_ || (has_vt && (bb < 0 || bb >= nt_)) || (has_vn && (cc < 0 || cc >= nn_))) return fail("index out of range");
                    idx.push_back((int32_t)a); idx.push_back((int32_t)bb); idx.push_back((int32_t)cc);
                }
                const int nc = (int)(idx.size() / 3);
                auto emit = [&](int k) { m.corners.insert(m.corners.end(), idx.begin() + 3 * k, idx.begin() + 3 * k + 3); };
                emit(0); emit(1); emit(2);
                for (int j = 3; j < nc; ++j) { emit(j); emit(0); emit(j - 1); }
            } else if (tok_is(key, "usemtl")) {
                cur = find_or_add(tok.size() > 1 ? std::string(tok[1].p, tok[1].n) : std::string());
            } else if (tok_is(key, "mtllib")) {
                if (tok.size() < 2) return fail("mtllib without a file name");
                if (!read_mtl(dir + std::string(tok[1].p, tok[1].n))) return false;
            }
        }
        return true;
    }
};

extern "C" {

const char* tr_obj_last_error(void) { return g_obj_error.c_str(); }

int tr_obj_open(const char* path, tr_obj** out) {
    if (!path || !out) { g_obj_error = "tr_obj_open: NULL argument"; return TR_ERR_INVALID; }
    tr_obj* o = new tr_obj();
    if (!o->parse(path)) { delete o; *out = nullptr; return TR_ERR_INVALID; }
    *out = o;
    return TR_OK;
}
void tr_obj_close(tr_obj* o) { delete o; }
int tr_obj_material_count(const tr_obj* o) { return o ? (int)o->mats.size() : 0; }

int tr_obj_material(const tr_obj* o, int k, char* name, int name_cap, double props[9], int64_t* n_vertices, int* has_vt, int* has_vn) {
    if (!o || k < 0 || k >= (int)o->mats.size()) { g_obj_error = "tr_obj_material: bad index"; return TR_ERR_INVALID; }
    const ObjMat& m = o->mats[k];
    if (name && name_cap > 0) { snprintf(name, (size_t)name_cap, "%s", m.name.c_str()); }
    if (props) { for (int c = 0; c < 3; ++c) { props[c] = m.diffuse[c]; props[3 + c] = m.emissive[c]; } props[6] = m.transparency; props[7] = m.shininess; props[8] = m.optical_density; }
    if (n_vertices) *n_vertices = (int64_t)(m.corners.size() / 3);
    if (has_vt) *has_vt = m.has_vt > 0; if (has_vn) *has_vn = m.has_vn > 0;
    return TR_OK;
}

int tr_obj_material_vertices(const tr_obj* o, int k, double* rows) {
    if (!o || k < 0 || k >= (int)o->mats.size() || !rows) { g_obj_error = "tr_obj_material_vertices: bad argument"; return TR_ERR_INVALID; }
    const ObjMat& m = o->mats[k];
    const size_t n = m.corners.size() / 3;
    for (size_t i = 0; i < n; ++i) {
        double* r = rows + i * 9;
        const int32_t a = m.corners[3 * i], b = m.corners[3 * i + 1], c = m.corners[3 * i + 2];
        r[0] = o->P[3 * (size_t)a]; r[1] = o->P[3 * (size_t)a + 1]; r[2] = o->P[3 * (size_t)a + 2];
        if (m.has_vn > 0) { r[3] = o->N[3 * (size_t)c]; r[4] = o->N[3 * (size_t)c + 1]; r[5] = o->N[3 * (size_t)c + 2]; } else { r[3] = r[4] = r[5] = 0.0; }
        if (m.has_vt > 0) { r[6] = o->T[2 * (size_t)b]; r[7] = o->T[2 * (size_t)b + 1]; } else { r[6] = r[7] = 0.0; }
        r[8] = 0.0;
    }
    return TR_OK;
}

}  // extern "C"
