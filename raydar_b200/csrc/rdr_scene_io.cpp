// rdr_scene_io.cpp -- host scene pipeline behind the C ABI: .rscn reader, camera matrices, default scene,
// PNG writer.  Pure host C++ (the reference's equivalents are serde_json + cgmath + image, none of which
// exist here).
//
//   .rscn = serde_json of `Scene` (scene/mod.rs:13-18), loaded as in cli/mod.rs:32-40: every field,
//           INCLUDING the four stored camera matrices, is taken from the file as-is.
//   update_matrices (scene/camera.rs:210-231) restated from cgmath 0.18.0: Matrix4::look_at_lh,
//           cgmath::perspective / ortho, Matrix4::invert (cofactor form).  The restatement reproduces
//           the matrices stored in scenes/default.rscn and scenes/benchmark.rscn bit for bit
//           (tests/test_scene_io.py), which pins it to the reference's own output.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "raydar_cuda.h"

namespace rdr { int api_fail(RdrRenderer *r, int status, const char *msg); }

namespace {

// ---- minimal JSON ------------------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const char *key) const
    {
        for (const auto &kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const char *p, *end;
    std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool fail(const char *m) { if (err.empty()) err = m; return false; }
    bool parse_string(std::string &out)
    {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': if (end - p < 5) return fail("bad \\u escape"); out += '?'; p += 4; break;
                default: out += *p;
                }
                ++p;
            } else out += *p++;
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool parse(Json &v, int depth = 0)
    {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end of input");
        if (*p == '{') {
            v.kind = Json::Object; ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string key;
                if (!parse_string(key)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                Json child;
                if (!parse(child, depth + 1)) return false;
                v.obj.emplace_back(std::move(key), std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = Json::Array; ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                Json child;
                if (!parse(child, depth + 1)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = Json::String; return parse_string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = Json::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = Json::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = Json::Null; p += 4; return true; }
        char *num_end = nullptr;
        std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
        double d = strtod(tmp.c_str(), &num_end);
        if (num_end == tmp.c_str()) return fail("unexpected character");
        v.kind = Json::Number; v.num = d; p += num_end - tmp.c_str();
        return true;
    }
};

// ---- cgmath 0.18 restatements (f32, one rounding per operation; built with -ffp-contract=off) --------
struct V3 { float x, y, z; };
float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
V3 normalize(V3 a) { float k = 1.0f / std::sqrt(dot(a, a)); return {a.x * k, a.y * k, a.z * k}; }
V3 cross(V3 a, V3 b) { return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)}; }

// Matrix4::look_at_lh(eye, center, up) = look_to_rh(eye, -(center - eye), up); m is column-major
void look_at_lh(const float eye[3], const float center[3], const float up[3], float m[16])
{
    V3 e{eye[0], eye[1], eye[2]};
    V3 dir{-(center[0] - eye[0]), -(center[1] - eye[1]), -(center[2] - eye[2])};
    V3 f = normalize(dir);
    V3 s = normalize(cross(f, V3{up[0], up[1], up[2]}));
    V3 u = cross(s, f);
    const float v[16] = {s.x, u.x, -f.x, 0.0f, s.y, u.y, -f.y, 0.0f, s.z, u.z, -f.z, 0.0f, -dot(e, s), -dot(e, u), dot(e, f), 1.0f};
    memcpy(m, v, sizeof v);
}

// cgmath::perspective(Deg(fov), aspect, near, far): f = cot(fovy / 2)
void perspective(float fov_deg, float aspect, float near, float far, float m[16])
{
    const float rad = fov_deg * (float)(3.14159265358979323846 / 180.0);
    const float f = 1.0f / (float)std::tan((double)(rad / 2.0f));     // correctly rounded f32 tan
    memset(m, 0, 16 * sizeof(float));
    m[0] = f / aspect;
    m[5] = f;
    m[10] = (far + near) / (near - far);
    m[11] = -1.0f;
    m[14] = (2.0f * far * near) / (near - far);
}

// cgmath::ortho(left, right, bottom, top, near, far)
void ortho(float left, float right, float bottom, float top, float near, float far, float m[16])
{
    memset(m, 0, 16 * sizeof(float));
    m[0] = 2.0f / (right - left);
    m[5] = 2.0f / (top - bottom);
    m[10] = -2.0f / (far - near);
    m[12] = -(right + left) / (right - left);
    m[13] = -(top + bottom) / (top - bottom);
    m[14] = -(far + near) / (far - near);
    m[15] = 1.0f;
}

// Matrix3::determinant, c[col][row]
float det3(const float c[3][3])
{
    return c[0][0] * (c[1][1] * c[2][2] - c[2][1] * c[1][2]) - c[1][0] * (c[0][1] * c[2][2] - c[2][1] * c[0][2]) +
           c[2][0] * (c[0][1] * c[1][2] - c[1][1] * c[0][2]);
}

// Matrix4::determinant: expansion along the first row of each column
float det4(const float m[16])
{
    auto M = [&](int col, int row) { return m[col * 4 + row]; };
    float d[4];
    for (int skip = 0; skip < 4; ++skip) {
        float c[3][3];
        for (int k = 0; k < 3; ++k) {               // k-th column of the Matrix3 = row k+1 of the kept columns
            int cc = 0;
            for (int col = 0; col < 4; ++col) { if (col == skip) continue; c[k][cc++] = M(col, k + 1); }
        }
        d[skip] = det3(c);
    }
    return M(0, 0) * d[0] - M(1, 0) * d[1] + M(2, 0) * d[2] - M(3, 0) * d[3];
}

// Matrix4::invert (cofactor form).  Returns false when det == 0 (the reference unwraps and panics).
bool invert4(const float m[16], float out[16])
{
    const float det = det4(m);
    if (det == 0.0f) return false;
    const float inv_det = 1.0f / det;
    float t[4][4];                                    // transpose: t[col][row] = m[row][col]
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) t[c][r] = m[r * 4 + c];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float c3[3][3];
            int cc = 0;
            for (int col = 0; col < 4; ++col) {
                if (col == i) continue;
                int rr = 0;
                for (int row = 0; row < 4; ++row) { if (row == j) continue; c3[cc][rr++] = t[col][row]; }
                ++cc;
            }
            const float sign = ((i + j) & 1) ? -1.0f : 1.0f;
            out[i * 4 + j] = det3(c3) * sign * inv_det;
        }
    return true;
}

}  // namespace

// ---- Scene (scene/mod.rs:13-18, camera.rs:14-30) ---------------------------------------------------------
struct RdrScene {
    float position[3] = {0, 0, 0}, target[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    uint32_t width = 0, height = 0;
    bool perspective_proj = true;
    float fov_or_size = 90.0f;
    float near_clip = 0.01f, far_clip = 1000.0f;
    float view[16], proj[16], inv_view[16], inv_proj[16];
    uint32_t world_kind = RDR_WORLD_SKY;
    float world_a[3] = {0, 0, 0}, world_b[3] = {0, 0, 0};
    std::vector<uint32_t> kind;
    std::vector<float> geom, material;
};

namespace {

thread_local std::string g_scene_err;

int scene_fail(int status, const std::string &msg) { rdr::api_fail(nullptr, status, msg.c_str()); return status; }

bool num(const Json *j, float &out)
{
    if (!j || j->kind != Json::Number) return false;
    out = (float)j->num;               // serde_json parses to f64, then `as f32`
    return true;
}
bool vec3(const Json *j, float out[3])
{
    return j && j->kind == Json::Object && num(j->get("x"), out[0]) && num(j->get("y"), out[1]) && num(j->get("z"), out[2]);
}
// cgmath Matrix4 serialises as {x: column0, y: column1, z: column2, w: column3}, a column as {x,y,z,w}
bool mat4(const Json *j, float out[16])
{
    if (!j || j->kind != Json::Object) return false;
    static const char *names[4] = {"x", "y", "z", "w"};
    for (int c = 0; c < 4; ++c) {
        const Json *col = j->get(names[c]);
        if (!col || col->kind != Json::Object) return false;
        for (int r = 0; r < 4; ++r) if (!num(col->get(names[r]), out[c * 4 + r])) return false;
    }
    return true;
}

// Camera::update_matrices, camera.rs:210-231
bool update_matrices(RdrScene &s)
{
    look_at_lh(s.position, s.target, s.up, s.view);
    const float aspect = (float)s.width / (float)s.height;
    if (s.perspective_proj) perspective(s.fov_or_size, aspect, s.near_clip, s.far_clip, s.proj);
    else { const float z = s.fov_or_size; ortho(-z * aspect, z * aspect, -z, z, s.near_clip, s.far_clip, s.proj); }
    return invert4(s.view, s.inv_view) && invert4(s.proj, s.inv_proj);
}

void push_object(RdrScene &s, uint32_t kind, float cx, float cy, float cz, float size, const float mat[RDR_MAT_STRIDE])
{
    s.kind.push_back(kind);
    s.geom.insert(s.geom.end(), {cx, cy, cz, size});
    s.material.insert(s.material.end(), mat, mat + RDR_MAT_STRIDE);
}

}  // namespace

extern "C" {

int rdr_scene_load_rscn(const char *path, RdrScene **out)
{
    if (!path || !out) return scene_fail(RDR_ERR_INVALID, "NULL argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return scene_fail(RDR_ERR_IO, std::string("Cannot open scene file: ") + path);
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    fclose(f);

    Parser ps{text.data(), text.data() + text.size(), {}};
    Json root;
    if (!ps.parse(root) || root.kind != Json::Object)
        return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: " + (ps.err.empty() ? std::string("not an object") : ps.err));

    std::unique_ptr<RdrScene> s(new RdrScene());
    const Json *cam = root.get("camera"), *world = root.get("world"), *objects = root.get("objects");
    if (!cam || !world || !objects || objects->kind != Json::Array) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: missing camera/world/objects");

    float rx = 0, ry = 0;
    bool ok = vec3(cam->get("position"), s->position) && vec3(cam->get("target"), s->target) && vec3(cam->get("up"), s->up) &&
              num(cam->get("resolution_x"), rx) && num(cam->get("resolution_y"), ry) &&
              num(cam->get("near_clip"), s->near_clip) && num(cam->get("far_clip"), s->far_clip) &&
              mat4(cam->get("view_matrix"), s->view) && mat4(cam->get("proj_matrix"), s->proj) &&
              mat4(cam->get("inverse_view_matrix"), s->inv_view) && mat4(cam->get("inverse_proj_matrix"), s->inv_proj);
    const Json *projection = cam->get("projection");
    if (ok && projection && projection->kind == Json::Object) {
        if (const Json *p = projection->get("Perspective")) { s->perspective_proj = true; ok = num(p->get("fov"), s->fov_or_size); }
        else if (const Json *o = projection->get("Orthographic")) { s->perspective_proj = false; ok = num(o->get("size"), s->fov_or_size); }
        else ok = false;
    } else ok = false;
    if (!ok || rx < 0 || ry < 0) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad camera");
    s->width = (uint32_t)rx; s->height = (uint32_t)ry;

    if (world->kind == Json::String && world->str == "Transparent") s->world_kind = RDR_WORLD_TRANSPARENT;
    else if (world->kind == Json::Object && world->get("SkyColor")) {
        const Json *sky = world->get("SkyColor");
        s->world_kind = RDR_WORLD_SKY;
        if (!vec3(sky->get("top_color"), s->world_a) || !vec3(sky->get("bottom_color"), s->world_b)) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad SkyColor");
    } else if (world->kind == Json::Object && world->get("SolidColor")) {
        s->world_kind = RDR_WORLD_SOLID;
        if (!vec3(world->get("SolidColor"), s->world_a)) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad SolidColor");
    } else return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad world");

    for (const Json &o : objects->arr) {
        const Json *g = o.get("geometry"), *m = o.get("material");
        if (!g || !m || g->kind != Json::Object || m->kind != Json::Object) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad object");
        float c[3], size = 0, mat[RDR_MAT_STRIDE];
        uint32_t kind;
        if (const Json *sp = g->get("Sphere")) { kind = RDR_SPHERE; ok = vec3(sp->get("center"), c) && num(sp->get("radius"), size); }
        else if (const Json *cu = g->get("Cube")) { kind = RDR_CUBE; ok = vec3(cu->get("center"), c) && num(cu->get("side_length"), size); }
        else return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: unknown geometry");
        ok = ok && vec3(m->get("albedo"), mat + 0) && num(m->get("roughness"), mat[3]) && num(m->get("metallic"), mat[4]) &&
             vec3(m->get("emission_color"), mat + 5) && num(m->get("emission_strength"), mat[8]) &&
             num(m->get("transmission"), mat[9]) && num(m->get("ior"), mat[10]);
        if (!ok) return scene_fail(RDR_ERR_PARSE, "Cannot parse scene file: bad object fields");
        push_object(*s, kind, c[0], c[1], c[2], size, mat);
    }
    *out = s.release();
    return RDR_OK;
}

// Scene::default(), scene/mod.rs:20-68 (Material::default, material.rs:15-27)
int rdr_scene_default(RdrScene **out)
{
    if (!out) return scene_fail(RDR_ERR_INVALID, "NULL argument");
    std::unique_ptr<RdrScene> s(new RdrScene());
    const float pos[3] = {-3.09f, 0.03f, -1.16f};
    memcpy(s->position, pos, sizeof pos);
    s->width = 854; s->height = 480; s->near_clip = 0.01f; s->far_clip = 1000.0f;
    s->perspective_proj = true; s->fov_or_size = 90.0f;
    s->world_kind = RDR_WORLD_SKY;
    const float top[3] = {0.53f, 0.8f, 0.92f}, bottom[3] = {1.0f, 1.0f, 1.0f};
    memcpy(s->world_a, top, sizeof top); memcpy(s->world_b, bottom, sizeof bottom);
    //                         albedo            rough metal emission      estr  trans ior
    const float glass[11]  = {1.0f, 1.0f, 1.0f, 0.2f, 0.0f, 0, 0, 0,       0.0f, 1.0f, 1.5f};
    const float ground[11] = {0.34f, 0.34f, 0.44f, 0.5f, 0.0f, 0, 0, 0,    0.0f, 0.0f, 1.5f};
    const float light[11]  = {0.8f, 0.8f, 0.8f, 0.5f, 0.0f, 0.8f, 0.5f, 0.2f, 30.0f, 0.0f, 1.5f};
    push_object(*s, RDR_SPHERE, 0.0f, 0.001f, 0.0f, 1.0f, glass);
    push_object(*s, RDR_CUBE, 0.0f, -101.0f, 0.0f, 200.0f, ground);
    push_object(*s, RDR_CUBE, 7.0f, 3.0f, 0.0f, 1.8f, light);
    if (!update_matrices(*s)) return scene_fail(RDR_ERR_INVALID, "camera matrix is singular");
    *out = s.release();
    return RDR_OK;
}

int rdr_scene_set_resolution(RdrScene *s, uint32_t width, uint32_t height)
{
    if (!s || width == 0 || height == 0) return scene_fail(RDR_ERR_INVALID, "bad resolution");
    s->width = width; s->height = height;
    if (!update_matrices(*s)) return scene_fail(RDR_ERR_INVALID, "camera matrix is singular (the reference panics: camera.rs:229-230)");
    return RDR_OK;
}

int rdr_scene_override_resolution(RdrScene *s, uint32_t width, uint32_t height)
{
    if (!s) return scene_fail(RDR_ERR_INVALID, "NULL scene");
    s->width = width; s->height = height;
    return RDR_OK;
}

int rdr_scene_flat(const RdrScene *s, RdrSceneFlat *out)
{
    if (!s || !out) return scene_fail(RDR_ERR_INVALID, "NULL argument");
    out->width = s->width; out->height = s->height;
    memcpy(out->inv_proj, s->inv_proj, sizeof out->inv_proj);
    memcpy(out->inv_view, s->inv_view, sizeof out->inv_view);
    memcpy(out->cam_pos, s->position, sizeof out->cam_pos);
    out->world_kind = s->world_kind;
    memcpy(out->world_a, s->world_a, sizeof out->world_a);
    memcpy(out->world_b, s->world_b, sizeof out->world_b);
    out->n_objects = (uint32_t)s->kind.size();
    out->kind = s->kind.data(); out->geom = s->geom.data(); out->material = s->material.data();
    return RDR_OK;
}

void rdr_scene_free(RdrScene *s) { delete s; }

// test hook: the four camera matrices (view, proj, inverse view, inverse proj), 64 floats
int rdr_debug_scene_matrices(const RdrScene *s, float out[64])
{
    if (!s || !out) return RDR_ERR_INVALID;
    memcpy(out, s->view, 64); memcpy(out + 16, s->proj, 64); memcpy(out + 32, s->inv_view, 64); memcpy(out + 48, s->inv_proj, 64);
    return RDR_OK;
}

// ---- PNG (RgbaImage::save, main.rs:19): 8-bit RGBA, zlib stream of stored blocks --------------------------
static uint32_t crc32_update(uint32_t crc, const uint8_t *d, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xff] ^ (crc >> 8);
    return crc;
}

static void put_u32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

static void put_chunk(std::vector<uint8_t> &png, const char type[4], const std::vector<uint8_t> &data)
{
    put_u32(png, (uint32_t)data.size());
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    put_u32(png, crc32_update(0xffffffffu, png.data() + start, png.size() - start) ^ 0xffffffffu);
}

int rdr_write_png(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height)
{
    if (!path || !rgba8 || width == 0 || height == 0) return scene_fail(RDR_ERR_INVALID, "bad image");
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * (1 + (size_t)width * 4));
    for (uint32_t y = 0; y < height; ++y) { raw.push_back(0); raw.insert(raw.end(), rgba8 + (size_t)y * width * 4, rgba8 + (size_t)(y + 1) * width * 4); }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += n;
    }
    put_u32(z, (b << 16) | a);

    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_u32(ihdr, width); put_u32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(png, "IHDR", ihdr);
    put_chunk(png, "IDAT", z);
    put_chunk(png, "IEND", {});
    FILE *f = fopen(path, "wb");
    if (!f) return scene_fail(RDR_ERR_IO, std::string("Cannot save image: ") + path);
    const bool ok = fwrite(png.data(), 1, png.size(), f) == png.size();
    fclose(f);
    return ok ? RDR_OK : scene_fail(RDR_ERR_IO, std::string("Cannot save image: ") + path);
}

}  // extern "C"
