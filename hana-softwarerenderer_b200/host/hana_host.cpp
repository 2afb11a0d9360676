/*
 * hana_host.cpp — host-side mirror of the caller of the hot path: the per-frame
 * uniform block DrawModel::draw builds (scene.h:55-71), the orbit camera that
 * feeds it (camera.cpp:44-92) and the matrix builders both use (maths.cpp,
 * matrix.h, gameobject.cpp:12-17). Plain C++ over flat float arrays, written
 * to the reference's evaluation order so that the uniforms are bit-identical
 * to the ones the reference's own Camera/DrawModel produce (tests compare them
 * with oracle/_ref). With libstdc++ the reference's unqualified sin/cos/tan/acos/atan2 on float
 * arguments resolve to the float overloads (sinf ...), pow(0.95, dolly) to the double one;
 * std::sin(float) etc. below select the same functions. No CUDA here; compiled into libhana_b200.so with
 * -ffp-contract=off.
 */
#include <cmath>
#include <cstring>

#include "../../include/hana_b200.h"

namespace {

/* The reference's matrix builders live in their own translation unit (maths.cpp), so their libm calls
 * happen at run time with glibc's results; GCC would otherwise fold e.g. tanf(const) at compile time with
 * correct rounding, which differs from glibc's tanf in the last bit. */
#define HANA_RUNTIME __attribute__((noipa))

const float kEps = 1e-5f;         /* EPSILON maths.h:6 */
const float kPi = 3.1415927f;     /* PI maths.h:7 */

struct V3 {
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { /* vector.h:64-67: components from last to first (order is immaterial per component) */
    return V3{a.x - b.x, a.y - b.y, a.z - b.z};
}
inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mulf(V3 a, float f) { return V3{a.x * f, a.y * f, a.z * f}; }
inline float dot(V3 a, V3 b) { /* vector.h:69-73: from the last component, starting at 0 */
    float r = 0.f;
    r += a.z * b.z;
    r += a.y * b.y;
    r += a.x * b.x;
    return r;
}
inline float norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); } /* vector.h:41 */
inline V3 normalized(V3 a) { return mulf(a, 1.f / norm(a)); }                    /* vector.h:42 */
inline V3 cross(V3 a, V3 b) { /* vector.h:97-99 */
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct M4 {
    float m[4][4];
};
M4 identity() {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.f : 0.f;
    return r;
}
M4 mul(const M4& a, const M4& b) { /* matrix.h:118-123 with the dot product of vector.h:69-73 */
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.f;
            for (int k = 4; k--;) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
/* determinant by cofactor expansion along row 0, columns from last to first: matrix.h:11-24,69-86 */
float det_n(const float* a, int n) { /* a: n x n row-major */
    if (n == 1) return a[0];
    float ret = 0.f;
    for (int i = n; i--;) {
        float minor[9];
        int mn = n - 1;
        for (int r = 0; r < mn; r++)
            for (int c = 0; c < mn; c++) minor[r * mn + c] = a[(r + 1) * n + (c < i ? c : c + 1)];
        float cof = det_n(minor, mn) * (float)((i % 2) ? -1 : 1);
        ret += a[i] * cof;
    }
    return ret;
}
float cofactor4(const M4& a, int row, int col) { /* matrix.h:84-86 */
    float minor[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) minor[r * 3 + c] = a.m[r < row ? r : r + 1][c < col ? c : c + 1];
    return det_n(minor, 3) * (float)(((row + col) % 2) ? -1 : 1);
}
M4 invert(const M4& a) { /* matrix.h:88-109: adjugate / (adjugate row 0 . row 0), transposed */
    M4 adj;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) adj.m[i][j] = cofactor4(a, i, j);
    float tmp = 0.f;
    for (int k = 4; k--;) tmp += adj.m[0][k] * a.m[0][k];
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.m[j][i] = adj.m[i][j] / tmp;
    return r;
}
M4 translate(float x, float y, float z) { /* maths.cpp:36-42 */
    M4 m = identity();
    m.m[0][3] = x;
    m.m[1][3] = y;
    m.m[2][3] = z;
    return m;
}
M4 scale(float x, float y, float z) { /* maths.cpp:52-59 */
    M4 m = identity();
    m.m[0][0] = x;
    m.m[1][1] = y;
    m.m[2][2] = z;
    return m;
}
HANA_RUNTIME M4 rotate_x(float a) { /* maths.cpp:105-114 */
    float c = std::cos(a), s = std::sin(a);
    M4 m = identity();
    m.m[1][1] = c; m.m[1][2] = -s; m.m[2][1] = s; m.m[2][2] = c;
    return m;
}
HANA_RUNTIME M4 rotate_y(float a) { /* maths.cpp:124-133 */
    float c = std::cos(a), s = std::sin(a);
    M4 m = identity();
    m.m[0][0] = c; m.m[0][2] = s; m.m[2][0] = -s; m.m[2][2] = c;
    return m;
}
HANA_RUNTIME M4 rotate_z(float a) { /* maths.cpp:143-152 */
    float c = std::cos(a), s = std::sin(a);
    M4 m = identity();
    m.m[0][0] = c; m.m[0][1] = -s; m.m[1][0] = s; m.m[1][1] = c;
    return m;
}
M4 lookat(V3 eye, V3 target, V3 up) { /* maths.cpp:180-195 */
    V3 z = normalized(sub(eye, target));
    V3 x = normalized(cross(up, z));
    V3 y = cross(z, x);
    M4 m = identity();
    m.m[0][0] = x.x; m.m[0][1] = x.y; m.m[0][2] = x.z;
    m.m[1][0] = y.x; m.m[1][1] = y.y; m.m[1][2] = y.z;
    m.m[2][0] = z.x; m.m[2][1] = z.y; m.m[2][2] = z.z;
    m.m[0][3] = -dot(x, eye);
    m.m[1][3] = -dot(y, eye);
    m.m[2][3] = -dot(z, eye);
    return m;
}
M4 orthographic(float right, float top, float near, float far) { /* maths.cpp:214-223 */
    float z_range = far - near;
    M4 m = identity();
    m.m[0][0] = 1 / right;
    m.m[1][1] = 1 / top;
    m.m[2][2] = -2 / z_range;
    m.m[2][3] = -(near + far) / z_range;
    return m;
}
HANA_RUNTIME M4 perspective(float fovy, float aspect, float near, float far) { /* maths.cpp:243-255 */
    float z_range = far - near;
    M4 m = identity();
    m.m[1][1] = 1 / std::tan(fovy / 2);
    m.m[0][0] = m.m[1][1] / aspect;
    m.m[2][2] = -(near + far) / z_range;
    m.m[2][3] = -2 * near * far / z_range;
    m.m[3][2] = -1;
    m.m[3][3] = 0;
    return m;
}
inline float to_radians(float deg) { return (kPi / 180) * deg; } /* maths.h:9 */
inline float clampf(float f, float lo, float hi) { return f < lo ? lo : (f > hi ? hi : f); }

void store(float* dst, const M4& m) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) dst[i * 4 + j] = m.m[i][j];
}

}  // namespace

extern "C" {

/* Camera (camera.h:13-33): position, target, aspect. */
HANA_API int hana_camera_init(HanaCamera* cam, const float position[3], const float target[3], float aspect) {
    if (!cam || !position || !target) return HANA_E_INVALID;
    for (int i = 0; i < 3; i++) {
        cam->position[i] = position[i];
        cam->target[i] = target[i];
    }
    cam->aspect = aspect;
    return HANA_OK;
}

/* Camera::update_transform (camera.cpp:63-70) with calculate_pan (:31-42) and calculate_offset (:44-61). */
HANA_API int hana_camera_update(HanaCamera* cam, float orbit_x, float orbit_y, float pan_x, float pan_y, float dolly) {
    if (!cam) return HANA_E_INVALID;
    const float fovy = to_radians(60);
    V3 position{cam->position[0], cam->position[1], cam->position[2]};
    V3 target{cam->target[0], cam->target[1], cam->target[2]};
    V3 from_target = sub(position, target);
    V3 from_camera = sub(target, position);
    /* pan */
    V3 forward = normalized(from_camera);
    V3 up0{0, 1, 0};
    V3 left = cross(up0, forward);
    V3 up = cross(forward, left);
    float distance = norm(from_camera);
    float factor = distance * std::tan(fovy / 2) * 2;
    V3 delta_x = mulf(mulf(left, pan_x), factor);
    V3 delta_y = mulf(mulf(up, pan_y), factor);
    V3 pan = add(delta_x, delta_y);
    /* orbit + dolly in spherical coordinates */
    float radius = norm(from_target);
    float theta = std::atan2(from_target.x, from_target.z);
    float phi = std::acos(from_target.y / radius);
    float two_pi = kPi * 2;
    radius *= (float)std::pow(0.95, (double)dolly);
    theta -= orbit_x * two_pi;
    phi -= orbit_y * two_pi;
    phi = clampf(phi, kEps, kPi - kEps);
    V3 offset;
    offset.x = radius * std::sin(phi) * std::sin(theta);
    offset.y = radius * std::cos(phi);
    offset.z = radius * std::sin(phi) * std::cos(theta);
    target = add(target, pan);
    position = add(target, offset);
    cam->position[0] = position.x; cam->position[1] = position.y; cam->position[2] = position.z;
    cam->target[0] = target.x; cam->target[1] = target.y; cam->target[2] = target.z;
    return HANA_OK;
}

HANA_API int hana_scene_defaults(HanaSceneDesc* s) {
    if (!s) return HANA_E_INVALID;
    memset(s, 0, sizeof(*s));
    s->light_pos[0] = s->light_pos[1] = s->light_pos[2] = 2.f;          /* scene.cpp:7 */
    s->model_scale[0] = s->model_scale[1] = s->model_scale[2] = 1.f;    /* gameobject.h defaults */
    s->light_color[0] = 255.f / 255; s->light_color[1] = 244.f / 255; s->light_color[2] = 214.f / 255; /* scene.h:9 */
    s->ambient[0] = 54.f / 255; s->ambient[1] = 58.f / 255; s->ambient[2] = 66.f / 255;                /* scene.h:8 */
    s->light_color[3] = s->ambient[3] = 255.f;                          /* Color default alpha color.h:10 */
    for (int i = 0; i < 3; i++) s->mat_color[i] = s->mat_specular[i] = 1.f; /* Color::White scene.cpp:81-82 */
    s->mat_color[3] = s->mat_specular[3] = 255.f;
    s->gloss = 50.f;                                                     /* scene.cpp:83 */
    s->bump_scale = 1.f;                                                 /* scene.cpp:84 */
    return HANA_OK;
}

/* The ShaderData DrawModel::draw fills before its passes (scene.h:55-71). */
HANA_API int hana_scene_uniforms(const HanaCamera* cam, const HanaSceneDesc* s, int width, int height, int enable_shadow,
                                 HanaUniforms* out) {
    if (!cam || !s || !out || width <= 0 || height <= 0) return HANA_E_INVALID;
    memset(out, 0, sizeof(*out));
    V3 position{cam->position[0], cam->position[1], cam->position[2]};
    V3 target{cam->target[0], cam->target[1], cam->target[2]};
    V3 up{0, 1, 0};
    M4 view = lookat(position, target, up);                                   /* camera.cpp:82-87 */
    M4 proj = perspective(to_radians(60), cam->aspect, 0.1f, 10000.f);        /* camera.cpp:11-14,89-92 */
    M4 m_t = translate(s->model_pos[0], s->model_pos[1], s->model_pos[2]);    /* gameobject.cpp:12-17 */
    M4 m_r = mul(mul(rotate_z(to_radians(s->model_rot_deg[2])), rotate_x(to_radians(s->model_rot_deg[0]))),
                 rotate_y(to_radians(s->model_rot_deg[1])));
    M4 m_s = scale(s->model_scale[0], s->model_scale[1], s->model_scale[2]);
    M4 model = mul(mul(m_t, m_r), m_s);
    M4 model_I = invert(model);
    V3 light{s->light_pos[0], s->light_pos[1], s->light_pos[2]};
    V3 origin{0, 0, 0}; /* Camera::get_target_position is hard-wired to Zero: camera.cpp:94-97 */
    V3 light_dir = normalized(sub(light, origin));
    float aspect = (float)width / (float)height;
    M4 light_vp = mul(orthographic(aspect, 1, 0, 5), lookat(light, origin, up)); /* scene.h:68-69 */
    M4 camera_vp = mul(proj, view);                                              /* scene.h:70 */
    store(out->model, model);
    store(out->model_I, model_I);
    store(out->camera_vp, camera_vp);
    store(out->light_vp, light_vp);
    out->view_pos[0] = position.x; out->view_pos[1] = position.y; out->view_pos[2] = position.z;
    out->light_dir[0] = light_dir.x; out->light_dir[1] = light_dir.y; out->light_dir[2] = light_dir.z;
    out->gloss = s->gloss;
    out->bump_scale = s->bump_scale;
    for (int i = 0; i < 4; i++) {
        out->light_color[i] = s->light_color[i];
        out->ambient[i] = s->ambient[i];
        out->mat_color[i] = s->mat_color[i];
        out->mat_specular[i] = s->mat_specular[i];
    }
    out->enable_shadow = enable_shadow ? 1 : 0;
    return HANA_OK;
}

/* BASELINE.json configs[2]: frame k = Camera(CAMERA_POSITION, CAMERA_TARGET, W/H) advanced k times by
 * update_transform(orbit = (1/frames_per_turn, 0)). Fills frames [first, first + count). */
HANA_API int hana_orbit_sweep_uniforms(const HanaSceneDesc* s, int width, int height, int enable_shadow, int first, int count,
                                       int frames_per_turn, HanaUniforms* out) {
    if (!s || !out || first < 0 || count < 0 || frames_per_turn <= 0) return HANA_E_INVALID;
    HanaCamera cam;
    const float pos[3] = {0, 0, 2.f}, tgt[3] = {0, 0, 0}; /* camera.h:8-9 */
    hana_camera_init(&cam, pos, tgt, (float)width / (float)height);
    const float step = 1.f / (float)frames_per_turn;
    for (int k = 0; k < first + count; k++) {
        if (k >= first) {
            int r = hana_scene_uniforms(&cam, s, width, height, enable_shadow, out + (k - first));
            if (r != HANA_OK) return r;
        }
        hana_camera_update(&cam, step, 0.f, 0.f, 0.f, 0.f);
    }
    return HANA_OK;
}

}  // extern "C"
