/*
 * hana_oracle.c — CPU restatement of Hana-SoftwareRenderer's rasterisation
 * hot path, in plain C99, over the same flat arrays the CUDA C ABI takes.
 *
 * TEST INFRASTRUCTURE ONLY: this is the parity checker. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it. The product (libhana_b200.so) never links or
 * executes anything in oracle/ and fails loudly without a CUDA device.
 *
 * PARITY PINNED: the reference has no tests or golden vectors of its own
 * (SURVEY.md §4), so this restatement is pinned against the reference ITSELF:
 * oracle/build_ref.sh compiles the reference's unmodified sources into
 * oracle/_ref/libhana_ref_inst.so, and tests/test_oracle_vs_reference.py
 * requires bit-identical colour (RGB) + depth + primitive-ID frames and
 * bit-identical per-stage outputs (vertex, clip, barycentric, varyings,
 * fragment) on the bundled scenes and on seeded synthetic ones; the frames the
 * reference produced are committed as fixtures under tests/golden/.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/Hana-SoftwareRenderer/). Arithmetic is IEEE float32 in the
 * reference's evaluation order; compile with -ffp-contract=off and without
 * -ffast-math (oracle/Makefile does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hana_b200.h"

typedef struct HOracleTex {
    const uint8_t* data; /* TGAImage::data layout, may be NULL */
    int32_t w, h, bpp;
} HOracleTex;

typedef struct HOracleCounters {
    uint64_t corners, tris_raster, bbox_pixels, inside, zpass;
} HOracleCounters;

/* v2f field offsets in floats: shader_struct_v2f IShader.h:41-47 */
enum { V_CLIP = 0, V_WPOS = 4, V_WNRM = 7, V_UV = 10, V_INT = 12, V2F_N = 13 };

/* ---- vector.h:69-73: dot product accumulates from the LAST component ---- */
static float dot4(const float* a, const float* b) {
    float r = 0.f;
    r += a[3] * b[3];
    r += a[2] * b[2];
    r += a[1] * b[1];
    r += a[0] * b[0];
    return r;
}
static float dot3(const float* a, const float* b) {
    float r = 0.f;
    r += a[2] * b[2];
    r += a[1] * b[1];
    r += a[0] * b[0];
    return r;
}
static float dot2(const float* a, const float* b) {
    float r = 0.f;
    r += a[1] * b[1];
    r += a[0] * b[0];
    return r;
}
/* matrix.h:112-116 */
static void mat4_vec4(const float* m, const float* v, float* out) {
    for (int i = 0; i < 4; i++) out[i] = dot4(m + 4 * i, v);
}
/* matrix.h:118-123 */
void horacle_mat4_mul(const float* a, const float* b, float* out) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float col[4] = {b[j], b[4 + j], b[8 + j], b[12 + j]};
            out[4 * i + j] = dot4(a + 4 * i, col);
        }
}
/* vector.h:41-42: v * (1 / sqrt((x*x + y*y) + z*z)) */
static void normalize3(float* v) {
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float s = 1.f / len;
    v[2] *= s;
    v[1] *= s;
    v[0] *= s;
}
/* maths.cpp:7-9 */
static float saturate(float f) { return f < 0 ? 0 : (f > 1 ? 1 : f); }
static float clamp01(float x) { /* std::min(std::max(0.f, x), 1.f) color.cpp:42,52 */
    float m = (0.f < x) ? x : 0.f;
    return (1.f < m) ? 1.f : m;
}

/* ---- vertex stage ------------------------------------------------------- */
/* IShader.h:55-75 + the vertex() bodies IShader.cpp:5-10,23-28,47-52,65-71,85-92,117-124,170-174.
 * mvp = camera_vp*model and lmvp = light_vp*model are the per-vertex matrix
 * products of IShader.h:56,60, hoisted (same operator, same bits). Fields a
 * shader leaves unset (indeterminate in the reference, App. D6) are 0 here. */
static void vertex_stage(int shader, const HanaUniforms* u, const float* mvp, const float* lmvp,
                         const float* a, float* v) {
    float p[4] = {a[0], a[1], a[2], 1.f};
    float n[4] = {a[3], a[4], a[5], 1.f}; /* embed<4>(normal) fills with 1: vector.h:85 */
    memset(v, 0, V2F_N * sizeof(float));
    mat4_vec4(shader == HANA_SHADER_SHADOW ? lmvp : mvp, p, v + V_CLIP);
    if (shader == HANA_SHADER_SHADOW) return;
    float wn[3];
    for (int j = 0; j < 3; j++) { /* (row vector n) * model_I: matrix.h:118-123 with R1 = 1 */
        float col[4] = {u->model_I[j], u->model_I[4 + j], u->model_I[8 + j], u->model_I[12 + j]};
        wn[j] = dot4(n, col);
    }
    if (shader == HANA_SHADER_BLINN || shader == HANA_SHADER_NORMALMAP) {
        float wp[4];
        mat4_vec4(u->model, p, wp);
        v[V_WPOS] = wp[0]; v[V_WPOS + 1] = wp[1]; v[V_WPOS + 2] = wp[2];
    }
    if (shader == HANA_SHADER_BLINN || shader == HANA_SHADER_NORMALMAP || shader == HANA_SHADER_TEXTURE_LIGHT) {
        v[V_WNRM] = wn[0]; v[V_WNRM + 1] = wn[1]; v[V_WNRM + 2] = wn[2];
    }
    if (shader == HANA_SHADER_GROUND || shader == HANA_SHADER_TOON)
        v[V_INT] = saturate(dot3(wn, u->light_dir));
    else {
        v[V_UV] = a[6]; v[V_UV + 1] = a[7];
    }
}

void horacle_vertex(int shader, const HanaUniforms* u, const float* a2v8, float* v2f13) {
    float mvp[16], lmvp[16];
    horacle_mat4_mul(u->camera_vp, u->model, mvp);
    horacle_mat4_mul(u->light_vp, u->model, lmvp);
    vertex_stage(shader, u, mvp, lmvp, a2v8, v2f13);
}

/* ---- homogeneous clipping: graphics.cpp:29-161 --------------------------- */
static int inside_plane(const float* c, int plane) { /* graphics.cpp:29-49, EPSILON maths.h:6 */
    switch (plane) {
        case 0: return c[3] >= 1e-5f;
        case 1: return c[0] <= +c[3];
        case 2: return c[0] >= -c[3];
        case 3: return c[1] <= +c[3];
        case 4: return c[1] >= -c[3];
        case 5: return c[2] <= +c[3];
        default: return c[2] >= -c[3];
    }
}
static float intersect_t(const float* p, const float* c, int plane) { /* graphics.cpp:51-71 */
    switch (plane) {
        case 0: return (p[3] - 1e-5f) / (p[3] - c[3]);
        case 1: return (p[3] - p[0]) / ((p[3] - p[0]) - (c[3] - c[0]));
        case 2: return (p[3] + p[0]) / ((p[3] + p[0]) - (c[3] + c[0]));
        case 3: return (p[3] - p[1]) / ((p[3] - p[1]) - (c[3] - c[1]));
        case 4: return (p[3] + p[1]) / ((p[3] + p[1]) - (c[3] + c[1]));
        case 5: return (p[3] - p[2]) / ((p[3] - p[2]) - (c[3] - c[2]));
        default: return (p[3] + p[2]) / ((p[3] + p[2]) - (c[3] + c[2]));
    }
}
static int clip_plane(int plane, int n, const float* in, float* out) { /* graphics.cpp:73-110 */
    int m = 0;
    for (int i = 0; i < n; i++) {
        const float* prev = in + V2F_N * ((i - 1 + n) % n);
        const float* cur = in + V2F_N * i;
        int pi = inside_plane(prev, plane), ci = inside_plane(cur, plane);
        if (pi != ci) {
            float t = intersect_t(prev, cur, plane);
            float* d = out + V2F_N * m++;
            for (int k = 0; k < V2F_N; k++) d[k] = prev[k] + (cur[k] - prev[k]) * t; /* graphics.cpp:15 */
        }
        if (ci) memcpy(out + V2F_N * m++, cur, V2F_N * sizeof(float));
    }
    return m;
}
/* graphics.cpp:132-161. in: 3 vertices; out: capacity 10 vertices. Returns vertex count (0 or >= 3). */
int horacle_clip(const float* in39, float* out) {
    int vis = 1;
    for (int k = 0; k < 3; k++) {
        const float* c = in39 + V2F_N * k;
        vis &= (fabsf(c[0]) <= c[3] && fabsf(c[1]) <= c[3] && fabsf(c[2]) <= c[3]);
    }
    if (vis) {
        memcpy(out, in39, 3 * V2F_N * sizeof(float));
        return 3;
    }
    float a[10 * V2F_N], b[10 * V2F_N];
    memcpy(a, in39, 3 * V2F_N * sizeof(float));
    int n = 3;
    float *src = a, *dst = b;
    for (int plane = 0; plane < 7; plane++) { /* W, +X, -X, +Y, -Y, +Z, -Z: graphics.cpp:152-158 */
        n = clip_plane(plane, n, src, dst);
        if (n < 3) return 0;
        float* t = src; src = dst; dst = t;
    }
    memcpy(out, src, (size_t)n * V2F_N * sizeof(float));
    return n;
}

/* ---- texture + shadow fetches ------------------------------------------- */
/* TGAImage::get tgaimage.cpp:248-253 + TGAColor(p,bpp) tgaimage.h:46-53: B,G,R,A bytes, zeros outside. */
static void tga_get(const HOracleTex* t, int x, int y, uint8_t bgra[4]) {
    bgra[0] = bgra[1] = bgra[2] = bgra[3] = 0;
    if (!t || !t->data || x < 0 || y < 0 || x >= t->w || y >= t->h) return;
    const uint8_t* p = t->data + ((size_t)x + (size_t)y * t->w) * t->bpp;
    for (int i = 0; i < t->bpp && i < 4; i++) bgra[i] = p[i];
}
/* IShader.h:85-89 + Color(TGAColor) color.cpp:5 */
static void tex_diffuse(const HOracleTex* t, const float* uv, float rgb[3]) {
    uint8_t c[4];
    int w = t ? t->w : 0, h = t ? t->h : 0;
    tga_get(t, (int)(uv[0] * w), (int)(uv[1] * h), c);
    rgb[0] = c[2] / 255.f; rgb[1] = c[1] / 255.f; rgb[2] = c[0] / 255.f;
}
/* IShader.h:91-99 */
static void tex_normal(const HOracleTex* t, const float* uv, float res[3]) {
    uint8_t c[4];
    int w = t ? t->w : 0, h = t ? t->h : 0;
    tga_get(t, (int)(uv[0] * w), (int)(uv[1] * h), c);
    for (int i = 0; i < 3; i++) res[2 - i] = (float)c[i] / 255.f * 2.f - 1.f;
}
/* IShader.h:107-129. Returns 1 = lit. shadow: RGBA8 colour plane of the shadow map or NULL. */
static int lit_test(const HanaUniforms* u, const uint8_t* shadow, int sw, int sh, const float* dp, float ndl) {
    if (!(u->enable_shadow && shadow)) return 1;
    float width = (float)sw, height = (float)sh;
    float nx = dp[0] / dp[3], ny = dp[1] / dp[3];
    float px = (nx + 1) * 0.5f * (float)(int)width;  /* maths.cpp:21-22 */
    float py = (ny + 1) * 0.5f * (float)(int)height;
    float bias = 0.05f * (1 - ndl);
    if (bias < 0.005f) bias = 0.01f;
    float cur = dp[2] - bias;
    if (px < 0 || py < 0 || px >= width || py >= height) return 1;
    float closest = shadow[((size_t)(int)py * sw + (int)px) * 4] / 255.f; /* renderbuffer.cpp:46-50 */
    return cur < closest;
}

/* ---- fragment stage ----------------------------------------------------- */
/* Colour algebra color.cpp:38-64: '+' and '*float' clamp to [0,1], '*Color' does not.
 * Returns the three RGB floats the reference hands to set_color. */
static void lit_colour(const HanaUniforms* u, const float* albedo_tex, const float* N, const float* wpos,
                       const uint8_t* shadow, int sw, int sh, float rgb[3]) {
    /* shared tail of BlinnShader::fragment IShader.cpp:96-107 and NormalMapShader::fragment :149-160 */
    float ndl = saturate(dot3(N, u->light_dir));
    float V[3] = {u->view_pos[0] - wpos[0], u->view_pos[1] - wpos[1], u->view_pos[2] - wpos[2]};
    normalize3(V);
    float Hh[3] = {V[0] + u->light_dir[0], V[1] + u->light_dir[1], V[2] + u->light_dir[2]};
    normalize3(Hh);
    float sp = powf(saturate(dot3(N, Hh)), u->gloss);
    float wp4[4] = {wpos[0], wpos[1], wpos[2], 1.f}, dp[4];
    mat4_vec4(u->light_vp, wp4, dp);
    float shadow_f = (float)lit_test(u, shadow, sw, sh, dp, ndl);
    float i_ndl = ndl > 1.f ? 1.f : (ndl < 0.f ? 0.f : ndl);
    float i_sp = sp > 1.f ? 1.f : (sp < 0.f ? 0.f : sp);
    float i_sh = shadow_f > 1.f ? 1.f : (shadow_f < 0.f ? 0.f : shadow_f);
    for (int k = 0; k < 3; k++) {
        float albedo = albedo_tex[k] * u->mat_color[k];
        float ambient = u->ambient[k] * albedo;
        float diffuse = clamp01(u->light_color[k] * albedo * i_ndl);
        float spec = clamp01(u->light_color[k] * u->mat_specular[k] * i_sp);
        float sum = clamp01(diffuse + spec);
        rgb[k] = clamp01(ambient + clamp01(sum * i_sh));
    }
}

static void fragment_stage(int shader, const HanaUniforms* u, const float* v, const HOracleTex* diffuse,
                           const HOracleTex* normal, const uint8_t* shadow, int sw, int sh, float rgb[3]) {
    switch (shader) {
        case HANA_SHADER_SHADOW: { /* IShader.cpp:176-180: White * clip_pos.z */
            float f = v[V_CLIP + 2];
            f = f > 1.f ? 1.f : (f < 0.f ? 0.f : f);
            rgb[0] = rgb[1] = rgb[2] = clamp01(1.f * f);
            return;
        }
        case HANA_SHADER_GROUND: { /* IShader.cpp:12-15 */
            float f = v[V_INT];
            f = f > 1.f ? 1.f : (f < 0.f ? 0.f : f);
            rgb[0] = rgb[1] = rgb[2] = clamp01(1.f * f);
            return;
        }
        case HANA_SHADER_TOON: { /* IShader.cpp:30-39: thresholds are DOUBLE compares (App. A.8) */
            float in = v[V_INT];
            if ((double)in > .85) in = 1;
            else if ((double)in > .60) in = (float).80;
            else if ((double)in > .45) in = (float).60;
            else if ((double)in > .30) in = (float).45;
            else if ((double)in > .15) in = (float).30;
            float f = in > 1.f ? 1.f : (in < 0.f ? 0.f : in);
            rgb[0] = clamp01(1.f * f);
            rgb[1] = clamp01((155 / 255.f) * f);
            rgb[2] = clamp01(0.f * f);
            return;
        }
        case HANA_SHADER_TEXTURE: /* IShader.cpp:54-57 */
            tex_diffuse(diffuse, v + V_UV, rgb);
            return;
        case HANA_SHADER_TEXTURE_LIGHT: { /* IShader.cpp:73-77 */
            float f = saturate(dot3(v + V_WNRM, u->light_dir));
            float t[3];
            tex_diffuse(diffuse, v + V_UV, t);
            f = f > 1.f ? 1.f : (f < 0.f ? 0.f : f);
            for (int k = 0; k < 3; k++) rgb[k] = clamp01(t[k] * f);
            return;
        }
        case HANA_SHADER_BLINN: { /* IShader.cpp:94-109 */
            float N[3] = {v[V_WNRM], v[V_WNRM + 1], v[V_WNRM + 2]};
            normalize3(N);
            float t[3];
            tex_diffuse(diffuse, v + V_UV, t);
            lit_colour(u, t, N, v + V_WPOS, shadow, sw, sh, rgb);
            return;
        }
        default: { /* NormalMapShader::fragment IShader.cpp:126-162 */
            float x = v[V_WNRM], y = v[V_WNRM + 1], z = v[V_WNRM + 2];
            float T[3] = {x * y / sqrtf(x * x + z * z), sqrtf(x * x + z * z), z * y / sqrtf(x * x + z * z)};
            float B[3] = {y * T[2] - z * T[1], z * T[0] - x * T[2], x * T[1] - y * T[0]}; /* cross(normal,t) vector.h:97-99 */
            float bump[3];
            tex_normal(normal, v + V_UV, bump);
            bump[0] = bump[0] * u->bump_scale;
            bump[1] = bump[1] * u->bump_scale;
            bump[2] = (float)sqrt(1.0 - (double)saturate(dot2(bump, bump))); /* DOUBLE sqrt: IShader.cpp:144 */
            float r0[3] = {T[0], B[0], x}, r1[3] = {T[1], B[1], y}, r2[3] = {T[2], B[2], z};
            float N[3] = {dot3(r0, bump), dot3(r1, bump), dot3(r2, bump)};
            normalize3(N);
            float t[3];
            tex_diffuse(diffuse, v + V_UV, t);
            lit_colour(u, t, N, v + V_WPOS, shadow, sw, sh, rgb);
            return;
        }
    }
}

void horacle_fragment(int shader, const HanaUniforms* u, const float* v2f13, const HOracleTex* diffuse,
                      const HOracleTex* normal, const uint8_t* shadow, int sw, int sh, float* rgb3) {
    fragment_stage(shader, u, v2f13, diffuse, normal, shadow, sw, sh, rgb3);
}

/* ---- rasterisation: graphics.cpp:314-376 -------------------------------- */
/* barycentric graphics.cpp:222-233 at integer pixel (px,py). Returns 0 if the pixel is rejected. */
int horacle_barycentric(const float* A, const float* B, const float* C, int px, int py, float* w) {
    float Px = (float)px, Py = (float)py; /* Vector2f(Vector2i) vector.cpp:6 */
    float s0[3] = {C[0] - A[0], B[0] - A[0], A[0] - Px};
    float s1[3] = {C[1] - A[1], B[1] - A[1], A[1] - Py};
    float ux = s0[1] * s1[2] - s0[2] * s1[1];
    float uy = s0[2] * s1[0] - s0[0] * s1[2];
    float uz = s0[0] * s1[1] - s0[1] * s1[0];
    if ((double)fabsf(uz) > 1e-2) {
        w[0] = 1.f - (ux + uy) / uz;
        w[1] = uy / uz;
        w[2] = ux / uz;
    } else {
        w[0] = -1; w[1] = 1; w[2] = 1;
    }
    return !(w[0] < 0 || w[1] < 0 || w[2] < 0);
}

static void raster_triangle(int shader, const HanaUniforms* u, const float* tri /* 3 x v2f */, const HOracleTex* diffuse,
                            const HOracleTex* normal, const uint8_t* shadow, int sw, int sh, int W, int H,
                            uint8_t* color, float* depth, uint32_t* primid, uint32_t order, HOracleCounters* cnt) {
    float ndc[3][3], sc[3][2], sd[3], rw[3];
    for (int k = 0; k < 3; k++) { /* graphics.cpp:317: true divisions by w */
        const float* c = tri + V2F_N * k;
        ndc[k][0] = c[0] / c[3]; ndc[k][1] = c[1] / c[3]; ndc[k][2] = c[2] / c[3];
    }
    /* is_back_facing graphics.cpp:172-180 */
    float area = ndc[0][0] * ndc[1][1] - ndc[0][1] * ndc[1][0] + ndc[1][0] * ndc[2][1] - ndc[1][1] * ndc[2][0] +
                 ndc[2][0] * ndc[0][1] - ndc[2][1] * ndc[0][0];
    if (area <= 0) return;
    if (cnt) cnt->tris_raster++;
    for (int k = 0; k < 3; k++) { /* viewport_transform maths.cpp:20-25 */
        sc[k][0] = (ndc[k][0] + 1) * 0.5f * (float)W;
        sc[k][1] = (ndc[k][1] + 1) * 0.5f * (float)H;
        sd[k] = (ndc[k][2] + 1) * 0.5f;
        rw[k] = 1 / tri[V2F_N * k + 3]; /* graphics.cpp:336 */
    }
    float bmin[2] = {3.402823466e+38f, 3.402823466e+38f}, bmax[2] = {-3.402823466e+38f, -3.402823466e+38f};
    float lim[2] = {(float)(W - 1), (float)(H - 1)};
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 2; j++) { /* graphics.cpp:342-347 */
            float mn = sc[k][j] < bmin[j] ? sc[k][j] : bmin[j];
            bmin[j] = 0.f < mn ? mn : 0.f;
            float mx = bmax[j] < sc[k][j] ? sc[k][j] : bmax[j];
            bmax[j] = mx < lim[j] ? mx : lim[j];
        }
    for (int px = (int)bmin[0]; (float)px <= bmax[0]; px++)     /* x OUTER: graphics.cpp:350 */
        for (int py = (int)bmin[1]; (float)py <= bmax[1]; py++) { /* y inner */
            float w[3];
            if (cnt) cnt->bbox_pixels++;
            if (!horacle_barycentric(sc[0], sc[1], sc[2], px, py, w)) continue;
            if (cnt) cnt->inside++;
            float z = dot3(sd, w); /* interpolate_depth graphics.cpp:186-194 */
            size_t idx = (size_t)py * W + px;
            if (z > depth[idx]) continue; /* graphics.cpp:359 */
            /* interpolate_varyings graphics.cpp:205-220 */
            float w0 = rw[0] * w[0], w1 = rw[1] * w[1], w2 = rw[2] * w[2];
            float norm = 1 / (w0 + w1 + w2);
            float v[V2F_N];
            for (int i = 0; i < V2F_N; i++) {
                float sum = tri[i] * w0 + tri[V2F_N + i] * w1 + tri[2 * V2F_N + i] * w2;
                v[i] = sum * norm;
            }
            float rgb[3];
            fragment_stage(shader, u, v, diffuse, normal, shadow, sw, sh, rgb);
            depth[idx] = z;                              /* renderbuffer.cpp:27-30 */
            color[idx * 4 + 0] = (uint8_t)(rgb[0] * 255); /* renderbuffer.cpp:38-44: alpha untouched */
            color[idx * 4 + 1] = (uint8_t)(rgb[1] * 255);
            color[idx * 4 + 2] = (uint8_t)(rgb[2] * 255);
            if (primid) primid[idx] = order;
            if (cnt) cnt->zpass++;
        }
}

/* graphics_draw_triangle graphics.cpp:378-407 over flat arrays.
 * a2v: ncorners x 8 floats; color/depth: in/out W*H RGBA8 / f32 (y up);
 * shadow: RGBA8 colour plane of the shadow map (or NULL); primid/cnt optional. */
void horacle_draw(int shader, const HanaUniforms* u, const float* a2v, int ncorners, const HOracleTex* diffuse,
                  const HOracleTex* normal, const uint8_t* shadow, int sw, int sh, int W, int H, uint8_t* color,
                  float* depth, uint32_t* primid, HOracleCounters* cnt) {
    float mvp[16], lmvp[16];
    horacle_mat4_mul(u->camera_vp, u->model, mvp);
    horacle_mat4_mul(u->light_vp, u->model, lmvp);
    for (int f = 0; f < ncorners / 3; f++) {
        float in[3 * V2F_N], poly[10 * V2F_N], tri[3 * V2F_N];
        for (int j = 0; j < 3; j++) {
            vertex_stage(shader, u, mvp, lmvp, a2v + (size_t)(f * 3 + j) * 8, in + V2F_N * j);
            if (cnt) cnt->corners++;
        }
        int n = horacle_clip(in, poly);
        for (int j = 0; j + 2 < n; j++) { /* fan (0, j+1, j+2): graphics.cpp:394-405 */
            memcpy(tri, poly, V2F_N * sizeof(float));
            memcpy(tri + V2F_N, poly + V2F_N * (j + 1), V2F_N * sizeof(float));
            memcpy(tri + 2 * V2F_N, poly + V2F_N * (j + 2), V2F_N * sizeof(float));
            raster_triangle(shader, u, tri, diffuse, normal, shadow, sw, sh, W, H, color, depth, primid,
                            (uint32_t)f * 8u + (uint32_t)j, cnt);
        }
    }
}

/* DrawModel::draw scene.h:53-99 over caller-owned buffers: optional shadow
 * pass into (shadow_color, shadow_depth) [both W*H, pre-cleared by the caller],
 * main pass into (color, depth), then the shadow map is cleared to
 * (0,0,0,a=clear_alpha) / FLT_MAX as scene.h:94-98 does. cnt[0] = shadow pass,
 * cnt[1] = main pass (optional). */
void horacle_draw_model(int shader, const HanaUniforms* u, const float* a2v, int ncorners, const HOracleTex* diffuse,
                        const HOracleTex* normal, int W, int H, uint8_t* color, float* depth, uint8_t* shadow_color,
                        float* shadow_depth, uint32_t* primid, HOracleCounters* cnt2) {
    const uint8_t* sm = NULL;
    if (u->enable_shadow) {
        horacle_draw(HANA_SHADER_SHADOW, u, a2v, ncorners, NULL, NULL, NULL, 0, 0, W, H, shadow_color, shadow_depth,
                     NULL, cnt2 ? &cnt2[0] : NULL);
        sm = shadow_color;
    }
    horacle_draw(shader, u, a2v, ncorners, diffuse, normal, sm, W, H, W, H, color, depth, primid,
                 cnt2 ? &cnt2[1] : NULL);
    if (u->enable_shadow) {
        for (size_t i = 0; i < (size_t)W * H; i++) {
            shadow_color[i * 4 + 0] = 0; shadow_color[i * 4 + 1] = 0; shadow_color[i * 4 + 2] = 0;
            shadow_color[i * 4 + 3] = 1; /* (uchar)(255*255.f): App. D5, observed value on x86-64 */
            shadow_depth[i] = 3.402823466e+38f;
        }
    }
}
