// ref_driver.cpp — headless C-ABI driver around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY. This file is compiled together with the
// reference's own .cpp files (read in place from /root/reference, never copied
// into this repository) into oracle/_ref/libhana_ref.so (timing build) and
// oracle/_ref/libhana_ref_inst.so (instrumented build: primitive-ID buffer,
// workload counters, stage-level entry points). Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load
// them. The product path (libhana_b200.so) never does.
//
// It replaces win32.cpp + main.cpp (SURVEY.md §2 "Platform"/"App main loop")
// and mirrors what SingleModelScene does (scene.cpp:74-103, :115-125) with the
// scene pieces exposed so that tests can pick the shader, move the camera with
// the reference's own Camera::update_transform (camera.cpp:63-70) and read the
// uniforms DrawModel::draw built (scene.h:55-71).
#include "scene.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/hana_b200.h"

// ---- platform stubs needed by camera.cpp (platform.h:23,30,32,36) ----------
void input_query_cursor(window_t*, float* xpos, float* ypos) { *xpos = 0; *ypos = 0; }
int input_key_pressed(window_t*, keycode_t) { return 0; }
void* window_get_userdata(window_t*) { return nullptr; }
float platform_get_time(void) { return 0.f; }

// ---- instrumentation hooks (called from the sed-instrumented graphics.cpp) --
struct HrefCounters {
    uint64_t corners;       // vertex() invocations
    uint64_t tris_raster;   // rasterize_triangle calls
    uint64_t bbox_pixels;   // barycentric evaluations (graphics.cpp:352)
    uint64_t inside;        // passed the inside test (graphics.cpp:353)
    uint64_t zpass;         // fragments shaded + written (graphics.cpp:371)
};
static HrefCounters g_cnt;
static uint32_t g_cur_prim = 0xFFFFFFFFu;
static uint32_t* g_primid = nullptr;   // W*H of the buffer currently drawn into
static int g_primid_w = 0;
static float* g_a2v_rec = nullptr;     // ncorners*8 floats, filled when non-null

extern "C" {
void href_hook_a2v(int face, int nth, const shader_struct_a2v* a) {
    g_cnt.corners++;
    if (g_a2v_rec) {
        float* d = g_a2v_rec + (size_t)(face * 3 + nth) * 8;
        d[0] = a->obj_pos.x; d[1] = a->obj_pos.y; d[2] = a->obj_pos.z;
        d[3] = a->obj_normal.x; d[4] = a->obj_normal.y; d[5] = a->obj_normal.z;
        d[6] = a->uv.x; d[7] = a->uv.y;
    }
}
void href_hook_prim(int face, int fan) {
    g_cnt.tris_raster++;
    g_cur_prim = (uint32_t)face * 8u + (uint32_t)fan;
}
void href_hook_bbox(void) { g_cnt.bbox_pixels++; }
void href_hook_inside(void) { g_cnt.inside++; }
void href_hook_frag(int x, int y) {
    g_cnt.zpass++;
    if (g_primid) g_primid[(size_t)y * g_primid_w + x] = g_cur_prim;
}
}

struct HrefScene {
    RenderBuffer* fb;
    Camera* camera;
    GameObject* light;
    GameObject_StaticModel* go;
    Material* material;
    IShader* shaders[HANA_SHADER_COUNT];
    DrawModel* dm;
    int shader_id;
    std::vector<uint32_t> primid;
};

static IShader* make_shader(int id) {
    switch (id) {
        case HANA_SHADER_SHADOW: return new ShadowShader();
        case HANA_SHADER_BLINN: return new BlinnShader();
        case HANA_SHADER_NORMALMAP: return new NormalMapShader();
        case HANA_SHADER_GROUND: return new GroundShader();
        case HANA_SHADER_TOON: return new ToonShader();
        case HANA_SHADER_TEXTURE: return new TextureShader();
        case HANA_SHADER_TEXTURE_LIGHT: return new TextureWithLightShader();
    }
    return nullptr;
}

static void copy_m(float* dst, const Matrix4x4& m) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) dst[i * 4 + j] = m[i][j];
}

extern "C" {

int href_instrumented(void) {
#ifdef HREF_INSTRUMENTED
    return 1;
#else
    return 0;
#endif
}

// Mirrors Scene::Scene (scene.cpp:3-9) + SingleModelScene ctor (scene.cpp:74-103).
HrefScene* href_scene_create(const char* obj_path, int w, int h, int shader_id) {
    if (shader_id < 0 || shader_id >= HANA_SHADER_COUNT) return nullptr;
    HrefScene* s = new HrefScene();
    s->fb = new RenderBuffer(w, h);
    float aspect = (float)w / (float)h;
    s->camera = new Camera(CAMERA_POSITION, CAMERA_TARGET, aspect);
    s->light = new GameObject(Vector3f(2, 2, 2));
    s->go = new GameObject_StaticModel(obj_path);
    s->material = new Material();
    s->material->diffuse_map = s->go->model->get_diffuse_map();
    s->material->normal_map = s->go->model->get_normal_map();
    s->material->specular_map = s->go->model->get_specular_map();
    s->material->color = Color::White;
    s->material->specular = Color::White;
    s->material->gloss = 50;
    s->material->bump_scale = 1;
    for (int i = 0; i < HANA_SHADER_COUNT; i++) s->shaders[i] = make_shader(i);
    s->shader_id = shader_id;
    s->dm = new DrawModel(s->light, s->go, s->material, s->shaders[shader_id]);
    return s;
}

void href_scene_destroy(HrefScene* s) {
    if (!s) return;
    delete s->dm;
    delete s->material;
    delete s->go;
    for (int i = 0; i < HANA_SHADER_COUNT; i++) delete s->shaders[i];
    delete s->camera;
    delete s->light;
    delete s->fb;
    delete s;
}

int href_scene_nfaces(HrefScene* s) { return s->go->model->nfaces(); }

void href_scene_set_shader(HrefScene* s, int shader_id) {
    // same as the Q key (scene.cpp:134-139): a fresh DrawModel with the other shader
    delete s->dm;
    s->shader_id = shader_id;
    s->dm = new DrawModel(s->light, s->go, s->material, s->shaders[shader_id]);
}

void href_camera_set(HrefScene* s, const float* pos, const float* target) {
    s->camera->set_transform(Vector3f(pos[0], pos[1], pos[2]), Vector3f(target[0], target[1], target[2]));
}
// Camera::update_transform (camera.cpp:63-70) with an explicit Motion.
void href_camera_motion(HrefScene* s, float orbit_x, float orbit_y, float pan_x, float pan_y, float dolly) {
    Motion m;
    m.orbit = Vector2f(orbit_x, orbit_y);
    m.pan = Vector2f(pan_x, pan_y);
    m.dolly = dolly;
    s->camera->update_transform(m);
}
void href_camera_get(HrefScene* s, float* pos) {
    Vector3f p = s->camera->get_position();
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
}
void href_light_set(HrefScene* s, const float* pos) { s->light->transform.position = Vector3f(pos[0], pos[1], pos[2]); }
void href_material_set(HrefScene* s, const float* color, const float* specular, float gloss, float bump) {
    s->material->color = Color(color[0], color[1], color[2], color[3]);
    s->material->specular = Color(specular[0], specular[1], specular[2], specular[3]);
    s->material->gloss = gloss;
    s->material->bump_scale = bump;
}
void href_model_transform(HrefScene* s, const float* pos, const float* rot_deg, const float* scale) {
    s->go->transform.position = Vector3f(pos[0], pos[1], pos[2]);
    s->go->transform.rotation = Vector3f(rot_deg[0], rot_deg[1], rot_deg[2]);
    s->go->transform.scale = Vector3f(scale[0], scale[1], scale[2]);
}

// One frame: the per-frame clear of main.cpp:152-153, then DrawModel::draw
// (scene.h:53-99). Returns the wall time of draw() alone, in seconds.
// clear == 0 leaves the frame and shadow map as they are (e.g. the ctor state
// of the reference's very first frame: colour (0,0,0,255), depth 1.0).
double href_render(HrefScene* s, int enable_shadow, int clear) {
    if (clear) {
        s->fb->renderbuffer_clear_color(Color::Black);
        s->fb->renderbuffer_clear_depth(std::numeric_limits<float>::max());
    }
    if (clear && enable_shadow && s->dm->shdaow_map) {
        // D8: make frame 0 start from the same shadow-map state as every later frame
        s->dm->shdaow_map->renderbuffer_clear_color(Color::Black);
        s->dm->shdaow_map->renderbuffer_clear_depth(std::numeric_limits<float>::max());
    }
#ifdef HREF_INSTRUMENTED
    s->primid.assign((size_t)s->fb->width * s->fb->height, 0xFFFFFFFFu);
    g_primid = nullptr;  // the shadow pass (if any) must not write the frame's id buffer
    g_primid_w = s->fb->width;
#endif
    auto t0 = std::chrono::steady_clock::now();
    s->dm->draw(s->camera, s->fb, enable_shadow != 0);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The first draw() with shadows lazily creates the shadow map with depth 1.0
// (renderbuffer.cpp:7); call this once so every measured frame sees FLT_MAX.
void href_warmup(HrefScene* s, int enable_shadow) { href_render(s, enable_shadow, 1); }

void href_get_frame(HrefScene* s, const uint8_t** color, const float** depth, int* w, int* h) {
    *color = s->fb->color_buffer;
    *depth = s->fb->depth_buffer;
    *w = s->fb->width;
    *h = s->fb->height;
}

// The uniforms exactly as DrawModel::draw left them in ShaderData (valid after a render).
void href_get_uniforms(HrefScene* s, HanaUniforms* u) {
    ShaderData* d = s->dm->shader_data;
    memset(u, 0, sizeof(*u));
    copy_m(u->model, d->model_matrix);
    copy_m(u->model_I, d->model_matrix_I);
    copy_m(u->camera_vp, d->camera_vp_matrix);
    copy_m(u->light_vp, d->light_vp_matrix);
    u->view_pos[0] = d->view_Pos.x; u->view_pos[1] = d->view_Pos.y; u->view_pos[2] = d->view_Pos.z;
    u->light_dir[0] = d->light_dir.x; u->light_dir[1] = d->light_dir.y; u->light_dir[2] = d->light_dir.z;
    u->gloss = d->matrial->gloss;
    u->bump_scale = d->matrial->bump_scale;
    for (int i = 0; i < 4; i++) {
        u->light_color[i] = d->light_color[i];
        u->ambient[i] = d->ambient[i];
        u->mat_color[i] = d->matrial->color[i];
        u->mat_specular[i] = d->matrial->specular[i];
    }
    u->enable_shadow = d->enable_shadow ? 1 : 0;
}

// Gathers the a2v stream through the reference's accessors in the order
// graphics.cpp:380-386 does. NOTE: Model::normal re-normalises in place
// (model.cpp:108-111), so this mutates the model the same way one pass does.
int href_model_export_a2v(HrefScene* s, float* out, int capacity_corners) {
    Model* m = s->go->model;
    int n = m->nfaces() * 3;
    if (!out) return n;
    if (capacity_corners < n) return -1;
    for (int i = 0; i < m->nfaces(); i++)
        for (int j = 0; j < 3; j++) {
            Vector3f p = m->vert(i, j);
            Vector3f nn = m->normal(i, j);
            Vector2f t = m->uv(i, j);
            float* d = out + (size_t)(i * 3 + j) * 8;
            d[0] = p.x; d[1] = p.y; d[2] = p.z; d[3] = nn.x; d[4] = nn.y; d[5] = nn.z; d[6] = t.x; d[7] = t.y;
        }
    return n;
}

// which: 0 diffuse, 1 normal (tangent), 2 specular.
int href_texture_info(HrefScene* s, int which, int* w, int* h, int* bpp, const uint8_t** data) {
    TGAImage* t = which == 0 ? s->material->diffuse_map : which == 1 ? s->material->normal_map : s->material->specular_map;
    if (!t) return -1;
    *w = t->get_width();
    *h = t->get_height();
    *bpp = t->get_bytespp();
    *data = t->buffer();
    return 0;
}

// --- the reference's own codecs, for the asset / output tests ---
// TGAImage::write_tga_file (tgaimage.cpp:145) of `data` (rows in file order).
int href_tga_write(const char* path, const uint8_t* data, int w, int h, int bpp, int rle) {
    TGAImage img(w, h, bpp);
    memcpy(img.buffer(), data, (size_t)w * h * bpp);
    return img.write_tga_file(path, rle != 0) ? 0 : -1;
}
// TGAImage::read_tga_file (tgaimage.cpp:40) [+ Model::load_texture's flip_vertically, model.cpp:81].
// out == NULL: returns the size only.
int href_tga_read(const char* path, int model_flip, uint8_t* out, int* w, int* h, int* bpp) {
    TGAImage img;
    if (!img.read_tga_file(path)) return -1;
    if (model_flip) img.flip_vertically();
    *w = img.get_width();
    *h = img.get_height();
    *bpp = img.get_bytespp();
    if (out) memcpy(out, img.buffer(), (size_t)(*w) * (*h) * (*bpp));
    return 0;
}
// A fresh Model (model.cpp:6) walked `pass` times the way graphics.cpp:380-386 walks it; the a2v stream of the last walk.
int href_obj_a2v(const char* path, int pass, float* out, int capacity_corners) {
    Model m(path);
    int n = m.nfaces() * 3;
    if (!out) return n;
    if (capacity_corners < n) return -1;
    for (int k = 1; k <= pass; k++)
        for (int i = 0; i < m.nfaces(); i++)
            for (int j = 0; j < 3; j++) {
                Vector3f p = m.vert(i, j);
                Vector3f nn = m.normal(i, j);
                Vector2f t = m.uv(i, j);
                float* d = out + (size_t)(i * 3 + j) * 8;
                d[0] = p.x; d[1] = p.y; d[2] = p.z; d[3] = nn.x; d[4] = nn.y; d[5] = nn.z; d[6] = t.x; d[7] = t.y;
            }
    return n;
}

#ifdef HREF_INSTRUMENTED
void href_counters_reset(void) { memset(&g_cnt, 0, sizeof(g_cnt)); }
void href_counters_get(uint64_t* out5) {
    out5[0] = g_cnt.corners; out5[1] = g_cnt.tris_raster; out5[2] = g_cnt.bbox_pixels;
    out5[3] = g_cnt.inside; out5[4] = g_cnt.zpass;
}
// Record the a2v stream exactly as the NEXT pass sees it (ncorners*8 floats), or NULL to stop.
void href_record_a2v(float* dst) { g_a2v_rec = dst; }

// A single graphics_draw_triangle pass (graphics.cpp:378) into a caller-owned
// target, with the ShaderData of the last href_render (call that first).
// shader_id selects the shader object; the shadow map the main-pass shaders
// read is whatever `shadow_color` holds (NULL -> lit everywhere, IShader.h:109).
// color/depth are in/out: existing contents take part in the depth test.
// primid (optional, W*H) receives face*8+fan of the last writer.
void href_draw_pass(HrefScene* s, int shader_id, int w, int h, uint8_t* color, float* depth,
                    const uint8_t* shadow_color, int shadow_w, int shadow_h, uint32_t* primid) {
    RenderBuffer target(w, h);
    memcpy(target.color_buffer, color, (size_t)w * h * 4);
    memcpy(target.depth_buffer, depth, (size_t)w * h * 4);
    ShaderData* d = s->dm->shader_data;
    RenderBuffer* saved_shadow = d->shadow_map;
    bool saved_enable = d->enable_shadow;
    RenderBuffer* sm = nullptr;
    if (shadow_color) {
        sm = new RenderBuffer(shadow_w, shadow_h);
        memcpy(sm->color_buffer, shadow_color, (size_t)shadow_w * shadow_h * 4);
        d->shadow_map = sm;
        d->enable_shadow = true;
    } else {
        d->shadow_map = nullptr;
    }
    IShader* sh = s->shaders[shader_id];
    sh->shader_data = d;
    DrawData dd;
    dd.model = s->go->model;
    dd.shader = sh;
    dd.render_buffer = &target;
    g_primid = primid;
    g_primid_w = w;
    graphics_draw_triangle(&dd);
    g_primid = nullptr;
    memcpy(color, target.color_buffer, (size_t)w * h * 4);
    memcpy(depth, target.depth_buffer, (size_t)w * h * 4);
    d->shadow_map = saved_shadow;
    d->enable_shadow = saved_enable;
    delete sm;
}

// Override the uniforms the stage entry points / href_draw_pass see.
void href_set_uniforms(HrefScene* s, const HanaUniforms* u) {
    ShaderData* d = s->dm->shader_data;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            d->model_matrix[i][j] = u->model[i * 4 + j];
            d->model_matrix_I[i][j] = u->model_I[i * 4 + j];
            d->camera_vp_matrix[i][j] = u->camera_vp[i * 4 + j];
            d->light_vp_matrix[i][j] = u->light_vp[i * 4 + j];
        }
    d->view_Pos = Vector3f(u->view_pos[0], u->view_pos[1], u->view_pos[2]);
    d->light_dir = Vector3f(u->light_dir[0], u->light_dir[1], u->light_dir[2]);
    d->light_color = Color(u->light_color[0], u->light_color[1], u->light_color[2], u->light_color[3]);
    d->ambient = Color(u->ambient[0], u->ambient[1], u->ambient[2], u->ambient[3]);
    d->matrial->color = Color(u->mat_color[0], u->mat_color[1], u->mat_color[2], u->mat_color[3]);
    d->matrial->specular = Color(u->mat_specular[0], u->mat_specular[1], u->mat_specular[2], u->mat_specular[3]);
    d->matrial->gloss = u->gloss;
    d->matrial->bump_scale = u->bump_scale;
    d->enable_shadow = u->enable_shadow != 0;
}

// IShader::vertex (IShader.h:52) on one a2v record (8 floats) -> v2f (13 floats).
void href_stage_vertex(HrefScene* s, int shader_id, const float* a2v8, float* v2f13) {
    IShader* sh = s->shaders[shader_id];
    sh->shader_data = s->dm->shader_data;
    shader_struct_a2v a;
    a.obj_pos = Vector3f(a2v8[0], a2v8[1], a2v8[2]);
    a.obj_normal = Vector3f(a2v8[3], a2v8[4], a2v8[5]);
    a.uv = Vector2f(a2v8[6], a2v8[7]);
    shader_struct_v2f v;
    memset((void*)&v, 0, sizeof(v));
    v = sh->vertex(&a);
    memcpy(v2f13, &v, sizeof(v));
}

// IShader::fragment (IShader.h:53) on one v2f -> Color rgba floats. The shadow
// map read by is_in_shadow is `shadow_color` (RGBA8, may be NULL).
int href_stage_fragment(HrefScene* s, int shader_id, const float* v2f13, float* rgba,
                        const uint8_t* shadow_color, int shadow_w, int shadow_h) {
    IShader* sh = s->shaders[shader_id];
    ShaderData* d = s->dm->shader_data;
    sh->shader_data = d;
    RenderBuffer* saved = d->shadow_map;
    RenderBuffer sm(shadow_color ? shadow_w : 1, shadow_color ? shadow_h : 1);
    if (shadow_color) {
        memcpy(sm.color_buffer, shadow_color, (size_t)shadow_w * shadow_h * 4);
        d->shadow_map = &sm;
    } else {
        d->shadow_map = nullptr;
    }
    shader_struct_v2f v;
    memcpy((void*)&v, v2f13, sizeof(v));
    Color c;
    bool discard = sh->fragment(&v, c);
    rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
    d->shadow_map = saved;
    return discard ? 1 : 0;
}
#endif  // HREF_INSTRUMENTED

}  // extern "C"
