/*
 * graphics_dropin.cpp — drop-in replacement for the reference's graphics.cpp.
 *
 * Defines the reference's one public draw entry point with its exact signature
 *     void graphics_draw_triangle(DrawData* app_data);          (graphics.h:15, graphics.cpp:378-407)
 * so that the reference's own scene code (DrawModel::draw, scene.h:53-99) links and runs unchanged on
 * top of the CUDA path: compile this file INSTEAD of graphics.cpp, against the reference's headers, and
 * link libhana_b200.so. It only marshals: Model accessors -> a2v stream, TGAImage
 * buffers -> textures, ShaderData/Material -> HanaUniforms, RenderBuffer host planes -> C-ABI calls.
 * All rendering happens in the library; there is no CPU fallback — an IShader subclass outside the
 * closed device set, or a missing GPU, aborts with a message (the reference's entry point returns void).
 *
 * Semantics per call are those of the reference: the target's existing colour/depth take part in the
 * depth test and only covered pixels change. The host RenderBuffer stays the source of truth (it is
 * uploaded before and downloaded after the pass), because the reference's callers read and clear it
 * directly (scene.h:96-97, main.cpp:152-153, IShader.h:124, win32.cpp:361).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <typeinfo>
#include <vector>

#include "graphics.h" /* the reference's header: DrawData, Model, IShader, RenderBuffer */

#include "../../include/hana_b200.h"

namespace {

struct CachedModel {
    hana_model* dev = nullptr;
    std::vector<float> a2v; /* what the device copy holds */
};
struct CachedTexture {
    hana_texture* dev = nullptr;
    const unsigned char* buffer = nullptr;
    int w = 0, h = 0, bpp = 0;
    unsigned long long fingerprint = 0;
};
struct DeviceCache {
    hana_ctx* ctx = nullptr;
    std::map<Model*, CachedModel> models;
    std::map<TGAImage*, CachedTexture> textures;
    hana_rb* target = nullptr;
    hana_rb* shadow = nullptr;
    std::vector<float> a2v;
};
DeviceCache g;

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "hana_b200 graphics_draw_triangle: %s: %s\n", what, hana_last_error());
    std::abort();
}
#define CK(call)                 \
    do {                         \
        if ((call) != HANA_OK) die(#call); \
    } while (0)

int shader_id_of(IShader* s) { /* the closed device set: IShader.h:132-165 */
    if (dynamic_cast<ShadowShader*>(s)) return HANA_SHADER_SHADOW;
    if (dynamic_cast<BlinnShader*>(s)) return HANA_SHADER_BLINN;
    if (dynamic_cast<NormalMapShader*>(s)) return HANA_SHADER_NORMALMAP;
    if (dynamic_cast<GroundShader*>(s)) return HANA_SHADER_GROUND;
    if (dynamic_cast<ToonShader*>(s)) return HANA_SHADER_TOON;
    if (dynamic_cast<TextureShader*>(s)) return HANA_SHADER_TEXTURE;
    if (dynamic_cast<TextureWithLightShader*>(s)) return HANA_SHADER_TEXTURE_LIGHT;
    return -1;
}

void copy_matrix(float* dst, const Matrix4x4& m) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) dst[i * 4 + j] = m[i][j];
}

/* FNV-1a over 4096 evenly spaced bytes plus the last one: catches TGAImage flips, scale, clear and a reallocated image at
 * the same address without reading megabytes per draw. An in-place edit of a few texels between two draws can escape it;
 * hana_dropin_invalidate() (below) drops every cached texture for callers that paint into their textures. */
unsigned long long fingerprint_of(const unsigned char* p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    const size_t step = n > 4096 ? n / 4096 : 1;
    for (size_t i = 0; i < n; i += step) h = (h ^ p[i]) * 1099511628211ull;
    if (n) h = (h ^ p[n - 1]) * 1099511628211ull;
    return h;
}

hana_texture* texture_of(TGAImage* img) {
    if (!img || !img->buffer() || img->get_width() <= 0 || img->get_height() <= 0) return nullptr; /* fetches return 0 */
    const int w = img->get_width(), h = img->get_height(), bpp = img->get_bytespp();
    const unsigned long long fp = fingerprint_of(img->buffer(), (size_t)w * h * bpp);
    CachedTexture& c = g.textures[img];
    if (c.dev && c.buffer == img->buffer() && c.w == w && c.h == h && c.bpp == bpp && c.fingerprint == fp) return c.dev;
    if (c.dev) hana_texture_destroy(c.dev); /* the image changed under the same TGAImage*: upload it again */
    c.dev = nullptr;
    CK(hana_texture_upload(g.ctx, img->buffer(), w, h, bpp, &c.dev));
    c.buffer = img->buffer();
    c.w = w;
    c.h = h;
    c.bpp = bpp;
    c.fingerprint = fp;
    return c.dev;
}

hana_rb* rb_of(hana_rb*& slot, int w, int h) {
    int cw = 0, ch = 0;
    if (slot) hana_rb_size(slot, &cw, &ch);
    if (!slot || cw != w || ch != h) {
        if (slot) hana_rb_destroy(slot);
        slot = nullptr;
        CK(hana_rb_create(g.ctx, w, h, &slot));
    }
    return slot;
}

}  // namespace

/* Drops every cached device texture and model (they are uploaded again at the next draw). For hosts that edit texels or
 * vertices in place between draws in ways the cheap change detection above cannot see. */
extern "C" void hana_dropin_invalidate(void) {
    for (auto& kv : g.textures)
        if (kv.second.dev) hana_texture_destroy(kv.second.dev);
    g.textures.clear();
    for (auto& kv : g.models)
        if (kv.second.dev) hana_model_destroy(kv.second.dev);
    g.models.clear();
}

void graphics_draw_triangle(DrawData* draw_data) {
    if (!g.ctx) {
        const char* dev = std::getenv("HANA_DEVICE");
        CK(hana_ctx_create(dev ? std::atoi(dev) : 0, &g.ctx));
    }
    Model* model = draw_data->model;
    IShader* shader = draw_data->shader;
    RenderBuffer* rb = draw_data->render_buffer;
    ShaderData* sd = shader->shader_data;
    const int shader_id = shader_id_of(shader);
    if (shader_id < 0) {
        std::fprintf(stderr, "hana_b200 graphics_draw_triangle: IShader subclass %s is outside the device shader set "
                             "(user-defined vertex()/fragment() cannot run on the GPU; there is no CPU fallback)\n",
                     typeid(*shader).name());
        std::abort();
    }
    /* the a2v stream, gathered exactly as graphics.cpp:380-386 does (Model::normal re-normalises in place) */
    const int nfaces = model->nfaces();
    g.a2v.resize((size_t)nfaces * 3 * 8);
    for (int i = 0; i < nfaces; i++)
        for (int j = 0; j < 3; j++) {
            Vector3f p = model->vert(i, j);
            Vector3f n = model->normal(i, j);
            Vector2f t = model->uv(i, j);
            float* d = g.a2v.data() + ((size_t)i * 3 + j) * 8;
            d[0] = p.x; d[1] = p.y; d[2] = p.z; d[3] = n.x; d[4] = n.y; d[5] = n.z; d[6] = t.x; d[7] = t.y;
        }
    /* The gather above must run on every pass (the reference's normals change under it until they reach their fixed
     * point); the upload only when the stream differs from what the device holds. */
    CachedModel& cm = g.models[model];
    if (cm.dev && hana_model_ncorners(cm.dev) != nfaces * 3) {
        hana_model_destroy(cm.dev);
        cm.dev = nullptr;
    }
    if (!cm.dev) {
        CK(hana_model_upload(g.ctx, g.a2v.data(), nfaces * 3, &cm.dev));
        cm.a2v = g.a2v;
    } else if (cm.a2v.size() != g.a2v.size() || std::memcmp(cm.a2v.data(), g.a2v.data(), g.a2v.size() * sizeof(float)) != 0) {
        CK(hana_model_update(cm.dev, g.a2v.data(), nfaces * 3));
        cm.a2v = g.a2v;
    }
    hana_model* dm = cm.dev;

    HanaUniforms u;
    std::memset(&u, 0, sizeof(u));
    copy_matrix(u.model, sd->model_matrix);
    copy_matrix(u.model_I, sd->model_matrix_I);
    copy_matrix(u.camera_vp, sd->camera_vp_matrix);
    copy_matrix(u.light_vp, sd->light_vp_matrix);
    u.view_pos[0] = sd->view_Pos.x; u.view_pos[1] = sd->view_Pos.y; u.view_pos[2] = sd->view_Pos.z;
    u.light_dir[0] = sd->light_dir.x; u.light_dir[1] = sd->light_dir.y; u.light_dir[2] = sd->light_dir.z;
    for (int i = 0; i < 4; i++) {
        u.light_color[i] = sd->light_color[i];
        u.ambient[i] = sd->ambient[i];
    }
    hana_texture *diffuse = nullptr, *normal = nullptr;
    if (sd->matrial) {
        u.gloss = sd->matrial->gloss;
        u.bump_scale = sd->matrial->bump_scale;
        for (int i = 0; i < 4; i++) {
            u.mat_color[i] = sd->matrial->color[i];
            u.mat_specular[i] = sd->matrial->specular[i];
        }
        diffuse = texture_of(sd->matrial->diffuse_map);
        normal = texture_of(sd->matrial->normal_map);
    }
    u.enable_shadow = sd->enable_shadow ? 1 : 0;

    hana_rb* target = rb_of(g.target, rb->width, rb->height);
    CK(hana_rb_upload(target, rb->color_buffer, rb->depth_buffer));
    hana_rb* shadow = nullptr;
    if (shader_id != HANA_SHADER_SHADOW && sd->enable_shadow && sd->shadow_map) { /* IShader.h:109 */
        shadow = rb_of(g.shadow, sd->shadow_map->width, sd->shadow_map->height);
        CK(hana_rb_upload(shadow, sd->shadow_map->color_buffer, nullptr)); /* only the colour plane is read: IShader.h:124 */
    }
    CK(hana_draw(g.ctx, target, dm, shader_id, &u, diffuse, normal, shadow));
    CK(hana_rb_download(target, rb->color_buffer, rb->depth_buffer));
}
