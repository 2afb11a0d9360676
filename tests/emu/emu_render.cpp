// emu_render.cpp — TEST INFRASTRUCTURE. Compiles hana_core.cuh (the __host__ __device__
// arithmetic the CUDA kernels are made of) with the host compiler and walks it with the
// same control flow as the kernels (setup records -> per-pixel coverage -> min-depth /
// max-key resolve -> shade the winner), so that the device arithmetic and the
// order-independent resolve can be checked against the CPU oracle without a GPU.
// It is never linked into, or called by, the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../hana-softwarerenderer_b200/csrc/hana_core.cuh"

using namespace hana;

struct EmuTri {
    TriRecord r;
    float attr[24];
};

template <int SHADER>
static void shade(const DevUniforms& u, const EmuTri& t, float w0, float w1, float w2, const DevTexture& dt,
                  const DevTexture& nt, const DevShadow& sm, float rgb[3]) {
    VaryingWeights vw = varying_weights(w0, w1, w2, t.r.rw0, t.r.rw1, t.r.rw2);
    float attr[8];
    int na = shader_nattr(SHADER);
    for (int k = 0; k < na; k++) attr[k] = interp(vw, t.attr[3 * k], t.attr[3 * k + 1], t.attr[3 * k + 2]);
    fragment_shader<SHADER>(u.frag, attr, dt, nt, sm, rgb);
}

extern "C" {

void emu_prepare(const HanaUniforms* u, DevUniforms* d) { prepare_uniforms(*u, *d); }

void emu_vertex(int shader, const HanaUniforms* u, const float* a2v8, float* v2f13) {
    DevUniforms d;
    prepare_uniforms(*u, d);
    vertex_shader(shader, d, a2v8, v2f13);
}

int emu_coverage(const float* abc6, int W, int H, int px, int py, float* w3) {
    // builds a record from screen coords directly (no bbox): returns inside flag
    float s0x = abc6[4] - abc6[0], s0y = abc6[2] - abc6[0];
    float s1x = abc6[5] - abc6[1], s1y = abc6[3] - abc6[1];
    float uz = s0x * s1y - s0y * s1x;
    if (!(fabsf(uz) > 0.01f)) return 0;
    float ux, uy, s;
    bool in = coverage_test(abc6[0], abc6[1], s0x, s0y, s1x, s1y, uz, fabsf(uz) * 5.9604644775390625e-08f, (float)px, (float)py,
                            ux, uy, s);
    barycentric_weights(ux, uy, s, uz, 1.f / uz, w3[0], w3[1], w3[2]);
    return in ? 1 : 0;
}

// Binning's conservative triangle/tile test against exhaustive coverage of the tile's 256 pixels: returns
// (tile_may_touch ? 1 : 0) | (any pixel of the tile inside ? 2 : 0). 2 without 1 would be a dropped fragment.
int emu_tile_touch(const float* abc6, int X0, int Y0) {
    float s0x = abc6[4] - abc6[0], s0y = abc6[2] - abc6[0];
    float s1x = abc6[5] - abc6[1], s1y = abc6[3] - abc6[1];
    float uz = s0x * s1y - s0y * s1x;
    if (!(fabsf(uz) > 0.01f)) return 0;
    float thr = fabsf(uz) * 5.9604644775390625e-08f;
    int any = 0;
    for (int y = Y0; y < Y0 + 16 && !any; y++)
        for (int x = X0; x < X0 + 16; x++) {
            float ux, uy, s;
            if (coverage_test(abc6[0], abc6[1], s0x, s0y, s1x, s1y, uz, thr, (float)x, (float)y, ux, uy, s)) {
                any = 1;
                break;
            }
        }
    return (tile_may_touch(abc6[0], abc6[1], s0x, s0y, s1x, s1y, uz, thr, (float)X0, (float)Y0) ? 1 : 0) | (any ? 2 : 0);
}

// graphics_draw_triangle semantics over flat arrays, kernel control flow.
// tex: u32 B|G<<8|R<<16|A<<24 texels (as hana_texture_upload packs them).
void emu_draw(int shader, const HanaUniforms* hu, const float* a2v, int ncorners, const uint32_t* dtex, int dw, int dh,
              const uint32_t* ntex, int nw, int nh, const uint8_t* shadow, int sw, int sh, int spitch, int sstride, int W,
              int H, uint8_t* color, float* depth, uint32_t* primid, uint64_t* counts /* tris, covered */) {
    DevUniforms u;
    prepare_uniforms(*hu, u);
    DevTexture dt{dtex, dw, dh}, nt{ntex, nw, nh};
    DevShadow sm{shadow, sw, sh, spitch, sstride};
    std::vector<EmuTri> tris;
    for (int f = 0; f < ncorners / 3; f++) {
        float v[10 * V2F_N];
        for (int j = 0; j < 3; j++) vertex_shader(shader, u, a2v + (size_t)(f * 3 + j) * 8, v + V2F_N * j);
        int n = 3;
        if (!clip_trivial_accept(v)) n = clip_polygon(v);
        for (int j = 0; j + 2 < n; j++) {
            const float* a = v;
            const float* b = v + V2F_N * (j + 1);
            const float* c = v + V2F_N * (j + 2);
            EmuTri t;
            if (!triangle_setup(a, b, c, W, H, (uint32_t)f * 8u + (uint32_t)j, t.r)) continue;
            int na = shader_nattr(shader);
            for (int k = 0; k < na; k++) {
                int src = shader_attr_src(shader, k);
                t.attr[3 * k] = a[src];
                t.attr[3 * k + 1] = b[src];
                t.attr[3 * k + 2] = c[src];
            }
            tris.push_back(t);
        }
    }
    if (counts) counts[0] = tris.size();
    size_t n = (size_t)W * H;
    std::vector<int> bkey(n, -1);
    std::vector<uint32_t> bidx(n, 0);
    std::vector<float> bw(n * 3, 0.f);
    // process triangles in REVERSE order to prove the resolve is order-free
    for (size_t ti = tris.size(); ti-- > 0;) {
        const TriRecord& r = tris[ti].r;
        int x0 = r.bbx & 0xFFFF, x1 = r.bbx >> 16, y0 = r.bby & 0xFFFF, y1 = r.bby >> 16;
        for (int py = y0; py <= y1; py++)
            for (int px = x0; px <= x1; px++) {
                float ux, uy, s;
                if (!coverage_test(r.ax, r.ay, r.s0x, r.s0y, r.s1x, r.s1y, r.uz, r.thr, (float)px, (float)py, ux, uy, s)) continue;
                float w0, w1, w2;
                barycentric_weights(ux, uy, s, r.uz, r.ruz, w0, w1, w2);
                float z = interpolate_depth(r.d0, r.d1, r.d2, w0, w1, w2);
                size_t i = (size_t)py * W + px;
                int key = (int)r.key;
                bool win = (bkey[i] < 0) ? !(z > depth[i]) : (z < depth[i] || (z == depth[i] && key > bkey[i]));
                if (win) {
                    depth[i] = z;
                    bkey[i] = key;
                    bidx[i] = (uint32_t)ti;
                    bw[i * 3] = w0; bw[i * 3 + 1] = w1; bw[i * 3 + 2] = w2;
                }
            }
    }
    uint64_t covered = 0;
    for (size_t i = 0; i < n; i++) {
        if (bkey[i] < 0) continue;
        covered++;
        const EmuTri& t = tris[bidx[i]];
        float rgb[3];
        switch (shader) {
            case HANA_SHADER_SHADOW: shade<HANA_SHADER_SHADOW>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            case HANA_SHADER_BLINN: shade<HANA_SHADER_BLINN>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            case HANA_SHADER_NORMALMAP: shade<HANA_SHADER_NORMALMAP>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            case HANA_SHADER_GROUND: shade<HANA_SHADER_GROUND>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            case HANA_SHADER_TOON: shade<HANA_SHADER_TOON>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            case HANA_SHADER_TEXTURE: shade<HANA_SHADER_TEXTURE>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
            default: shade<HANA_SHADER_TEXTURE_LIGHT>(u, t, bw[i*3], bw[i*3+1], bw[i*3+2], dt, nt, sm, rgb); break;
        }
        uint32_t c = colour_bytes(rgb);
        color[i * 4] = c & 255; color[i * 4 + 1] = (c >> 8) & 255; color[i * 4 + 2] = (c >> 16) & 255;
        if (primid) primid[i] = (uint32_t)bkey[i];
    }
    if (counts) counts[1] = covered;
}
}
