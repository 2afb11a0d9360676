/*
 * hana_assets.cpp — the data formats either side of the rasterisation path, host side:
 *
 *   hana_obj_load   Wavefront OBJ -> the a2v corner stream graphics.cpp:380-386 gathers (model.cpp:6-48, 66-111)
 *   hana_tga_load   TGA (raw / RLE, 8/24/32 bpp) -> the byte image Model holds (tgaimage.cpp:40-143, model.cpp:74-84)
 *   hana_tga_write  byte image -> TGA file, raw or RLE, byte-identical to TGAImage::write_tga_file (tgaimage.cpp:145-246)
 *
 * Written from the formats and the reference's observable behaviour (which lines it accepts, what it does with
 * malformed faces, when it flips), not from its code: one pass over a memory-mapped-style buffer with strtof/strtol
 * instead of iostreams (the reference spends seconds in istringstream on a 10 M-triangle OBJ; this parser is bounded
 * by the file read). Citations are relative to /root/reference/Hana-SoftwareRenderer/.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hana_b200.h"

extern "C" int hana_set_error(int code, const char* msg); /* hana_b200.cu */

namespace {

bool read_file(const char* path, std::vector<char>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) {
        fclose(f);
        return false;
    }
    out.resize((size_t)n + 1);
    size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    out[got] = 0;
    out.resize(got + 1);
    return true;
}

/* `iss >> float` semantics for the plain decimal numbers OBJ files hold: skip blanks, parse, stop at the first
 * character that is not part of the number. A field that does not parse leaves the value at 0 (the stream fails and
 * the reference's default-constructed component stays 0). */
const char* parse_float(const char* p, const char* end, float* v) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    if (p >= end) return nullptr; /* strtof would skip the newline and read the next line */
    char* q = nullptr;
    float x = strtof(p, &q);
    if (q == p || q > end) return nullptr;
    *v = x;
    return q;
}
const char* parse_int(const char* p, const char* end, int* v) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    if (p >= end) return nullptr;
    char* q = nullptr;
    long x = strtol(p, &q, 10);
    if (q == p || q > end) return nullptr;
    *v = (int)x;
    return q;
}

/* `iss >> char`: skips blanks, then consumes one character */
const char* skip_sep(const char* p, const char* end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    return p < end ? p + 1 : end;
}

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};

/* Vector3f::normalize vector.h:41-42: v * (1 / sqrt(x*x + y*y + z*z)), float throughout, left to right */
void normalize_in_place(V3& v) {
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    float s = 1.f / len;
    v.x = v.x * s;
    v.y = v.y * s;
    v.z = v.z * s;
}

#pragma pack(push, 1)
struct TgaHeader { /* the 18-byte TGA file header */
    uint8_t idlength, colormaptype, datatypecode;
    uint16_t colormaporigin, colormaplength;
    uint8_t colormapdepth;
    uint16_t x_origin, y_origin, width, height;
    uint8_t bitsperpixel, imagedescriptor;
};
#pragma pack(pop)
static_assert(sizeof(TgaHeader) == 18, "TGA header is 18 bytes");

void flip_rows(uint8_t* d, int w, int h, int bpp) {
    const size_t line = (size_t)w * bpp;
    std::vector<uint8_t> tmp(line);
    for (int j = 0; j < h / 2; j++) {
        uint8_t* a = d + (size_t)j * line;
        uint8_t* b = d + (size_t)(h - 1 - j) * line;
        memcpy(tmp.data(), a, line);
        memcpy(a, b, line);
        memcpy(b, tmp.data(), line);
    }
}
void flip_columns(uint8_t* d, int w, int h, int bpp) {
    uint8_t t[4];
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w / 2; i++) {
            uint8_t* a = d + ((size_t)j * w + i) * bpp;
            uint8_t* b = d + ((size_t)j * w + (w - 1 - i)) * bpp;
            memcpy(t, a, bpp);
            memcpy(a, b, bpp);
            memcpy(b, t, bpp);
        }
}

}  // namespace

/* OBJ -> a2v. Lines "v ", "vn ", "vt ", "f " (model.cpp:15-41); a face is a run of v/vt/vn index triples, 1-based;
 * the draw path reads corners 0..2 of every face (graphics.cpp:381). `normal_pass`: Model::normal() re-normalises the
 * stored normal in place on every access (model.cpp:108-111), so the normal a corner sees depends on
 * how many times its index was visited before; the stream returned is the one pass number `normal_pass` (1-based) of
 * the draw loop sees, i.e. what the reference feeds its shaders in that pass of its life. */
extern "C" int hana_obj_load(const char* path, int normal_pass, float** out_a2v, int* out_ncorners) {
    if (!path || !out_a2v || !out_ncorners) return hana_set_error(HANA_E_INVALID, "NULL argument");
    if (normal_pass < 1) normal_pass = 1;
    *out_a2v = nullptr;
    *out_ncorners = 0;
    std::vector<char> buf;
    if (!read_file(path, buf)) return hana_set_error(HANA_E_INVALID, (std::string("cannot open ") + path).c_str());
    std::vector<V3> verts, norms;
    std::vector<V2> uvs;
    std::vector<int> faces; /* 9 ints per face: (v, vt, vn) x 3 corners; -1 where the face had fewer than 3 */
    const char* p = buf.data();
    const char* end = p + buf.size() - 1;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol - p >= 2 && p[0] == 'v' && p[1] == ' ') {
            V3 v = {0.f, 0.f, 0.f};
            const char* q = p + 1;
            float* c[3] = {&v.x, &v.y, &v.z};
            for (int i = 0; i < 3 && q; i++) q = parse_float(q, eol, c[i]);
            verts.push_back(v);
        } else if (eol - p >= 3 && p[0] == 'v' && p[1] == 'n' && p[2] == ' ') {
            V3 v = {0.f, 0.f, 0.f};
            const char* q = p + 2;
            float* c[3] = {&v.x, &v.y, &v.z};
            for (int i = 0; i < 3 && q; i++) q = parse_float(q, eol, c[i]);
            norms.push_back(v);
        } else if (eol - p >= 3 && p[0] == 'v' && p[1] == 't' && p[2] == ' ') {
            V2 v = {0.f, 0.f};
            const char* q = p + 2;
            q = parse_float(q, eol, &v.x);
            if (q) parse_float(q, eol, &v.y);
            uvs.push_back(v);
        } else if (eol - p >= 2 && p[0] == 'f' && p[1] == ' ') {
            int f[9];
            for (int& x : f) x = -1;
            int n = 0;
            const char* q = p + 1;
            while (q && q < eol) { /* "a/b/c" triples until one fails to parse (model.cpp:36) */
                int t[3];
                const char* r = parse_int(q, eol, &t[0]);
                if (!r || r >= eol) break;
                r = parse_int(skip_sep(r, eol), eol, &t[1]); /* one separator character, whatever it is */
                if (!r || r >= eol) break;
                r = parse_int(skip_sep(r, eol), eol, &t[2]);
                if (!r) break;
                if (n < 3)
                    for (int k = 0; k < 3; k++) f[3 * n + k] = t[k] - 1;
                n++;
                q = r;
            }
            faces.insert(faces.end(), f, f + 9);
        }
        p = eol + 1;
    }
    const size_t nfaces = faces.size() / 9;
    float* a2v = (float*)malloc(std::max<size_t>(nfaces * 3 * 8, 1) * sizeof(float));
    if (!a2v) return hana_set_error(HANA_E_INVALID, "out of host memory");
    auto in_range = [](int i, size_t n) { return i >= 0 && (size_t)i < n; };
    for (size_t i = 0; i < nfaces * 9; i++) {
        const int kind = (int)(i % 3);
        const size_t lim = kind == 0 ? verts.size() : kind == 1 ? uvs.size() : norms.size();
        if (!in_range(faces[i], lim)) {
            free(a2v);
            return hana_set_error(HANA_E_INVALID, "OBJ face with fewer than 3 corners or an index out of range "
                                                  "(the reference reads out of bounds here)");
        }
    }
    for (int pass = 1; pass <= normal_pass; pass++)
        for (size_t i = 0; i < nfaces; i++)
            for (int j = 0; j < 3; j++) {
                const int* f = &faces[i * 9 + j * 3];
                V3& n = norms[(size_t)f[2]];
                normalize_in_place(n);
                if (pass == normal_pass) {
                    float* d = a2v + (i * 3 + j) * 8;
                    const V3& v = verts[(size_t)f[0]];
                    const V2& t = uvs[(size_t)f[1]];
                    d[0] = v.x; d[1] = v.y; d[2] = v.z;
                    d[3] = n.x; d[4] = n.y; d[5] = n.z;
                    d[6] = t.x; d[7] = t.y;
                }
            }
    *out_a2v = a2v;
    *out_ncorners = (int)(nfaces * 3);
    return HANA_OK;
}

/* TGA -> bytes as TGAImage holds them after read_tga_file (rows top-down: bottom-origin files are flipped on read,
 * tgaimage.cpp:85-90), and after Model::load_texture's extra flip_vertically (model.cpp:81) when model_flip != 0,
 * which is the orientation tex_diffuse/tex_normal index (v up) and hana_texture_upload expects. */
extern "C" int hana_tga_load(const char* path, int model_flip, uint8_t** out_data, int* out_w, int* out_h, int* out_bpp) {
    if (!path || !out_data || !out_w || !out_h || !out_bpp) return hana_set_error(HANA_E_INVALID, "NULL argument");
    *out_data = nullptr;
    std::vector<char> buf;
    if (!read_file(path, buf)) return hana_set_error(HANA_E_INVALID, (std::string("cannot open ") + path).c_str());
    const size_t size = buf.size() - 1;
    if (size < sizeof(TgaHeader)) return hana_set_error(HANA_E_INVALID, "TGA: truncated header");
    TgaHeader h;
    memcpy(&h, buf.data(), sizeof(h));
    const int w = (int16_t)h.width, ht = (int16_t)h.height, bpp = h.bitsperpixel >> 3;
    if (w <= 0 || ht <= 0 || (bpp != 1 && bpp != 3 && bpp != 4)) return hana_set_error(HANA_E_INVALID, "TGA: bad size or depth");
    const size_t nbytes = (size_t)w * ht * bpp;
    uint8_t* d = (uint8_t*)malloc(nbytes);
    if (!d) return hana_set_error(HANA_E_INVALID, "out of host memory");
    /* the reference does not skip the id field / colour map either (it reads pixels right after the header) */
    const uint8_t* s = (const uint8_t*)buf.data() + sizeof(TgaHeader);
    const uint8_t* send = (const uint8_t*)buf.data() + size;
    if (h.datatypecode == 2 || h.datatypecode == 3) {
        if ((size_t)(send - s) < nbytes) {
            free(d);
            return hana_set_error(HANA_E_INVALID, "TGA: truncated pixel data");
        }
        memcpy(d, s, nbytes);
    } else if (h.datatypecode == 10 || h.datatypecode == 11) {
        size_t px = 0;
        const size_t npx = (size_t)w * ht;
        while (px < npx) {
            if (s >= send) {
                free(d);
                return hana_set_error(HANA_E_INVALID, "TGA: truncated RLE stream");
            }
            int c = *s++;
            if (c < 128) { /* raw packet of c+1 pixels */
                size_t n = (size_t)c + 1;
                if (px + n > npx || (size_t)(send - s) < n * bpp) {
                    free(d);
                    return hana_set_error(HANA_E_INVALID, "TGA: bad raw packet");
                }
                memcpy(d + px * bpp, s, n * bpp);
                s += n * bpp;
                px += n;
            } else { /* run packet of c-127 copies */
                size_t n = (size_t)c - 127;
                if (px + n > npx || (size_t)(send - s) < (size_t)bpp) {
                    free(d);
                    return hana_set_error(HANA_E_INVALID, "TGA: bad run packet");
                }
                for (size_t k = 0; k < n; k++) memcpy(d + (px + k) * bpp, s, bpp);
                s += bpp;
                px += n;
            }
        }
    } else {
        free(d);
        return hana_set_error(HANA_E_INVALID, "TGA: unsupported image type");
    }
    bool flipped = false;
    if (!(h.imagedescriptor & 0x20)) flipped = !flipped; /* bottom-origin file -> top-down rows */
    if (model_flip) flipped = !flipped;                  /* Model::load_texture flips again */
    if (flipped) flip_rows(d, w, ht, bpp);
    if (h.imagedescriptor & 0x10) flip_columns(d, w, ht, bpp);
    *out_data = d;
    *out_w = w;
    *out_h = ht;
    *out_bpp = bpp;
    return HANA_OK;
}

/* bytes (rows in file order, top-left origin flagged: tgaimage.cpp:166) -> TGA file. RLE packets are formed exactly as
 * the reference forms them (runs of up to 128 equal pixels; a raw packet ends one pixel before a run starts), so the
 * files are byte-identical to TGAImage::write_tga_file's. */
extern "C" int hana_tga_write(const char* path, const uint8_t* data, int w, int h, int bpp, int rle) {
    if (!path || !data) return hana_set_error(HANA_E_INVALID, "NULL argument");
    if (w <= 0 || h <= 0 || w > 32767 || h > 32767 || (bpp != 1 && bpp != 3 && bpp != 4))
        return hana_set_error(HANA_E_INVALID, "TGA: bad size or depth");
    FILE* f = fopen(path, "wb");
    if (!f) return hana_set_error(HANA_E_INVALID, (std::string("cannot create ") + path).c_str());
    TgaHeader hd;
    memset(&hd, 0, sizeof(hd));
    hd.bitsperpixel = (uint8_t)(bpp << 3);
    hd.width = (uint16_t)w;
    hd.height = (uint16_t)h;
    hd.datatypecode = (uint8_t)(bpp == 1 ? (rle ? 11 : 3) : (rle ? 10 : 2));
    hd.imagedescriptor = 0x20;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    const size_t npx = (size_t)w * h;
    if (ok && !rle) {
        ok = fwrite(data, 1, npx * bpp, f) == npx * bpp;
    } else if (ok) {
        std::vector<uint8_t> out;
        out.reserve(npx * bpp / 2 + 1024);
        size_t cur = 0;
        while (cur < npx) {
            /* length of the packet starting at cur: either a run of equal pixels or a raw stretch that stops
             * right before the next pair of equal pixels; at most 128 pixels either way */
            const uint8_t* px0 = data + cur * bpp;
            size_t len = 1;
            bool raw = true;
            while (cur + len < npx && len < 128) {
                const bool eq = memcmp(px0 + (len - 1) * bpp, px0 + len * bpp, bpp) == 0;
                if (len == 1) raw = !eq;
                if (raw && eq) { /* the pair belongs to the next (run) packet */
                    len--;
                    break;
                }
                if (!raw && !eq) break;
                len++;
            }
            out.push_back((uint8_t)(raw ? len - 1 : len + 127));
            out.insert(out.end(), px0, px0 + (raw ? len * bpp : (size_t)bpp));
            cur += len;
        }
        ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    }
    static const uint8_t tail[26] = {0, 0, 0, 0, 0, 0, 0, 0, 'T', 'R', 'U', 'E', 'V', 'I', 'S', 'I', 'O',
                                     'N', '-', 'X', 'F', 'I', 'L', 'E', '.', 0};
    if (ok) ok = fwrite(tail, 1, sizeof(tail), f) == sizeof(tail);
    if (fclose(f) != 0) ok = false;
    if (!ok) return hana_set_error(HANA_E_INVALID, "TGA: write failed");
    return HANA_OK;
}

extern "C" void hana_free(void* p) { free(p); }
