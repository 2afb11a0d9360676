// emu_tga.cpp — TEST INFRASTRUCTURE. Compiles hana_tga_core.cuh (the __host__ __device__ arithmetic of the device-side
// RLE TGA packetiser) with the host compiler and walks it with the control flow of the kernels in hana_tga.cuh
// (equal-neighbour bits -> per-span structure passes with scans between them -> per-word byte counts -> prefix sum ->
// per-pixel writes), so the packetiser can be checked against the sequential algorithm without a GPU.
// Never linked into, or called by, the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../hana-softwarerenderer_b200/csrc/hana_tga_core.cuh"

using namespace hana;

extern "C" {

// px: n pixels in file order, value = B | G << 8 | R << 16. threads = virtual threads of the structure kernel.
// out must hold 4 * n + 64 bytes. Returns the payload length.
size_t emu_tga_payload(const uint32_t* px, int n, int threads, uint8_t* out) {
    const int nw = (n + 31) / 32, nbatch = (nw + 31) / 32;
    std::vector<uint32_t> E((size_t)nbatch * 32 + TGA_E_PAD, 0u);
    for (int i = 0; i + 1 < n; i++)
        if ((px[i] & 0xFFFFFFu) == (px[i + 1] & 0xFFFFFFu)) E[i >> 5] |= 1u << (i & 31);
    // ---- tga_structure_kernel
    const int span = ((nw + threads - 1) / threads + 3) & ~3; // a multiple of 4 words, as in the kernel
    std::vector<int> la(threads), lt(threads), a_in(threads), t_in(threads), x_in(threads);
    std::vector<unsigned> F(threads);
    auto W0 = [&](int t) { return std::min(t * span, nw); };
    auto W1 = [&](int t) { return std::min(W0(t) + span, nw); };
    for (int t = 0; t < threads; t++) tga_span_last(E.data(), W0(t), W1(t), la[t], lt[t]);
    int ca = -1, ct = -1;
    for (int t = 0; t < threads; t++) {
        a_in[t] = ca;
        t_in[t] = ct;
        ca = std::max(ca, la[t]);
        ct = std::max(ct, lt[t]);
    }
    for (int t = 0; t < threads; t++) F[t] = tga_span_function(E.data(), W0(t), W1(t), a_in[t], t_in[t]);
    unsigned P = 2u;
    for (int t = 0; t < threads; t++) {
        x_in[t] = (int)(P & 1u);
        P = tga_compose(P, F[t]);
    }
    std::vector<TgaRec> R(std::max(nw, 1));
    for (int t = 0; t < threads; t++) tga_span_records(E.data(), W0(t), W1(t), a_in[t], t_in[t], x_in[t], R.data());
    // ---- tga_count_kernel
    std::vector<unsigned> cnt((size_t)nbatch * 32, 0u);
    std::vector<int> cls((size_t)nbatch * 32, 0);
    for (int w = 0; w < nw; w++) {
        const uint32_t cur = E[w], eprev = w > 0 ? E[w - 1] >> 31 : 0u, enext = E[w + 1] & 1u;
        cls[w] = tga_word_class(w, nw, n, cur, eprev, enext);
        const TgaMasks m = tga_masks(cur, eprev, enext);
        cnt[w] = tga_word_bytes(R[w], m, w, n); // the count kernel: closed form, whatever the class
        unsigned c = 0;
        for (int j = 0; j < 32; j++) c += tga_role_bytes(tga_lane_role(R[w], m, w, j, n));
        if (c != cnt[w]) return (size_t)-1 - (size_t)w; // closed form and per-pixel roles disagree
        if (cls[w] != TGA_W_GENERAL && tga_closed_word_bytes(cls[w], R[w], w) != c) return (size_t)-1 - (size_t)w;
        const TgaWordEmit we = tga_word_emit(R[w], m, w, n); // the write kernel's form of the role, pixel by pixel
        for (int j = 0; j < 32; j++)
            if (tga_lane_emit(m, R[w].xm, we, w, j, n) != tga_lane_role(R[w], m, w, j, n)) return (size_t)-1 - (size_t)w;
    }
    // ---- scan + tga_write_kernel
    size_t off = 0;
    for (int w = 0; w < nw; w++) {
        const uint32_t cur = E[w], eprev = w > 0 ? E[w - 1] >> 31 : 0u, enext = E[w + 1] & 1u;
        uint8_t* o = out + off;
        if (cls[w] == TGA_W_RUN) {
            const int j = tga_run_word_end(R[w], w);
            if (j < 32) {
                const uint32_t v = px[w * 32 + j];
                o[0] = 255;
                o[1] = (uint8_t)v;
                o[2] = (uint8_t)(v >> 8);
                o[3] = (uint8_t)(v >> 16);
            }
        } else {
            const TgaMasks m = tga_masks(cur, eprev, enext);
            size_t pos = 0;
            for (int j = 0; j < 32; j++) {
                unsigned inf;
                if (cls[w] == TGA_W_RAW) { // the kernel's shortcut
                    const int k = (tga_raw_word_idx0(R[w], w) + j) & 127;
                    inf = 2u | ((unsigned)k << 2) | ((k == 127 || w * 32 + j == n - 1) ? 1u << 9 : 0u);
                } else {
                    inf = tga_lane_role(R[w], m, w, j, n);
                }
                const unsigned role = inf & 3u, k = (inf >> 2) & 127u;
                if (!role) continue;
                const uint32_t v = px[w * 32 + j];
                if (role == 1u) {
                    o[pos] = (uint8_t)(k + 128u);
                    o[pos + 1] = (uint8_t)v;
                    o[pos + 2] = (uint8_t)(v >> 8);
                    o[pos + 3] = (uint8_t)(v >> 16);
                    pos += 4;
                } else {
                    const size_t cpos = pos + (k == 0u ? 1 : 0);
                    o[cpos] = (uint8_t)v;
                    o[cpos + 1] = (uint8_t)(v >> 8);
                    o[cpos + 2] = (uint8_t)(v >> 16);
                    if (inf & (1u << 9)) *(o + cpos - 3 * (size_t)k - 1) = (uint8_t)k;
                    pos = cpos + 3;
                }
            }
            if (pos != cnt[w]) return (size_t)-1 - (size_t)w; // the count kernel and the write kernel disagree
        }
        off += cnt[w];
    }
    return off;
}
}
