/*
 * hana_tga.cuh — RLE-compressed TGA files made on the device (the output side of the path).
 *
 * Replaces TGAImage::write_tga_file(rle = true) + unload_rle_data (tgaimage.cpp:145-246) for the frames of a sweep: the
 * file bytes are produced in HBM, so a frame leaves the GPU as ~1.1 MB instead of 8.3 MB of raw colour (the bundled
 * scenes are ~80 % background), which is what the PCIe-bound end-to-end path needs. Output is BYTE-IDENTICAL to the
 * reference's writer (and to hana_tga_write): same packets, same header, same footer.
 *
 * The reference's packetiser is sequential and greedy. Its parallel formulation (hana_tga_core.cuh; derived and checked
 * against the sequential one in tests/rle_model.py and, function by function, in tests/emu/emu_tga.cpp) rests on one
 * observation: with e[i] = (pixel i == pixel i+1), everything a stretch of pixels (from one "e turns true" to the next)
 * inherits from the pixels before it is ONE bit x. So the only sequential part works on 1 bit per pixel:
 *
 *   tga_ebits_kernel      e[] packed 32 pixels to a word: the one pass over the frame's colour (4 B per pixel read)
 *   tga_structure_kernel  ONE CTA per frame walks the words (a thread per contiguous span, three passes with a block scan
 *                         between them) and leaves, per word, the state (last T-start, last tail, x) at its first pixel
 *   tga_count_kernel      bytes each word emits, a lane per word, in closed form (popcounts + the two positions mod 128 that
 *                         can fall into the stretch a word continues), and their sums per batch of 32 words
 *   tga_scan_kernel       one CTA per frame: offsets of the batches, file size
 *   tga_offsets_kernel    file sizes -> 16-byte aligned starts of the files in ONE buffer (one device-to-host copy)
 *   tga_write_kernel      a lane per pixel: role, position, bytes — written where the file lies, no staging copy
 *
 * (Round 2's first encoder chained 4096-pixel chunks of a frame through mailboxes: 18.7 us per 1080p frame in batches,
 * 0.29 ms alone; it paid 11 M warp instructions per frame for per-pixel structure work that 1 bit per pixel answers.)
 */
#ifndef HANA_TGA_CUH
#define HANA_TGA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "hana_tga_core.cuh"

namespace hana {

constexpr int TGA_E_THREADS = 256;  /* tga_ebits_kernel: 8 warps, 32 words (1024 pixels) each */
constexpr int TGA_B_THREADS = 1024; /* tga_structure_kernel */
constexpr int TGA_C_THREADS = 256;  /* count / write: 8 warps, a batch of 32 words each */
constexpr int TGA_S_THREADS = 1024; /* tga_scan_kernel */

/* per-frame strides of the encoder's scratch, in elements */
struct TgaLayout {
    int n, nw, nbatch;     /* pixels, words (32 pixels), batches (32 words) */
    size_t estride;        /* e-bit words per frame: nbatch * 32 + TGA_E_PAD, zero beyond the frame */
    size_t rstride;        /* records per frame */
};
__host__ __device__ inline TgaLayout tga_layout(int W, int H) {
    TgaLayout L;
    L.n = W * H;
    L.nw = (L.n + 31) / 32;
    L.nbatch = (L.nw + 31) / 32;
    L.estride = (size_t)L.nbatch * 32 + TGA_E_PAD;
    L.rstride = (size_t)L.nbatch * 32;
    return L;
}
/* worst case of a frame's file: 4 bytes per pixel (a one-pixel raw packet between runs) + header + footer, 16-byte aligned */
__host__ __device__ inline size_t tga_worst_bytes(size_t npix) { return (npix * 4 + TGA_HEADER + TGA_FOOTER + 15) / 16 * 16; }

/* file-order pixel i (rows top-down: file row r is buffer row H-1-r, win32.cpp:358 / tgaimage.cpp:150) -> buffer index */
__device__ __forceinline__ size_t tga_src_index(int row, int col, int W, int H) { return (size_t)(H - 1 - row) * W + col; }

/* ---- e bits ---------------------------------------------------------------------------------------------------------- */
template <bool VEC>
__global__ void __launch_bounds__(TGA_E_THREADS) tga_ebits_kernel(const uint32_t* __restrict__ color, size_t frame_stride, int first, int W, int H,
                                                                   uint32_t* __restrict__ Eall, TgaLayout L) {
    constexpr unsigned FULL = 0xFFFFFFFFu, RGB = 0x00FFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.y;
    const uint32_t* src = color + (size_t)(first + f) * frame_stride;
    uint32_t* E = Eall + (size_t)f * L.estride;
    const int wbase = (blockIdx.x * (TGA_E_THREADS / 32) + wid) * 32; /* the warp's first word */
    if ((size_t)wbase >= L.estride) return;
    const int n = L.n;
    if (VEC) { /* W % 4 == 0 and W >= 128: a lane's four pixels lie in one row, 16-byte aligned; 128 pixels on cross at most one row end */
        /* the warp's 1024 pixels as 8 coalesced 128-bit loads per lane, all in flight together */
        const int base = wbase * 32;
        const bool inner = base + 1024 < n; /* every pixel and every successor exists: no checks */
        uint4 q[8];
        uint32_t x9 = 0; /* the pixel behind the warp's last */
        if (base < n) {
            const int row0 = base / W; /* one division per warp */
            int col = base - row0 * W + lane * 4;
            int idx = (H - 1 - row0) * W + col; /* buffer index of file pixel (row0, col): the next file row lies W in front */
            if (col >= W) col -= W, idx -= 2 * W;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                q[it] = (inner || base + it * 128 + lane * 4 < n) ? __ldg(reinterpret_cast<const uint4*>(src + idx)) : make_uint4(0, 0, 0, 0);
                col += 128, idx += 128;
                if (col >= W) col -= W, idx -= 2 * W;
            }
            if (inner && lane == 0) x9 = __ldg(src + idx); /* lane 0 has walked on to pixel base + 1024 */
        } else {
#pragma unroll
            for (int it = 0; it < 8; it++) q[it] = make_uint4(0, 0, 0, 0);
        }
        uint32_t V = 0; /* nibble `it` = the e bits of the lane's four pixels of load `it` */
#pragma unroll
        for (int it = 0; it < 8; it++) {
            /* lane l is read by lane l - 1 only, lane 0 by lane 31, which needs the first pixel of the NEXT load */
            const uint32_t send = lane == 0 ? (it < 7 ? q[it < 7 ? it + 1 : 7].x : x9) : q[it].x;
            const uint32_t nx = __shfl_sync(FULL, send, (lane + 1) & 31);
            uint32_t nib = ((((q[it].x ^ q[it].y) & RGB) == 0u) ? 1u : 0u) | ((((q[it].y ^ q[it].z) & RGB) == 0u) ? 2u : 0u) |
                           ((((q[it].z ^ q[it].w) & RGB) == 0u) ? 4u : 0u) | ((((q[it].w ^ nx) & RGB) == 0u) ? 8u : 0u);
            if (!inner) {
                const int i4 = base + it * 128 + lane * 4;
                if (i4 + 4 >= n) nib &= 7u; /* the last pixel of the frame has no successor */
                if (i4 >= n) nib = 0u;
            }
            V |= nib << (4 * it);
        }
        /* 8 x 8 nibble transpose inside each group of 8 lanes: lane l of a group ends up with load l's word of the group */
        {
            uint32_t P = __shfl_xor_sync(FULL, V, 4);
            V = (lane & 4) ? (((P >> 16) & 0x0000FFFFu) | (V & 0xFFFF0000u)) : ((V & 0x0000FFFFu) | ((P & 0x0000FFFFu) << 16));
            P = __shfl_xor_sync(FULL, V, 2);
            V = (lane & 2) ? (((P >> 8) & 0x00FF00FFu) | (V & 0xFF00FF00u)) : ((V & 0x00FF00FFu) | ((P & 0x00FF00FFu) << 8));
            P = __shfl_xor_sync(FULL, V, 1);
            V = (lane & 1) ? (((P >> 4) & 0x0F0F0F0Fu) | (V & 0xF0F0F0F0u)) : ((V & 0x0F0F0F0Fu) | ((P & 0x0F0F0F0Fu) << 4));
        }
        const int widx = wbase + (lane & 7) * 4 + (lane >> 3); /* the warp's 32 words: one coalesced store */
        if ((size_t)widx < L.estride) E[widx] = V;
    } else {
        uint32_t keep = 0;
        int i = wbase * 32 + lane;
        int row = i < n ? i / W : 0, col = i < n ? i - row * W : 0;
        for (int it = 0; it < 32; it++) {
            uint32_t p = 0, pn = 0;
            if (i < n) p = __ldg(src + tga_src_index(row, col, W, H)) & RGB;
            pn = __shfl_down_sync(FULL, p, 1);
            if (lane == 31 && i + 1 < n) {
                int r2 = row, c2 = col + 1;
                if (c2 >= W) {
                    c2 = 0;
                    r2++;
                }
                pn = __ldg(src + tga_src_index(r2, c2, W, H)) & RGB;
            }
            const uint32_t word = __ballot_sync(FULL, i + 1 < n && p == pn);
            if (lane == it) keep = word;
            i += 32;
            col += 32;
            while (col >= W) {
                col -= W;
                row++;
            }
        }
        if ((size_t)(wbase + lane) < L.estride) E[wbase + lane] = keep;
    }
}

/* ---- structure: one CTA per frame -------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(TGA_B_THREADS, 1) tga_structure_kernel(const uint32_t* __restrict__ Eall, TgaRec* __restrict__ Rall, TgaLayout L) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int NW = TGA_B_THREADS / 32;
    __shared__ int s_a[NW], s_t[NW];
    __shared__ unsigned s_f[NW];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* E = Eall + (size_t)blockIdx.x * L.estride;
    TgaRec* R = Rall + (size_t)blockIdx.x * L.rstride;
    const int nw = L.nw;
    const int span = ((nw + TGA_B_THREADS - 1) / TGA_B_THREADS + 3) & ~3; /* 16-byte aligned spans */
    const int w0 = min(tid * span, nw), w1 = min(w0 + span, nw);

    /* pass 0: last T-start / last tail in front of every span */
    int la, lt;
    tga_span_last(E, w0, w1, la, lt);
    int ia = la, it = lt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int oa = __shfl_up_sync(FULL, ia, d), ot = __shfl_up_sync(FULL, it, d);
        if (lane >= d) {
            ia = max(ia, oa);
            it = max(it, ot);
        }
    }
    if (lane == 31) {
        s_a[wid] = ia;
        s_t[wid] = it;
    }
    __syncthreads();
    int pa = __shfl_up_sync(FULL, ia, 1), pt = __shfl_up_sync(FULL, it, 1);
    if (lane == 0) pa = pt = -1;
    for (int w = 0; w < wid; w++) {
        pa = max(pa, s_a[w]);
        pt = max(pt, s_t[w]);
    }
    /* pass 1: x behind each span as a function of x in front of it, composed over the spans in front */
    const unsigned F = tga_span_function(E, w0, w1, pa, pt);
    unsigned P = F;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(FULL, P, d);
        if (lane >= d) P = tga_compose(o, P);
    }
    if (lane == 31) s_f[wid] = P;
    __syncthreads();
    unsigned Pex = __shfl_up_sync(FULL, P, 1);
    if (lane == 0) Pex = 2u;
    unsigned Pw = 2u;
    for (int w = 0; w < wid; w++) Pw = tga_compose(Pw, s_f[w]);
    Pex = tga_compose(Pw, Pex);
    /* pass 2: the records (a frame starts with x = 0) */
    tga_span_records(E, w0, w1, pa, pt, (int)(Pex & 1u), R);
}

/* ---- a batch of 32 words per warp: what a lane knows about its word ------------------------------------------------- */
struct TgaLaneWord {
    uint32_t cur, eprev, enext;
    TgaRec rec;
    int cls;
};
__device__ __forceinline__ TgaLaneWord tga_load_lane_word(const uint32_t* __restrict__ E, const TgaRec* __restrict__ R, int w, int lane, const TgaLayout& L) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    TgaLaneWord q;
    q.cur = __ldg(E + w); /* zero beyond the frame (estride covers the last batch) */
    uint32_t before = __shfl_up_sync(FULL, q.cur, 1), behind = __shfl_down_sync(FULL, q.cur, 1);
    if (lane == 0) before = w > 0 ? __ldg(E + w - 1) : 0u;
    if (lane == 31) behind = __ldg(E + w + 1);
    q.eprev = before >> 31;
    q.enext = behind & 1u;
    q.cls = tga_word_class(w, L.nw, L.n, q.cur, q.eprev, q.enext);
    if (q.cls != TGA_W_EMPTY) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(R + w));
        q.rec.a = (int)r.x, q.rec.t = (int)r.y, q.rec.xm = r.z, q.rec.pad = 0u;
    } else {
        q.rec.a = q.rec.t = -1, q.rec.xm = q.rec.pad = 0u;
    }
    return q;
}

/* ---- bytes per word (closed form, a lane per word), sums per batch ---------------------------------------------------- */
__global__ void __launch_bounds__(TGA_C_THREADS) tga_count_kernel(const uint32_t* __restrict__ Eall, const TgaRec* __restrict__ Rall, uint8_t* __restrict__ counts,
                                                                   uint32_t* __restrict__ batch_sums, TgaLayout L) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int batch = blockIdx.x * (TGA_C_THREADS / 32) + (threadIdx.x >> 5);
    if (batch >= L.nbatch) return;
    const int f = blockIdx.y;
    const uint32_t* E = Eall + (size_t)f * L.estride;
    const TgaRec* R = Rall + (size_t)f * L.rstride;
    const int w = batch * 32 + lane;
    const TgaLaneWord q = tga_load_lane_word(E, R, w, lane, L);
    unsigned cnt = 0;
    if (q.cls == TGA_W_RUN) cnt = tga_closed_word_bytes(TGA_W_RUN, q.rec, w); /* the background: nothing else to look at */
    else if (q.cls != TGA_W_EMPTY) cnt = tga_word_bytes(q.rec, tga_masks(q.cur, q.eprev, q.enext), w, L.n);
    counts[(size_t)f * L.rstride + w] = (uint8_t)cnt; /* <= 128: at most 4 bytes per pixel */
    unsigned sum = cnt;
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
    if (lane == 0) batch_sums[(size_t)f * L.nbatch + batch] = sum;
}

/* ---- offsets of the batches of a frame, file size: one CTA per frame -------------------------------------------------- */
__global__ void __launch_bounds__(TGA_S_THREADS) tga_scan_kernel(const uint32_t* __restrict__ batch_sums, unsigned long long* __restrict__ batch_offs,
                                                                  unsigned long long* __restrict__ sizes, int nbatch) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int NW = TGA_S_THREADS / 32;
    __shared__ unsigned long long s_w[NW];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* in = batch_sums + (size_t)blockIdx.x * nbatch;
    unsigned long long* out = batch_offs + (size_t)blockIdx.x * nbatch;
    if (tid == 0) s_carry = 0ull;
    __syncthreads();
    for (int base = 0; base < nbatch; base += TGA_S_THREADS) {
        const int k = base + tid;
        const unsigned long long v = k < nbatch ? (unsigned long long)in[k] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (int q = 0; q < wid; q++) pre += s_w[q];
        if (k < nbatch) out[k] = pre + inc - v;
        __syncthreads();
        if (tid == TGA_S_THREADS - 1) s_carry = pre + inc;
        __syncthreads();
    }
    if (tid == 0) sizes[blockIdx.x] = (unsigned long long)TGA_HEADER + s_carry + TGA_FOOTER;
}

/* sizes -> 16-byte aligned offsets of the files in the packed buffer; offsets[n_frames] = end of the last file */
__global__ void tga_offsets_kernel(const unsigned long long* __restrict__ sizes, unsigned long long* __restrict__ offsets, int n_frames) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    /* one warp, 32 files per round */
    const int lane = threadIdx.x;
    unsigned long long carry = 0;
    for (int base = 0; base < n_frames; base += 32) {
        const int f = base + lane;
        const unsigned long long v = f < n_frames ? (sizes[f] + 15ull) & ~15ull : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (f < n_frames) offsets[f] = carry + inc - v;
        carry += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) offsets[n_frames] = carry;
}

/* ---- the bytes ------------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(TGA_C_THREADS) tga_write_kernel(const uint32_t* __restrict__ color, size_t frame_stride, int first, int W, int H,
                                                                   const uint32_t* __restrict__ Eall, const TgaRec* __restrict__ Rall,
                                                                   const uint8_t* __restrict__ counts, const unsigned long long* __restrict__ batch_offs,
                                                                   const unsigned long long* __restrict__ sizes, const unsigned long long* __restrict__ offsets,
                                                                   uint8_t* __restrict__ packed, TgaLayout L) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int batch = blockIdx.x * (TGA_C_THREADS / 32) + (threadIdx.x >> 5);
    if (batch >= L.nbatch) return;
    const int f = blockIdx.y;
    const uint32_t* E = Eall + (size_t)f * L.estride;
    const TgaRec* R = Rall + (size_t)f * L.rstride;
    const uint32_t* src = color + (size_t)(first + f) * frame_stride;
    uint8_t* file = packed + offsets[f];
    const int n = L.n;
    const int w = batch * 32 + lane;
    const TgaLaneWord q = tga_load_lane_word(E, R, w, lane, L);
    const unsigned cnt = counts[(size_t)f * L.rstride + w];
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    /* payload offset of the lane's word */
    const unsigned long long woff = batch_offs[(size_t)f * L.nbatch + batch] + (inc - cnt);
    uint8_t* const payload = file + TGA_HEADER;

    if (batch == 0) { /* TGA_Header, tgaimage.cpp:157-163: type 10 (RLE true colour), 24 bits, top-left origin */
        if (lane < TGA_HEADER) {
            uint8_t hb = 0;
            if (lane == 2) hb = 10;
            if (lane == 12) hb = (uint8_t)(W & 255);
            if (lane == 13) hb = (uint8_t)(W >> 8);
            if (lane == 14) hb = (uint8_t)(H & 255);
            if (lane == 15) hb = (uint8_t)(H >> 8);
            if (lane == 16) hb = 24;
            if (lane == 17) hb = 0x20;
            file[lane] = hb;
        }
    }
    if (batch == L.nbatch - 1) { /* developer area ref, extension area ref, "TRUEVISION-XFILE.\0": tgaimage.cpp:146-148 */
        if (lane < TGA_FOOTER) {
            const char sig[18] = {'T', 'R', 'U', 'E', 'V', 'I', 'S', 'I', 'O', 'N', '-', 'X', 'F', 'I', 'L', 'E', '.', 0};
            file[sizes[f] - TGA_FOOTER + lane] = lane < 8 ? (uint8_t)0 : (uint8_t)sig[lane - 8];
        }
    }

    /* buffer index of the word's first pixel, its column */
    int col = 0, idx = 0;
    if (q.cls != TGA_W_EMPTY) {
        const int row = (w * 32) / W;
        col = w * 32 - row * W;
        idx = (H - 1 - row) * W + col;
    }
    /* a word in the middle of a run: at most one packet ends in it, [255, B, G, R]; all its pixels are equal. The pixel is
     * requested here and stored behind the loop over the other words, so nobody waits for it. */
    const bool run_end = q.cls == TGA_W_RUN && tga_run_word_end(q.rec, w) < 32;
    uint32_t run_px = 0;
    if (run_end) run_px = __ldg(src + idx);
    auto store_run_end = [&]() {
        if (run_end) {
            uint8_t* o = payload + woff;
            o[0] = 255;
            o[1] = (uint8_t)(run_px >> 16); /* the file wants B,G,R: the buffer's u32 is R | G << 8 | B << 16 */
            o[2] = (uint8_t)(run_px >> 8);
            o[3] = (uint8_t)run_px;
        }
    };
    /* every other word: a lane per pixel. What the lanes need of word g travels in seven shuffles: its e bits, its xm, the
     * two masks and two numbers of tga_word_emit, its byte offset + neighbour bits, buffer index / column of its first pixel */
    unsigned gm = __ballot_sync(FULL, q.cls == TGA_W_GENERAL || q.cls == TGA_W_RAW);
    if (!gm) {
        store_run_end();
        return;
    }
    const unsigned lt_mask = (1u << lane) - 1u;
    const TgaWordEmit we = tga_word_emit(q.rec, tga_masks(q.cur, q.eprev, q.enext), w, n);
    const uint32_t misc = (inc - cnt) | (q.eprev << 13) | (q.enext << 14) | (we.rc << 15) | (we.r0c << 22); /* offset in the batch <= 4096 */
    uint8_t* const bpay = payload + batch_offs[(size_t)f * L.nbatch + batch];
    /* the pixel of word g for this lane; the load of the NEXT word's pixels is in flight while a word is worked on */
    auto fetch = [&](int g) -> uint32_t {
        int c2 = __shfl_sync(FULL, col, g) + lane, i2 = __shfl_sync(FULL, idx, g) + lane;
        while (c2 >= W) c2 -= W, i2 -= 2 * W; /* into the next file row = the buffer row in front */
        return (batch * 32 + g) * 32 + lane < n ? __ldg(src + i2) : 0u;
    };
    int g = __ffs((int)gm) - 1;
    gm &= gm - 1;
    uint32_t px = fetch(g);
    while (true) {
        int g_next = -1;
        uint32_t px_next = 0;
        if (gm) {
            g_next = __ffs((int)gm) - 1;
            gm &= gm - 1;
            px_next = fetch(g_next);
        }
        const int wg = batch * 32 + g;
        const uint32_t gmisc = __shfl_sync(FULL, misc, g);
        const TgaMasks m = tga_masks(__shfl_sync(FULL, q.cur, g), (gmisc >> 13) & 1u, (gmisc >> 14) & 1u);
        TgaWordEmit e;
        e.alone = __shfl_sync(FULL, we.alone, g);
        e.ends = __shfl_sync(FULL, we.ends, g);
        e.rc = (gmisc >> 15) & 127u;
        e.r0c = (gmisc >> 22) & 127u;
        const unsigned inf = tga_lane_emit(m, __shfl_sync(FULL, q.rec.xm, g), e, wg, lane, n);
        const unsigned role = inf & 3u, k = (inf >> 2) & 127u;
        const unsigned m1 = __ballot_sync(FULL, role == 1u), m2 = __ballot_sync(FULL, role == 2u), mh = __ballot_sync(FULL, role == 2u && k == 0u);
        const unsigned pos = (gmisc & 0x1FFFu) + 4u * __popc(m1 & lt_mask) + 3u * __popc(m2 & lt_mask) + __popc(mh & lt_mask);
        if (role) {
            /* one path for both roles: a byte in front of the colour when the pixel ends a run packet ([k + 128]) or opens a
             * raw packet (its header, written by the packet's last pixel, 3 k + 1 bytes in front of that pixel's colour) */
            const bool run = role == 1u;
            uint8_t* c = bpay + (pos + ((run || k == 0u) ? 1u : 0u));
            c[0] = (uint8_t)(px >> 16); /* the file wants B,G,R: the buffer's u32 is R | G << 8 | B << 16 */
            c[1] = (uint8_t)(px >> 8);
            c[2] = (uint8_t)px;
            if (run || (inf & (1u << 9))) *(c - 1 - (run ? 0 : 3 * (int)k)) = (uint8_t)(k + (run ? 128u : 0u));
        }
        if (g_next < 0) break;
        g = g_next;
        px = px_next;
    }
    store_run_end();
}

}  // namespace hana
#endif /* HANA_TGA_CUH */
