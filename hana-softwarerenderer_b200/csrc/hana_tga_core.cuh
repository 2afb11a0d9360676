/*
 * hana_tga_core.cuh — the arithmetic of the device-side RLE TGA packetiser (hana_tga.cuh), as __host__ __device__
 * functions: the kernels are loops and scans around these, and tests/emu/emu_tga.cpp walks the very same functions on the
 * CPU box against the sequential algorithm (TGAImage::unload_rle_data, tgaimage.cpp:206-246).
 *
 * Statement (tests/rle_model.py, `words_packets`): with e[i] = (pixel i == pixel i+1) packed 32 pixels to a word,
 *   T-start  e[i] && !e[i-1]      tail  !e[i] && e[i-1]      stretch end  !e[i] && e[i+1]  (pixel i+1 is a T-start)
 * a stretch = the pixels from one T-start up to the next. All a stretch inherits from the pixels in front of it is ONE bit
 * x: its first pixel was taken, unseen, as the 128th pixel of the raw packet in front (the reference's loop tests the
 * length before it looks at the pixel). A pixel's packet, its position in it and the bytes it emits follow from
 * (a, t, x) = (last T-start, last tail, x of the stretch) at that pixel:
 *   run pixel, idx = i - (a + x): the last pixel of its packet (idx % 128 == 127, or the tail) emits [idx % 128 + 128, B, G, R];
 *   raw pixel, rawidx = position in its raw group: emits [B, G, R], preceded by a header byte when rawidx % 128 == 0; the
 *   last pixel of a raw packet writes the length into that header.
 */
#ifndef HANA_TGA_CORE_CUH
#define HANA_TGA_CORE_CUH

#include <stdint.h>

#ifdef __CUDACC__
#define TGA_HD __host__ __device__ __forceinline__
#else
#define TGA_HD inline
#endif

namespace hana {

constexpr int TGA_HEADER = 18, TGA_FOOTER = 26;

/* per 32-pixel word: the state at its first pixel. xm bit 0 = x there; bit j > 0 = x of the stretch that starts at pixel j
 * of the word (meaningful where pixel j is a T-start) */
struct alignas(16) TgaRec {
    int a, t;
    uint32_t xm, pad;
};

TGA_HD int tga_hi(uint32_t m) { /* highest set bit, m != 0 */
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}
TGA_HD int tga_lo(uint32_t m) { /* lowest set bit, m != 0 */
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

struct TgaMasks {
    uint32_t cur, sh, up; /* bit j: e(j), e(j-1), e(j+1) */
    uint32_t ts, tl, se;  /* T-starts, tails, stretch ends */
};
TGA_HD TgaMasks tga_masks(uint32_t cur, uint32_t eprev, uint32_t enext) { /* eprev: e of the pixel in front of the word, enext: of the one behind */
    TgaMasks m;
    m.cur = cur;
    m.sh = (cur << 1) | (eprev & 1u);
    m.up = (cur >> 1) | ((enext & 1u) << 31);
    m.ts = cur & ~m.sh;
    m.tl = ~cur & m.sh;
    m.se = ~cur & m.up;
    return m;
}

/* one-bit functions as two bits (bit v = f(v)); identity = 0b10 */
TGA_HD unsigned tga_compose(unsigned first, unsigned then) {
    return ((then >> (first & 1u)) & 1u) | (((then >> ((first >> 1) & 1u)) & 1u) << 1);
}

/* x of the next stretch, given the one that ends at pixel i (its T-start a, its tail t, its own x) */
TGA_HD int tga_next_x(int i, int a, int t, int x) {
    const int rawlen = a < 0 ? i + 1 : (i - t) + ((((t - a - x + 1) & 127) == 1) ? 1 : 0);
    return (rawlen & 127) == 127 ? 1 : 0;
}

/* ---- structure pass over a span of words [w0, w1) (one thread of tga_structure_kernel). E holds 16 zero words more than
 * the frame's batches have, w0 is a multiple of 4 and the frame's words are 16-byte aligned: a thread reads its span four
 * words at a time, two loads ahead of the word it works on. ---- */
constexpr int TGA_E_PAD = 16;
struct TgaQuad {
    uint32_t v[4];
};
TGA_HD TgaQuad tga_load4(const uint32_t* __restrict__ p) {
    TgaQuad q;
#ifdef __CUDA_ARCH__
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    q.v[0] = u.x, q.v[1] = u.y, q.v[2] = u.z, q.v[3] = u.w;
#else
    q.v[0] = p[0], q.v[1] = p[1], q.v[2] = p[2], q.v[3] = p[3];
#endif
    return q;
}
template <class Op>
TGA_HD void tga_walk_span(const uint32_t* __restrict__ E, int w0, int w1, Op& op) {
    if (w0 >= w1) return;
    uint32_t prev = w0 > 0 ? E[w0 - 1] : 0u;
    TgaQuad c = tga_load4(E + w0), n1 = tga_load4(E + w0 + 4);
    for (int w = w0; w < w1; w += 4) {
        const TgaQuad n2 = tga_load4(E + w + 8);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) {
            if (w + k < w1) {
                const uint32_t cur = c.v[k];
                op(w + k, tga_masks(cur, prev >> 31, k < 3 ? c.v[k + 1] : n1.v[0]));
                prev = cur;
            }
        }
        c = n1;
        n1 = n2;
    }
}

/* pass 0: last T-start / last tail inside the span (-1 = none) */
struct TgaLastOp {
    int la, lt;
    TGA_HD void operator()(int w, const TgaMasks& m) {
        if (m.ts) la = w * 32 + tga_hi(m.ts);
        if (m.tl) lt = w * 32 + tga_hi(m.tl);
    }
};
TGA_HD void tga_span_last(const uint32_t* __restrict__ E, int w0, int w1, int& la, int& lt) {
    TgaLastOp op;
    op.la = op.lt = -1;
    tga_walk_span(E, w0, w1, op);
    la = op.la;
    lt = op.lt;
}

/* Passes 1 and 2 need x' only at stretch ends, and a stretch end whose tail lies in the same word has a raw part shorter
 * than 32 pixels: x' = 0 whatever came before (x' = 1 needs rawlen % 128 == 127). So per word at most ONE stretch end — the
 * lowest, and only when no tail of the word lies at or below it; no T-start of the word does then either — needs the
 * general formula, with (a, t) from in front of the word: both passes are a handful of bit operations per word, no loop
 * over events. */

/* pass 1: x behind the span as a function of x in front of it; (a, t) = last T-start / tail in front of the span */
struct TgaFuncOp {
    int a, t;
    unsigned F;
    TGA_HD void operator()(int w, const TgaMasks& m) {
        if (m.se) {
            const int jl = tga_hi(m.se); /* the last stretch end of the word decides what leaves it */
            if (m.tl & (0xFFFFFFFFu >> (31 - jl))) {
                F = 0u; /* constant 0 */
            } else {
                const int i = w * 32 + jl;
                F = tga_compose(F, (unsigned)tga_next_x(i, a, t, 0) | ((unsigned)tga_next_x(i, a, t, 1) << 1));
            }
        }
        if (m.ts) a = w * 32 + tga_hi(m.ts);
        if (m.tl) t = w * 32 + tga_hi(m.tl);
    }
};
TGA_HD unsigned tga_span_function(const uint32_t* __restrict__ E, int w0, int w1, int a, int t) {
    TgaFuncOp op;
    op.a = a, op.t = t, op.F = 2u;
    tga_walk_span(E, w0, w1, op);
    return op.F;
}

/* pass 2: the per-word records, given the state in front of the span */
struct TgaRecOp {
    int a, t, x;
    TgaRec* __restrict__ R;
    TGA_HD void operator()(int w, const TgaMasks& m) {
        TgaRec r;
        r.a = a;
        r.t = t;
        r.xm = (uint32_t)x;
        r.pad = 0u;
        if (m.se) {
            const int js = tga_lo(m.se), jl = tga_hi(m.se);
            int xs = 0;
            if (!(m.tl & (0xFFFFFFFFu >> (31 - js)))) { /* the one stretch end that reaches back in front of the word */
                xs = tga_next_x(w * 32 + js, a, t, x);
                if (js < 31) r.xm |= (uint32_t)xs << (js + 1); /* x of the stretch that starts at the next pixel */
            }
            x = (m.tl & (0xFFFFFFFFu >> (31 - jl))) ? 0 : xs; /* no tail at or below the last stretch end: it is that one */
        }
        if (m.ts) a = w * 32 + tga_hi(m.ts);
        if (m.tl) t = w * 32 + tga_hi(m.tl);
        R[w] = r;
    }
};
TGA_HD void tga_span_records(const uint32_t* __restrict__ E, int w0, int w1, int a, int t, int x, TgaRec* __restrict__ R) {
    TgaRecOp op;
    op.a = a, op.t = t, op.x = x, op.R = R;
    tga_walk_span(E, w0, w1, op);
}

/* ---- per pixel (one lane of tga_count_kernel / tga_write_kernel) ---- */

/* (a, t, x) at pixel j of word w, from the word's record and masks */
TGA_HD void tga_lane_state(const TgaRec& r, const TgaMasks& m, int w, int j, int& a, int& t, int& x) {
    const uint32_t upto = 0xFFFFFFFFu >> (31 - j);
    const uint32_t ma = m.ts & upto, mt = m.tl & upto;
    const int ha = ma ? tga_hi(ma) : 0; /* xm bit 0 is the x of the stretch the word starts in, and of a T-start at bit 0 */
    a = ma ? w * 32 + ha : r.a;
    x = (int)((r.xm >> ha) & 1u);
    t = mt ? w * 32 + tga_hi(mt) : r.t;
}

/* what pixel i emits: bits 0-1 role (0 nothing, 1 last pixel of a run packet, 2 raw pixel), bits 2-8 k = position in the
 * packet, bit 9 (raw): last pixel of its packet, which writes the header. Straight-line: every lane of a warp runs it. */
TGA_HD unsigned tga_role(int i, int n, int a, int t, int x, bool is_e, bool e_prev, bool e_next) {
    const bool is_tail = !is_e && e_prev;
    const bool in_group = (is_e || is_tail) && a >= 0; /* the equal-pixel part of a stretch */
    const int r = a + x, idx = i - r;
    const bool swallowed = in_group && idx < 0;                      /* taken as the 128th pixel of the raw packet in front */
    const bool alone = in_group && is_tail && (idx & 127) == 0;      /* a tail alone in its packet: first pixel of the raw group behind */
    const bool run_pixel = in_group && !swallowed && !alone;
    const bool run_end = run_pixel && ((idx & 127) == 127 || is_tail);
    const int y = (((t - r + 1) & 127) == 1) ? 1 : 0;
    int rawidx = i - t - 1 + y;
    if (a < 0) rawidx = i;
    if (swallowed) rawidx = 127;
    if (alone) rawidx = 0;
    const int k = rawidx & 127;
    const bool nxt_tstart = e_next && !is_e;
    const bool last = k == 127 || i == n - 1 || (nxt_tstart && k != 126);
    const unsigned raw = 2u | ((unsigned)k << 2) | (last ? 1u << 9 : 0u);
    const unsigned run = run_end ? (1u | ((unsigned)(idx & 127) << 2)) : 0u;
    return run_pixel ? run : raw;
}
TGA_HD unsigned tga_lane_role(const TgaRec& r, const TgaMasks& m, int w, int j, int n) {
    const int i = w * 32 + j;
    int a, t, x;
    tga_lane_state(r, m, w, j, a, t, x);
    const unsigned inf = tga_role(i, n, a, t, x, (m.cur >> j) & 1u, (m.sh >> j) & 1u, (m.up >> j) & 1u);
    return i < n ? inf : 0u;
}
TGA_HD unsigned tga_role_bytes(unsigned inf) {
    const unsigned role = inf & 3u;
    return role == 1u ? 4u : role == 2u ? 3u + (((inf >> 2) & 127u) == 0u ? 1u : 0u) : 0u;
}

/* word classes with closed forms: the middle of a run (the background) and the middle of a raw stretch */
enum { TGA_W_EMPTY = 0, TGA_W_RUN = 1, TGA_W_RAW = 2, TGA_W_GENERAL = 3 };
TGA_HD int tga_word_class(int w, int nw, int n, uint32_t cur, uint32_t eprev, uint32_t enext) {
    if (w >= nw) return TGA_W_EMPTY;
    const bool full = (w + 1) * 32 <= n;
    if (full && cur == 0xFFFFFFFFu && (eprev & 1u)) return TGA_W_RUN;
    if (full && cur == 0u && !(eprev & 1u) && !(enext & 1u)) return TGA_W_RAW;
    return TGA_W_GENERAL;
}
/* TGA_W_RUN: every pixel is a run pixel with idx = i - r; the pixel that ends a packet (idx % 128 == 127), if any: j < 32 */
TGA_HD int tga_run_word_end(const TgaRec& r, int w) { return (127 - (w * 32 - (r.a + (int)(r.xm & 1u)))) & 127; }
/* TGA_W_RAW: consecutive raw pixels, rawidx = rawidx0 + j */
TGA_HD int tga_raw_word_idx0(const TgaRec& r, int w) {
    if (r.a < 0) return w * 32;
    return (w * 32 - r.t - 1) + ((((r.t - (r.a + (int)(r.xm & 1u)) + 1) & 127) == 1) ? 1 : 0);
}
TGA_HD unsigned tga_closed_word_bytes(int cls, const TgaRec& r, int w) {
    if (cls == TGA_W_RUN) return tga_run_word_end(r, w) < 32 ? 4u : 0u;
    if (cls == TGA_W_RAW) return 96u + ((((-tga_raw_word_idx0(r, w)) & 127) < 32) ? 1u : 0u);
    return 0u;
}


TGA_HD int tga_popc(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

/* Bytes the pixels of word w emit, in closed form (one lane of tga_count_kernel; equals the sum of tga_role_bytes over the
 * word's pixels, which tests/emu checks word by word). The word splits at its first T-start: the pixels in front of it
 * continue the stretch of the record (one packet boundary of the run part and one of the raw part can fall among them:
 * positions mod 128); the stretches that start inside the word are shorter than 32 pixels, so their run packets end at
 * their tails and their raw groups begin behind them — popcounts. */
TGA_HD unsigned tga_word_bytes(const TgaRec& r, const TgaMasks& m, int w, int n) {
    const int base = w * 32, nv = n - base;
    if (nv <= 0) return 0u;
    const uint32_t VM = nv >= 32 ? 0xFFFFFFFFu : ((1u << nv) - 1u);
    const uint32_t G = m.cur | m.tl; /* pixels of the equal-pixel parts */
    const uint32_t NR = ~G & VM;
    const uint32_t L = m.ts ? ~((m.ts & (0u - m.ts)) - 1u) : 0u; /* from the first T-start on */
    const uint32_t C = VM & ~L;
    /* stretches that start in the word */
    const uint32_t XS = m.ts & r.xm;           /* T-starts taken as the 128th pixel of the raw packet in front: raw, no header */
    const uint32_t LT = m.tl & L;
    const uint32_t alone = LT & (XS << 1);     /* a tail right behind such a T-start: alone in its packet -> first pixel of the raw group */
    const uint32_t ends = LT & ~alone;         /* run packets end at the other tails */
    unsigned bytes = 4u * (unsigned)tga_popc(ends) + 3u * (unsigned)(tga_popc(NR & L) + tga_popc(XS)) + 4u * (unsigned)tga_popc(alone) +
                     (unsigned)tga_popc((ends << 1) & NR); /* a raw pixel right behind a run packet opens a raw packet */
    /* the pixels that continue the record's stretch */
    if (C) {
        if (r.a < 0) {
            bytes += 3u * (unsigned)tga_popc(C) + ((w & 3) == 0 ? 1u : 0u); /* rawidx = i */
        } else {
            const int r0 = r.a + (int)(r.xm & 1u);
            const uint32_t RG = G & C, tcm = m.tl & C;
            int tt = r.t;
            if (RG) {
                const int j127 = (127 - (base - r0)) & 127;
                if (j127 < 32 && ((RG >> j127) & 1u)) bytes += 4u; /* a full packet ends here */
                if (tcm) {
                    const int tc = tga_lo(tcm), m7 = (base + tc - r0) & 127;
                    if (m7 != 127) bytes += 4u; /* m7 == 0: alone, raw with header; else the run's last packet; 127: counted above */
                    tt = base + tc;
                }
            }
            const uint32_t RW = NR & C;
            if (RW) {
                const int y = (((tt - r0 + 1) & 127) == 1) ? 1 : 0;
                const int j0 = (tt + 1 - y - base) & 127; /* where rawidx % 128 == 0 */
                bytes += 3u * (unsigned)tga_popc(RW) + ((j0 < 32 && ((RW >> j0) & 1u)) ? 1u : 0u);
            }
        }
    }
    return bytes;
}


/* ---- the write kernel's form of the per-pixel role. What the pixels in FRONT of a word's first T-start need of the record
 * (they continue its stretch) is folded by the word's lane into two masks and two small numbers; with them a pixel's role
 * follows from the word's masks alone — straight-line code, no (a, t) arithmetic per pixel (tests/emu checks it against
 * tga_lane_role pixel by pixel). ---- */
struct TgaWordEmit {
    uint32_t alone; /* tails that are alone in their packet: raw, first pixel of the raw group behind */
    uint32_t ends;  /* last pixels of run packets */
    uint32_t rc;    /* rawidx % 128 of (virtual) pixel 0 of the continued raw part */
    uint32_t r0c;   /* idx % 128 of (virtual) pixel 0 of the continued run part */
};
TGA_HD TgaWordEmit tga_word_emit(const TgaRec& r, const TgaMasks& m, int w, int n) {
    const int base = w * 32, nv = n - base;
    const uint32_t VM = nv >= 32 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
    const uint32_t G = m.cur | m.tl;
    const uint32_t L = m.ts ? ~((m.ts & (0u - m.ts)) - 1u) : 0u; /* from the first T-start on */
    const uint32_t C = VM & ~L;
    const uint32_t XS = m.ts & r.xm; /* T-starts taken as the 128th pixel of the raw packet in front */
    TgaWordEmit e;
    e.alone = m.tl & (XS << 1);
    e.ends = 0u;
    e.rc = (uint32_t)base & 127u;
    e.r0c = 0u;
    if (C && r.a >= 0) {
        const int r0 = r.a + (int)(r.xm & 1u);
        e.r0c = (uint32_t)(base - r0) & 127u;
        const uint32_t RG = G & C, tcm = m.tl & C;
        int tt = r.t;
        if (RG) {
            const uint32_t j127 = (127u - e.r0c) & 127u;
            if (j127 < 32u) e.ends = RG & (1u << j127); /* a full packet of the continued run ends here */
            if (tcm) {
                const int tc = tga_lo(tcm);
                if (((e.r0c + (uint32_t)tc) & 127u) == 0u) e.alone |= tcm;
                tt = base + tc;
            }
        }
        const int y = (((tt - r0 + 1) & 127) == 1) ? 1 : 0;
        e.rc = (uint32_t)(base - tt - 1 + y) & 127u;
    }
    e.ends |= m.tl & ~e.alone;
    return e;
}
/* same result as tga_lane_role; `e` is the word's, j the pixel */
TGA_HD unsigned tga_lane_emit(const TgaMasks& m, uint32_t xm, const TgaWordEmit& e, int w, int j, int n) {
    const int i = w * 32 + j;
    const uint32_t bit = 1u << j, upto = 0xFFFFFFFFu >> (31 - j);
    const uint32_t XS = m.ts & xm;
    const uint32_t ts_upto = m.ts & upto, tl_upto = m.tl & upto;
    const bool is_end = (e.ends & bit) != 0u;
    const bool is_raw = !is_end && (((~(m.cur | m.tl) | XS | e.alone) & bit) != 0u);
    const int ha = tga_hi(ts_upto | 1u), ht = tga_hi(tl_upto | 1u); /* 0 when there is none: unused then */
    /* run packet end: k = idx % 128 */
    const int k_end = ts_upto ? j - ha - (int)((xm >> ha) & 1u) : (int)((e.r0c + (uint32_t)j) & 127u);
    /* raw: k = rawidx % 128 */
    int k_raw = tl_upto ? j - ht - 1 + (int)((e.alone >> ht) & 1u) : (int)((e.rc + (uint32_t)j) & 127u);
    if (e.alone & bit) k_raw = 0;
    if (XS & bit) k_raw = 127;
    const bool last = k_raw == 127 || i == n - 1 || ((m.se & bit) && k_raw != 126);
    unsigned out = 0u;
    if (is_raw) out = 2u | ((unsigned)k_raw << 2) | (last ? 1u << 9 : 0u);
    if (is_end) out = 1u | ((unsigned)k_end << 2);
    return i < n ? out : 0u;
}

}  // namespace hana
#endif /* HANA_TGA_CORE_CUH */
