/*
 * hana_tga.cuh — RLE-compressed TGA files made on the device (the output side of the path).
 *
 * Replaces TGAImage::write_tga_file(rle = true) + unload_rle_data (tgaimage.cpp:145-246) for the frames of a sweep: the
 * file bytes are produced in HBM, so a frame leaves the GPU as ~1.4 MB instead of 8.3 MB of raw colour (the bundled
 * scenes are ~80 % background), which is what the PCIe-bound end-to-end path needs. Output is BYTE-IDENTICAL to the
 * reference's writer (and to hana_tga_write): same packets, same header, same footer.
 *
 * The reference's packetiser is sequential and greedy. Its parallel formulation — derived and checked against the
 * sequential one in tests/rle_model.py, of which the kernel below is a transcription — rests on one observation: with
 * e[i] = (pixel i == pixel i+1), everything a stretch of pixels (from one "e turns true" to the next) inherits from
 * the pixels before it is ONE bit x (its first pixel was taken as the 128th pixel of the raw packet in front). x is a
 * prefix composition of one-bit functions over the stretches; packet boundaries, packet headers and byte offsets then
 * follow from per-pixel arithmetic and an ordinary prefix sum.
 *
 * One CTA per chunk of 4096 pixels; chunks of a frame are chained (each waits for its predecessor's carry), handed out
 * chunk-major / frame-minor by an atomic ticket so that a predecessor has always started. The carry travels in two
 * instalments so that the serial part of a hop stays short: the structure (last T-start, last tail, x) is published as
 * soon as the chunk's one-bit functions are composed, the byte offset later, when the chunk's own byte count is known;
 * a successor needs the offset only when it places its bytes. Bytes are assembled in shared memory and leave as
 * 16-byte stores.
 */
#ifndef HANA_TGA_CUH
#define HANA_TGA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace hana {

constexpr int TGA_THREADS = 256;
constexpr int TGA_PPT = 16;                         /* pixels per thread */
constexpr int TGA_CHUNK = TGA_THREADS * TGA_PPT;    /* pixels per CTA */
constexpr int TGA_HEADER = 18, TGA_FOOTER = 26;
#ifndef TGA_MIN_CTAS
#define TGA_MIN_CTAS 6 /* resident CTAs per SM the encoder is compiled for: the phases of a CTA are separated by barriers, other CTAs fill the gaps */
#endif

struct alignas(16) TgaCarry { /* per frame: what chunk c hands to chunk c+1, in two instalments of one 16-byte word each */
    int a_last, t_last;       /* last T-start / last tail so far (pixel index, -1 = none) */
    int x;                    /* x of the stretch the next pixel lies in */
    unsigned int seq_x;       /* chunks whose structure (a, t, x) has been published: the successor can assign roles */
    unsigned int off_lo, off_hi; /* payload bytes emitted so far */
    unsigned int seq_off;     /* chunks whose byte total has been published: the successor can place its bytes */
    int pad;
};

/* worst case of a frame's file: every packet raw (3 bytes per pixel + one header per 128) + header + footer, rounded */
__host__ __device__ inline size_t tga_slot_bytes(size_t npix) { return ((npix * 3 + npix / 64 + 1024 + TGA_HEADER + TGA_FOOTER) + 255) / 256 * 256; }

__device__ __forceinline__ int tga_pad(int k) { return k + (k >> 5); } /* bank-conflict-free stride-16 reads */
/* The two instalments of the carry are 16-byte words that carry their own sequence number, written and polled as ONE
 * 128-bit access each (an aligned 128-bit access is a single transaction: value and flag arrive together, so the hop
 * needs no fence; the same device CUB's decoupled look-back uses for its tile descriptors). */
__device__ __forceinline__ uint4 tga_poll(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tga_post(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
/* one-bit functions as two bits (bit v = f(v)); identity = 0b10 */
__device__ __forceinline__ unsigned tga_compose(unsigned first, unsigned then) {
    return ((then >> (first & 1u)) & 1u) | (((then >> ((first >> 1) & 1u)) & 1u) << 1);
}

__global__ void __launch_bounds__(TGA_THREADS, TGA_MIN_CTAS)
    tga_rle_kernel(const uint32_t* __restrict__ color, size_t frame_stride, int first, int W, int H, int n_frames,
                   uint8_t* __restrict__ slots, size_t slot_bytes, TgaCarry* carry, unsigned long long* __restrict__ sizes,
                   unsigned int* ticket) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int NW = TGA_THREADS / 32;
    __shared__ uint32_t s_px[TGA_CHUNK + 3 + (TGA_CHUNK + 3) / 32 + 1];
    __shared__ __align__(16) uint8_t s_out[TGA_CHUNK * 4 + 64];
    __shared__ int s_wa[NW], s_wt[NW];
    __shared__ unsigned s_wf[NW], s_ws[NW];
    __shared__ unsigned s_ticket;
    __shared__ TgaCarry s_carry;
    __shared__ int s_hole;
    __shared__ unsigned s_total;
    __shared__ unsigned long long s_base;
    __shared__ int s_first, s_first_a, s_first_t, s_vfirst, s_amax, s_tmax;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(ticket, 1u);
        s_hole = -1;
        s_first = TGA_CHUNK;
        s_first_a = s_first_t = -1;
    }
    __syncthreads();
    const unsigned tk = s_ticket;
    const int c = (int)(tk / (unsigned)n_frames), f = (int)(tk % (unsigned)n_frames);
    const int n = W * H;
    const int i0 = c * TGA_CHUNK;
    /* one mailbox per (frame, chunk): chunk c polls its own (written by chunk c-1) and posts into chunk c+1's. One shared
     * mailbox per frame made every waiting CTA of the frame poll the same 32 bytes, and that one L2 line then took ~4 us
     * to hand a value over (8 us per hop, measured, whatever the CTAs did in between). */
    const int n_chunks = (W * H + TGA_CHUNK - 1) / TGA_CHUNK;
    TgaCarry* const my_box = carry + (size_t)f * (n_chunks + 1) + c;
    TgaCarry* const next_box = my_box + 1;
    const uint32_t* src = color + (size_t)(first + f) * frame_stride;
    uint8_t* slot = slots + (size_t)f * slot_bytes;

    /* pixels i0-1 .. i0+CHUNK+1 in file order (rows top-down: file row r is buffer row H-1-r, win32.cpp:358), RGB only */
    {
        /* 16 coalesced loads per thread, all in flight before the first is used; one division per thread, not per pixel */
        uint32_t ld[TGA_PPT];
        int i = i0 - 1 + tid;
        int row = i >= 0 ? i / W : 0, col = i >= 0 ? i - row * W : -1;
#pragma unroll
        for (int q = 0; q < TGA_PPT; q++) {
            ld[q] = (i >= 0 && i < n) ? __ldg(src + (size_t)(H - 1 - row) * W + col) : 0u;
            i += TGA_THREADS;
            col += TGA_THREADS;
            while (col >= W) {
                col -= W;
                row++;
            }
        }
#pragma unroll
        for (int q = 0; q < TGA_PPT; q++) s_px[tga_pad(tid + q * TGA_THREADS)] = ld[q] & 0x00FFFFFFu;
        if (tid < 3) { /* the last three of the CHUNK + 3 pixels */
            const int k = TGA_CHUNK + tid, ih = i0 - 1 + k;
            uint32_t vh = 0;
            if (ih < n) {
                const int rh = ih / W, ch = ih - rh * W;
                vh = __ldg(src + (size_t)(H - 1 - rh) * W + ch) & 0x00FFFFFFu;
            }
            s_px[tga_pad(k)] = vh;
        }
    }
    __syncthreads();

    /* the thread's 16 pixels p0..p0+15 and the halo p0-1, p0+16, p0+17 */
    const int p0 = i0 + tid * TGA_PPT;
    uint32_t v[TGA_PPT + 3];
#pragma unroll
    for (int j = 0; j < TGA_PPT + 3; j++) v[j] = s_px[tga_pad(tid * TGA_PPT + j)]; /* v[j] = pixel p0 - 1 + j */
    /* e(m) for m = -1..16: pixel p0+m equals pixel p0+m+1 (both inside the image); bit m+1 of EB */
    unsigned EB = 0;
#pragma unroll
    for (int m = -1; m <= TGA_PPT; m++) {
        const int i = p0 + m;
        if (i >= 0 && i + 1 < n && v[m + 1] == v[m + 2]) EB |= 1u << (m + 1);
    }
#define TGA_E(m) ((EB >> ((m) + 1)) & 1u)

    /* ---- fast path: every pixel of the chunk equals its successor and the pixel in front of the chunk equals the first
     * (the middle of a long run: the background, ~80 % of the chunks of the bundled scenes). No T-start, no tail: the
     * structure passes through unchanged, every pixel is a run pixel with idx = i - r, and exactly CHUNK / 128 packets
     * end inside the chunk, each [255, B, G, R]. ---- */
    {
        const unsigned want = (p0 + TGA_PPT <= n) ? 0x1FFFFu : 0u; /* bits 0..16: e(-1)..e(15); a chunk that reaches the end is not all-equal (e[n-1] is false) */
        const int fast = __syncthreads_and(want != 0u && (EB & want) == want);
        if (fast) {
            if (tid == 0) {
                const uint4* cr = reinterpret_cast<const uint4*>(my_box);
                uint4* cw = reinterpret_cast<uint4*>(next_box);
                uint4 q0;
                do q0 = tga_poll(cr); while (q0.w != (unsigned)c); /* c > 0: the pixel in front of the chunk exists */
                tga_post(cw, make_uint4(q0.x, q0.y, q0.z, (unsigned)(c + 1))); /* a_last, t_last, x stay as they are */
                s_carry.a_last = (int)q0.x;
                s_carry.x = (int)q0.z;
                uint4 q1;
                do q1 = tga_poll(cr + 1); while (q1.z != (unsigned)c);
                const unsigned long long b0 = (unsigned long long)q1.x | ((unsigned long long)q1.y << 32);
                const unsigned long long b1 = b0 + (unsigned long long)(TGA_CHUNK / 128) * 4ull;
                tga_post(cw + 1, make_uint4((unsigned)b1, (unsigned)(b1 >> 32), (unsigned)(c + 1), 0u));
                s_base = b0;
            }
            __syncthreads();
            const int r = s_carry.a_last + s_carry.x;
            const int j = (127 - (p0 - r)) & 127; /* the thread's pixel that ends a packet, if j < 16 */
            if (j < TGA_PPT) {
                const int first_end = i0 + ((127 - (i0 - r)) & 127);
                const unsigned rank = (unsigned)(p0 + j - first_end) >> 7;
                const uint32_t px = v[j + 1];
                uint8_t* o = slot + TGA_HEADER + s_base + 4ull * rank;
                o[0] = 255;
                o[1] = (uint8_t)(px >> 16);
                o[2] = (uint8_t)(px >> 8);
                o[3] = (uint8_t)px;
            }
            return;
        }
    }

    /* ---- phase 1: last T-start / last tail before each thread (block max-scan) ---- */
    /* bit j of t_starts: e(j) && !e(j-1); of tails: !e(j) && e(j-1) (a pixel beyond the image has e = 0 and no e in front) */
    const unsigned t_starts = (EB >> 1) & ~EB & 0xFFFFu, tails = ~(EB >> 1) & EB & 0xFFFFu;
    const int la = t_starts ? p0 + 31 - __clz((int)t_starts) : -1, lt = tails ? p0 + 31 - __clz((int)tails) : -1;
    const unsigned stretch_ends = (EB >> 2) & ~(EB >> 1) & 0xFFFFu; /* bit j: pixel p0+j+1 is a T-start */
    if (stretch_ends) atomicMin(&s_first, tid * TGA_PPT + (__ffs(stretch_ends) - 1));
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
        s_wa[wid] = ia;
        s_wt[wid] = it;
    }
    __syncthreads();
    int pa = __shfl_up_sync(FULL, ia, 1), pt = __shfl_up_sync(FULL, it, 1);
    if (lane == 0) pa = pt = -1;
    for (int w = 0; w < wid; w++) {
        pa = max(pa, s_wa[w]);
        pt = max(pt, s_wt[w]);
    }
    /* ---- phase 2, before the carry is needed: x of the stretch each thread starts in, as a function of v = the x of the
     * chunk's first NEW stretch. A stretch end (pixel whose successor is a T-start) turns x into f(x), f from the stretch's
     * T-start and tail. Only the FIRST stretch end of a chunk can reach back into earlier chunks for those (every later
     * one follows a T-start inside the chunk), so all functions but that one are composed here, off the chain's critical
     * path; the first is evaluated by one thread when the carry arrives. ---- */
    const int first_end = s_first; /* chunk-relative pixel index of the first stretch end; TGA_CHUNK if there is none */
    const bool owns_first = first_end >= tid * TGA_PPT && first_end < (tid + 1) * TGA_PPT;
    unsigned F = 2u; /* identity: a thread without a stretch end (the usual case) hands x through */
    if (stretch_ends) {
        F = 0;
#pragma unroll
        for (int vx = 0; vx < 2; vx++) {
            int xc = vx, a = pa, t = pt; /* local prefix only: where it lacks a T-start or tail, the stretch end is the chunk's first */
#pragma unroll
            for (int j = 0; j < TGA_PPT; j++) {
                const int i = p0 + j;
                if (i < n) {
                    if (TGA_E(j) && !TGA_E(j - 1)) a = i;
                    if (!TGA_E(j) && TGA_E(j - 1)) t = i;
                    if (TGA_E(j + 1) && !TGA_E(j)) { /* pixel i ends its stretch */
                        if (tid * TGA_PPT + j == first_end) {
                            if (vx == 0) { /* what the chunk holds itself of the first stretch: T-start / tail at or before its end, or -1 */
                                s_first_a = a;
                                s_first_t = t;
                            }
                            xc = vx; /* left out here: v is defined as the x AFTER the first stretch end */
                        } else {
                            const int rawlen = (i - t) + ((((t - a - xc + 1) & 127) == 1) ? 1 : 0);
                            xc = ((rawlen & 127) == 127) ? 1 : 0;
                        }
                    }
                }
            }
            F |= (unsigned)xc << vx;
        }
    }
    unsigned P = F;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(FULL, P, d);
        if (lane >= d) P = tga_compose(o, P);
    }
    if (lane == 31) s_wf[wid] = P;
    if (tid == TGA_THREADS - 1) {
        s_amax = max(pa, la);
        s_tmax = max(pt, lt);
    }
    __syncthreads();
    unsigned Pex = __shfl_up_sync(FULL, P, 1);
    if (lane == 0) Pex = 2u; /* identity */
    unsigned Pw = 2u;
    for (int w = 0; w < wid; w++) Pw = tga_compose(Pw, s_wf[w]);
    Pex = tga_compose(Pw, Pex);
    const bool last_chunk = i0 + TGA_CHUNK >= n;

    /* ---- the predecessor's structure carry, and at once this chunk's (first instalment): the serial part of a hop ---- */
    if (tid == 0) {
        int ca = -1, ct = -1, cx = 0;
        if (c > 0) {
            uint4 q0;
            do q0 = tga_poll(reinterpret_cast<const uint4*>(my_box)); while (q0.w != (unsigned)c);
            ca = (int)q0.x;
            ct = (int)q0.y;
            cx = (int)q0.z;
        }
        int vfirst = cx, xout = cx;
        if (first_end < TGA_CHUNK) {
            const int is = i0 + first_end;
            const int a = s_first_a >= 0 ? s_first_a : ca, t = s_first_t >= 0 ? s_first_t : ct;
            const int rawlen = a < 0 ? is + 1 : (is - t) + ((((t - a - cx + 1) & 127) == 1) ? 1 : 0);
            vfirst = ((rawlen & 127) == 127) ? 1 : 0;
            unsigned Ptot = 2u;
#pragma unroll
            for (int w = 0; w < NW; w++) Ptot = tga_compose(Ptot, s_wf[w]);
            xout = (int)((Ptot >> vfirst) & 1u);
        }
        if (!last_chunk)
            tga_post(reinterpret_cast<uint4*>(next_box),
                     make_uint4((unsigned)max(s_amax, ca), (unsigned)max(s_tmax, ct), (unsigned)xout, (unsigned)(c + 1)));
        s_carry.a_last = ca;
        s_carry.t_last = ct;
        s_carry.x = cx;
        s_vfirst = vfirst;
    }
    __syncthreads();
    const int a_in = max(pa, s_carry.a_last), t_in = max(pt, s_carry.t_last);
    /* a thread that starts at or before the first stretch end is still in the stretch the chunk began in */
    const int x_start = (tid * TGA_PPT > first_end) ? (int)((Pex >> (s_vfirst & 1)) & 1u) : s_carry.x;
    (void)owns_first;

    /* ---- phase 3: what each pixel emits; info[j]: bits 0-1 role (0 nothing, 1 last pixel of a run packet, 2 raw),
     * bits 2-8 position in the packet (k), bit 9 raw: last pixel of its packet ---- */
    uint16_t info[TGA_PPT];
    unsigned S = 0;
    const unsigned ebits = EB & 0x1FFFFu; /* e(-1)..e(15) */
    if (p0 + TGA_PPT <= n && ebits == 0x1FFFFu) {
        /* the thread lies inside a run (most background threads): idx = i - r, at most one of its pixels ends a packet */
        const int r = a_in + x_start;
        const int je = (127 - (p0 - r)) & 127;
#pragma unroll
        for (int j = 0; j < TGA_PPT; j++) info[j] = (uint16_t)(j == je ? (1u | (127u << 2)) : 0u);
        if (je < TGA_PPT) S = 4;
    } else if (ebits == 0u) {
        /* no pixel of the thread (nor the one in front) equals its successor (most threads inside the model): consecutive
         * raw pixels; only the last one can end its stretch */
        int rawidx0 = p0;
        if (a_in >= 0) rawidx0 = (p0 - t_in - 1) + ((((t_in - (a_in + x_start) + 1) & 127) == 1) ? 1 : 0);
#pragma unroll
        for (int j = 0; j < TGA_PPT; j++) {
            unsigned inf = 0;
            if (p0 + j < n) {
                const unsigned k = (unsigned)(rawidx0 + j) & 127u;
                const bool last = k == 127u || p0 + j == n - 1 || (j == TGA_PPT - 1 && TGA_E(TGA_PPT) && ((k + 1u) & 127u) != 127u);
                inf = 2u | (k << 2) | (last ? 1u << 9 : 0u);
                S += 3 + (k == 0u ? 1 : 0);
            }
            info[j] = (uint16_t)inf;
        }
    } else {
        int xc = x_start, a = a_in, t = t_in;
#pragma unroll
        for (int j = 0; j < TGA_PPT; j++) {
            const int i = p0 + j;
            unsigned inf = 0;
            if (i < n) {
                if (TGA_E(j) && !TGA_E(j - 1)) a = i;
                if (!TGA_E(j) && TGA_E(j - 1)) t = i;
                const bool is_e = TGA_E(j), is_tail = !is_e && TGA_E(j - 1);
                const bool nxt_tstart = TGA_E(j + 1) && !is_e;
                int rawidx = -1;
                if (a < 0) {
                    rawidx = i;
                } else if (is_e || is_tail) { /* the equal-pixel part of the stretch */
                    const int r = a + xc;
                    if (i < r) {
                        rawidx = 127; /* taken as the 128th pixel of the raw packet in front */
                    } else {
                        const int idx = i - r;
                        if (is_tail && ((idx + 1) & 127) == 1) {
                            rawidx = 0; /* alone in its packet: first pixel of the raw group that follows */
                        } else if ((idx & 127) == 127 || is_tail) {
                            inf = 1u | ((unsigned)(idx & 127) << 2);
                            S += 4;
                        }
                    }
                } else {
                    const int r = a + xc;
                    const int y = (((t - r + 1) & 127) == 1) ? 1 : 0;
                    rawidx = (i - t - 1) + y;
                }
                if (rawidx >= 0) {
                    const int k = rawidx & 127;
                    const bool last = k == 127 || i == n - 1 || (nxt_tstart && ((k + 1) & 127) != 127);
                    inf = 2u | ((unsigned)k << 2) | (last ? 1u << 9 : 0u);
                    S += 3 + (k == 0 ? 1 : 0);
                }
                if (nxt_tstart) { /* x of the next stretch (the same update as in phase 2) */
                    const int rawlen = a < 0 ? i + 1 : (i - t) + ((((t - a - xc + 1) & 127) == 1) ? 1 : 0);
                    xc = ((rawlen & 127) == 127) ? 1 : 0;
                }
            }
            info[j] = (uint16_t)inf;
        }
    }
    unsigned Si = S;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(FULL, Si, d);
        if (lane >= d) Si += o;
    }
    if (lane == 31) s_ws[wid] = Si;
    __syncthreads();
    unsigned so = Si - S;
    for (int w = 0; w < wid; w++) so += s_ws[w];
    if (tid == TGA_THREADS - 1) {
        /* ---- the predecessor's byte offset, and the second instalment: this chunk's, before its bytes are written ---- */
        unsigned long long b0 = 0;
        uint4* cw = reinterpret_cast<uint4*>(next_box) + 1;
        if (c > 0) {
            uint4 q1;
            do q1 = tga_poll(reinterpret_cast<const uint4*>(my_box) + 1); while (q1.z != (unsigned)c);
            b0 = (unsigned long long)q1.x | ((unsigned long long)q1.y << 32);
        }
        if (!last_chunk) {
            const unsigned long long b1 = b0 + so + S;
            tga_post(cw, make_uint4((unsigned)b1, (unsigned)(b1 >> 32), (unsigned)(c + 1), 0u));
        }
        s_base = b0;
        s_total = so + S;
    }
    __syncthreads();
    const unsigned total = s_total;
    const unsigned long long base = s_base;

    /* ---- bytes into shared memory; a packet header that lies in the predecessor's range goes straight to HBM ---- */
    const unsigned shift = (unsigned)((TGA_HEADER + base) & 15ull); /* s_out[shift + k] <-> payload byte base + k */
    {
        unsigned pos = so; /* payload offset of the thread's next byte, relative to base */
#pragma unroll
        for (int j = 0; j < TGA_PPT; j++) {
            const unsigned inf = info[j];
            const unsigned role = inf & 3u, k = (inf >> 2) & 127u;
            const uint32_t px = v[j + 1];
            if (role == 1u) {
                uint8_t* o = s_out + shift + pos;
                o[0] = (uint8_t)(k + 128u);
                o[1] = (uint8_t)(px >> 16); /* the file wants B,G,R: the buffer's u32 is R | G << 8 | B << 16 */
                o[2] = (uint8_t)(px >> 8);
                o[3] = (uint8_t)px;
                pos += 4;
            } else if (role == 2u) {
                const unsigned cpos = pos + (k == 0 ? 1u : 0u);
                uint8_t* o = s_out + shift + cpos;
                o[0] = (uint8_t)(px >> 16);
                o[1] = (uint8_t)(px >> 8);
                o[2] = (uint8_t)px;
                const long long hrel = (long long)cpos - 3ll * k - 1ll; /* the packet's header, relative to base */
                if (inf & (1u << 9)) {
                    if (hrel >= 0) s_out[shift + hrel] = (uint8_t)k;
                    else slot[TGA_HEADER + base + hrel] = (uint8_t)k;
                } else if (p0 + j == min(i0 + TGA_CHUNK, n) - 1 && hrel >= 0) {
                    s_hole = (int)(shift + hrel); /* the chunk ends inside this packet: its header is the successor's to write */
                }
                pos = cpos + 3;
            }
        }
    }
    __syncthreads();

    /* ---- copy out: 16-byte stores where a block is complete, bytes at the ragged ends and around the hole ---- */
    {
        uint8_t* dst = slot + (TGA_HEADER + base - shift); /* 16-byte aligned */
        const unsigned lo = shift, hi = shift + total;
        const int hole = s_hole;
        for (unsigned b = tid; b * 16u < hi; b += TGA_THREADS) {
            const unsigned b0 = b * 16u, b1 = b0 + 16u;
            if (b0 >= lo && b1 <= hi && !(hole >= (int)b0 && hole < (int)b1)) {
                *reinterpret_cast<uint4*>(dst + b0) = *reinterpret_cast<const uint4*>(s_out + b0);
            } else {
                for (unsigned q = max(b0, lo); q < min(b1, hi); q++)
                    if ((int)q != hole) dst[q] = s_out[q];
            }
        }
    }
    if (c == 0 && tid < TGA_HEADER) { /* TGA_Header, tgaimage.cpp:157-163: type 10 (RLE true colour), 24 bits, top-left origin */
        uint8_t hb = 0;
        if (tid == 2) hb = 10;
        if (tid == 12) hb = (uint8_t)(W & 255);
        if (tid == 13) hb = (uint8_t)(W >> 8);
        if (tid == 14) hb = (uint8_t)(H & 255);
        if (tid == 15) hb = (uint8_t)(H >> 8);
        if (tid == 16) hb = 24;
        if (tid == 17) hb = 0x20;
        slot[tid] = hb;
    }
    if (last_chunk) { /* developer area ref, extension area ref, "TRUEVISION-XFILE.\0": tgaimage.cpp:146-148 */
        const unsigned long long end = TGA_HEADER + base + total;
        if (tid < TGA_FOOTER) {
            const char sig[18] = {'T', 'R', 'U', 'E', 'V', 'I', 'S', 'I', 'O', 'N', '-', 'X', 'F', 'I', 'L', 'E', '.', 0};
            slot[end + tid] = tid < 8 ? (uint8_t)0 : (uint8_t)sig[tid - 8];
        }
        if (tid == 0) {
            sizes[f] = end + TGA_FOOTER;
        }
    }
#undef TGA_E
}

/* sizes -> 16-byte aligned offsets of the files in the packed buffer; offsets[n_frames] = end of the last file */
__global__ void tga_offsets_kernel(const unsigned long long* __restrict__ sizes, unsigned long long* __restrict__ offsets, int n_frames) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long o = 0;
        for (int f = 0; f < n_frames; f++) {
            offsets[f] = o;
            o += (sizes[f] + 15ull) & ~15ull;
        }
        offsets[n_frames] = o;
    }
}
/* files from their worst-case slots to where they lie back to back (16-byte aligned starts) */
__global__ void __launch_bounds__(256) tga_pack_kernel(const uint8_t* __restrict__ slots, size_t slot_bytes,
                                                       const unsigned long long* __restrict__ sizes,
                                                       const unsigned long long* __restrict__ offsets, uint8_t* __restrict__ packed) {
    const int f = blockIdx.y;
    const unsigned long long n16 = (sizes[f] + 15ull) >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(slots + (size_t)f * slot_bytes);
    uint4* d = reinterpret_cast<uint4*>(packed + offsets[f]);
    for (unsigned long long k = (unsigned long long)blockIdx.x * 256ull + threadIdx.x; k < n16; k += (unsigned long long)gridDim.x * 256ull) d[k] = s[k];
}

}  // namespace hana
#endif /* HANA_TGA_CUH */
