/*
 * hana_kernels.cuh — the sm_100a kernels of the rasterisation path.
 *
 *   begin_kernel   HanaUniforms -> DevUniforms (matrix products hoisted), counters zeroed
 *   setup_kernel   vertex shading + homogeneous clip + cull + triangle setup, warp-scan
 *                  compaction of the surviving triangles
 *   pairs_kernel   <false>: per-tile counts of the triangles that can touch the tile; <true>: see fill
 *   scan_kernel    exclusive scan of the tile counts -> list offsets + non-empty tile queue
 *   pairs_kernel<true>  ("fill") triangle raster records copied into the per-tile lists (no indirection in the rasteriser)
 *   raster_kernel  persistent CTAs, one 16x16 tile at a time: coverage + depth resolve in
 *                  registers, shading of the winning fragment, tile flush by TMA store;
 *                  empty tiles are cleared by fire-and-forget TMA stores from a constant tile
 *   vertex_kernel  stage-level entry point (one thread per corner)
 *   fill32/checksum/count helpers
 *
 * Replaces graphics.cpp:378-407 (graphics_draw_triangle) and everything it
 * calls. Visibility: the reference writes a fragment iff !(z > stored)
 * (graphics.cpp:359) while walking primitives in submission order, no shader
 * discards and nothing blends, so the surviving fragment of a pixel is the one
 * with the smallest depth and, among equal depths, the largest submission key
 * (SURVEY.md §0). The rasteriser resolves exactly that pair, which makes the
 * per-tile lists order-free, and shades only the winner.
 */
#ifndef HANA_KERNELS_CUH
#define HANA_KERNELS_CUH

#include <cuda.h>
#include <type_traits>
#include <cuda_runtime.h>

#include "hana_core.cuh"
#include "hana_pack.cuh"

namespace hana {

constexpr int TILE = 16;             /* screen tile edge (pixels) */
constexpr int TILE_PIX = TILE * TILE;
constexpr int TILE_BITS = 20;        /* work item = frame << 20 | tile_y << 10 | tile_x */
constexpr uint32_t TILE_MASK = (1u << TILE_BITS) - 1u;
constexpr uint32_t WORK_INVALID = 0xFFFFFFFFu;
#ifndef HANA_SETUP_THREADS
#define HANA_SETUP_THREADS 256
#endif
constexpr int SETUP_THREADS = HANA_SETUP_THREADS;
#ifndef HANA_SETUP_CTAS
#define HANA_SETUP_CTAS 3 /* resident CTAs per SM setup_kernel is compiled for */
#endif
#ifndef HANA_MICRO_EXTENT
#define HANA_MICRO_EXTENT 5 /* configs[3], ms per frame: 3 -> 3.37 (793 k list records left), 4 -> 3.16 (18 k), 5 -> 3.13 (172), 8 -> 3.09 */
#endif
static_assert(HANA_MICRO_EXTENT >= 1 && HANA_MICRO_EXTENT <= 16, "a micro-triangle's pixel range must fit 2 x 2 tiles");
constexpr int MICRO_EXTENT = HANA_MICRO_EXTENT; /* a triangle of a dense mesh whose pixel range is at most this many pixels on a side takes the visibility-buffer path (at most MICRO_EXTENT^2 pixel tests by its thread in setup_kernel) instead of the tile lists */
constexpr int TRI_COUNT_WAYS = 32;   /* per-frame statistics counters: one atomic per warp, spread so that they do not queue on one address */
constexpr uint32_t DEAD_BBY = 0x0000FFFFu; /* bby of a slot that holds no triangle (y0 = 0xFFFF > y1 = 0: an empty range for every consumer) */
constexpr int SCAN_THREADS = 1024;

enum RasterMode {
    MODE_CLEAR_FOLD = 0, /* target holds (clear colour, clear depth) before the draw; every tile is written once */
    MODE_RMW = 1,        /* target has contents that take part in the depth test; only covered pixels change */
    MODE_SHADOW_R8 = 2,  /* ShadowShader into the sweep's internal 1-byte-per-texel maps, cleared to 0 */
    MODE_SHADOW_R8_WIDE = 3 /* the same for passes that may emit more than R8_SLOT_LIMIT triangles per frame: the shadow byte lives
                               beside the state word instead of in its top 8 bits, so the triangle slot keeps all 32 bits */
};
constexpr int N_RASTER_MODES = 4;
constexpr uint32_t R8_SLOT_LIMIT = 0x00FFFFFEu; /* largest triangle capacity MODE_SHADOW_R8's 24-bit slot field can address */
__host__ __device__ constexpr bool mode_is_r8(int mode) { return mode == MODE_SHADOW_R8 || mode == MODE_SHADOW_R8_WIDE; }

/* Device-side bookkeeping of one pass (all frames of a batch). */
struct PassCounters {
    uint32_t pool_used;     /* tile-list pool entries reserved by scan_kernel */
    uint32_t n_work;        /* non-empty (frame,tile) items queued */
    uint32_t work_cursor;   /* raster queue head */
    uint32_t clear_cursor;  /* clear queue head (all (frame,tile) slots) */
    uint32_t tri_needed;    /* max over frames of triangles emitted (capacity check) */
    uint32_t tiles_touched;
    uint32_t pad[2];
};

/* Capacity needs of the last render of a sweep, kept across both passes (atomicMax): lets the host launch a whole
 * batch without reading anything back in the middle and check afterwards that nothing was dropped. */
struct OverflowRecord {
    uint32_t tri_needed;  /* max over passes and frames of triangles emitted */
    uint32_t pool_needed; /* max over passes of tile-list records */
    uint32_t pad[2];
};

struct PassParams {
    /* geometry */
    const float4* posu;   /* per corner: obj_pos.xyz, uv.x */
    const float4* nrmv;   /* per corner: obj_normal.xyz, uv.y */
    int nfaces;
    int n_frames;
    int W, H, tiles_x, tiles_y, n_tiles;
    int band_y0, band_y1; /* tile rows [band_y0, band_y1) this pass renders: the whole frame, or one GPU's band of a split frame */
    const DevUniforms* uniforms; /* [n_frames] */
    /* scratch */
    float4* tri_rec;      /* [n_frames][tri_cap][4]: raster records */
    float4* tri_attr;     /* [n_frames][tri_cap][attr_quads] */
    uint2* tri_bbox;      /* [n_frames][tri_cap] pixel range (bbx, bby) of the slot's triangle as the PAIR kernels see it: DEAD_BBY
                             in .y for a slot that is not to be listed (no triangle, or a micro-triangle). 8 bytes per slot
                             instead of a 16-byte read out of every 64-byte record */
    uint32_t tri_cap;     /* slots per frame: slot = face index for an unclipped face, slots >= nfaces for what clipping emits */
    uint32_t* tri_count;  /* [n_frames][TRI_COUNT_WAYS] triangles emitted (statistics only; spread over several counters) */
    uint32_t* tri_extra;  /* [n_frames] slots allocated beyond nfaces by the clipped path */
    uint32_t* tile_count; /* [n_frames][tile_pad], indexed through tile_slot(): see there */
    uint32_t* tile_offset; /* [n_frames][n_tiles] */
    uint32_t* tile_cursor; /* [n_frames][tile_pad], tile_slot() */
    uint32_t* tile_micro;  /* [n_frames][tile_pad], tile_slot(): != 0 where the visibility buffer holds fragments of the tile */
    unsigned long long* vis; /* optional [n_frames][H][W]: (depth bits << 32 | ~order key) of the best micro-triangle fragment
                                per pixel, all ones where there is none (dense meshes: see setup_kernel) */
    uint32_t* group_listed;  /* optional (passes with a visibility buffer) [n_frames][(tri_cap + 31) / 32]: != 0 where one of the 32
                                slots holds a triangle the pair kernels must list; a dense mesh is almost all micro-triangles,
                                and their warps leave at once instead of reading 32 pixel ranges to find nothing */
    int tile_pad, tile_rows; /* tile_pad = 32 * tile_rows >= n_tiles */
    float4* tile_recs;    /* pool of raster records (4 x float4 each), grouped per (frame, tile) */
    uint32_t pool_cap;    /* in records */
    uint4* work;          /* [n_frames*n_tiles] non-empty tiles: {item, count, offset, 0} */
    PassCounters* counters;
    OverflowRecord* overflow; /* optional */
    float* dbg_v2f;       /* optional: [tri_cap][39] post-clip v2f of frame 0 (stage tests) */
};

/* Where tile t of a frame keeps its counter / list cursor. The L2 serialises atomics that fall into the same line, and
 * the triangles of a warp touch neighbouring tiles, so neighbours are kept 32 "rows" apart instead of 4 bytes apart. */
__device__ __forceinline__ int tile_slot(const PassParams& p, int t) { return (t & 31) * p.tile_rows + (t >> 5); }

struct RasterParams {
    PassParams p;
    /* target */
    uint32_t* color;      /* RGBA8 as u32, (y*W+x), frame stride below */
    float* depth;
    size_t frame_stride;  /* pixels between consecutive frames */
    uint8_t* shadow_out;  /* MODE_SHADOW_R8: [frame][y*pitch + x] */
    int shadow_out_pitch;
    size_t shadow_out_frame_stride; /* bytes */
    uint32_t clear_color; /* RGBA packed */
    float clear_depth;
    int use_tma;
    /* shading inputs */
    DevTexture diffuse, normal;
    DevShadow shadow;            /* frame 0; further frames at + shadow_frame_stride bytes */
    size_t shadow_frame_stride;
    uint32_t* primid;            /* optional, frame 0 only: W*H order keys */
    uint32_t* pixels_covered;    /* optional: [n_frames] */
};

/* ---- small PTX wrappers: TMA (cp.async.bulk.tensor) + mbarrier ------------ */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm),
                 "r"(smem_u32(smem)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, void* smem, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

/* ---- begin: uniforms + counters ------------------------------------------- */
__global__ void begin_kernel(const HanaUniforms* __restrict__ in, DevUniforms* __restrict__ out, int n_frames) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_frames) prepare_uniforms(in[f], out[f]);
}

/* ---- vertex stage only (hana_stage_vertex) -------------------------------- */
__global__ void vertex_kernel(const float4* __restrict__ posu, const float4* __restrict__ nrmv, int ncorners, int shader,
                              const DevUniforms* __restrict__ u, float* __restrict__ out13) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncorners) return;
    float4 p = __ldg(posu + i), n = __ldg(nrmv + i); /* coalesced 128-bit loads over the SoA streams */
    float a[8] = {p.x, p.y, p.z, n.x, n.y, n.z, p.w, n.w};
    float v[V2F_N];
    vertex_shader(shader, *u, a, v);
    for (int k = 0; k < V2F_N; k++) out13[(size_t)i * V2F_N + k] = v[k];
}

/* ---- setup ------------------------------------------------------------------ */
/* attribute floats a shader's fragment() reads / float4 chunks of the per-triangle attribute block
 * (chunk 0 = the three 1/w of graphics.cpp:336, then 3 floats per attribute) */
template <int SHADER>
struct ShaderAttrs {
    static constexpr int NA = (SHADER == HANA_SHADER_SHADOW || SHADER == HANA_SHADER_GROUND || SHADER == HANA_SHADER_TOON)
                                  ? 1
                                  : (SHADER == HANA_SHADER_TEXTURE ? 2 : (SHADER == HANA_SHADER_TEXTURE_LIGHT ? 5 : 8));
    static constexpr int NQ = 1 + (3 * NA + 3) / 4;
    static constexpr bool LIT = (SHADER == HANA_SHADER_BLINN || SHADER == HANA_SHADER_NORMALMAP);
    static constexpr bool READS_UNIFORMS = LIT || SHADER == HANA_SHADER_TEXTURE_LIGHT; /* in fragment() */
};
constexpr int MAX_ATTR_QUADS = 7;

/* Raster record in memory, 4 x float4:
 *   q0 = ax, ay, s0x, s0y      q1 = s1x, s1y, uz, |uz| * 2^-24      (all the coverage test reads)
 *   q2 = bbx, bby, triangle index (attribute block), order key      q3 = d0, d1, d2, 1/uz   (covered pixels only) */
template <int SHADER>
__device__ __forceinline__ void store_triangle(const PassParams& p, int f, uint32_t slot, const TriRecord& r,
                                               const float* v0, const float* v1, const float* v2, bool with_range = true) {
    constexpr int NA = ShaderAttrs<SHADER>::NA;
    constexpr int NQ = ShaderAttrs<SHADER>::NQ;
    float4* dst = p.tri_rec + ((size_t)f * p.tri_cap + slot) * 4;
    dst[0] = make_float4(r.ax, r.ay, r.s0x, r.s0y);
    dst[1] = make_float4(r.s1x, r.s1y, r.uz, r.thr);
    dst[2] = make_float4(__uint_as_float(r.bbx), __uint_as_float(r.bby), __uint_as_float(slot), __uint_as_float(r.key));
    dst[3] = make_float4(r.d0, r.d1, r.d2, r.ruz);
    if (with_range) p.tri_bbox[(size_t)f * p.tri_cap + slot] = make_uint2(r.bbx, r.bby); /* a micro-triangle's slot gets the empty range instead */
    float a[(NQ - 1) * 4];
#pragma unroll
    for (int k = 0; k < NA; k++) {
        const int src = shader_attr_src(SHADER, k);
        if (ShaderAttrs<SHADER>::LIT) { /* pairwise: (v0.a, v0.b, v1.a, v1.b, v2.a, v2.b) per attribute pair (interp_lit_packed) */
            const int b = 6 * (k >> 1) + (k & 1);
            a[b] = v0[src];
            a[b + 2] = v1[src];
            a[b + 4] = v2[src];
        } else {
            a[3 * k + 0] = v0[src];
            a[3 * k + 1] = v1[src];
            a[3 * k + 2] = v2[src];
        }
    }
#pragma unroll
    for (int k = 3 * NA; k < (NQ - 1) * 4; k++) a[k] = 0.f;
    float4* ad = p.tri_attr + ((size_t)f * p.tri_cap + slot) * NQ;
    ad[0] = make_float4(r.rw0, r.rw1, r.rw2, 0.f);
#pragma unroll
    for (int q = 1; q < NQ; q++) ad[q] = make_float4(a[4 * q - 4], a[4 * q - 3], a[4 * q - 2], a[4 * q - 1]);
    if (p.dbg_v2f && f == 0) {
        float* d = p.dbg_v2f + (size_t)slot * 39;
        for (int k = 0; k < V2F_N; k++) {
            d[k] = v0[k];
            d[13 + k] = v1[k];
            d[26 + k] = v2[k];
        }
    }
}

/* A triangle whose pixel range lies inside ONE tile (of the pass's band) is counted right here; the others are left to
 * pairs_kernel<false>, which skips these (dense meshes are almost all single-tile triangles: BASELINE.json configs[3]). */
/* the counter slot of the single tile a triangle's pixel range lies in, or -1 */
__device__ __forceinline__ int single_tile_slot(const PassParams& p, const TriRecord& r) {
    const int tx0 = (int)(r.bbx & 0xFFFFu) >> 4, tx1 = (int)(r.bbx >> 16) >> 4;
    const int ty0 = max((int)(r.bby & 0xFFFFu) >> 4, p.band_y0), ty1 = min((int)(r.bby >> 16) >> 4, p.band_y1 - 1);
    return (tx0 == tx1 && ty0 == ty1) ? tile_slot(p, ty0 * p.tiles_x + tx0) : -1;
}
__device__ __forceinline__ void count_single_tile(const PassParams& p, int f, const TriRecord& r) {
    const int s = single_tile_slot(p, r);
    if (s >= 0) atomicAdd(p.tile_count + (size_t)f * p.tile_pad + s, 1u);
}
/* Warp-aggregated: the lanes of a warp that count into the same tile (neighbouring faces of a dense mesh) send ONE
 * atomic. Every lane of the warp calls this (key < 0: nothing to count). */
__device__ __forceinline__ void count_aggregated(uint32_t* counters, int key) {
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
    if (key >= 0 && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(counters + key, (uint32_t)__popc(peers));
}

/* Rare path: the face is not trivially accepted. Sutherland-Hodgman in local
 * memory, then the fan (0, j+1, j+2) of graphics.cpp:394-405. */
template <int SHADER>
__device__ __noinline__ void setup_clipped(const PassParams& p, int f, int face, const float* v39) {
    float poly[10 * V2F_N];
    for (int i = 0; i < 3 * V2F_N; i++) poly[i] = v39[i];
    int n = clip_polygon(poly);
    for (int j = 0; j + 2 < n; j++) {
        const float* a = poly;
        const float* b = poly + V2F_N * (j + 1);
        const float* c = poly + V2F_N * (j + 2);
        TriRecord r;
        if (!triangle_setup(a, b, c, p.W, p.H, (uint32_t)face * 8u + (uint32_t)j, r)) continue;
        const uint32_t slot = (uint32_t)p.nfaces + atomicAdd(p.tri_extra + f, 1u); /* a clipped face's triangles live behind the per-face slots */
        atomicAdd(p.tri_count + (size_t)f * TRI_COUNT_WAYS + (face & (TRI_COUNT_WAYS - 1)), 1u);
        if (slot < p.tri_cap) {
            store_triangle<SHADER>(p, f, slot, r, a, b, c);
            count_single_tile(p, f, r);
            if (p.group_listed) p.group_listed[(size_t)f * ((p.tri_cap + 31u) >> 5) + (slot >> 5)] = 1u;
        }
    }
}

template <int SHADER>
/* `p` is __grid_constant__: the rare clipped path takes its address (a __noinline__ call), and without the qualifier
 * every thread first copies all of PassParams to local memory. For the same reason the 39 vertex-stage floats are
 * copied to a second array only on that path, so the common path keeps them in registers. */
__global__ void __launch_bounds__(SETUP_THREADS, HANA_SETUP_CTAS) setup_kernel(const __grid_constant__ PassParams p) {
    const int f = blockIdx.y;
    const int face = blockIdx.x * SETUP_THREADS + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const DevUniforms& u = p.uniforms[f];
    bool emit = false;
    TriRecord r;
    float v[3 * V2F_N];
    if (face < p.nfaces) {
#pragma unroll
        for (int j = 0; j < 3; j++) { /* graphics.cpp:381-387 */
            float4 pu = __ldg(p.posu + (size_t)face * 3 + j);
            float4 nv = __ldg(p.nrmv + (size_t)face * 3 + j);
            float a[8] = {pu.x, pu.y, pu.z, nv.x, nv.y, nv.z, pu.w, nv.w};
            vertex_shader(SHADER, u, a, v + V2F_N * j);
        }
        if (clip_trivial_accept(v)) {
            emit = triangle_setup(v, v + V2F_N, v + 2 * V2F_N, p.W, p.H, (uint32_t)face * 8u, r);
        } else {
            float vc[3 * V2F_N];
#pragma unroll
            for (int k = 0; k < 3 * V2F_N; k++) vc[k] = v[k];
            setup_clipped<SHADER>(p, f, face, vc);
        }
    }
    /* The triangle of an unclipped face lives in slot = face index: no slot allocation, hence no atomic and no CTA
     * barrier (round 1 compacted the survivors with a block scan; its two barriers held 40 % of this kernel's stall
     * samples on a 10 M-face mesh). A face that emits nothing here (culled, degenerate, or clipped: its fan lives in
     * slots >= nfaces) marks its slot empty; the pair kernels skip it by its empty pixel range. */
    int key = -1;
    bool listed = false; /* the pair kernels have to look at this slot */
    if (face < p.nfaces && (uint32_t)face < p.tri_cap) {
        if (emit) {
            const bool micro = p.vis && (int)(r.bbx >> 16) - (int)(r.bbx & 0xFFFFu) < MICRO_EXTENT && (int)(r.bby >> 16) - (int)(r.bby & 0xFFFFu) < MICRO_EXTENT && r.uz < 0.f;
            store_triangle<SHADER>(p, f, (uint32_t)face, r, v, v + V2F_N, v + 2 * V2F_N, !micro);
            /* Micro-triangles (pixel range at most MICRO_EXTENT on a side: a dense mesh has millions, BASELINE.json configs[3]) do not go
             * through the tile lists, where the tile's warp would walk them one by one. Their pixels are tested right
             * here with the scalar statement of the coverage / weight / depth arithmetic (hana_core.cuh: the bits the tile
             * rasteriser's packed form produces) and resolved with one 64-bit atomicMin per covered pixel on the
             * visibility buffer: smallest depth first, then the largest order key (graphics.cpp:359 in submission order).
             * The tile rasteriser starts from that buffer and recomputes the winner's weights when it shades. */
            const int x0 = (int)(r.bbx & 0xFFFFu), x1 = (int)(r.bbx >> 16), y0 = (int)(r.bby & 0xFFFFu), y1 = (int)(r.bby >> 16);
            if (micro) {
                uint32_t touched = 0; /* tiles (at most 2 x 2: MICRO_EXTENT <= 16) that got a fragment: bit (ty - ty0) * 2 + (tx - tx0) */
                const int tx_o = x0 >> 4, ty_o = y0 >> 4;
                for (int y = y0; y <= y1; y++) {
                    if ((y >> 4) < p.band_y0 || (y >> 4) >= p.band_y1) continue;
                    for (int x = x0; x <= x1; x++) {
                        float ux, uy, su;
                        if (!coverage_test(r.ax, r.ay, r.s0x, r.s0y, r.s1x, r.s1y, r.uz, r.thr, (float)x, (float)y, ux, uy, su)) continue;
                        float w0, w1, w2;
                        barycentric_weights(ux, uy, su, r.uz, r.ruz, w0, w1, w2);
                        const float z = interpolate_depth(r.d0, r.d1, r.d2, w0, w1, w2);
                        if (!(z == z)) continue; /* a NaN depth never wins */
                        atomicMin(p.vis + ((size_t)f * p.H + y) * p.W + x,
                                  ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(0xFFFFFFFFu - r.key));
                        touched |= 1u << ((((y >> 4) - ty_o) << 1) | ((x >> 4) - tx_o));
                    }
                }
                /* one flag store per touched tile and triangle, behind the loop: nothing in the loop waits for memory */
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((touched >> k) & 1u) p.tile_micro[(size_t)f * p.tile_pad + tile_slot(p, (ty_o + (k >> 1)) * p.tiles_x + tx_o + (k & 1))] = 1u;
                /* not listed: the pair kernels skip a record whose pixel range is empty */
                p.tri_bbox[(size_t)f * p.tri_cap + (uint32_t)face] = make_uint2(0u, DEAD_BBY);
            } else {
                key = single_tile_slot(p, r);
                listed = true;
            }
        } else {
            p.tri_bbox[(size_t)f * p.tri_cap + (uint32_t)face] = make_uint2(0u, DEAD_BBY);
        }
    }
    if (p.group_listed) { /* warp-uniform */
        const unsigned any = __ballot_sync(0xFFFFFFFFu, listed);
        if (lane == 0 && any && (uint32_t)face < p.tri_cap) p.group_listed[(size_t)f * ((p.tri_cap + 31u) >> 5) + ((uint32_t)face >> 5)] = 1u;
    }
    const unsigned live = __ballot_sync(0xFFFFFFFFu, emit);
    if (lane == 0 && live) atomicAdd(p.tri_count + (size_t)f * TRI_COUNT_WAYS + (blockIdx.x & (TRI_COUNT_WAYS - 1)), (uint32_t)__popc(live));
    if (__ballot_sync(0xFFFFFFFFu, key >= 0)) count_aggregated(p.tile_count + (size_t)f * p.tile_pad, key); /* warp-uniform */
}

/* ---- scan: per frame, tile counts -> offsets into the pool + work queue ---- */
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t val, uint32_t* total, uint32_t* warp_sums /* [32] */) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t incl = val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if ((int)lane >= d) wi += t;
        }
        warp_sums[lane] = wi - w; /* exclusive */
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + incl - val;
    __syncthreads();
    return r;
}

/* One CTA per chunk of SCAN_CHUNK consecutive tiles of a frame (4 per thread): chunk totals by a block scan, ONE
 * reservation of pool records and of work-queue entries per chunk, then the offsets. Lists need no particular order
 * in the pool and the queue none among its entries, so chunks do not wait for each other. */
constexpr int SCAN_CHUNK = SCAN_THREADS * 4;
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(PassParams p) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_total_refs, s_total_ne, s_base_refs, s_base_work;
    const int f = blockIdx.y;
    const int t0 = blockIdx.x * SCAN_CHUNK + (int)threadIdx.x * 4;
    const uint32_t* tc = p.tile_count + (size_t)f * p.tile_pad;
    const uint32_t* tm = p.tile_micro + (size_t)f * p.tile_pad;
    uint32_t c[4], m[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        c[k] = (t0 + k < p.n_tiles) ? tc[tile_slot(p, t0 + k)] : 0u;
        m[k] = (p.vis && t0 + k < p.n_tiles) ? tm[tile_slot(p, t0 + k)] : 0u; /* the visibility buffer holds fragments of the tile */
    }
    const uint32_t refs = c[0] + c[1] + c[2] + c[3];
    const uint32_t ne = ((c[0] | m[0]) != 0u) + ((c[1] | m[1]) != 0u) + ((c[2] | m[2]) != 0u) + ((c[3] | m[3]) != 0u);
    const uint32_t ex_refs = block_exclusive_scan(refs, &s_total_refs, warp_sums);
    const uint32_t ex_ne = block_exclusive_scan(ne, &s_total_ne, warp_sums);
    if (threadIdx.x == 0) {
        s_base_refs = atomicAdd(&p.counters->pool_used, s_total_refs);
        s_base_work = atomicAdd(&p.counters->n_work, s_total_ne);
        atomicAdd(&p.counters->tiles_touched, s_total_ne);
        if (p.overflow) atomicMax(&p.overflow->pool_needed, s_base_refs + s_total_refs);
        if (blockIdx.x == 0) {
            const uint32_t slots = (uint32_t)p.nfaces + p.tri_extra[f];
            atomicMax(&p.counters->tri_needed, slots);
            if (p.overflow) atomicMax(&p.overflow->tri_needed, slots);
        }
    }
    __syncthreads();
    uint32_t off = s_base_refs + ex_refs;
    uint32_t wi = s_base_work + ex_ne;
    uint32_t* to = p.tile_offset + (size_t)f * p.n_tiles;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int t = t0 + k;
        if (t < p.n_tiles) {
            to[t] = off;
            if (c[k] | m[k]) {
                const uint32_t tx = (uint32_t)(t % p.tiles_x), ty = (uint32_t)(t / p.tiles_x);
                p.work[wi++] = make_uint4(((uint32_t)f << TILE_BITS) | (ty << 10) | tx, c[k], off, m[k] ? 1u : 0u);
            }
            off += c[k];
        }
    }
}

/* ---- count / fill: the (triangle, tile) pairs ------------------------------------
 * pairs_kernel<false> counts, for every 16x16 tile, the triangles that can touch it (tile_may_touch over the tiles of
 * the triangle's pixel range, clamped to the pass's band); pairs_kernel<true>, after the scan, copies each raster
 * record into those tiles' lists. Both enumerate the pairs the same way: a lane owns one triangle, but the pairs of the
 * whole warp are flattened by a shuffle prefix sum and dealt out round-robin, so the atomics of a round are independent
 * (a triangle covering 100 tiles costs its warp 4 rounds instead of one lane 100 dependent ones); gridDim.z warps share
 * the rounds of the same 32 triangles, which is what keeps the SMs busy when a few thousand triangles cover hundreds of
 * tiles each (BASELINE.json configs[4]). */
template <bool FILL>
__global__ void __launch_bounds__(256) pairs_kernel(PassParams p) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int f = blockIdx.y;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t n = (uint32_t)p.nfaces + p.tri_extra[f];      /* slots in use */
    if (n > p.tri_cap) n = p.tri_cap;
    if ((i & ~31u) >= n) return;                            /* whole warp past the end */
    if (p.group_listed && p.group_listed[(size_t)f * ((p.tri_cap + 31u) >> 5) + (i >> 5)] == 0u) return; /* nothing to list in these 32 slots */
    if (FILL && p.counters->pool_used > p.pool_cap) return; /* host re-runs the pass with a larger pool */
    const float4* warp_rec = p.tri_rec + ((size_t)f * p.tri_cap + (i & ~31u)) * 4;
    int tx0 = 0, ty0 = 0, ntx = 1, nt = 0;
    if (i < n) {
        const uint2 bb = __ldg(p.tri_bbox + (size_t)f * p.tri_cap + i);
        const uint32_t bbx = bb.x, bby = bb.y;
        tx0 = (int)(bbx & 0xFFFFu) >> 4;
        ty0 = max((int)(bby & 0xFFFFu) >> 4, p.band_y0);
        ntx = ((int)(bbx >> 16) >> 4) - tx0 + 1;
        nt = ntx * max(min((int)(bby >> 16) >> 4, p.band_y1 - 1) - ty0 + 1, 0);
        if (!FILL && nt == 1) nt = 0; /* counted by setup_kernel (count_single_tile) */
    }
    int incl = nt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, d);
        if ((int)lane >= d) incl += t;
    }
    const int excl = incl - nt;
    const int total = __shfl_sync(FULL, incl, 31);
    uint32_t* cnt = p.tile_count + (size_t)f * p.tile_pad;
    uint32_t* cur = p.tile_cursor + (size_t)f * p.tile_pad;
    const uint32_t* to = p.tile_offset + (size_t)f * p.n_tiles;
    for (int base = (int)blockIdx.z * 32; base < total; base += 32 * (int)gridDim.z) {
        const int k = base + (int)lane;
        /* owner of pair k: the last lane whose exclusive prefix is <= k (it has nt > 0 whenever k < total) */
        int j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int e = __shfl_sync(FULL, excl, (j + step) & 31);
            if (e <= k) j += step;
        }
        const int local = k - __shfl_sync(FULL, excl, j);
        const int jtx0 = __shfl_sync(FULL, tx0, j), jty0 = __shfl_sync(FULL, ty0, j), jntx = __shfl_sync(FULL, ntx, j);
        const bool single = __shfl_sync(FULL, nt, j) == 1; /* a range inside one tile: the triangle touches it or covers nothing */
        int key = -1, t = 0; /* key: counter slot of the pair's tile if the triangle can touch it */
        float4 q0, q1;
        const float4* rec = warp_rec + j * 4;
        if (k < total) {
            q0 = __ldg(rec);
            q1 = __ldg(rec + 1);
            const int row = local / jntx;
            const int tx = jtx0 + (local - row * jntx), ty = jty0 + row;
            t = ty * p.tiles_x + tx;
            /* the same operands in both instantiations: the count and the fill agree */
            if (single || tile_may_touch(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, (float)(tx * TILE), (float)(ty * TILE)))
                key = tile_slot(p, t);
        }
        if (!FILL) {
            if (key >= 0) atomicAdd(cnt + key, 1u); /* pairs of one round are mostly different tiles: not worth aggregating */
        } else {
            /* the lanes of a round that append to the same list (neighbouring small triangles) reserve their slots with ONE atomic */
            const unsigned peers = __match_any_sync(FULL, key);
            const int leader = __ffs(peers) - 1;
            uint32_t s = 0;
            if (key >= 0 && (int)lane == leader) s = atomicAdd(cur + key, (uint32_t)__popc(peers));
            s = __shfl_sync(FULL, s, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (key >= 0) {
                const float4 q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
                float4* d = p.tile_recs + ((size_t)__ldg(to + t) + s) * 4;
                d[0] = q0;
                d[1] = q1;
                d[2] = q2;
                d[3] = q3;
            }
        }
    }
}

/* ---- raster ------------------------------------------------------------------
 * One WARP per 16x16 tile, no CTA-wide barrier anywhere in the tile loop. A lane owns 8 pixels of the
 * tile: pixel (lane & 7, lane >> 3) of each of the eight 8x4 sub-blocks (2 across, 4 down; sub-block
 * sb = row quarter * 2 + column half). The tile's records are staged 32 at a time in the warp's
 * private shared memory by lane j for record j, which also computes which sub-blocks the record's pixel
 * range meets, so the triangle loop is warp-uniform and skips rows of sub-blocks without a per-pixel
 * instruction.
 *
 * Arithmetic is packed two pixels at a time (hana_pack.cuh): the two pixels of a lane that share a column
 * half and sit in row quarters 2k and 2k+1 go through every multiply / add of graphics.cpp:222-233 and
 * :186-194 as the two halves of one FFMA2, with the record's scalars as broadcast operands, so one
 * instruction covers an 8x8 block of the tile (pairing the two column halves of a row instead, 16x4 pixels per
 * instruction, measured 6 % slower in the shadow pass: small triangles leave more of a long thin block
 * empty). Per record: 3 packed operations for the x-dependent terms, 3 per pair of row quarters for the
 * y-dependent ones, then 5 per 8x8 block and one 3-input max + one compare per pixel for the inside test;
 * covered pixels take 14 more packed operations for the exact weights and the depth. The (min depth, max key) resolve lives in registers as
 * {depth, triangle slot} per pixel; the three weights of a fragment that wins are parked in shared memory
 * (they cost three stores where the fragment wins, which happens ~1.01 times per visible pixel), so the
 * shading stage neither re-fetches a record nor recomputes a weight. The warp's tile is flushed with TMA
 * stores issued by lane 0; warps never wait for each other.
 *
 * Every kept triangle has u.z < 0 (u.z = -2 x the screen area of a counter-clockwise triangle, and
 * graphics.cpp:320 culled the others), except slivers whose two area computations round to different
 * signs. Those are staged with B and C exchanged in the coverage constants, which negates u.x<->u.y, their
 * sum, u.z and the threshold difference exactly, so one code path (the u.z < 0 one) serves every record;
 * their two quotients swap places again when the weights are formed. */
constexpr int RW_WARPS = 4;
#ifndef HANA_OCC_LIT
#define HANA_OCC_LIT 6 /* resident CTAs per SM the lit shaders' rasteriser is compiled for (register cap 80) */
#endif
#ifndef HANA_OCC_OTHER
#define HANA_OCC_OTHER 7
#endif
constexpr int RW_THREADS = RW_WARPS * 32;
#ifndef HANA_RW_CHUNK
#define HANA_RW_CHUNK 16 /* records staged per round (and attribute blocks staged per tile): 5 + NQ float4 of shared memory each */
#endif
constexpr int RW_CHUNK = HANA_RW_CHUNK;
constexpr int RW_REC_Q = 5; /* float4 per staged record */
constexpr uint32_t ORD_NONE = 0xFFFFFFFFu;

/* Staged record (shared memory), 5 x float4:
 *   q0 = ax, ay, s0x, s0y         q1 = s1x, s1y, uz (< 0), thr
 *   q2 = bbox centre x, centre y, half extent x, half extent y   (exact: sums of two 16-bit integers, halved)
 *   q3 = d0, d1, d2, 1/uz         q4 = triangle slot, order key, B/C exchanged flag, sub-block mask */
template <int MODE, int NQ>
struct alignas(128) WarpTile {
    /* TMA sources/destinations first: each 128-byte aligned */
    uint32_t color[mode_is_r8(MODE) ? 32 : TILE_PIX]; /* box 16x16 u32; CLEAR_FOLD: holds the parked w0 of a pixel until it is shaded */
    float depth[mode_is_r8(MODE) ? 32 : TILE_PIX];
    uint8_t r8[mode_is_r8(MODE) ? TILE_PIX : 128];    /* box 16x16 u8 */
    uint32_t ord[mode_is_r8(MODE) ? 32 : TILE_PIX];   /* winning triangle slot per pixel, parked for shading */
    float pw1[mode_is_r8(MODE) ? 32 : TILE_PIX];      /* parked weights of the winning fragment */
    float pw2[mode_is_r8(MODE) ? 32 : TILE_PIX];
    float pw0_own[MODE == MODE_RMW ? TILE_PIX : 32];        /* RMW: color[] holds the target's pixels */
    float4 tri[RW_CHUNK * RW_REC_Q];                        /* staged raster records */
    float4 sattr[mode_is_r8(MODE) ? RW_CHUNK * 2 : RW_CHUNK * NQ]; /* SHADOW_R8: the records' (1/w, clip z) blocks; otherwise, for a
                                                                      tile whose list fits one chunk, the records' attribute blocks
                                                                      (NQ float4 each): the shading stage reads them with LDS instead
                                                                      of one dependent L2 round trip per sub-block */
    FragUniforms fu;                                        /* the tile's frame: what fragment() reads */
    alignas(8) uint64_t bar;
};
template <int MODE, int NQ>
struct alignas(128) RasterSmem {
    WarpTile<MODE, NQ> w[RW_WARPS];
};

/* Clear duty. Tiles no triangle touches are never rasterised, so somebody has to give them the clear values (the clear
 * is folded into the pass: one write per pixel, no separate pass over the target). Work item i of a pass = (frame, tile
 * row, group of CLEAR_GROUP consecutive tiles of that row); a warp takes one item per tile it rasterises until the queue
 * is empty. The lanes read the group's tile counts, and the empty tiles are written with plain 128-bit stores, the
 * whole warp side by side: lane = (tile of the group, 16-byte quarter of a pixel row) for the 4-byte planes, lane = tile
 * for the 1-byte shadow maps — 32 (16) store instructions per item. (Round 1 issued one TMA box store per empty tile
 * and plane from a constant shared-memory tile; UTMASTG takes its operands from uniform registers, so the compiler
 * serialised the warp lane by lane: ~24 instructions per store against 4 per tile here.) */
template <int MODE>
struct ClearGeom {
    static constexpr int GROUP = mode_is_r8(MODE) ? 32 : 8;
};
template <int MODE>
__device__ __forceinline__ void clear_item(const RasterParams& q, uint32_t item, int groups_x) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int G = ClearGeom<MODE>::GROUP;
    const PassParams& p = q.p;
    const unsigned lane = threadIdx.x & 31u;
    const int gx = (int)(item % (uint32_t)groups_x);
    const uint32_t r = item / (uint32_t)groups_x;
    const int ty = (int)(r % (uint32_t)p.tiles_y), f = (int)(r / (uint32_t)p.tiles_y);
    if (ty < p.band_y0 || ty >= p.band_y1) return; /* warp-uniform */
    const int tx_l = gx * G + (int)lane;
    bool empty = false;
    if ((int)lane < G && tx_l < p.tiles_x) {
        const size_t ts = (size_t)f * p.tile_pad + tile_slot(p, ty * p.tiles_x + tx_l);
        empty = p.tile_count[ts] == 0u && (!p.vis || p.tile_micro[ts] == 0u);
    }
    const unsigned mask = __ballot_sync(FULL, empty);
    if (!mask) return;
    if (mode_is_r8(MODE)) {
        /* rows and columns beyond the frame exist in the padded maps (pitch and height are multiples of 16) */
        if (empty) {
            uint8_t* dst = q.shadow_out + (size_t)f * q.shadow_out_frame_stride + (size_t)(ty * TILE) * q.shadow_out_pitch + (size_t)tx_l * TILE;
#pragma unroll
            for (int yy = 0; yy < TILE; yy++) *reinterpret_cast<uint4*>(dst + (size_t)yy * q.shadow_out_pitch) = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        const int t8 = (int)(lane >> 2), px = (gx * G + t8) * TILE + (int)(lane & 3u) * 4;
        if (((mask >> t8) & 1u) && px < p.W) {
            const int rows = min(TILE, p.H - ty * TILE);
            const size_t o = (size_t)f * q.frame_stride + (size_t)(ty * TILE) * p.W + px;
            if ((p.W & 3) == 0) { /* rows are 16-byte aligned */
                const uint4 cc = make_uint4(q.clear_color, q.clear_color, q.clear_color, q.clear_color);
                const float4 dd = make_float4(q.clear_depth, q.clear_depth, q.clear_depth, q.clear_depth);
                for (int yy = 0; yy < rows; yy++) {
                    *reinterpret_cast<uint4*>(q.color + o + (size_t)yy * p.W) = cc;
                    *reinterpret_cast<float4*>(q.depth + o + (size_t)yy * p.W) = dd;
                }
            } else {
                const int nx = min(4, p.W - px);
                for (int yy = 0; yy < rows; yy++)
                    for (int xx = 0; xx < nx; xx++) {
                        q.color[o + (size_t)yy * p.W + xx] = q.clear_color;
                        q.depth[o + (size_t)yy * p.W + xx] = q.clear_depth;
                    }
            }
        }
    }
}

/* word with byte K replaced by the low byte of b (one PRMT) */
template <int K>
__device__ __forceinline__ uint32_t put_byte(uint32_t word, uint32_t b) {
    return __byte_perm(word, b, K == 0 ? 0x3214 : (K == 1 ? 0x3240 : (K == 2 ? 0x3410 : 0x4210)));
}

/* atomicAdd whose result is NOT needed yet. nvcc turns an atomicAdd under `if (lane == 0)` into its warp-aggregated
 * form (vote + one atomic + a shuffle that reads the result at once), which stalls the warp for the whole L2 round trip
 * and defeats the software pipelining of the work queue below; the PTX form is left alone. */
__device__ __forceinline__ uint32_t atomic_add_async(uint32_t* addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ uint4 fetch_work(const PassParams& p, uint32_t idx, uint32_t n_work) {
    if (idx < n_work) return __ldg(p.work + idx);
    return make_uint4(WORK_INVALID, 0u, 0u, 0u);
}

/* Sub-blocks (bit r*2+c: column half c, row quarter r) of the tile at (X0,Y0) that the pixel range meets.
 * The range is known to meet the tile (that is why the record is in this tile's list). */
__device__ __forceinline__ uint32_t subblock_mask(int x0, int x1, int y0, int y1, int X0, int Y0) {
    const int cx0 = max(x0 - X0, 0) >> 3, cx1 = min(x1 - X0, TILE - 1) >> 3;
    const int ry0 = max(y0 - Y0, 0) >> 2, ry1 = min(y1 - Y0, TILE - 1) >> 2;
    const uint32_t cm = (2u << cx1) - (1u << cx0);        /* 2 bits */
    const uint32_t rm = (2u << ry1) - (1u << ry0);        /* 4 bits */
    const uint32_t spread = (rm & 1u) | ((rm & 2u) << 1) | ((rm & 4u) << 2) | ((rm & 8u) << 3);
    return cm * spread;
}

/* graphics.cpp:362-373 for one fragment: interpolate the attributes the shader reads, run fragment(), return the R,G,B
 * bytes set_color stores (renderbuffer.cpp:38-44). The fragment stage's uniforms arrive as 128-bit loads. FAST: sqrt /
 * reciprocal fast paths with one shared range check (qsqrt); *bad tells the caller that an operand left their range. */
template <int SHADER, bool FAST, class Src>
__device__ __forceinline__ uint32_t shade_fragment_t(const FragUniforms& fu, const Src ap, float w0, float w1, float w2,
                                                     const DevTexture& diffuse, const DevTexture& normal, const DevShadow& sh,
                                                     const void* safe, bool& bad) {
    constexpr int NA = ShaderAttrs<SHADER>::NA;
    constexpr int NQ = ShaderAttrs<SHADER>::NQ;
    bad = false;
    bool* pb = FAST ? &bad : nullptr;
    if (ShaderAttrs<SHADER>::LIT) {
        const LitAttrs la = interp_lit_packed(ap, w0, w1, w2, pb);
        return fragment_lit_packed<SHADER>(fu, la, diffuse, normal, sh, safe, pb);
    }
    const float4 rw = ap.load(0);
    VaryingWeights vw = varying_weights(w0, w1, w2, rw.x, rw.y, rw.z, pb);
    float a[(NQ - 1) * 4];
#pragma unroll
    for (int k = 0; k < NQ - 1; k++) {
        float4 v = ap.load(1 + k);
        a[4 * k] = v.x;
        a[4 * k + 1] = v.y;
        a[4 * k + 2] = v.z;
        a[4 * k + 3] = v.w;
    }
    float attr[NA];
#pragma unroll
    for (int k = 0; k < NA; k++) attr[k] = interp(vw, a[3 * k], a[3 * k + 1], a[3 * k + 2]);
    float rgb[3];
    fragment_shader<SHADER>(fu, attr, diffuse, normal, sh, rgb, pb);
    return colour_bytes(rgb);
}
/* The rare re-evaluation with the full sqrt / reciprocal functions. */
template <int SHADER>
__device__ __noinline__ uint32_t shade_fragment_exact(const FragUniforms* fu, const float4* ap, float w0, float w1, float w2,
                                                      const DevTexture* diffuse, const DevTexture* normal, const DevShadow* sh,
                                                      const void* safe) {
    bool bad;
    return shade_fragment_t<SHADER, false>(*fu, AttrGeneric{ap}, w0, w1, w2, *diffuse, *normal, *sh, safe, bad);
}
/* safe: any readable global address (shadow_probe loads from it unconditionally when there is no texel to fetch) */
template <int SHADER, class Src>
__device__ __forceinline__ uint32_t shade_fragment(const FragUniforms& fu, const Src ap, float w0, float w1, float w2,
                                                   const DevTexture& diffuse, const DevTexture& normal, const DevShadow& sh,
                                                   const void* safe) {
    bool bad;
    if (ShaderAttrs<SHADER>::LIT) { /* several sqrt/reciprocal sites: worth the shared range check */
        const uint32_t c = shade_fragment_t<SHADER, true>(fu, ap, w0, w1, w2, diffuse, normal, sh, safe, bad);
        if (bad) return shade_fragment_exact<SHADER>(&fu, ap.generic(), w0, w1, w2, &diffuse, &normal, &sh, safe); /* an operand left the fast paths' range */
        return c;
    }
    return shade_fragment_t<SHADER, false>(fu, ap, w0, w1, w2, diffuse, normal, sh, safe, bad);
}

/* VIS: the pass has a visibility buffer (micro-triangle path of dense meshes); a variant of its own so that the code and
 * registers it takes stay out of the kernels every other pass runs */
template <int SHADER, int MODE, bool VIS>
__global__ void __launch_bounds__(RW_THREADS, ShaderAttrs<SHADER>::LIT ? HANA_OCC_LIT : HANA_OCC_OTHER)
    raster_kernel(const __grid_constant__ RasterParams q, const __grid_constant__ CUtensorMap tm_color, const __grid_constant__ CUtensorMap tm_depth,
                  const __grid_constant__ CUtensorMap tm_r8) {
    constexpr int NQ = ShaderAttrs<SHADER>::NQ;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    /* SHADOW_R8 (always the ShadowShader): its one-attribute fragment() is evaluated right where a fragment wins the
     * resolve, so nothing is parked or re-fetched afterwards. The byte rides in the top 8 bits of the state word, which
     * leaves 24 bits for the triangle slot: run_pass routes a pass whose per-frame triangle capacity exceeds
     * R8_SLOT_LIMIT to the WIDE variant, which keeps the eight bytes of a lane packed in two registers of their own
     * (replaced with one PRMT per win) and the slot at 32 bits — two more registers, ~3 % slower, no limit. */
    constexpr bool INLOOP = mode_is_r8(MODE);
    constexpr bool WIDE = (MODE == MODE_SHADOW_R8_WIDE);
    constexpr uint32_t SLOT_MASK = (INLOOP && !WIDE) ? 0x00FFFFFFu : 0xFFFFFFFFu;
    __shared__ RasterSmem<MODE, NQ> sm;
    const PassParams& p = q.p;
    const unsigned lane = threadIdx.x & 31u, wid = __shfl_sync(FULL, threadIdx.x >> 5, 0); /* the broadcast lets ptxas keep the warp's shared-memory base in a uniform register instead of rebuilding it from S2R in the record loop */
    WarpTile<MODE, NQ>& wt = sm.w[wid];
    const int lx = (int)(lane & 7u), ly = (int)(lane >> 3);
    const bool tma = q.use_tma != 0;
    float* const pw0 = MODE == MODE_RMW ? wt.pw0_own : reinterpret_cast<float*>(wt.color);

    if (p.counters->pool_used > p.pool_cap) return; /* lists are incomplete: host re-runs with a larger pool */
    const uint32_t n_work = p.counters->n_work;
    const int clear_groups_x = (p.tiles_x + ClearGeom<MODE>::GROUP - 1) / ClearGeom<MODE>::GROUP;
    const uint32_t n_clear = (uint32_t)p.n_frames * (uint32_t)p.tiles_y * (uint32_t)clear_groups_x; /* clear work items */

    if (MODE == MODE_RMW && lane == 0) {
        mbar_init(&wt.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads(); /* the only CTA-wide barrier */

    /* Work queue, software-pipelined on lane 0: the atomic for item k+2 and the entry load for item k+1 are in
     * flight while item k is processed and are consumed only at the end of it, so neither latency is exposed. */
    uint32_t i_next = 0;
    uint4 e_cur = make_uint4(WORK_INVALID, 0u, 0u, 0u);
    if (lane == 0) {
        /* the first two items of every warp are its own by position (the queue's cursor counts from behind them): no 3 552
         * atomics on one address before the first tile of a launch (-0.035 ms per 1024-frame step) */
        const uint32_t i0 = 2u * (blockIdx.x * RW_WARPS + wid);
        e_cur = fetch_work(p, i0, n_work);
        i_next = i0 + 1u;
    }
    bool clear_done = (MODE == MODE_RMW);
    uint32_t load_phase = 0;

    while (true) {
        uint4 cur;
        cur.x = __shfl_sync(FULL, e_cur.x, 0);
        if (cur.x == WORK_INVALID) break;
        cur.y = __shfl_sync(FULL, e_cur.y, 0);
        cur.z = __shfl_sync(FULL, e_cur.z, 0);
        /* the tile has fragments in the visibility buffer (only passes that were given one: warp-uniform and false otherwise) */
        const bool has_micro = VIS && MODE != MODE_RMW && __shfl_sync(FULL, e_cur.w, 0) != 0u;
        __syncwarp(); /* every lane is done with the previous tile's shared memory (fu, ord, tri) */
        uint4 e_nxt = make_uint4(WORK_INVALID, 0u, 0u, 0u);
        uint32_t i_nn = 0, clr_base = 0;
        if (lane == 0) {
            e_nxt = fetch_work(p, i_next, n_work);
            i_nn = atomic_add_async(&p.counters->work_cursor, 1u) + 2u * gridDim.x * RW_WARPS;
            if (MODE != MODE_RMW && !clear_done) clr_base = atomic_add_async(&p.counters->clear_cursor, 1u);
        }

        /* -- one non-empty tile -- */
        const int f = (int)(cur.x >> TILE_BITS);
        const int tx = (int)(cur.x & 1023u), ty = (int)((cur.x >> 10) & 1023u);
        const int X0 = tx * TILE, Y0 = ty * TILE;
        const uint32_t cnt = cur.y;
        const float4* list = p.tile_recs + (size_t)cur.z * 4;
        const float4* frame_rec = p.tri_rec + (size_t)f * p.tri_cap * 4; /* the frame's records by slot: tie-breaks, primitive ids */
        const float4* frame_attr = p.tri_attr + (size_t)f * p.tri_cap * NQ; /* ... and their attribute blocks */
        if (ShaderAttrs<SHADER>::READS_UNIFORMS && lane < sizeof(FragUniforms) / 16) /* read again only after the __syncwarp()s of the record loop */
            reinterpret_cast<float4*>(&wt.fu)[lane] = __ldg(reinterpret_cast<const float4*>(&p.uniforms[f].frag) + lane);
        const float fpx0 = (float)(X0 + lx), fpx1 = (float)(X0 + 8 + lx);
        const float fpy0 = (float)(Y0 + ly), fpy1 = (float)(Y0 + 4 + ly), fpy2 = (float)(Y0 + 8 + ly), fpy3 = (float)(Y0 + 12 + ly);
        const f2 FPX = f2_make(fpx0, fpx1), FPY01 = f2_make(fpy0, fpy1), FPY23 = f2_make(fpy2, fpy3);
        const int ipx0 = X0 + lx, ipy0 = Y0 + ly;
        int pix0 = ly * TILE + lx; /* the lane's pixel in sub-block 0; sub-block sb adds (sb >> 1) * 64 + (sb & 1) * 8 */
        asm volatile("" : "+r"(pix0)); /* opaque: ptxas otherwise recomputes it from %tid (S2R + four ALU ops) at every parked win */

        float bz[8];
        uint32_t bj[8];
        uint32_t bb[2] = {0u, 0u}; /* WIDE: the ShadowShader byte of sub-block sb in byte sb & 3 of bb[sb >> 2]; 0 where nothing was drawn */
#pragma unroll
        for (int sb = 0; sb < 8; sb++) {
            bz[sb] = q.clear_depth;
            bj[sb] = ORD_NONE;
        }
        if (MODE == MODE_RMW) {
            if (tma) {
                if (lane == 0) {
                    tma_wait_read0(); /* the previous tile's stores have read color/depth */
                    mbar_expect_tx(&wt.bar, 2u * TILE_PIX * 4u);
                    tma_load_3d(&tm_color, wt.color, &wt.bar, X0, Y0, f);
                    tma_load_3d(&tm_depth, wt.depth, &wt.bar, X0, Y0, f);
                }
                mbar_wait(&wt.bar, load_phase);
                load_phase ^= 1u;
#pragma unroll
                for (int sb = 0; sb < 8; sb++) bz[sb] = wt.depth[pix0 + (sb >> 1) * 64 + (sb & 1) * 8];
            } else {
#pragma unroll
                for (int sb = 0; sb < 8; sb++) {
                    const int px = ipx0 + (sb & 1) * 8, py = ipy0 + (sb >> 1) * 4;
                    if (px < p.W && py < p.H) bz[sb] = q.depth[(size_t)f * q.frame_stride + (size_t)py * p.W + px];
                }
            }
        } else if (MODE == MODE_CLEAR_FOLD && tma) {
            if (lane == 0) tma_wait_read0(); /* the previous tile's stores have drained the staging tile: weights are parked in it below */
            __syncwarp();
        }

        /* Micro-triangles were resolved per pixel by setup_kernel: the resolve starts from their best fragment. Their slot is
         * their face index (unclipped faces only), i.e. order key >> 3. */
        const unsigned long long* vis_frame = has_micro ? p.vis + (size_t)f * p.H * p.W : nullptr;
        if (has_micro) {
#pragma unroll
            for (int sb = 0; sb < 8; sb++) {
                const int px = ipx0 + (sb & 1) * 8, py = ipy0 + (sb >> 1) * 4;
                if (px < p.W && py < p.H) {
                    const unsigned long long vv = __ldg(vis_frame + (size_t)py * p.W + px);
                    if (vv != ~0ull) {
                        bz[sb] = __uint_as_float((uint32_t)(vv >> 32));
                        bj[sb] = (0xFFFFFFFFu - (uint32_t)vv) >> 3;
                    }
                }
            }
        }

        /* Equal depths (rare): graphics.cpp:359 is "skip if z > stored" in submission order, so a fragment as deep as the
         * target's value passes (LEQUAL) and of two equally deep fragments the later submission wins. */
        /* A tile whose whole list fits one chunk (the common case) parks the record's index in the chunk instead of the
         * triangle's slot in the frame: its staged record (order key) and staged attribute block are then one LDS away. */
        const bool single = !INLOOP && cnt <= (uint32_t)RW_CHUNK && !has_micro; /* warp-uniform; visibility-buffer winners are known by their slot in the frame */
        auto wins_tie = [&](const int sb, const uint32_t key) -> bool {
            if (bj[sb] == ORD_NONE) return true;
            const uint32_t kb = single ? __float_as_uint(wt.tri[bj[sb] * RW_REC_Q + 4].y)
                                       : __float_as_uint(__ldg(frame_rec + (size_t)(bj[sb] & SLOT_MASK) * 4 + 2).w);
            return key > kb;
        };

        for (uint32_t c0 = 0; c0 < cnt; c0 += RW_CHUNK) {
            const int n = (int)min((uint32_t)RW_CHUNK, cnt - c0);
            __syncwarp(); /* previous chunk fully consumed */
            if ((int)lane < n) { /* lane j stages record j */
                const float4* rec = list + ((size_t)c0 + lane) * 4;
                float4 a0 = __ldg(rec), a1 = __ldg(rec + 1);
                const float4 a2 = __ldg(rec + 2), a3 = __ldg(rec + 3);
                const uint32_t bbx = __float_as_uint(a2.x), bby = __float_as_uint(a2.y);
                const int x0 = (int)(bbx & 0xFFFFu), x1 = (int)(bbx >> 16), y0 = (int)(bby & 0xFFFFu), y1 = (int)(bby >> 16);
                const uint32_t mask = subblock_mask(x0, x1, y0, y1, X0, Y0);
                float swapped = 0.f;
                if (a1.z > 0.f) { /* sliver whose u.z rounded positive: exchange B and C in the coverage constants */
                    a0 = make_float4(a0.x, a0.y, a0.w, a0.z);
                    a1 = make_float4(a1.y, a1.x, -a1.z, a1.w);
                    swapped = 1.f;
                }
                float4* dst = wt.tri + lane * RW_REC_Q;
                dst[0] = a0;
                dst[1] = a1;
                dst[2] = make_float4((float)(x0 + x1) * 0.5f, (float)(y0 + y1) * 0.5f, (float)(x1 - x0) * 0.5f, (float)(y1 - y0) * 0.5f);
                dst[3] = swapped != 0.f ? make_float4(a3.x, a3.y, a3.z, -a3.w) : a3;
                dst[4] = make_float4(single ? __uint_as_float(lane) : a2.z, a2.w, swapped, __uint_as_float(mask));
                if (single) {
                    const float4* ap = frame_attr + (size_t)__float_as_uint(a2.z) * NQ;
#pragma unroll
                    for (int k = 0; k < NQ; k++) wt.sattr[lane * NQ + k] = __ldg(ap + k);
                }
                if (INLOOP) {
                    const float4* ap = p.tri_attr + ((size_t)f * p.tri_cap + __float_as_uint(a2.z)) * 2;
                    float4 rw = __ldg(ap);
                    /* all three 1/w exactly 1 (an orthographic light: lmvp's last row is (0,0,0,1)): flag it in the unused fourth float */
                    rw.w = (rw.x == 1.f && rw.y == 1.f && rw.z == 1.f) ? 1.f : 0.f;
                    wt.sattr[lane * 2] = rw;
                    wt.sattr[lane * 2 + 1] = __ldg(ap + 1);
                }
            }
            __syncwarp();
            /* last record first: the resolve is order-free (smallest depth, then largest key), and a list is filled roughly in
             * submission order, so scenes drawn back to front (BASELINE.json configs[4]) meet their nearest layer first and
             * the layers behind it fail the depth test instead of each winning, parking weights and being overwritten */
            for (int j = n - 1; j >= 0; j--) {
                const float4 r0 = wt.tri[j * RW_REC_Q + 0]; /* ax, ay, s0x, s0y */
                const float4 r1 = wt.tri[j * RW_REC_Q + 1]; /* s1x, s1y, uz, thr */
                const float4 r4 = wt.tri[j * RW_REC_Q + 4]; /* slot, key, exchanged, mask */
                const uint32_t m = __float_as_uint(r4.w);
                /* Two pixels of the lane, A and B (state indices ia, ib; coordinates (fxA, fyA), (fxB, fyB); shared-memory
                 * pixel indices pixA, pixB), whose u.x and u.y arrive as the halves of UX and UY. */
                auto pixel_pair = [&](const f2 UX, const f2 UY, auto ia_, auto ib_, const float fxA, const float fxB, const float fyA,
                                      const float fyB, const int pixA, const int pixB) {
                    constexpr int ia = decltype(ia_)::value, ib = decltype(ib_)::value;
                    const f2 S = f2_add(UX, UY);
                    const f2 D = f2_sub(S, f2_dup(r1.z));
                    const f2 E = f2_sub(f2_dup(-r1.w), D);
                    /* coverage_test() for u.z < 0: u.x <= 0, u.y <= 0, d >= -thr */
                    const bool cA = fmax3(f2_lo(UX), f2_lo(UY), f2_lo(E)) <= 0.f;
                    const bool cB = fmax3(f2_hi(UX), f2_hi(UY), f2_hi(E)) <= 0.f;
                    if (cA || cB) {
                        /* the reference only visits pixels of its clamped bounding box: graphics.cpp:339-351 */
                        const float4 r2 = wt.tri[j * RW_REC_Q + 2];
                        const bool inA = cA && fabsf(fyA - r2.y) <= r2.w && fabsf(fxA - r2.x) <= r2.z;
                        const bool inB = cB && fabsf(fyB - r2.y) <= r2.w && fabsf(fxB - r2.x) <= r2.z;
                        if (inA || inB) {
                            const float4 r3 = wt.tri[j * RW_REC_Q + 3]; /* d0, d1, d2, 1/uz */
                            const f2 RUZ = f2_dup(r3.w), NUZ = f2_dup(-r1.z);
                            /* (1 - (u.x+u.y)/u.z, u.y/u.z, u.x/u.z): graphics.cpp:231; three independent Markstein
                             * quotients (f2_div_by_recip), written interleaved */
                            const f2 qs0 = f2_mul(S, RUZ), qy0 = f2_mul(UY, RUZ), qx0 = f2_mul(UX, RUZ);
                            const f2 rs = f2_fma(qs0, NUZ, S), ry = f2_fma(qy0, NUZ, UY), rx = f2_fma(qx0, NUZ, UX);
                            const f2 qs = f2_fma(rs, RUZ, qs0);
                            f2 W1 = f2_fma(ry, RUZ, qy0);
                            f2 W2 = f2_fma(rx, RUZ, qx0);
                            const f2 W0 = f2_fma(qs, f2_negone(), f2_one());
                            if (r4.z != 0.f) { /* B and C were exchanged: so were u.x and u.y */
                                const f2 t = W1;
                                W1 = W2;
                                W2 = t;
                            }
                            /* interpolate_depth graphics.cpp:186-194 */
                            f2 Z = f2_mul_from_zero(f2_dup(r3.z), W2);
                            Z = f2_add(Z, f2_mul(f2_dup(r3.y), W1));
                            Z = f2_add(Z, f2_mul(f2_dup(r3.x), W0));
                            const uint32_t slot = __float_as_uint(r4.x);
                            const float zA = f2_lo(Z), zB = f2_hi(Z);
                            bool winA = inA && zA < bz[ia], winB = inB && zB < bz[ib];
                            const bool tieA = inA && zA == bz[ia], tieB = inB && zB == bz[ib];
                            if (tieA || tieB) { /* rare */
                                const uint32_t key = __float_as_uint(r4.y);
                                if (tieA) winA = wins_tie(ia, key);
                                if (tieB) winB = wins_tie(ib, key);
                            }
                            if (winA) {
                                bz[ia] = zA;
                                bj[ia] = slot;
                            }
                            if (winB) {
                                bz[ib] = zB;
                                bj[ib] = slot;
                            }
                            if (INLOOP) {
                                if (winA || winB) { /* ShadowShader::fragment IShader.cpp:176-180 on the winning fragments */
                                    const float4 rw = wt.sattr[j * 2], a = wt.sattr[j * 2 + 1];
                                    /* interpolate_varyings graphics.cpp:205-220 for clip_pos.z */
                                    f2 V0 = W0, V1 = W1, V2 = W2, NORM; /* 1/w == 1: the products are the weights themselves */
                                    if (rw.w != 0.f) { /* warp-uniform. The weights of a covered pixel are >= 0 and sum to 1 within a few ulps: no range check */
                                        NORM = f2_rcp_normal(f2_add(f2_add(V0, V1), V2));
                                    } else {
                                        V0 = f2_mul(f2_dup(rw.x), W0);
                                        V1 = f2_mul(f2_dup(rw.y), W1);
                                        V2 = f2_mul(f2_dup(rw.z), W2);
                                        NORM = f2_rcp(f2_add(f2_add(V0, V1), V2));
                                    }
                                    const f2 AT = f2_mul(f2_add(f2_add(f2_mul(f2_dup(a.x), V0), f2_mul(f2_dup(a.y), V1)), f2_mul(f2_dup(a.z), V2)), NORM);
                                    if (WIDE) {
                                        if (winA) bb[ia >> 2] = put_byte<ia & 3>(bb[ia >> 2], shadow_byte(f2_lo(AT)));
                                        if (winB) bb[ib >> 2] = put_byte<ib & 3>(bb[ib >> 2], shadow_byte(f2_hi(AT)));
                                    } else {
                                        if (winA) bj[ia] |= shadow_byte(f2_lo(AT)) << 24;
                                        if (winB) bj[ib] |= shadow_byte(f2_hi(AT)) << 24;
                                    }
                                }
                            } else {
                                if (winA) {
                                    pw0[pixA] = f2_lo(W0);
                                    wt.pw1[pixA] = f2_lo(W1);
                                    wt.pw2[pixA] = f2_lo(W2);
                                }
                                if (winB) {
                                    pw0[pixB] = f2_hi(W0);
                                    wt.pw1[pixB] = f2_hi(W1);
                                    wt.pw2[pixB] = f2_hi(W2);
                                }
                            }
                        }
                    }
                };
                /* x-dependent terms of cross(s0, s1) for both column halves: graphics.cpp:224-229 */
                const f2 S0Z = f2_sub(f2_dup(r0.x), FPX);          /* A.x - P.x */
                const f2 T2 = f2_mul(S0Z, f2_dup(r1.y));           /* s0.z * s1.y */
                const f2 T3 = f2_mul(S0Z, f2_dup(r1.x));           /* s0.z * s1.x */
#pragma unroll
                for (int rp = 0; rp < 2; rp++) {
                    if (m & (0xFu << (4 * rp))) { /* warp-uniform */
                        /* y-dependent terms for two row quarters */
                        const f2 FPY = rp ? FPY23 : FPY01;
                        const f2 S1Z = f2_sub(f2_dup(r0.y), FPY); /* A.y - P.y */
                        const f2 T1 = f2_mul(f2_dup(r0.w), S1Z);  /* s0.y * s1.z */
                        const f2 T4 = f2_mul(f2_dup(r0.z), S1Z);  /* s0.x * s1.z */
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            /* h = column half; the pair is the lane's two pixels of that column in row quarters 2rp, 2rp+1:
                             * an 8x8 block of the tile per instruction */
                            if (m & (5u << (4 * rp + h))) { /* warp-uniform */
                                const f2 UX = f2_sub(T1, f2_dup(f2_half(T2, h)));
                                const f2 UY = f2_sub(f2_dup(f2_half(T3, h)), T4);
                                const float fx = f2_half(FPX, h);
                                if (rp == 0 && h == 0) pixel_pair(UX, UY, std::integral_constant<int, 0>(), std::integral_constant<int, 2>(), fx, fx, f2_lo(FPY), f2_hi(FPY), pix0, pix0 + 64);
                                if (rp == 0 && h == 1) pixel_pair(UX, UY, std::integral_constant<int, 1>(), std::integral_constant<int, 3>(), fx, fx, f2_lo(FPY), f2_hi(FPY), pix0 + 8, pix0 + 72);
                                if (rp == 1 && h == 0) pixel_pair(UX, UY, std::integral_constant<int, 4>(), std::integral_constant<int, 6>(), fx, fx, f2_lo(FPY), f2_hi(FPY), pix0 + 128, pix0 + 192);
                                if (rp == 1 && h == 1) pixel_pair(UX, UY, std::integral_constant<int, 5>(), std::integral_constant<int, 7>(), fx, fx, f2_lo(FPY), f2_hi(FPY), pix0 + 136, pix0 + 200);
                            }
                        }
                    }
                }
            }
        }

        /* which of the lane's pixels are still owned by a visibility-buffer fragment (no list record beat it): those have no
         * parked weights / no ShadowShader byte yet. A micro-triangle is never listed, so the slot tells. */
        uint32_t micro_mask = 0;
        if (has_micro) {
#pragma unroll
            for (int sb = 0; sb < 8; sb++) {
                const int px = ipx0 + (sb & 1) * 8, py = ipy0 + (sb >> 1) * 4;
                if (px < p.W && py < p.H) {
                    const unsigned long long vv = __ldg(vis_frame + (size_t)py * p.W + px);
                    if (vv != ~0ull && bj[sb] != ORD_NONE && (bj[sb] & SLOT_MASK) == ((0xFFFFFFFFu - (uint32_t)vv) >> 3) &&
                        (!INLOOP || WIDE || (bj[sb] >> 24) == 0u) && __float_as_uint(bz[sb]) == (uint32_t)(vv >> 32))
                        micro_mask |= 1u << sb;
                }
            }
        }
        /* graphics.cpp:222-233 + :186-194 for one known-covered pixel of triangle `slot`, from its raster record: the scalar
         * statement of what the record loop computes two pixels at a time */
        auto micro_weights = [&](uint32_t slot, float fx, float fy, float& w0, float& w1, float& w2) {
            const float4 q0 = __ldg(frame_rec + (size_t)slot * 4), q1 = __ldg(frame_rec + (size_t)slot * 4 + 1),
                         q3 = __ldg(frame_rec + (size_t)slot * 4 + 3);
            float ux, uy, su;
            coverage_test(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, fx, fy, ux, uy, su);
            barycentric_weights(ux, uy, su, q1.z, q3.w, w0, w1, w2);
        };
        if (tma && mode_is_r8(MODE)) {
            if (lane == 0) tma_wait_read0(); /* the previous tile's store has drained the staging tile */
            __syncwarp();
        }
        uint32_t covered_acc = 0;
        if (INLOOP) {
#pragma unroll
            for (int sb = 0; sb < 8; sb++) {
                const int pix = pix0 + (sb >> 1) * 64 + (sb & 1) * 8;
                uint8_t v = WIDE ? (uint8_t)(bb[sb >> 2] >> (8 * (sb & 3))) : (uint8_t)(bj[sb] == ORD_NONE ? 0u : (bj[sb] >> 24));
                if (micro_mask & (1u << sb)) { /* ShadowShader::fragment IShader.cpp:176-180 for a visibility-buffer winner */
                    const uint32_t slot = bj[sb] & SLOT_MASK;
                    float w0, w1, w2;
                    micro_weights(slot, (float)(ipx0 + (sb & 1) * 8), (float)(ipy0 + (sb >> 1) * 4), w0, w1, w2);
                    const float4 rw = __ldg(frame_attr + (size_t)slot * 2), at = __ldg(frame_attr + (size_t)slot * 2 + 1);
                    const VaryingWeights vw = varying_weights(w0, w1, w2, rw.x, rw.y, rw.z);
                    v = (uint8_t)shadow_byte(interp(vw, at.x, at.y, at.z));
                }
                if (q.pixels_covered) covered_acc += __popc(__ballot_sync(FULL, bj[sb] != ORD_NONE));
                if (tma) {
                    wt.r8[pix] = v;
                } else {
                    const int px = ipx0 + (sb & 1) * 8, py = ipy0 + (sb >> 1) * 4;
                    if (px < p.W && py < p.H)
                        q.shadow_out[(size_t)f * q.shadow_out_frame_stride + (size_t)py * q.shadow_out_pitch + px] = v;
                }
            }
        } else {
        /* -- park the resolve state, shade the winners one sub-block at a time (graphics.cpp:362-373) -- */
#pragma unroll
        for (int sb = 0; sb < 8; sb++) {
            const int pix = pix0 + (sb >> 1) * 64 + (sb & 1) * 8;
            wt.ord[pix] = bj[sb];
            wt.depth[pix] = bz[sb];
        }
        DevShadow sh = q.shadow;
        if (sh.base) sh.base += (size_t)f * q.shadow_frame_stride;
#pragma unroll 1
        for (int sb = 0; sb < 8; sb++) {
            const int pix = pix0 + (sb >> 1) * 64 + (sb & 1) * 8;
            const int px = ipx0 + (sb & 1) * 8, py = ipy0 + (sb >> 1) * 4;
            const bool in_frame = px < p.W && py < p.H;
            const uint32_t slot = wt.ord[pix];
            uint32_t col = q.clear_color;
            if (MODE == MODE_RMW) {
                if (tma) col = wt.color[pix];
                else if (in_frame && slot != ORD_NONE) col = q.color[(size_t)f * q.frame_stride + (size_t)py * p.W + px];
            }
            if (slot != ORD_NONE) {
                float w0 = pw0[pix], w1 = wt.pw1[pix], w2 = wt.pw2[pix];
                if (micro_mask & (1u << sb)) micro_weights(slot, (float)px, (float)py, w0, w1, w2); /* never parked: recompute */
                uint32_t rgb;
                if (single) rgb = shade_fragment<SHADER>(wt.fu, AttrShared{wt.sattr + slot * NQ}, w0, w1, w2, q.diffuse, q.normal, sh, frame_attr);
                else rgb = shade_fragment<SHADER>(wt.fu, AttrGlobal{frame_attr + (size_t)slot * NQ}, w0, w1, w2, q.diffuse, q.normal, sh, frame_attr);
                col = (col & 0xFF000000u) | rgb; /* alpha is never written: renderbuffer.cpp:38-44 */
                if (q.primid && f == 0)
                    q.primid[(size_t)py * p.W + px] = single ? __float_as_uint(wt.tri[slot * RW_REC_Q + 4].y)
                                                             : __float_as_uint(__ldg(frame_rec + (size_t)slot * 4 + 2).w);
            }
            if (q.pixels_covered) covered_acc += __popc(__ballot_sync(FULL, slot != ORD_NONE));
            if (tma) {
                wt.color[pix] = col;
            } else if (in_frame) {
                if (MODE == MODE_CLEAR_FOLD || slot != ORD_NONE) {
                    const size_t o = (size_t)f * q.frame_stride + (size_t)py * p.W + px;
                    q.color[o] = col;
                    q.depth[o] = wt.depth[pix];
                }
            }
        }
        }
        if (q.pixels_covered && lane == 0 && covered_acc) atomicAdd(q.pixels_covered + f, covered_acc);

        /* -- flush: lane 0 hands the warp's tile to the TMA engine -- */
        if (tma) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (mode_is_r8(MODE)) {
                    tma_store_3d(&tm_r8, wt.r8, X0, Y0, f);
                } else {
                    tma_store_3d(&tm_color, wt.color, X0, Y0, f);
                    tma_store_3d(&tm_depth, wt.depth, X0, Y0, f);
                }
                tma_commit();
            }
        }
        /* clear duty while clear work remains: one item (a group of tiles of one tile row) per tile rasterised */
        if (MODE != MODE_RMW && !clear_done) {
            const uint32_t item = __shfl_sync(FULL, clr_base, 0);
            if (item >= n_clear) clear_done = true;
            else clear_item<MODE>(q, item, clear_groups_x);
        }
        /* advance the queue */
        e_cur = e_nxt;
        i_next = i_nn;
    }
    /* raster queue drained: every warp helps with what is left of the clear queue */
    if (MODE != MODE_RMW) {
        while (!clear_done) {
            uint32_t item = 0;
            if (lane == 0) item = atomicAdd(&p.counters->clear_cursor, 1u);
            item = __shfl_sync(FULL, item, 0);
            if (item >= n_clear) clear_done = true;
            else clear_item<MODE>(q, item, clear_groups_x);
        }
    }
    /* shared memory must outlive the bulk stores that read it */
    tma_wait_read0();
}

/* ---- helpers ---------------------------------------------------------------- */
/* Small device -> pinned-host read-backs of the asynchronous sweep path (capacity needs, counters, file offsets) as a
 * KERNEL that stores into the mapped host allocation. A cudaMemcpyAsync would do, but it is a copy-engine operation:
 * queued behind a frame download of the previous batch in the engine's FIFO, it held up this stream's kernels until that
 * download had finished (measured: 512-frame batches took render + encode + copy = 25.8 ms instead of max(16.5, 9.3)). */
__global__ void post_words_kernel(uint32_t* __restrict__ dst_host, const uint32_t* __restrict__ src, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst_host[i] = src[i];
    __threadfence_system();
}

__global__ void fill32_kernel(uint32_t* __restrict__ dst, uint32_t value, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t n4 = n / 4;
    uint4 v = make_uint4(value, value, value, value);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t k = i; k < n4; k += stride) d4[k] = v;
    for (size_t k = n4 * 4 + i; k < n; k += stride) dst[k] = value;
}

/* ---- present ------------------------------------------------------
 * window_draw_buffer win32.cpp:348-370: row r of the render buffer (y up) lands in row H-1-r of the surface and R, B
 * change places. format 0: 4 bytes per pixel B,G,R,255; format 1: 3 bytes per pixel B,G,R (a TGA payload, rows
 * top-down as tgaimage.cpp:166 flags them). One thread per 4 pixels of a row: 128-bit loads, 128- or 3x32-bit stores. */
__global__ void __launch_bounds__(256) present_kernel(const uint32_t* __restrict__ color, size_t frame_stride, int first,
                                                      int W, int H, int format, uint8_t* __restrict__ dst) {
    const int f = blockIdx.z;
    const int y = blockIdx.y;
    const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x4 >= W) return;
    const uint32_t* src = color + (size_t)(first + f) * frame_stride + (size_t)y * W + x4;
    const int bpp = format == 0 ? 4 : 3;
    uint8_t* out = dst + ((size_t)f * H + (size_t)(H - 1 - y)) * (size_t)W * bpp + (size_t)x4 * bpp;
    uint32_t c[4];
    const int n = min(4, W - x4);
    if (n == 4 && (W & 3) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
        for (int k = 0; k < 4; k++) c[k] = k < n ? __ldg(src + k) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = ((c[k] >> 16) & 0xFFu) | (c[k] & 0xFF00u) | ((c[k] & 0xFFu) << 16); /* 0x00RRGGBB: B,G,R bytes */
    if (format == 0) {
        if (n == 4 && (W & 3) == 0) {
            *reinterpret_cast<uint4*>(out) = make_uint4(c[0] | 0xFF000000u, c[1] | 0xFF000000u, c[2] | 0xFF000000u, c[3] | 0xFF000000u);
        } else {
            for (int k = 0; k < n; k++) reinterpret_cast<uint32_t*>(out)[k] = c[k] | 0xFF000000u;
        }
    } else if (n == 4 && (W & 3) == 0) { /* 12 bytes, 4-byte aligned because W*3 is a multiple of 4 */
        uint32_t* o = reinterpret_cast<uint32_t*>(out);
        o[0] = c[0] | (c[1] << 24);
        o[1] = (c[1] >> 8) | (c[2] << 16);
        o[2] = (c[2] >> 16) | (c[3] << 8);
    } else {
        for (int k = 0; k < n; k++) {
            out[3 * k] = (uint8_t)c[k];
            out[3 * k + 1] = (uint8_t)(c[k] >> 8);
            out[3 * k + 2] = (uint8_t)(c[k] >> 16);
        }
    }
}

/* Order-independent 64-bit frame checksum: sum over pixels of mix(index, RGB, depth bits).
 * tests/ and the multi-GPU sharding check recompute it with numpy. */
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
__global__ void checksum_kernel(const uint32_t* __restrict__ color, const float* __restrict__ depth, size_t frame_stride,
                                size_t npix, unsigned long long* __restrict__ out) {
    const int f = blockIdx.y;
    const uint32_t* c = color + (size_t)f * frame_stride;
    const float* d = depth + (size_t)f * frame_stride;
    uint64_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = ((uint64_t)(c[i] & 0x00FFFFFFu) << 32) | (uint64_t)__float_as_uint(d[i]);
        acc += mix64(v ^ mix64((uint64_t)i + 0x9E3779B97F4A7C15ULL));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + f, (unsigned long long)acc);
}

/* pixels whose depth differs from the clear depth (sweep statistics) */
__global__ void count_written_kernel(const float* __restrict__ depth, size_t frame_stride, size_t npix, float clear_depth,
                                     uint32_t* __restrict__ out) {
    const int f = blockIdx.y;
    const float* d = depth + (size_t)f * frame_stride;
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
        acc += (__float_as_uint(d[i]) != __float_as_uint(clear_depth));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + f, acc);
}

}  // namespace hana
#endif /* HANA_KERNELS_CUH */
