/*
 * hana_b200.cu — the C ABI of include/hana_b200.h over the kernels in
 * hana_kernels.cuh. Host side only: device memory, tensor maps, pass
 * sequencing, capacity management. No rendering arithmetic happens here and
 * there is no CPU fallback: without a CUDA device every entry point fails
 * with HANA_E_NODEVICE.
 *
 * Pass sequencing mirrors DrawModel::draw (scene.h:53-99): ShadowShader pass,
 * then the shaded pass reading its colour plane, then the shadow-map clear.
 */
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "hana_kernels.cuh"
#include "hana_tga.cuh"

using namespace hana;

#define HANA_VERSION_NUM 100

/* ---- errors ------------------------------------------------------------------ */
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
extern "C" int hana_set_error(int code, const char* msg) { return fail(code, msg ? msg : ""); } /* for host/ sources */
#define CU_TRY(expr)                                                                                             \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? HANA_E_NODEVICE : HANA_E_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                                    \
    } while (0)
#define HANA_TRY(expr)            \
    do {                          \
        int r__ = (expr);         \
        if (r__ != HANA_OK) return r__; \
    } while (0)

/* ---- objects ------------------------------------------------------------------ */
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum ProfKind { PROF_BEGIN = 0, PROF_SETUP, PROF_SCAN, PROF_FILL, PROF_RASTER_SHADOW, PROF_RASTER_MAIN, PROF_OTHER, PROF_KINDS };

struct Scratch {
    float4* tri_rec = nullptr; /* 4 float4 per record */
    float4* tri_attr = nullptr;
    uint2* tri_bbox = nullptr;
    size_t tri_total = 0; /* records allocated (n_frames * tri_cap) */
    size_t attr_total = 0;
    size_t bbox_total = 0;
    uint32_t* tri_count = nullptr; /* [frames][TRI_COUNT_WAYS] emitted triangles, then [frames] extra slots */
    size_t frames_cap = 0;
    uint32_t* tile_arrays = nullptr; /* count | cursor | micro | offset, each n_frames * n_tiles (the first three padded) */
    size_t tile_arr_cap = 0;
    unsigned long long* vis = nullptr; /* visibility buffer of the micro-triangle path (dense meshes only): n_frames * W * H */
    size_t vis_cap = 0;
    uint32_t* group_listed = nullptr;  /* with a visibility buffer: n_frames * ceil(tri_cap / 32) flags (PassParams) */
    size_t group_cap = 0;
    float4* tile_recs = nullptr; /* pool of raster records: 4 float4 each */
    size_t pool_cap = 0;         /* float4 elements */
    uint4* work = nullptr;
    size_t work_cap = 0;
    PassCounters* counters = nullptr;
    PassCounters* counters_host = nullptr; /* pinned */
    PassParams saved_p;                    /* what the binning phase of a split pass left for its raster phase */
};

struct hana_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; /* device->host frame copies, so that they overlap the next batch's kernels */
    std::vector<hana_sweep*> sweeps;
    uint64_t launches = 0;
    bool use_tma = true;
    EncodeTiledFn encode = nullptr;
    Scratch sc;
    Scratch sc2;                          /* second set: the main pass is binned on side_stream beside the shadow pass */
    Scratch sc3, sc4;                     /* the same two for every other submission of a pipelined sweep (below) */
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    /* Pipelined sweeps: the two passes of submission k+1 are binned (bin_stream, side_stream; scratch sets of the other
     * parity) while the rasterisers of submission k still run on `stream`, so the fixed cost of a submission — eight
     * small dependent kernels in front of the first rasteriser, the rasterisers' tails — is hidden behind its neighbours.
     * A rasteriser waits for its pass's binning (ev_bin), a binning for the rasterisers that last read its scratch set
     * and uniform block (ev_raster_done of the same parity), and for whatever serial work used sets 0/1 since (ev_serial). */
    cudaStream_t bin_stream = nullptr;
    cudaEvent_t ev_bin[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; /* [parity][pass] */
    cudaEvent_t ev_raster_done[2] = {nullptr, nullptr};
    bool raster_done_valid[2] = {false, false};
    cudaEvent_t ev_serial = nullptr;
    bool serial_dirty = false;
    int parity = 0;
    bool pipeline = true;                 /* HANA_NO_PIPELINE=1 in the environment: one submission after the other */
    uint32_t tri_cap_hint = 0;
    size_t pool_hint = 0;                 /* list records a batch has needed so far: what a scratch set is created with */
    HanaUniforms* u_raw = nullptr;  /* device, 1 */
    DevUniforms* u_dev = nullptr;   /* device, 1 */
    uint32_t* stat_pixels = nullptr; /* device, 1 */
    HanaStats last_stats;
    bool profile = false;
    struct ProfEv {
        cudaEvent_t a, b;
        int kind;
    };
    std::vector<ProfEv> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[PROF_KINDS] = {0};
    uint64_t prof_n[PROF_KINDS] = {0};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    /* host-buffer path cache */
    hana_sweep* host_sweep = nullptr;
    hana_rb* host_frame = nullptr;
    hana_rb* host_shadow = nullptr;
    int occ[HANA_SHADER_COUNT][2 * N_RASTER_MODES]; /* [mode] without, [N_RASTER_MODES + mode] with a visibility buffer */
    uint32_t r8_slot_limit = R8_SLOT_LIMIT; /* HANA_R8_SLOT_LIMIT in the environment lowers it (tests force the WIDE variant) */
    uint64_t wide_r8_launches = 0;
};

static Scratch& scratch_of(hana_ctx* ctx, int i) { return i == 0 ? ctx->sc : i == 1 ? ctx->sc2 : i == 2 ? ctx->sc3 : ctx->sc4; }

struct hana_model {
    hana_ctx* ctx;
    float4* posu;
    float4* nrmv;
    int ncorners;
};
struct hana_texture {
    hana_ctx* ctx;
    uint32_t* texels;
    int w, h;
};
struct hana_rb {
    hana_ctx* ctx;
    int w, h;
    uint32_t* color;
    float* depth;
    bool tma_ok;
    CUtensorMap tm_color, tm_depth;
};
struct hana_sweep {
    hana_ctx* ctx;
    int w, h, max_frames;
    uint32_t* color;
    float* depth;
    uint8_t* shadow_r8;
    int shadow_pitch;
    size_t shadow_frame_bytes;
    bool tma_ok;
    CUtensorMap tm_color, tm_depth, tm_r8;
    HanaUniforms* u_raw;
    DevUniforms* u_dev;            /* the block the last submission used: u_dev_buf + parity * max_frames */
    DevUniforms* u_dev_buf;        /* device [2][max_frames]: consecutive pipelined submissions alternate */
    unsigned long long* checksums;
    uint32_t* pix_counts;
    int last_frames;
    float last_clear_depth;
    HanaStats last_stats; /* frame-independent part */
    std::vector<uint32_t> last_tri_counts[2];
    /* asynchronous rendering: a render is launched without any read-back; what it needed is checked (and the batch
     * re-rendered with larger scratch if something was dropped) at the sweep's next synchronisation point */
    OverflowRecord* overflow;      /* device: the record the last submission used (overflow_buf + parity) */
    OverflowRecord* overflow_buf;  /* device [2] */
    /* A render that is superseded by the next one before anybody synchronised with it (back-to-back submissions) is
     * not forgotten: its needs land in a slot of their own and are examined, without blocking, at later calls; a batch
     * that ran out of scratch is counted (hana_sweep_overflow_count) and the scratch grows for the batches that follow. */
    enum { CHECK_RING = 8, SLOT_PASSES_OK = CHECK_RING + 1, NEED_SLOTS = CHECK_RING + 2 };
    struct Pinned {
        OverflowRecord need[NEED_SLOTS];
        PassCounters counters[2];
    }* pin;                        /* pinned host */
    cudaEvent_t ev_check[CHECK_RING + 1];
    struct Check {
        int slot;
        uint32_t tri_cap, pool_cap;
    } checks[CHECK_RING];
    int check_tail, n_checks, next_slot;
    uint64_t overflow_batches;
    uint32_t* tri_counts_pin;      /* pinned host: [2][max_frames] */
    cudaEvent_t ev_render, ev_copy;
    bool copy_in_flight;
    int band[2][2];                /* tile rows [first, first+count) of the shadow / main pass; count 0 = all (hana_sweep_set_bands) */
    bool shadow_reuse = false;     /* hana_sweep_set_shadow_reuse */
    bool shadow_shared = false;    /* the last batch rendered ONE shadow map for all its frames */
    uint64_t shadow_shared_batches = 0;
    uint8_t* present_buf;          /* device: presented frames (hana_sweep_present) */
    size_t present_cap;
    /* RLE TGA files made on the device (hana_sweep_encode_tga / hana_sweep_fetch_tga) */
    struct Tga {
        uint8_t* packed = nullptr;     /* the files back to back (16-byte aligned starts), written in place by tga_write_kernel */
        size_t cap_frames = 0, worst_bytes = 0;
        uint32_t* ebits = nullptr;     /* [count][estride] equal-neighbour bits, 32 pixels per word */
        TgaRec* recs = nullptr;        /* [count][rstride] state at the first pixel of every word */
        uint8_t* counts = nullptr;     /* [count][rstride] bytes each word emits */
        uint32_t* batch_sums = nullptr;            /* [count][nbatch] */
        unsigned long long* batch_offs = nullptr;  /* [count][nbatch] */
        unsigned long long* sizes = nullptr;   /* device [max_frames] */
        unsigned long long* offsets = nullptr; /* device [max_frames + 1] */
        unsigned long long* meta_pin = nullptr; /* pinned: offsets [max_frames + 1], then sizes [max_frames] */
        cudaEvent_t ev = nullptr;
        int first = 0, count = 0;      /* what the last encode covered */
        uint64_t rerender_seen = 0;
        bool valid = false;
    } tga;
    uint64_t rerender_count;       /* batches rendered again by sweep_verify */
    bool stats_lazy;               /* the last render read nothing back: statistics are still in scratch set stats_scratch */
    int stats_scratch;
    struct Pending {
        bool active = false;
        const hana_model* model = nullptr;
        int shader = 0, enable_shadow = 0, n_frames = 0;
        const hana_texture* diffuse = nullptr;
        const hana_texture* normal = nullptr;
        uint8_t clear_rgba[4] = {0, 0, 0, 0};
        float clear_depth = 0.f;
        uint32_t tri_cap = 0, pool_cap = 0;
        int slot = 0;
        bool share_shadow = false;
    } pending;
};

extern "C" const char* hana_last_error(void) { return g_err.c_str(); }
extern "C" int hana_version(void) { return HANA_VERSION_NUM; }
extern "C" int hana_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

/* ---- tensor maps ---------------------------------------------------------------- */
static int make_tensor_map(hana_ctx* ctx, CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, void* base, uint64_t w,
                           uint64_t h, uint64_t frames, uint64_t row_bytes, uint64_t frame_bytes) {
    memset(tm, 0, sizeof(*tm));
    if (!ctx->encode) return fail(HANA_E_CUDA, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[3] = {w, h, frames};
    cuuint64_t strides[2] = {row_bytes, frame_bytes};
    cuuint32_t box[3] = {TILE, TILE, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    (void)elem_bytes;
    CUresult r = ctx->encode(tm, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HANA_E_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return HANA_OK;
}

/* ---- context ------------------------------------------------------------------- */
static int use_device(hana_ctx* ctx) {
    CU_TRY(cudaSetDevice(ctx->device));
    return HANA_OK;
}

extern "C" int hana_ctx_create(int device, hana_ctx** out) {
    if (!out) return fail(HANA_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(HANA_E_NODEVICE, std::string("no CUDA device (there is no CPU fallback): ") +
                                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= n) return fail(HANA_E_INVALID, "device ordinal out of range");
    hana_ctx* ctx = new (std::nothrow) hana_ctx();
    if (!ctx) return fail(HANA_E_INVALID, "out of host memory");
    ctx->device = device;
    memset(&ctx->last_stats, 0, sizeof(ctx->last_stats));
    for (auto& a : ctx->occ)
        for (int& b : a) b = 0;
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete ctx;
        return fail(HANA_E_NODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                         "; this library contains sm_100a code only");
    }
    ctx->sm_count = prop.multiProcessorCount;
    {
        /* the stream the rasterisers run on outranks the streams that bin the next submission beside them: CTAs of a
         * waiting rasteriser take free SM slots first, binning fills what the rasterisers' tails leave idle */
        int prio_lo = 0, prio_hi = 0;
        CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const char* np = getenv("HANA_NO_STREAM_PRIO");
        CU_TRY(cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, (np && np[0] == '1') ? prio_lo : prio_hi));
    }
    CU_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
        ctx->encode = (EncodeTiledFn)fn;
    else
        cudaGetLastError();
    if (const char* lim = getenv("HANA_R8_SLOT_LIMIT")) {
        const long v = atol(lim);
        if (v >= 0 && (unsigned long)v < R8_SLOT_LIMIT) ctx->r8_slot_limit = (uint32_t)v;
    }
    const char* no_tma = getenv("HANA_NO_TMA");
    ctx->use_tma = ctx->encode != nullptr && !(no_tma && no_tma[0] == '1');
    CU_TRY(cudaMalloc(&ctx->u_raw, sizeof(HanaUniforms)));
    CU_TRY(cudaMalloc(&ctx->u_dev, sizeof(DevUniforms)));
    CU_TRY(cudaMalloc(&ctx->stat_pixels, sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&ctx->sc.counters, sizeof(PassCounters)));
    CU_TRY(cudaMallocHost(&ctx->sc.counters_host, sizeof(PassCounters)));
    for (int i = 1; i < 4; i++) {
        CU_TRY(cudaMalloc(&scratch_of(ctx, i).counters, sizeof(PassCounters)));
        CU_TRY(cudaMallocHost(&scratch_of(ctx, i).counters_host, sizeof(PassCounters)));
    }
    CU_TRY(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&ctx->bin_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU_TRY(cudaEventCreateWithFlags(&ctx->ev_bin[i][0], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ctx->ev_bin[i][1], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ctx->ev_raster_done[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&ctx->ev_serial, cudaEventDisableTiming));
    if (const char* np = getenv("HANA_NO_PIPELINE")) ctx->pipeline = !(np[0] == '1');
    CU_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CU_TRY(cudaEventCreate(&ctx->t0));
    CU_TRY(cudaEventCreate(&ctx->t1));
    *out = ctx;
    return HANA_OK;
}

extern "C" int hana_sweep_destroy(hana_sweep* s);
extern "C" int hana_rb_destroy(hana_rb* rb);

extern "C" int hana_ctx_destroy(hana_ctx* ctx) {
    if (!ctx) return HANA_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->host_sweep) hana_sweep_destroy(ctx->host_sweep);
    if (ctx->host_frame) hana_rb_destroy(ctx->host_frame);
    if (ctx->host_shadow) hana_rb_destroy(ctx->host_shadow);
    for (Scratch* sp : {&ctx->sc, &ctx->sc2, &ctx->sc3, &ctx->sc4}) {
        Scratch& s = *sp;
        cudaFree(s.tri_rec); cudaFree(s.tri_attr); cudaFree(s.tri_bbox); cudaFree(s.tri_count); cudaFree(s.tile_arrays);
        cudaFree(s.tile_recs); cudaFree(s.work); cudaFree(s.counters); cudaFreeHost(s.counters_host); cudaFree(s.vis); cudaFree(s.group_listed);
    }
    cudaStreamDestroy(ctx->side_stream);
    cudaStreamDestroy(ctx->bin_stream);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); cudaEventDestroy(ctx->ev_serial);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_bin[i][0]); cudaEventDestroy(ctx->ev_bin[i][1]); cudaEventDestroy(ctx->ev_raster_done[i]); }
    cudaFree(ctx->u_raw); cudaFree(ctx->u_dev); cudaFree(ctx->stat_pixels);
    for (auto& p : ctx->prof_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1);
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return HANA_OK;
}

extern "C" int hana_ctx_set_stream(hana_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(use_device(ctx));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return HANA_OK;
}
static int sweep_verify(hana_sweep* s);
static int sweep_poll_checks(hana_sweep* s, bool wait_all);
extern "C" int hana_sync(hana_ctx* ctx) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(use_device(ctx));
    for (hana_sweep* s : ctx->sweeps) {
        HANA_TRY(sweep_poll_checks(s, true));
        HANA_TRY(sweep_verify(s)); /* re-renders a batch that ran out of scratch */
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->bin_stream));
    CU_TRY(cudaStreamSynchronize(ctx->side_stream));
    CU_TRY(cudaStreamSynchronize(ctx->copy_stream));
    return HANA_OK;
}
extern "C" int hana_ctx_launch_count(hana_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return fail(HANA_E_INVALID, "NULL argument");
    *out = ctx->launches;
    return HANA_OK;
}
extern "C" int hana_ctx_uses_tma(hana_ctx* ctx) { return ctx && ctx->use_tma ? 1 : 0; }
extern "C" int hana_ctx_set_tma(hana_ctx* ctx, int enable) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    ctx->use_tma = enable && ctx->encode;
    return HANA_OK;
}
extern "C" int hana_ctx_sm_count(hana_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int hana_ctx_set_pipeline(hana_ctx* ctx, int enable) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(hana_sync(ctx));
    ctx->pipeline = enable != 0;
    return HANA_OK;
}
extern "C" int hana_ctx_wide_r8_launches(hana_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return fail(HANA_E_INVALID, "NULL argument");
    *out = ctx->wide_r8_launches;
    return HANA_OK;
}

/* ---- profiling (CUDA events around each kernel class, on the launching stream) ---- */
static void prof_begin(hana_ctx* ctx, int kind, cudaEvent_t* a, cudaEvent_t* b, cudaStream_t st = nullptr) {
    *a = *b = nullptr;
    if (!ctx->profile) return;
    auto get = [&]() {
        cudaEvent_t e = nullptr;
        if (!ctx->ev_pool.empty()) {
            e = ctx->ev_pool.back();
            ctx->ev_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    };
    *a = get();
    *b = get();
    cudaEventRecord(*a, st ? st : ctx->stream);
    (void)kind;
}
static void prof_end(hana_ctx* ctx, int kind, cudaEvent_t a, cudaEvent_t b, cudaStream_t st = nullptr) {
    if (!a) return;
    cudaEventRecord(b, st ? st : ctx->stream);
    ctx->prof_pending.push_back({a, b, kind});
}
static void prof_resolve(hana_ctx* ctx) {
    for (auto& p : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ctx->prof_ms[p.kind] += ms;
            ctx->prof_n[p.kind] += 1;
        }
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->prof_pending.clear();
}
extern "C" int hana_ctx_profile(hana_ctx* ctx, int enable) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(use_device(ctx));
    ctx->profile = enable != 0; /* no synchronisation: pending event pairs are resolved by profile_get / profile_reset */
    return HANA_OK;
}
extern "C" int hana_ctx_profile_reset(hana_ctx* ctx) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(use_device(ctx));
    prof_resolve(ctx);
    for (int i = 0; i < PROF_KINDS; i++) {
        ctx->prof_ms[i] = 0;
        ctx->prof_n[i] = 0;
    }
    return HANA_OK;
}
extern "C" int hana_ctx_profile_get(hana_ctx* ctx, int kind, double* total_ms, uint64_t* launches) {
    if (!ctx || kind < 0 || kind >= PROF_KINDS) return fail(HANA_E_INVALID, "bad argument");
    HANA_TRY(use_device(ctx));
    prof_resolve(ctx);
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    if (launches) *launches = ctx->prof_n[kind];
    return HANA_OK;
}
extern "C" int hana_timer_start(hana_ctx* ctx) {
    if (!ctx) return fail(HANA_E_INVALID, "ctx is NULL");
    HANA_TRY(use_device(ctx));
    CU_TRY(cudaEventRecord(ctx->t0, ctx->stream));
    return HANA_OK;
}
extern "C" int hana_timer_stop(hana_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(HANA_E_INVALID, "NULL argument");
    HANA_TRY(use_device(ctx));
    for (hana_sweep* s : ctx->sweeps) { /* the timed region ends when checked frames (and their copies) are complete */
        HANA_TRY(sweep_poll_checks(s, true));
        HANA_TRY(sweep_verify(s));
        if (s->copy_in_flight) CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
    }
    CU_TRY(cudaEventRecord(ctx->t1, ctx->stream));
    CU_TRY(cudaEventSynchronize(ctx->t1));
    CU_TRY(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return HANA_OK;
}

/* ---- inputs ----------------------------------------------------------------------- */
extern "C" int hana_model_upload(hana_ctx* ctx, const float* a2v, int ncorners, hana_model** out) {
    if (!ctx || !out || (!a2v && ncorners > 0)) return fail(HANA_E_INVALID, "NULL argument");
    if (ncorners < 0 || ncorners % 3 != 0) return fail(HANA_E_INVALID, "ncorners must be a non-negative multiple of 3");
    if (ncorners / 3 >= (1 << 28)) return fail(HANA_E_INVALID, "too many faces for the 32-bit order key");
    HANA_TRY(use_device(ctx));
    hana_model* m = new hana_model{ctx, nullptr, nullptr, ncorners};
    size_t n = (size_t)std::max(ncorners, 1);
    /* AoS a2v records -> two SoA float4 streams */
    std::vector<float4> posu(n), nrmv(n);
    for (int i = 0; i < ncorners; i++) {
        const float* a = a2v + (size_t)i * 8;
        posu[i] = make_float4(a[0], a[1], a[2], a[6]);
        nrmv[i] = make_float4(a[3], a[4], a[5], a[7]);
    }
    cudaError_t e1 = cudaMalloc(&m->posu, n * sizeof(float4));
    cudaError_t e2 = cudaMalloc(&m->nrmv, n * sizeof(float4));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaFree(m->posu); cudaFree(m->nrmv);
        delete m;
        cudaGetLastError();
        return fail(HANA_E_CUDA, "cudaMalloc failed for the model streams");
    }
    CU_TRY(cudaMemcpyAsync(m->posu, posu.data(), n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(m->nrmv, nrmv.data(), n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return HANA_OK;
}
extern "C" int hana_model_update(hana_model* m, const float* a2v, int ncorners) {
    if (!m || (!a2v && ncorners > 0)) return fail(HANA_E_INVALID, "NULL argument");
    if (ncorners != m->ncorners) return fail(HANA_E_INVALID, "corner count differs from the uploaded model");
    if (ncorners == 0) return HANA_OK;
    hana_ctx* ctx = m->ctx;
    HANA_TRY(use_device(ctx));
    std::vector<float4> posu(ncorners), nrmv(ncorners);
    for (int i = 0; i < ncorners; i++) {
        const float* a = a2v + (size_t)i * 8;
        posu[i] = make_float4(a[0], a[1], a[2], a[6]);
        nrmv[i] = make_float4(a[3], a[4], a[5], a[7]);
    }
    CU_TRY(cudaMemcpyAsync(m->posu, posu.data(), (size_t)ncorners * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(m->nrmv, nrmv.data(), (size_t)ncorners * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return HANA_OK;
}
extern "C" int hana_model_destroy(hana_model* m) {
    if (!m) return HANA_OK;
    cudaSetDevice(m->ctx->device);
    cudaFree(m->posu);
    cudaFree(m->nrmv);
    delete m;
    return HANA_OK;
}
extern "C" int hana_model_ncorners(const hana_model* m) { return m ? m->ncorners : 0; }

extern "C" int hana_texture_upload(hana_ctx* ctx, const uint8_t* data, int w, int h, int bytespp, hana_texture** out) {
    if (!ctx || !out || !data) return fail(HANA_E_INVALID, "NULL argument");
    if (w <= 0 || h <= 0 || (bytespp != 1 && bytespp != 3 && bytespp != 4))
        return fail(HANA_E_INVALID, "texture must be w,h > 0 with 1, 3 or 4 bytes per texel");
    HANA_TRY(use_device(ctx));
    size_t n = (size_t)w * h;
    std::vector<uint32_t> tex(n);
    for (size_t i = 0; i < n; i++) { /* TGAColor(p, bpp): bytes beyond bpp stay 0 (tgaimage.h:46-53) */
        uint32_t v = 0;
        for (int b = 0; b < bytespp; b++) v |= (uint32_t)data[i * bytespp + b] << (8 * b);
        tex[i] = v;
    }
    hana_texture* t = new hana_texture{ctx, nullptr, w, h};
    if (cudaMalloc(&t->texels, n * 4) != cudaSuccess) {
        delete t;
        cudaGetLastError();
        return fail(HANA_E_CUDA, "cudaMalloc failed for the texture");
    }
    cudaError_t ce = cudaMemcpyAsync(t->texels, tex.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) {
        cudaFree(t->texels);
        delete t;
        cudaGetLastError();
        return fail(HANA_E_CUDA, std::string("texture upload failed: ") + cudaGetErrorString(ce));
    }
    *out = t;
    return HANA_OK;
}
extern "C" int hana_texture_destroy(hana_texture* t) {
    if (!t) return HANA_OK;
    cudaSetDevice(t->ctx->device);
    cudaFree(t->texels);
    delete t;
    return HANA_OK;
}

/* ---- launches ------------------------------------------------------------------ */
/* nbytes (a multiple of 4) from device memory to a cudaMallocHost allocation, by a kernel: see post_words_kernel */
static int post_to_host(hana_ctx* ctx, void* dst_pinned, const void* src_dev, size_t nbytes, cudaStream_t st) {
    const uint32_t n = (uint32_t)(nbytes / 4);
    if (!n) return HANA_OK;
    post_words_kernel<<<(unsigned)std::min<uint32_t>((n + 255) / 256, 64u), 256, 0, st>>>((uint32_t*)dst_pinned, (const uint32_t*)src_dev, n);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return HANA_OK;
}

static int launch_fill32(hana_ctx* ctx, uint32_t* dst, uint32_t value, size_t n) {
    if (n == 0) return HANA_OK;
    int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)ctx->sm_count * 16);
    cudaEvent_t a, b;
    prof_begin(ctx, PROF_OTHER, &a, &b);
    fill32_kernel<<<blocks, 256, 0, ctx->stream>>>(dst, value, n);
    prof_end(ctx, PROF_OTHER, a, b);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return HANA_OK;
}

/* ---- render targets ---------------------------------------------------------------- */
extern "C" int hana_rb_create(hana_ctx* ctx, int width, int height, hana_rb** out) {
    if (!ctx || !out) return fail(HANA_E_INVALID, "NULL argument");
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535)
        return fail(HANA_E_INVALID, "render buffer size must be in 1..65535");
    HANA_TRY(use_device(ctx));
    hana_rb* rb = new hana_rb();
    rb->ctx = ctx;
    rb->w = width;
    rb->h = height;
    rb->color = nullptr;
    rb->depth = nullptr;
    size_t n = (size_t)width * height;
    if (cudaMalloc(&rb->color, n * 4) != cudaSuccess || cudaMalloc(&rb->depth, n * 4) != cudaSuccess) {
        cudaFree(rb->color);
        delete rb;
        cudaGetLastError();
        return fail(HANA_E_CUDA, "cudaMalloc failed for the render buffer");
    }
    rb->tma_ok = false;
    if (ctx->encode && width % 4 == 0) {
        int r1 = make_tensor_map(ctx, &rb->tm_color, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, rb->color, width, height, 1,
                                 (uint64_t)width * 4, (uint64_t)n * 4);
        int r2 = make_tensor_map(ctx, &rb->tm_depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rb->depth, width, height, 1,
                                 (uint64_t)width * 4, (uint64_t)n * 4);
        rb->tma_ok = (r1 == HANA_OK && r2 == HANA_OK);
    }
    /* RenderBuffer ctor: colour (0,0,0,255), depth 1.0 (renderbuffer.cpp:6-7,16-17) */
    HANA_TRY(launch_fill32(ctx, rb->color, 0xFF000000u, n));
    float one = 1.0f;
    uint32_t bits;
    memcpy(&bits, &one, 4);
    HANA_TRY(launch_fill32(ctx, reinterpret_cast<uint32_t*>(rb->depth), bits, n));
    *out = rb;
    return HANA_OK;
}
extern "C" int hana_rb_destroy(hana_rb* rb) {
    if (!rb) return HANA_OK;
    cudaSetDevice(rb->ctx->device);
    cudaStreamSynchronize(rb->ctx->stream);
    cudaFree(rb->color);
    cudaFree(rb->depth);
    delete rb;
    return HANA_OK;
}
extern "C" int hana_rb_size(const hana_rb* rb, int* width, int* height) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    if (width) *width = rb->w;
    if (height) *height = rb->h;
    return HANA_OK;
}
extern "C" int hana_rb_clear_color(hana_rb* rb, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    HANA_TRY(use_device(rb->ctx));
    uint32_t v = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
    return launch_fill32(rb->ctx, rb->color, v, (size_t)rb->w * rb->h);
}
extern "C" int hana_rb_clear_depth(hana_rb* rb, float depth) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    HANA_TRY(use_device(rb->ctx));
    uint32_t bits;
    memcpy(&bits, &depth, 4);
    return launch_fill32(rb->ctx, reinterpret_cast<uint32_t*>(rb->depth), bits, (size_t)rb->w * rb->h);
}
extern "C" int hana_rb_upload(hana_rb* rb, const uint8_t* color_rgba, const float* depth) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    HANA_TRY(use_device(rb->ctx));
    size_t n = (size_t)rb->w * rb->h * 4;
    if (color_rgba) CU_TRY(cudaMemcpyAsync(rb->color, color_rgba, n, cudaMemcpyHostToDevice, rb->ctx->stream));
    if (depth) CU_TRY(cudaMemcpyAsync(rb->depth, depth, n, cudaMemcpyHostToDevice, rb->ctx->stream));
    CU_TRY(cudaStreamSynchronize(rb->ctx->stream));
    return HANA_OK;
}
extern "C" int hana_rb_download(hana_rb* rb, uint8_t* color_rgba, float* depth) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    HANA_TRY(use_device(rb->ctx));
    size_t n = (size_t)rb->w * rb->h * 4;
    if (color_rgba) CU_TRY(cudaMemcpyAsync(color_rgba, rb->color, n, cudaMemcpyDeviceToHost, rb->ctx->stream));
    if (depth) CU_TRY(cudaMemcpyAsync(depth, rb->depth, n, cudaMemcpyDeviceToHost, rb->ctx->stream));
    CU_TRY(cudaStreamSynchronize(rb->ctx->stream));
    return HANA_OK;
}
extern "C" int hana_rb_device_ptrs(hana_rb* rb, void** color_dev, void** depth_dev) {
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    if (color_dev) *color_dev = rb->color;
    if (depth_dev) *depth_dev = rb->depth;
    return HANA_OK;
}

/* ---- one pass -------------------------------------------------------------------- */
struct PassDesc {
    int shader = 0, mode = MODE_RMW, n_frames = 1, W = 0, H = 0;
    const hana_model* model = nullptr;
    const DevUniforms* uniforms = nullptr;
    uint32_t* color = nullptr;
    float* depth = nullptr;
    size_t frame_stride = 0;
    const CUtensorMap* tm_color = nullptr;
    const CUtensorMap* tm_depth = nullptr;
    const CUtensorMap* tm_r8 = nullptr;
    bool tma_ok = false;
    uint8_t* shadow_out = nullptr;
    int shadow_out_pitch = 0;
    size_t shadow_out_frame_stride = 0;
    uint32_t clear_color = 0;
    float clear_depth = 0.f;
    DevTexture diffuse{nullptr, 0, 0}, normal{nullptr, 0, 0};
    DevShadow shadow{nullptr, 0, 0, 0, 0};
    size_t shadow_frame_stride = 0;
    uint32_t* primid = nullptr;
    uint32_t* pixels_covered = nullptr;
    float* dbg_v2f = nullptr;
    uint32_t dbg_cap = 0;
    bool setup_only = false;
    int prof_kind = PROF_RASTER_MAIN;
    std::vector<uint32_t>* tri_counts_out = nullptr; /* triangles emitted per frame */
    std::vector<uint32_t>* slots_out = nullptr;      /* record slots in use per frame (faces + what clipping added) */
    PassCounters* counters_out = nullptr;
    /* lazy mode (sweeps): nothing is read back inside the pass; capacities are the context's current ones and the
     * needs are accumulated in *overflow for the host to check at the sweep's next synchronisation point */
    bool lazy = false;
    OverflowRecord* overflow = nullptr;
    uint32_t* tri_counts_pinned = nullptr; /* lazy: [n_frames], filled asynchronously */
    PassCounters* counters_pinned = nullptr;
    uint32_t* tri_cap_used = nullptr;      /* lazy: capacities this pass ran with */
    uint32_t* pool_cap_used = nullptr;
    /* split passes (lazy only): phase 1 = setup + scan + fill on `stream` with scratch set `scratch`, phase 2 = raster */
    int phase = 0;
    int scratch = 0;      /* scratch set 0..3 */
    bool pipelined = false; /* part of a pipelined sweep submission: does not count as serial use of sets 0/1 */
    int band_first = 0, band_count = 0; /* tile rows of a split frame; count 0 = the whole frame */
    cudaStream_t stream = nullptr;
};

template <typename T>
static int grow(T** ptr, size_t* cap, size_t need, hana_ctx* ctx) {
    if (need <= *cap && *ptr) return HANA_OK;
    size_t n = std::max(need, *cap + *cap / 2);
    n = std::max<size_t>(n, 1);
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->side_stream));
    CU_TRY(cudaStreamSynchronize(ctx->bin_stream));
    if (*ptr) CU_TRY(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    CU_TRY(cudaMalloc(ptr, n * sizeof(T)));
    *cap = n;
    return HANA_OK;
}

template <int MODE, bool VIS>
static int launch_raster(cudaStream_t st, int shader, int blocks, const RasterParams& rp, const CUtensorMap& a,
                         const CUtensorMap& b, const CUtensorMap& c) {
    if constexpr (mode_is_r8(MODE)) { /* the 1-byte maps take the ShadowShader only */
        raster_kernel<HANA_SHADER_SHADOW, MODE, VIS><<<blocks, RW_THREADS, 0, st>>>(rp, a, b, c);
        return HANA_OK;
    }
#define HANA_RASTER_CASE(S) \
    case S: if constexpr (!mode_is_r8(MODE)) raster_kernel<S, MODE, VIS><<<blocks, RW_THREADS, 0, st>>>(rp, a, b, c); break;
    switch (shader) {
        HANA_RASTER_CASE(HANA_SHADER_SHADOW)
        HANA_RASTER_CASE(HANA_SHADER_BLINN)
        HANA_RASTER_CASE(HANA_SHADER_NORMALMAP)
        HANA_RASTER_CASE(HANA_SHADER_GROUND)
        HANA_RASTER_CASE(HANA_SHADER_TOON)
        HANA_RASTER_CASE(HANA_SHADER_TEXTURE)
        HANA_RASTER_CASE(HANA_SHADER_TEXTURE_LIGHT)
        default: return fail(HANA_E_UNSUPPORTED, "unknown shader id");
    }
#undef HANA_RASTER_CASE
    return HANA_OK;
}

template <int MODE, bool VIS>
static int raster_occupancy(int shader) {
    int occ = 0;
    if constexpr (mode_is_r8(MODE)) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, raster_kernel<HANA_SHADER_SHADOW, MODE, VIS>, RW_THREADS, 0);
        return occ;
    }
#define HANA_OCC_CASE(S) \
    case S: if constexpr (!mode_is_r8(MODE)) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, raster_kernel<S, MODE, VIS>, RW_THREADS, 0); break;
    switch (shader) {
        HANA_OCC_CASE(HANA_SHADER_SHADOW)
        HANA_OCC_CASE(HANA_SHADER_BLINN)
        HANA_OCC_CASE(HANA_SHADER_NORMALMAP)
        HANA_OCC_CASE(HANA_SHADER_GROUND)
        HANA_OCC_CASE(HANA_SHADER_TOON)
        HANA_OCC_CASE(HANA_SHADER_TEXTURE)
        HANA_OCC_CASE(HANA_SHADER_TEXTURE_LIGHT)
    }
#undef HANA_OCC_CASE
    return occ;
}

static int run_pass(hana_ctx* ctx, const PassDesc& d) {
    if (d.shader < 0 || d.shader >= HANA_SHADER_COUNT) return fail(HANA_E_UNSUPPORTED, "unknown shader id");
    if (mode_is_r8(d.mode) && d.shader != HANA_SHADER_SHADOW)
        return fail(HANA_E_INVALID, "R8 targets take the ShadowShader only");
    const int tiles_x = (d.W + TILE - 1) / TILE, tiles_y = (d.H + TILE - 1) / TILE;
    const size_t n_tiles = (size_t)tiles_x * tiles_y;
    if (tiles_x > 1024 || tiles_y > 1024) return fail(HANA_E_INVALID, "target larger than 16384 pixels on a side");
    if (d.n_frames < 1 || d.n_frames > 4096) return fail(HANA_E_INVALID, "1..4096 frames per batch");
    const int nfaces = d.model->ncorners / 3;
    Scratch& sc = scratch_of(ctx, d.scratch);
    cudaStream_t st = d.stream ? d.stream : ctx->stream;
    if (!d.pipelined) ctx->serial_dirty = true;
    const size_t tiles_total = n_tiles * d.n_frames;
    if (tiles_total > 0xFFFFFFF0ull) return fail(HANA_E_INVALID, "frames x tiles exceeds 32 bits");

    uint32_t tri_cap = std::max<uint32_t>(ctx->tri_cap_hint, (uint32_t)(nfaces + nfaces / 8 + 64));
    if (d.dbg_v2f) tri_cap = std::min(tri_cap, std::max<uint32_t>(d.dbg_cap, 1u));
    PassParams p;
    memset(&p, 0, sizeof(p));
    RasterParams rp;
    bool lists_ready = false;
    if (d.phase == 2) { /* raster phase of a split lazy pass: the lists are where phase 1 put them */
        p = sc.saved_p;
        tri_cap = p.tri_cap;
        PassCounters& c = *sc.counters_host;
        memset(&c, 0, sizeof(c));
        c.tri_needed = tri_cap;
        c.pool_used = 1;
        c.n_work = 0xFFFFFFFFu;
        lists_ready = true;
    }
    for (int attempt = 0; attempt < 4 && !lists_ready; attempt++) {
        /* capacities */
        size_t tri_total = (size_t)tri_cap * d.n_frames;
        HANA_TRY(grow(&sc.tri_rec, &sc.tri_total, tri_total * 4, ctx));
        HANA_TRY(grow(&sc.tri_attr, &sc.attr_total, tri_total * MAX_ATTR_QUADS, ctx));
        HANA_TRY(grow(&sc.tri_bbox, &sc.bbox_total, tri_total, ctx));
        HANA_TRY(grow(&sc.tri_count, &sc.frames_cap, (size_t)d.n_frames * (TRI_COUNT_WAYS + 1), ctx));
        const int tile_rows = (int)((n_tiles + 31) / 32);
        const size_t tiles_pad_total = (size_t)tile_rows * 32 * d.n_frames;
        HANA_TRY(grow(&sc.tile_arrays, &sc.tile_arr_cap, tiles_pad_total * 3 + tiles_total, ctx));
        /* Dense meshes (at least one face per 16 pixels: BASELINE.json configs[3] has 1.2 per pixel) resolve their micro-
         * triangles through a per-pixel visibility buffer instead of the tile lists (setup_kernel). Not for RenderBuffer
         * draws: the target's existing depth takes part there. */
        const bool use_vis = d.mode != MODE_RMW && !d.setup_only && (size_t)nfaces * 16 >= (size_t)d.W * d.H && !getenv("HANA_NO_VIS");
        if (use_vis) HANA_TRY(grow(&sc.vis, &sc.vis_cap, (size_t)d.W * d.H * d.n_frames, ctx));
        const size_t n_groups = ((size_t)tri_cap + 31) / 32 * d.n_frames;
        if (use_vis) HANA_TRY(grow(&sc.group_listed, &sc.group_cap, n_groups, ctx));
        HANA_TRY(grow(&sc.work, &sc.work_cap, tiles_total, ctx));
        if (!sc.tile_recs || sc.pool_cap / 4 < ctx->pool_hint)
            HANA_TRY(grow(&sc.tile_recs, &sc.pool_cap, std::max<size_t>(std::max<size_t>(tri_total * 2, 65536), ctx->pool_hint) * 4, ctx));

        p.posu = d.model->posu;
        p.nrmv = d.model->nrmv;
        p.nfaces = nfaces;
        p.n_frames = d.n_frames;
        p.W = d.W;
        p.H = d.H;
        p.tiles_x = tiles_x;
        p.tiles_y = tiles_y;
        p.n_tiles = (int)n_tiles;
        p.band_y0 = d.band_count > 0 ? std::min(d.band_first, tiles_y) : 0;
        p.band_y1 = d.band_count > 0 ? std::min(d.band_first + d.band_count, tiles_y) : tiles_y;
        p.uniforms = d.uniforms;
        p.tri_rec = sc.tri_rec;
        p.tri_attr = sc.tri_attr;
        p.tri_bbox = sc.tri_bbox;
        p.tri_cap = tri_cap;
        p.tri_count = sc.tri_count;
        p.tri_extra = sc.tri_count + (size_t)d.n_frames * TRI_COUNT_WAYS;
        p.tile_count = sc.tile_arrays;
        p.tile_cursor = sc.tile_arrays + tiles_pad_total;
        p.tile_micro = sc.tile_arrays + 2 * tiles_pad_total;
        p.tile_offset = sc.tile_arrays + 3 * tiles_pad_total;
        p.vis = use_vis ? sc.vis : nullptr;
        p.group_listed = use_vis ? sc.group_listed : nullptr;
        p.tile_rows = tile_rows;
        p.tile_pad = tile_rows * 32;
        p.tile_recs = sc.tile_recs;
        p.pool_cap = (uint32_t)std::min<size_t>(sc.pool_cap / 4, 0xFFFFFFFFull);
        p.work = sc.work;
        p.counters = sc.counters;
        p.overflow = d.overflow;
        p.dbg_v2f = d.dbg_v2f;

        /* zero: counters, per-frame triangle counts, tile counts + cursors (contiguous) */
        CU_TRY(cudaMemsetAsync(sc.counters, 0, sizeof(PassCounters), st));
        CU_TRY(cudaMemsetAsync(sc.tri_count, 0, sizeof(uint32_t) * d.n_frames * (TRI_COUNT_WAYS + 1), st));
        CU_TRY(cudaMemsetAsync(sc.tile_arrays, 0, sizeof(uint32_t) * 3 * tiles_pad_total, st));
        if (use_vis) CU_TRY(cudaMemsetAsync(sc.vis, 0xFF, sizeof(unsigned long long) * (size_t)d.W * d.H * d.n_frames, st));
        if (use_vis) CU_TRY(cudaMemsetAsync(sc.group_listed, 0, sizeof(uint32_t) * n_groups, st));

        cudaEvent_t ea, eb;
        if (nfaces > 0) {
            dim3 grid((nfaces + SETUP_THREADS - 1) / SETUP_THREADS, d.n_frames);
            prof_begin(ctx, PROF_SETUP, &ea, &eb, st);
#define HANA_SETUP_CASE(S) \
    case S: setup_kernel<S><<<grid, SETUP_THREADS, 0, st>>>(p); break;
            switch (d.shader) {
                HANA_SETUP_CASE(HANA_SHADER_SHADOW)
                HANA_SETUP_CASE(HANA_SHADER_BLINN)
                HANA_SETUP_CASE(HANA_SHADER_NORMALMAP)
                HANA_SETUP_CASE(HANA_SHADER_GROUND)
                HANA_SETUP_CASE(HANA_SHADER_TOON)
                HANA_SETUP_CASE(HANA_SHADER_TEXTURE)
                HANA_SETUP_CASE(HANA_SHADER_TEXTURE_LIGHT)
            }
#undef HANA_SETUP_CASE
            prof_end(ctx, PROF_SETUP, ea, eb, st);
            ctx->launches++;
            CU_TRY(cudaGetLastError());
        }
        /* (triangle, tile) pairs: gridDim.z warps share the rounds of 32 triangles when triangles cover many tiles each */
        const unsigned pair_z = (unsigned)std::min<size_t>(16, std::max<size_t>(1, n_tiles / (size_t)std::max(nfaces, 1) / 2));
        if (nfaces > 0) {
            dim3 grid((tri_cap + 255) / 256, d.n_frames, pair_z);
            prof_begin(ctx, PROF_SETUP, &ea, &eb, st);
            pairs_kernel<false><<<grid, 256, 0, st>>>(p);
            prof_end(ctx, PROF_SETUP, ea, eb, st);
            ctx->launches++;
            CU_TRY(cudaGetLastError());
        }
        prof_begin(ctx, PROF_SCAN, &ea, &eb, st);
        scan_kernel<<<dim3((unsigned)((n_tiles + SCAN_CHUNK - 1) / SCAN_CHUNK), (unsigned)d.n_frames), SCAN_THREADS, 0, st>>>(p);
        prof_end(ctx, PROF_SCAN, ea, eb, st);
        ctx->launches++;
        CU_TRY(cudaGetLastError());
        if (d.lazy) { /* no read-back: run with what we have, the sweep verifies afterwards */
            PassCounters& c = *sc.counters_host;
            memset(&c, 0, sizeof(c));
            c.tri_needed = tri_cap;
            c.pool_used = 1;
            c.n_work = 0xFFFFFFFFu;
            /* statistics (triangle counts, list sizes) are not read back per batch: hana_sweep_stats fetches them from the
             * scratch on demand */
            if (d.tri_cap_used) *d.tri_cap_used = tri_cap;
            if (d.pool_cap_used) *d.pool_cap_used = p.pool_cap;
            lists_ready = true;
            break;
        }
        CU_TRY(cudaMemcpyAsync(sc.counters_host, sc.counters, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const PassCounters& c = *sc.counters_host;
        if (c.tri_needed > tri_cap) {
            if (d.dbg_v2f) return fail(HANA_E_OVERFLOW, "stage output capacity too small: " + std::to_string(c.tri_needed));
            tri_cap = c.tri_needed + c.tri_needed / 8 + 64;
            ctx->tri_cap_hint = std::max(ctx->tri_cap_hint, tri_cap);
            continue; /* triangles were dropped: redo setup with room for all of them */
        }
        if ((size_t)c.pool_used > sc.pool_cap / 4) {
            HANA_TRY(grow(&sc.tile_recs, &sc.pool_cap, ((size_t)c.pool_used + c.pool_used / 8) * 4, ctx));
            p.tile_recs = sc.tile_recs;
            p.pool_cap = (uint32_t)std::min<size_t>(sc.pool_cap / 4, 0xFFFFFFFFull);
        }
        lists_ready = true;
    }
    if (!lists_ready) return fail(HANA_E_OVERFLOW, "triangle capacity still exceeded after retries");
    const PassCounters cnt = *sc.counters_host;
    if (d.counters_out) *d.counters_out = cnt;
    if (d.tri_counts_out && !d.lazy) {
        std::vector<uint32_t> ways((size_t)d.n_frames * (TRI_COUNT_WAYS + 1));
        CU_TRY(cudaMemcpyAsync(ways.data(), sc.tri_count, sizeof(uint32_t) * ways.size(), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        d.tri_counts_out->assign(d.n_frames, 0u);
        for (int fi = 0; fi < d.n_frames; fi++)
            for (int w = 0; w < TRI_COUNT_WAYS; w++) (*d.tri_counts_out)[fi] += ways[(size_t)fi * TRI_COUNT_WAYS + w];
        if (d.slots_out) {
            d.slots_out->assign(d.n_frames, 0u);
            for (int fi = 0; fi < d.n_frames; fi++) (*d.slots_out)[fi] = (uint32_t)nfaces + ways[(size_t)d.n_frames * TRI_COUNT_WAYS + fi];
        }
    }
    if (d.setup_only) return HANA_OK;

    cudaEvent_t ea, eb;
    if (cnt.pool_used > 0 && d.phase != 2) {
        const unsigned pair_z = (unsigned)std::min<size_t>(16, std::max<size_t>(1, n_tiles / (size_t)std::max(nfaces, 1) / 2));
        dim3 grid((std::min(cnt.tri_needed, tri_cap) + 255) / 256, d.n_frames, pair_z);
        if (grid.x > 0) {
            prof_begin(ctx, PROF_FILL, &ea, &eb, st);
            pairs_kernel<true><<<grid, 256, 0, st>>>(p);
            prof_end(ctx, PROF_FILL, ea, eb, st);
            ctx->launches++;
            CU_TRY(cudaGetLastError());
        }
    }
    if (d.phase == 1) {
        sc.saved_p = p;
        return HANA_OK;
    }
    if (d.mode == MODE_RMW && cnt.n_work == 0) return HANA_OK; /* nothing covers anything */

    memset(&rp, 0, sizeof(rp));
    rp.p = p;
    rp.color = d.color;
    rp.depth = d.depth;
    rp.frame_stride = d.frame_stride;
    rp.shadow_out = d.shadow_out;
    rp.shadow_out_pitch = d.shadow_out_pitch;
    rp.shadow_out_frame_stride = d.shadow_out_frame_stride;
    rp.clear_color = d.clear_color;
    rp.clear_depth = d.clear_depth;
    rp.use_tma = (ctx->use_tma && d.tma_ok) ? 1 : 0;
    rp.diffuse = d.diffuse;
    rp.normal = d.normal;
    rp.shadow = d.shadow;
    rp.shadow_frame_stride = d.shadow_frame_stride;
    rp.primid = d.primid;
    rp.pixels_covered = d.pixels_covered;

    /* MODE_SHADOW_R8 packs {shadow byte << 24 | triangle slot}: a pass that may emit more triangles per frame than 24 bits
     * address takes the WIDE variant (hana_kernels.cuh; ctx->r8_slot_limit is R8_SLOT_LIMIT unless a test lowered it) */
    const int mode = (d.mode == MODE_SHADOW_R8 && p.tri_cap > ctx->r8_slot_limit) ? (int)MODE_SHADOW_R8_WIDE : d.mode;
    const bool vis = p.vis != nullptr; /* never with MODE_RMW */
    int& occ = ctx->occ[d.shader][mode + (vis ? N_RASTER_MODES : 0)];
    if (occ == 0) {
        occ = mode == MODE_RMW          ? raster_occupancy<MODE_RMW, false>(d.shader)
              : mode == MODE_CLEAR_FOLD ? (vis ? raster_occupancy<MODE_CLEAR_FOLD, true>(d.shader) : raster_occupancy<MODE_CLEAR_FOLD, false>(d.shader))
              : mode == MODE_SHADOW_R8  ? (vis ? raster_occupancy<MODE_SHADOW_R8, true>(HANA_SHADER_SHADOW)
                                               : raster_occupancy<MODE_SHADOW_R8, false>(HANA_SHADER_SHADOW))
                                        : (vis ? raster_occupancy<MODE_SHADOW_R8_WIDE, true>(HANA_SHADER_SHADOW)
                                               : raster_occupancy<MODE_SHADOW_R8_WIDE, false>(HANA_SHADER_SHADOW));
        cudaGetLastError();
        if (occ <= 0) occ = 1;
    }
    size_t want = ((size_t)cnt.n_work + RW_WARPS - 1) / RW_WARPS; /* one tile per warp at a time */
    if (d.mode != MODE_RMW) want = std::max<size_t>(want, (tiles_total + 1023) / 1024); /* empty targets still need their clears */
    int blocks = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)ctx->sm_count * occ));
    static const CUtensorMap dummy = {};
    const CUtensorMap& ta = d.tm_color ? *d.tm_color : dummy;
    const CUtensorMap& tb = d.tm_depth ? *d.tm_depth : dummy;
    const CUtensorMap& tc = d.tm_r8 ? *d.tm_r8 : dummy;
    prof_begin(ctx, d.prof_kind, &ea, &eb, st);
    int r = mode == MODE_RMW          ? launch_raster<MODE_RMW, false>(st, d.shader, blocks, rp, ta, tb, tc)
            : mode == MODE_CLEAR_FOLD ? (vis ? launch_raster<MODE_CLEAR_FOLD, true>(st, d.shader, blocks, rp, ta, tb, tc)
                                             : launch_raster<MODE_CLEAR_FOLD, false>(st, d.shader, blocks, rp, ta, tb, tc))
            : mode == MODE_SHADOW_R8  ? (vis ? launch_raster<MODE_SHADOW_R8, true>(st, HANA_SHADER_SHADOW, blocks, rp, ta, tb, tc)
                                             : launch_raster<MODE_SHADOW_R8, false>(st, HANA_SHADER_SHADOW, blocks, rp, ta, tb, tc))
                                      : (vis ? launch_raster<MODE_SHADOW_R8_WIDE, true>(st, HANA_SHADER_SHADOW, blocks, rp, ta, tb, tc)
                                             : launch_raster<MODE_SHADOW_R8_WIDE, false>(st, HANA_SHADER_SHADOW, blocks, rp, ta, tb, tc));
    if (mode == MODE_SHADOW_R8_WIDE) ctx->wide_r8_launches++;
    prof_end(ctx, d.prof_kind, ea, eb, st);
    HANA_TRY(r);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return HANA_OK;
}

static int upload_uniforms(hana_ctx* ctx, const HanaUniforms* host, int n, HanaUniforms* raw_dev, DevUniforms* dev, cudaStream_t st = nullptr) {
    if (!st) st = ctx->stream;
    if (host) CU_TRY(cudaMemcpyAsync(raw_dev, host, sizeof(HanaUniforms) * n, cudaMemcpyHostToDevice, st));
    cudaEvent_t a, b;
    prof_begin(ctx, PROF_BEGIN, &a, &b, st);
    begin_kernel<<<(n + 63) / 64, 64, 0, st>>>(raw_dev, dev, n);
    prof_end(ctx, PROF_BEGIN, a, b, st);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return HANA_OK;
}

static DevTexture dev_tex(const hana_texture* t) {
    DevTexture d{nullptr, 0, 0};
    if (t) {
        d.texels = t->texels;
        d.w = t->w;
        d.h = t->h;
    }
    return d;
}

/* graphics_draw_triangle (graphics.cpp:378-407) into an existing target. */
static int draw_rmw(hana_ctx* ctx, hana_rb* rb, const hana_model* model, int shader, const hana_texture* diffuse,
                    const hana_texture* normal, const hana_rb* shadow_map, uint32_t* primid_dev, bool want_stats) {
    PassDesc d;
    d.shader = shader;
    d.mode = MODE_RMW;
    d.n_frames = 1;
    d.W = rb->w;
    d.H = rb->h;
    d.model = model;
    d.uniforms = ctx->u_dev;
    d.color = rb->color;
    d.depth = rb->depth;
    d.frame_stride = (size_t)rb->w * rb->h;
    d.tm_color = &rb->tm_color;
    d.tm_depth = &rb->tm_depth;
    d.tma_ok = rb->tma_ok;
    d.diffuse = dev_tex(diffuse);
    d.normal = dev_tex(normal);
    if (shadow_map) {
        d.shadow.base = reinterpret_cast<const uint8_t*>(shadow_map->color);
        d.shadow.w = shadow_map->w;
        d.shadow.h = shadow_map->h;
        d.shadow.pitch = shadow_map->w * 4;
        d.shadow.stride = 4;
    }
    d.primid = primid_dev;
    d.prof_kind = shader == HANA_SHADER_SHADOW ? PROF_RASTER_SHADOW : PROF_RASTER_MAIN;
    PassCounters cnt;
    std::vector<uint32_t> tri_counts;
    d.counters_out = &cnt;
    d.tri_counts_out = &tri_counts;
    if (want_stats) {
        CU_TRY(cudaMemsetAsync(ctx->stat_pixels, 0, 4, ctx->stream));
        d.pixels_covered = ctx->stat_pixels;
    }
    HANA_TRY(run_pass(ctx, d));
    HanaStats& s = ctx->last_stats;
    memset(&s, 0, sizeof(s));
    s.faces_in = (uint32_t)(model->ncorners / 3);
    s.tris_out = tri_counts.empty() ? 0 : tri_counts[0];
    s.tile_refs = cnt.pool_used;
    s.tiles_touched = cnt.tiles_touched;
    return HANA_OK;
}

static int check_draw_args(hana_ctx* ctx, const hana_model* model, const HanaUniforms* u, int shader) {
    if (!ctx || !model || !u) return fail(HANA_E_INVALID, "NULL argument");
    if (model->ctx != ctx) return fail(HANA_E_INVALID, "model belongs to another context");
    if (shader < 0 || shader >= HANA_SHADER_COUNT)
        return fail(HANA_E_UNSUPPORTED, "shader id outside the closed device set (IShader subclasses cannot run on the device)");
    return HANA_OK;
}

extern "C" int hana_draw(hana_ctx* ctx, hana_rb* rb, const hana_model* model, int shader_id, const HanaUniforms* uniforms,
                         const hana_texture* diffuse, const hana_texture* normal, const hana_rb* shadow_map) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!rb) return fail(HANA_E_INVALID, "rb is NULL");
    HANA_TRY(use_device(ctx));
    HANA_TRY(upload_uniforms(ctx, uniforms, 1, ctx->u_raw, ctx->u_dev));
    return draw_rmw(ctx, rb, model, shader_id, diffuse, normal, uniforms->enable_shadow ? shadow_map : nullptr, nullptr, true);
}

extern "C" int hana_draw_primid(hana_ctx* ctx, hana_rb* rb, const hana_model* model, int shader_id,
                                const HanaUniforms* uniforms, const hana_texture* diffuse, const hana_texture* normal,
                                const hana_rb* shadow_map, uint32_t* out_primid_host) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!rb || !out_primid_host) return fail(HANA_E_INVALID, "NULL argument");
    HANA_TRY(use_device(ctx));
    size_t n = (size_t)rb->w * rb->h;
    uint32_t* pid = nullptr;
    CU_TRY(cudaMalloc(&pid, n * 4));
    int r = launch_fill32(ctx, pid, 0xFFFFFFFFu, n);
    if (r == HANA_OK) r = upload_uniforms(ctx, uniforms, 1, ctx->u_raw, ctx->u_dev);
    if (r == HANA_OK)
        r = draw_rmw(ctx, rb, model, shader_id, diffuse, normal, uniforms->enable_shadow ? shadow_map : nullptr, pid, true);
    if (r == HANA_OK) {
        cudaError_t e = cudaMemcpyAsync(out_primid_host, pid, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) r = fail(HANA_E_CUDA, cudaGetErrorString(e));
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(pid);
    return r;
}

extern "C" int hana_draw_model(hana_ctx* ctx, hana_rb* frame, hana_rb* shadow_map, const hana_model* model, int shader_id,
                               const HanaUniforms* uniforms, const hana_texture* diffuse, const hana_texture* normal) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!frame) return fail(HANA_E_INVALID, "frame is NULL");
    if (uniforms->enable_shadow && !shadow_map) return fail(HANA_E_INVALID, "enable_shadow needs a shadow map target");
    HANA_TRY(use_device(ctx));
    HANA_TRY(upload_uniforms(ctx, uniforms, 1, ctx->u_raw, ctx->u_dev));
    if (uniforms->enable_shadow) { /* scene.h:73-88 */
        HANA_TRY(draw_rmw(ctx, shadow_map, model, HANA_SHADER_SHADOW, nullptr, nullptr, nullptr, nullptr, false));
    }
    HANA_TRY(draw_rmw(ctx, frame, model, shader_id, diffuse, normal, uniforms->enable_shadow ? shadow_map : nullptr, nullptr,
                      true)); /* scene.h:90-91 */
    if (uniforms->enable_shadow) { /* scene.h:94-98: Color::Black has a = 255 -> (uchar)(255*255) -> 1 on x86-64 (App. D5) */
        HANA_TRY(hana_rb_clear_color(shadow_map, 0, 0, 0, 1));
        HANA_TRY(hana_rb_clear_depth(shadow_map, 3.402823466e+38f));
    }
    return HANA_OK;
}

extern "C" int hana_last_stats(hana_ctx* ctx, HanaStats* out) {
    if (!ctx || !out) return fail(HANA_E_INVALID, "NULL argument");
    HANA_TRY(use_device(ctx));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t px = 0;
    CU_TRY(cudaMemcpy(&px, ctx->stat_pixels, 4, cudaMemcpyDeviceToHost));
    ctx->last_stats.pixels_covered = px;
    *out = ctx->last_stats;
    return HANA_OK;
}

/* ---- batched sweep ------------------------------------------------------------------ */
extern "C" int hana_sweep_create(hana_ctx* ctx, int width, int height, int max_frames, hana_sweep** out) {
    if (!ctx || !out) return fail(HANA_E_INVALID, "NULL argument");
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535 || max_frames < 1 || max_frames > 4096)
        return fail(HANA_E_INVALID, "sweep size out of range");
    HANA_TRY(use_device(ctx));
    hana_sweep* s = new hana_sweep();
    s->ctx = ctx;
    s->w = width;
    s->h = height;
    s->max_frames = max_frames;
    s->last_frames = 0;
    s->last_clear_depth = 0.f;
    memset(s->band, 0, sizeof(s->band));
    memset(&s->last_stats, 0, sizeof(s->last_stats));
    size_t n = (size_t)width * height;
    s->shadow_pitch = (width + 15) / 16 * 16;
    s->shadow_frame_bytes = (size_t)s->shadow_pitch * ((height + 15) / 16 * 16);
    s->color = nullptr; s->depth = nullptr; s->shadow_r8 = nullptr; s->u_raw = nullptr; s->u_dev = nullptr; s->u_dev_buf = nullptr; s->overflow_buf = nullptr;
    s->checksums = nullptr; s->pix_counts = nullptr; s->overflow = nullptr; s->pin = nullptr; s->tri_counts_pin = nullptr;
    s->ev_render = nullptr; s->ev_copy = nullptr; s->copy_in_flight = false; s->present_buf = nullptr; s->present_cap = 0;
    for (auto& e : s->ev_check) e = nullptr;
    s->check_tail = s->n_checks = s->next_slot = 0;
    s->overflow_batches = 0;
    s->rerender_count = 0;
    s->stats_lazy = false;
    s->stats_scratch = 0;
    cudaError_t e = cudaMalloc(&s->color, n * 4 * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->depth, n * 4 * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->shadow_r8, s->shadow_frame_bytes * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->u_raw, sizeof(HanaUniforms) * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->u_dev_buf, sizeof(DevUniforms) * max_frames * 2);
    s->u_dev = s->u_dev_buf;
    if (e == cudaSuccess) e = cudaMalloc(&s->checksums, sizeof(unsigned long long) * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->pix_counts, sizeof(uint32_t) * max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->overflow_buf, sizeof(OverflowRecord) * 2);
    s->overflow = s->overflow_buf;
    if (e == cudaSuccess) e = cudaMallocHost(&s->pin, sizeof(*s->pin));
    if (e == cudaSuccess) e = cudaMallocHost(&s->tri_counts_pin, sizeof(uint32_t) * 2 * max_frames * TRI_COUNT_WAYS);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_render, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming);
    for (auto& ev : s->ev_check)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        hana_sweep_destroy(s);
        return fail(HANA_E_CUDA, std::string("allocation failed for the frame ring: ") + cudaGetErrorString(e));
    }
    memset(s->pin, 0, sizeof(*s->pin));
    s->tma_ok = false;
    if (ctx->encode && width % 4 == 0) {
        int r1 = make_tensor_map(ctx, &s->tm_color, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, s->color, width, height, max_frames,
                                 (uint64_t)width * 4, (uint64_t)n * 4);
        int r2 = make_tensor_map(ctx, &s->tm_depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s->depth, width, height, max_frames,
                                 (uint64_t)width * 4, (uint64_t)n * 4);
        int r3 = make_tensor_map(ctx, &s->tm_r8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, s->shadow_r8, width, height, max_frames,
                                 (uint64_t)s->shadow_pitch, (uint64_t)s->shadow_frame_bytes);
        s->tma_ok = (r1 == HANA_OK && r2 == HANA_OK && r3 == HANA_OK);
    }
    ctx->sweeps.push_back(s);
    *out = s;
    return HANA_OK;
}
extern "C" int hana_sweep_destroy(hana_sweep* s) {
    if (!s) return HANA_OK;
    hana_ctx* ctx = s->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(s->color); cudaFree(s->depth); cudaFree(s->shadow_r8); cudaFree(s->u_raw); cudaFree(s->u_dev_buf);
    cudaFree(s->checksums); cudaFree(s->pix_counts); cudaFree(s->overflow_buf); cudaFree(s->present_buf);
    if (s->pin) cudaFreeHost(s->pin);
    if (s->tri_counts_pin) cudaFreeHost(s->tri_counts_pin);
    if (s->ev_render) cudaEventDestroy(s->ev_render);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    cudaFree(s->tga.packed); cudaFree(s->tga.ebits); cudaFree(s->tga.recs); cudaFree(s->tga.counts); cudaFree(s->tga.batch_sums);
    cudaFree(s->tga.batch_offs); cudaFree(s->tga.sizes); cudaFree(s->tga.offsets);
    if (s->tga.meta_pin) cudaFreeHost(s->tga.meta_pin);
    if (s->tga.ev) cudaEventDestroy(s->tga.ev);
    for (auto ev : s->ev_check)
        if (ev) cudaEventDestroy(ev);
    if (ctx->host_sweep == s) ctx->host_sweep = nullptr;
    ctx->sweeps.erase(std::remove(ctx->sweeps.begin(), ctx->sweeps.end(), s), ctx->sweeps.end());
    delete s;
    return HANA_OK;
}

/* Both passes of every frame of the batch (scene.h:73-91). lazy: launch everything without reading anything back. */
enum { SWEEP_PASS_SHADOW = 1, SWEEP_PASS_MAIN = 2, SWEEP_PASS_BOTH = 3 };
static int sweep_render_passes(hana_sweep* s, const hana_model* model, int shader_id, int enable_shadow, int n_frames,
                               const hana_texture* diffuse, const hana_texture* normal, const uint8_t clear_rgba[4],
                               float clear_depth, bool lazy, int passes = SWEEP_PASS_BOTH, int pipe_parity = -1, bool share_shadow = false) {
    hana_ctx* ctx = s->ctx;
    s->shadow_shared = share_shadow && enable_shadow && passes == SWEEP_PASS_BOTH;
    if (s->shadow_shared) s->shadow_shared_batches++;
    const bool pipe = pipe_parity >= 0; /* lazy, both passes: binned on bin_stream / side_stream with this parity's scratch sets */
    PassCounters cnt[2];
    memset(cnt, 0, sizeof(cnt));
    PassDesc shadow_desc;
    if (enable_shadow && (passes & SWEEP_PASS_SHADOW)) { /* scene.h:73-88, into the internal R8 maps */
        PassDesc d;
        d.shader = HANA_SHADER_SHADOW;
        d.mode = MODE_SHADOW_R8;
        d.n_frames = s->shadow_shared ? 1 : n_frames; /* every frame's light and model are frame 0's: one map */
        d.W = s->w;
        d.H = s->h;
        d.model = model;
        d.uniforms = s->u_dev;
        d.tm_r8 = &s->tm_r8;
        d.tma_ok = s->tma_ok;
        d.shadow_out = s->shadow_r8;
        d.shadow_out_pitch = s->shadow_pitch;
        d.shadow_out_frame_stride = s->shadow_frame_bytes;
        d.clear_depth = 3.402823466e+38f; /* scene.h:97 */
        d.prof_kind = PROF_RASTER_SHADOW;
        d.counters_out = &cnt[0];
        d.tri_counts_out = &s->last_tri_counts[0];
        d.lazy = lazy;
        d.overflow = s->overflow;
        d.tri_counts_pinned = s->tri_counts_pin;
        d.counters_pinned = &s->pin->counters[0];
        d.tri_cap_used = &s->pending.tri_cap;
        d.pool_cap_used = &s->pending.pool_cap;
        d.band_first = s->band[0][0];
        d.band_count = s->band[0][1];
        if (lazy && passes == SWEEP_PASS_BOTH) { /* fork: the main pass is binned on the side stream while this one is binned here */
            cudaStream_t bst = pipe ? ctx->bin_stream : ctx->stream;
            CU_TRY(cudaEventRecord(ctx->ev_fork, bst));
            CU_TRY(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
            d.phase = 1;
            if (pipe) {
                d.scratch = 2 * pipe_parity;
                d.stream = ctx->bin_stream;
                d.pipelined = true;
            }
        }
        HANA_TRY(run_pass(ctx, d));
        if (pipe) CU_TRY(cudaEventRecord(ctx->ev_bin[pipe_parity][0], ctx->bin_stream));
        shadow_desc = d;
    }
    if (!(passes & SWEEP_PASS_MAIN)) {
        s->last_frames = n_frames;
        return HANA_OK;
    }
    PassDesc d;
    d.shader = shader_id;
    d.band_first = s->band[1][0];
    d.band_count = s->band[1][1];
    d.mode = MODE_CLEAR_FOLD;
    d.n_frames = n_frames;
    d.W = s->w;
    d.H = s->h;
    d.model = model;
    d.uniforms = s->u_dev;
    d.color = s->color;
    d.depth = s->depth;
    d.frame_stride = (size_t)s->w * s->h;
    d.tm_color = &s->tm_color;
    d.tm_depth = &s->tm_depth;
    d.tma_ok = s->tma_ok;
    d.clear_color = (uint32_t)clear_rgba[0] | ((uint32_t)clear_rgba[1] << 8) | ((uint32_t)clear_rgba[2] << 16) |
                    ((uint32_t)clear_rgba[3] << 24);
    d.clear_depth = clear_depth;
    d.diffuse = dev_tex(diffuse);
    d.normal = dev_tex(normal);
    if (enable_shadow) {
        d.shadow.base = s->shadow_r8;
        d.shadow.w = s->w;
        d.shadow.h = s->h;
        d.shadow.pitch = s->shadow_pitch;
        d.shadow.stride = 1;
        d.shadow_frame_stride = s->shadow_shared ? 0 : s->shadow_frame_bytes;
    }
    d.prof_kind = PROF_RASTER_MAIN;
    d.counters_out = &cnt[1];
    d.tri_counts_out = &s->last_tri_counts[1];
    d.lazy = lazy;
    d.overflow = s->overflow;
    d.tri_counts_pinned = s->tri_counts_pin + (size_t)s->max_frames * TRI_COUNT_WAYS;
    d.counters_pinned = &s->pin->counters[1];
    uint32_t tc = 0, pc = 0;
    d.tri_cap_used = &tc;
    d.pool_cap_used = &pc;
    if (lazy && enable_shadow && passes == SWEEP_PASS_BOTH) {
        d.phase = 1;
        d.scratch = pipe ? 2 * pipe_parity + 1 : 1;
        d.stream = ctx->side_stream;
        d.pipelined = pipe;
        HANA_TRY(run_pass(ctx, d));
        cudaEvent_t joined = pipe ? ctx->ev_bin[pipe_parity][1] : ctx->ev_join;
        CU_TRY(cudaEventRecord(joined, ctx->side_stream));
        if (pipe) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_bin[pipe_parity][0], 0));
        shadow_desc.phase = 2;
        shadow_desc.stream = nullptr;
        HANA_TRY(run_pass(ctx, shadow_desc)); /* shadow rasteriser */
        CU_TRY(cudaStreamWaitEvent(ctx->stream, joined, 0));
        d.phase = 2;
        d.stream = nullptr;
        HANA_TRY(run_pass(ctx, d));           /* main rasteriser */
    } else {
        HANA_TRY(run_pass(ctx, d));
    }
    if (lazy) { /* the smaller of the two passes' capacities is what both must fit */
        if (!enable_shadow || tc < s->pending.tri_cap) s->pending.tri_cap = tc;
        if (!enable_shadow || pc < s->pending.pool_cap) s->pending.pool_cap = pc;
    }
    s->last_frames = n_frames;
    s->last_clear_depth = clear_depth;
    s->stats_lazy = lazy;
    s->stats_scratch = d.scratch;
    memset(&s->last_stats, 0, sizeof(s->last_stats));
    s->last_stats.faces_in = (uint32_t)(model->ncorners / 3);
    s->last_stats.tile_refs = cnt[1].pool_used;
    s->last_stats.tiles_touched = cnt[1].tiles_touched;
    return HANA_OK;
}

static int grow_scratch_for(hana_ctx* ctx, const OverflowRecord& need) {
    ctx->tri_cap_hint = std::max(ctx->tri_cap_hint, need.tri_needed + need.tri_needed / 8 + 64);
    ctx->pool_hint = std::max(ctx->pool_hint, (size_t)need.pool_needed + need.pool_needed / 4); /* headroom for the other frames of an orbit */
    for (Scratch* sp : {&ctx->sc, &ctx->sc2, &ctx->sc3, &ctx->sc4})
        if (ctx->pool_hint > sp->pool_cap / 4 && (sp == &ctx->sc || sp->tile_recs))
            HANA_TRY(grow(&sp->tile_recs, &sp->pool_cap, ctx->pool_hint * 4, ctx));
    return HANA_OK;
}

/* Renders that were superseded before anybody synchronised with them: examine those that have completed (all of them
 * if wait_all). Their frames are gone, so a batch that ran out of scratch cannot be rendered again; it is counted and
 * the scratch is grown so that the batches that follow fit. */
static int sweep_poll_checks(hana_sweep* s, bool wait_all) {
    while (s->n_checks > 0) {
        const hana_sweep::Check c = s->checks[s->check_tail];
        if (wait_all) {
            CU_TRY(cudaEventSynchronize(s->ev_check[c.slot]));
        } else {
            const cudaError_t q = cudaEventQuery(s->ev_check[c.slot]);
            if (q == cudaErrorNotReady) break;
            CU_TRY(q);
        }
        s->check_tail = (s->check_tail + 1) % hana_sweep::CHECK_RING;
        s->n_checks--;
        const OverflowRecord need = s->pin->need[c.slot];
        if (need.tri_needed > c.tri_cap || need.pool_needed > c.pool_cap) {
            s->overflow_batches++;
            HANA_TRY(grow_scratch_for(s->ctx, need));
        }
    }
    return HANA_OK;
}

/* The sweep's synchronisation point: waits for its last render, and if that render ran out of scratch (triangles
 * or tile-list records were dropped), grows the scratch and renders the batch again, this time with read-backs. */
static int sweep_verify(hana_sweep* s) {
    hana_ctx* ctx = s->ctx;
    HANA_TRY(sweep_poll_checks(s, false));
    if (!s->pending.active) return HANA_OK;
    hana_sweep::Pending pd = s->pending;
    CU_TRY(cudaEventSynchronize(s->ev_check[pd.slot]));
    s->pending.active = false;
    const OverflowRecord need = s->pin->need[pd.slot];
    if (need.tri_needed <= pd.tri_cap && need.pool_needed <= pd.pool_cap) return HANA_OK;
    HANA_TRY(grow_scratch_for(ctx, need));
    s->rerender_count++;
    return sweep_render_passes(s, pd.model, pd.shader, pd.enable_shadow, pd.n_frames, pd.diffuse, pd.normal, pd.clear_rgba,
                               pd.clear_depth, false, SWEEP_PASS_BOTH, -1, pd.share_shadow);
}

static int sweep_render_common(hana_sweep* s, const hana_model* model, int shader_id, const HanaUniforms* host_uniforms,
                               int enable_shadow, int n_frames, const hana_texture* diffuse, const hana_texture* normal,
                               const uint8_t clear_rgba[4], float clear_depth, const void* uniforms_dev = nullptr) {
    hana_ctx* ctx = s->ctx;
    /* the previous render's frames are about to be overwritten: it cannot be rendered again, but what it needed is
     * still examined once it has completed (sweep_poll_checks); its copies must be out */
    if (s->pending.active) {
        if (s->n_checks == hana_sweep::CHECK_RING) { /* ring full: wait for the oldest */
            CU_TRY(cudaEventSynchronize(s->ev_check[s->checks[s->check_tail].slot]));
        }
        HANA_TRY(sweep_poll_checks(s, false));
        hana_sweep::Check& c = s->checks[(s->check_tail + s->n_checks) % hana_sweep::CHECK_RING];
        c.slot = s->pending.slot;
        c.tri_cap = s->pending.tri_cap;
        c.pool_cap = s->pending.pool_cap;
        s->n_checks++;
        s->pending.active = false;
    } else {
        HANA_TRY(sweep_poll_checks(s, false));
    }
    if (s->copy_in_flight) {
        CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
        s->copy_in_flight = false;
    }
    /* two-pass submissions on the context's own stream are pipelined (hana_ctx: bin_stream); a caller's stream
     * (hana_ctx_set_stream) keeps everything in that stream's order */
    const bool pipe = enable_shadow && ctx->pipeline && ctx->stream == ctx->own_stream;
    int parity = -1;
    cudaStream_t ust = ctx->stream; /* where the uniforms are made ready */
    if (pipe) {
        parity = ctx->parity;
        ctx->parity ^= 1;
        ust = ctx->bin_stream;
        if (ctx->raster_done_valid[parity]) CU_TRY(cudaStreamWaitEvent(ust, ctx->ev_raster_done[parity], 0));
        if (ctx->serial_dirty) { /* draws / single-pass sweeps since the last pipelined submission used sets 0 / 1 on `stream` */
            CU_TRY(cudaEventRecord(ctx->ev_serial, ctx->stream));
            CU_TRY(cudaStreamWaitEvent(ust, ctx->ev_serial, 0));
            ctx->serial_dirty = false;
        }
    }
    s->u_dev = s->u_dev_buf + (size_t)(pipe ? parity : 0) * s->max_frames;
    s->overflow = s->overflow_buf + (pipe ? parity : 0);
    if (uniforms_dev && uniforms_dev != s->u_raw)
        CU_TRY(cudaMemcpyAsync(s->u_raw, uniforms_dev, sizeof(HanaUniforms) * n_frames, cudaMemcpyDeviceToDevice, ust));
    HANA_TRY(upload_uniforms(ctx, host_uniforms, n_frames, s->u_raw, s->u_dev, ust));
    CU_TRY(cudaMemsetAsync(s->overflow, 0, sizeof(OverflowRecord), ust));
    const bool dirty_before = ctx->serial_dirty;
    /* static light (hana_sweep_set_shadow_reuse): the ShadowShader pass reads light_vp * model only (IShader.cpp:170) */
    bool share = false;
    if (s->shadow_reuse && enable_shadow && host_uniforms && n_frames > 1) {
        share = true;
        for (int i = 1; i < n_frames && share; i++)
            share = memcmp(host_uniforms[i].light_vp, host_uniforms[0].light_vp, sizeof(host_uniforms[0].light_vp)) == 0 &&
                    memcmp(host_uniforms[i].model, host_uniforms[0].model, sizeof(host_uniforms[0].model)) == 0;
    }
    HANA_TRY(sweep_render_passes(s, model, shader_id, enable_shadow, n_frames, diffuse, normal, clear_rgba, clear_depth, true,
                                 SWEEP_PASS_BOTH, parity, share));
    if (pipe) ctx->serial_dirty = dirty_before;
    const int slot = s->next_slot;
    s->next_slot = (slot + 1) % (hana_sweep::CHECK_RING + 1);
    HANA_TRY(post_to_host(ctx, &s->pin->need[slot], s->overflow, sizeof(OverflowRecord), ctx->stream));
    CU_TRY(cudaEventRecord(s->ev_check[slot], ctx->stream));
    CU_TRY(cudaEventRecord(s->ev_render, ctx->stream));
    if (pipe) {
        CU_TRY(cudaEventRecord(ctx->ev_raster_done[parity], ctx->stream));
        ctx->raster_done_valid[parity] = true;
    }
    hana_sweep::Pending& pd = s->pending;
    pd.active = true;
    pd.slot = slot;
    pd.model = model;
    pd.shader = shader_id;
    pd.enable_shadow = enable_shadow;
    pd.n_frames = n_frames;
    pd.diffuse = diffuse;
    pd.normal = normal;
    memcpy(pd.clear_rgba, clear_rgba, 4);
    pd.clear_depth = clear_depth;
    pd.share_shadow = share;
    return HANA_OK;
}

extern "C" int hana_sweep_set_shadow_reuse(hana_sweep* s, int enable) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    s->shadow_reuse = enable != 0;
    return HANA_OK;
}

extern "C" int hana_sweep_render(hana_sweep* s, const hana_model* model, int shader_id, const HanaUniforms* uniforms,
                                 int n_frames, const hana_texture* diffuse, const hana_texture* normal,
                                 const uint8_t clear_rgba[4], float clear_depth) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    HANA_TRY(check_draw_args(s->ctx, model, uniforms, shader_id));
    if (!clear_rgba) return fail(HANA_E_INVALID, "clear_rgba is NULL");
    if (n_frames < 1 || n_frames > s->max_frames) return fail(HANA_E_INVALID, "n_frames outside 1..max_frames");
    for (int i = 1; i < n_frames; i++)
        if ((uniforms[i].enable_shadow != 0) != (uniforms[0].enable_shadow != 0))
            return fail(HANA_E_INVALID, "all frames of a batch must agree on enable_shadow");
    HANA_TRY(use_device(s->ctx));
    return sweep_render_common(s, model, shader_id, uniforms, uniforms[0].enable_shadow != 0, n_frames, diffuse, normal,
                               clear_rgba, clear_depth);
}

extern "C" int hana_sweep_render_dev(hana_sweep* s, const hana_model* model, int shader_id, const void* uniforms_dev,
                                     int enable_shadow, int n_frames, const hana_texture* diffuse,
                                     const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth) {
    if (!s || !uniforms_dev || !model || !clear_rgba) return fail(HANA_E_INVALID, "NULL argument");
    if (model->ctx != s->ctx) return fail(HANA_E_INVALID, "model belongs to another context");
    if (shader_id < 0 || shader_id >= HANA_SHADER_COUNT) return fail(HANA_E_UNSUPPORTED, "shader id outside the device set");
    if (n_frames < 1 || n_frames > s->max_frames) return fail(HANA_E_INVALID, "n_frames outside 1..max_frames");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    return sweep_render_common(s, model, shader_id, nullptr, enable_shadow != 0, n_frames, diffuse, normal, clear_rgba,
                               clear_depth, uniforms_dev);
}

extern "C" int hana_sweep_uniforms_dev(hana_sweep* s, void** out) {
    if (!s || !out) return fail(HANA_E_INVALID, "NULL argument");
    *out = s->u_raw;
    return HANA_OK;
}

/* ---- single frame split by screen tiles over several GPUs ---------------------------------------
 * Each GPU owns a band of tile rows per pass. Pass 1 leaves this GPU's band of the R8 shadow maps in HBM; the caller
 * exchanges the bands (the maps of all GPUs must be complete before any pass-2 fragment is shaded: IShader.h:124 looks
 * up arbitrary light-space texels), then pass 2 renders this GPU's band of the frame. The exchange itself is the
 * launcher's (sharding.py: NCCL all-gather over NVLink): this library has no inter-process state. */
extern "C" int hana_sweep_set_bands(hana_sweep* s, int shadow_row_first, int shadow_row_count, int main_row_first,
                                    int main_row_count) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (shadow_row_first < 0 || shadow_row_count < 0 || main_row_first < 0 || main_row_count < 0)
        return fail(HANA_E_INVALID, "negative tile row range");
    s->band[0][0] = shadow_row_first;
    s->band[0][1] = shadow_row_count;
    s->band[1][0] = main_row_first;
    s->band[1][1] = main_row_count;
    return HANA_OK;
}

extern "C" int hana_sweep_render_pass(hana_sweep* s, int pass, const hana_model* model, int shader_id,
                                      const HanaUniforms* uniforms, int n_frames, const hana_texture* diffuse,
                                      const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (pass != HANA_PASS_SHADOW && pass != HANA_PASS_MAIN) return fail(HANA_E_INVALID, "pass is HANA_PASS_SHADOW or HANA_PASS_MAIN");
    HANA_TRY(check_draw_args(s->ctx, model, uniforms, pass == HANA_PASS_SHADOW ? HANA_SHADER_SHADOW : shader_id));
    if (!clear_rgba) return fail(HANA_E_INVALID, "clear_rgba is NULL");
    if (n_frames < 1 || n_frames > s->max_frames) return fail(HANA_E_INVALID, "n_frames outside 1..max_frames");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s));
    s->pending.active = false;
    if (s->copy_in_flight) {
        CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
        s->copy_in_flight = false;
    }
    HANA_TRY(upload_uniforms(ctx, uniforms, n_frames, s->u_raw, s->u_dev));
    /* with read-backs (not lazy): a split frame is one large frame, not a stream of small ones */
    return sweep_render_passes(s, model, shader_id, uniforms[0].enable_shadow != 0, n_frames, diffuse, normal, clear_rgba, clear_depth,
                               false, pass == HANA_PASS_SHADOW ? SWEEP_PASS_SHADOW : SWEEP_PASS_MAIN);
}

/* The same pass queued without any read-back (scratch capacities are the context's current ones), so that a caller
 * whose collectives run on the context's stream (hana_ctx_set_stream) never synchronises the host inside a frame. */
extern "C" int hana_sweep_render_pass_async(hana_sweep* s, int pass, const hana_model* model, int shader_id,
                                            const HanaUniforms* uniforms, int n_frames, const hana_texture* diffuse,
                                            const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (pass != HANA_PASS_SHADOW && pass != HANA_PASS_MAIN) return fail(HANA_E_INVALID, "pass is HANA_PASS_SHADOW or HANA_PASS_MAIN");
    HANA_TRY(check_draw_args(s->ctx, model, uniforms, pass == HANA_PASS_SHADOW ? HANA_SHADER_SHADOW : shader_id));
    if (!clear_rgba) return fail(HANA_E_INVALID, "clear_rgba is NULL");
    if (n_frames < 1 || n_frames > s->max_frames) return fail(HANA_E_INVALID, "n_frames outside 1..max_frames");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    const bool shadowed = uniforms[0].enable_shadow != 0;
    const bool first = pass == HANA_PASS_SHADOW || !shadowed; /* first pass of the frame: new uniforms, needs start from zero */
    if (first) {
        HANA_TRY(sweep_verify(s));
        s->pending.active = false;
        if (s->copy_in_flight) {
            CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
            s->copy_in_flight = false;
        }
        HANA_TRY(upload_uniforms(ctx, uniforms, n_frames, s->u_raw, s->u_dev));
        CU_TRY(cudaMemsetAsync(s->overflow, 0, sizeof(OverflowRecord), ctx->stream));
        s->pending.tri_cap = s->pending.pool_cap = 0xFFFFFFFFu;
    }
    if (pass == HANA_PASS_SHADOW && !shadowed) return HANA_OK; /* scene.h:73: no shadow pass */
    HANA_TRY(sweep_render_passes(s, model, shader_id, shadowed, n_frames, diffuse, normal, clear_rgba, clear_depth, true,
                                 pass == HANA_PASS_SHADOW ? SWEEP_PASS_SHADOW : SWEEP_PASS_MAIN));
    s->pending.active = false; /* nothing to re-render behind the caller's back: hana_sweep_passes_ok reports instead */
    return HANA_OK;
}

/* Waits for the passes queued by hana_sweep_render_pass_async. *ok = 0 if one of them ran out of triangle or tile-list
 * scratch (fragments were dropped): the scratch has been grown, queue the frame again. */
extern "C" int hana_sweep_passes_ok(hana_sweep* s, int* ok) {
    if (!s || !ok) return fail(HANA_E_INVALID, "NULL argument");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    CU_TRY(cudaMemcpyAsync(&s->pin->need[hana_sweep::SLOT_PASSES_OK], s->overflow, sizeof(OverflowRecord), cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    const OverflowRecord need = s->pin->need[hana_sweep::SLOT_PASSES_OK];
    *ok = (need.tri_needed <= s->pending.tri_cap && need.pool_needed <= s->pending.pool_cap) ? 1 : 0;
    if (!*ok) HANA_TRY(grow_scratch_for(ctx, need));
    return HANA_OK;
}

/* Batches of this sweep that ran out of triangle or tile-list scratch AND were overwritten by a later submission before
 * they could be rendered again (only possible when renders are queued back to back without any synchronising call in
 * between). Waits for every render queued so far. 0 = every frame handed out so far was complete. */
extern "C" int hana_sweep_overflow_count(hana_sweep* s, uint64_t* out) {
    if (!s || !out) return fail(HANA_E_INVALID, "NULL argument");
    HANA_TRY(use_device(s->ctx));
    HANA_TRY(sweep_poll_checks(s, true));
    HANA_TRY(sweep_verify(s));
    *out = s->overflow_batches;
    return HANA_OK;
}

extern "C" int hana_sweep_shadow_ptrs(hana_sweep* s, void** r8_dev, int* pitch_bytes, size_t* frame_stride_bytes) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    HANA_TRY(use_device(s->ctx));
    HANA_TRY(sweep_verify(s));
    CU_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (r8_dev) *r8_dev = s->shadow_r8;
    if (pitch_bytes) *pitch_bytes = s->shadow_pitch;
    if (frame_stride_bytes) *frame_stride_bytes = s->shadow_frame_bytes;
    return HANA_OK;
}

extern "C" int hana_sweep_download(hana_sweep* s, int frame, uint8_t* color_rgba, float* depth) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (frame < 0 || frame >= s->max_frames) return fail(HANA_E_INVALID, "frame out of range");
    HANA_TRY(use_device(s->ctx));
    HANA_TRY(sweep_verify(s));
    size_t n = (size_t)s->w * s->h;
    if (color_rgba)
        CU_TRY(cudaMemcpyAsync(color_rgba, s->color + n * frame, n * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (depth) CU_TRY(cudaMemcpyAsync(depth, s->depth + n * frame, n * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU_TRY(cudaStreamSynchronize(s->ctx->stream));
    return HANA_OK;
}
/* Copies on the context's copy stream, after the sweep's render: the next batch (into another sweep) overlaps them.
 * Complete at hana_sync(). */
extern "C" int hana_sweep_download_async(hana_sweep* s, int first, int count, uint8_t* color_rgba_pinned, float* depth_pinned) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (first < 0 || count < 0 || first + count > s->max_frames) return fail(HANA_E_INVALID, "frame range out of bounds");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s)); /* host waits for the render (not for any copy) so that only checked frames leave */
    CU_TRY(cudaEventRecord(s->ev_render, ctx->stream));
    CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, s->ev_render, 0));
    size_t n = (size_t)s->w * s->h;
    if (color_rgba_pinned)
        CU_TRY(cudaMemcpyAsync(color_rgba_pinned, s->color + n * first, n * 4 * count, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (depth_pinned)
        CU_TRY(cudaMemcpyAsync(depth_pinned, s->depth + n * first, n * 4 * count, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU_TRY(cudaEventRecord(s->ev_copy, ctx->copy_stream));
    s->copy_in_flight = true;
    return HANA_OK;
}
/* The presentable surface of frames [first, first+count) (win32.cpp:348-370): rows top-down, B,G,R[,255]. The
 * conversion runs on the device after the sweep's render; the copy to dst_host (pinned for overlap) runs on the copy
 * stream and is complete after hana_sync(). dst_dev_out (optional) receives the device address of the surfaces. */
extern "C" int hana_sweep_present(hana_sweep* s, int first, int count, int format, uint8_t* dst_host, void** dst_dev_out) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (first < 0 || count < 1 || first + count > s->max_frames) return fail(HANA_E_INVALID, "frame range out of bounds");
    if (format != HANA_PRESENT_BGRA8 && format != HANA_PRESENT_BGR8) return fail(HANA_E_INVALID, "unknown present format");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s));
    const size_t bytes = (size_t)s->w * s->h * (format == HANA_PRESENT_BGRA8 ? 4 : 3) * count;
    if (s->copy_in_flight) { /* the previous copy reads present_buf / the ring */
        CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
        s->copy_in_flight = false;
    }
    if (bytes > s->present_cap) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->copy_stream));
        cudaFree(s->present_buf);
        s->present_buf = nullptr;
        s->present_cap = 0;
        CU_TRY(cudaMalloc(&s->present_buf, bytes));
        s->present_cap = bytes;
    }
    dim3 grid((unsigned)((s->w + 1023) / 1024), (unsigned)s->h, (unsigned)count);
    cudaEvent_t a, b;
    prof_begin(ctx, PROF_OTHER, &a, &b);
    present_kernel<<<grid, 256, 0, ctx->stream>>>(s->color, (size_t)s->w * s->h, first, s->w, s->h, format, s->present_buf);
    prof_end(ctx, PROF_OTHER, a, b);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    if (dst_dev_out) *dst_dev_out = s->present_buf;
    if (dst_host) {
        CU_TRY(cudaEventRecord(s->ev_render, ctx->stream));
        CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, s->ev_render, 0));
        CU_TRY(cudaMemcpyAsync(dst_host, s->present_buf, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CU_TRY(cudaEventRecord(s->ev_copy, ctx->copy_stream));
        s->copy_in_flight = true;
    }
    return HANA_OK;
}
/* ---- RLE TGA files of the frames, made on the device (hana_tga.cuh) ---------------------------------------------- */
static int tga_encode_launch(hana_sweep* s, int first, int count) {
    hana_ctx* ctx = s->ctx;
    hana_sweep::Tga& t = s->tga;
    const size_t npix = (size_t)s->w * s->h;
    const size_t worst = tga_worst_bytes(npix);
    const TgaLayout L = tga_layout(s->w, s->h);
    if (s->copy_in_flight) { /* the previous fetch reads the packed buffer */
        CU_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copy, 0));
        s->copy_in_flight = false;
    }
    if (!t.sizes) {
        CU_TRY(cudaMalloc(&t.sizes, sizeof(unsigned long long) * s->max_frames));
        CU_TRY(cudaMalloc(&t.offsets, sizeof(unsigned long long) * (s->max_frames + 1)));
        CU_TRY(cudaMallocHost(&t.meta_pin, sizeof(unsigned long long) * (2 * (size_t)s->max_frames + 1)));
        CU_TRY(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
    }
    if ((size_t)count > t.cap_frames || worst != t.worst_bytes) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->copy_stream));
        cudaFree(t.packed); cudaFree(t.ebits); cudaFree(t.recs); cudaFree(t.counts); cudaFree(t.batch_sums); cudaFree(t.batch_offs);
        t.packed = nullptr; t.ebits = nullptr; t.recs = nullptr; t.counts = nullptr; t.batch_sums = nullptr; t.batch_offs = nullptr;
        t.cap_frames = 0;
        CU_TRY(cudaMalloc(&t.packed, worst * count));
        CU_TRY(cudaMalloc(&t.ebits, sizeof(uint32_t) * L.estride * count));
        CU_TRY(cudaMalloc(&t.recs, sizeof(TgaRec) * L.rstride * count));
        CU_TRY(cudaMalloc(&t.counts, L.rstride * count));
        CU_TRY(cudaMalloc(&t.batch_sums, sizeof(uint32_t) * (size_t)L.nbatch * count));
        CU_TRY(cudaMalloc(&t.batch_offs, sizeof(unsigned long long) * (size_t)L.nbatch * count));
        t.cap_frames = (size_t)count;
        t.worst_bytes = worst;
    }
    cudaEvent_t a, b;
    prof_begin(ctx, PROF_OTHER, &a, &b);
    const dim3 egrid((unsigned)((L.estride + 32 * (TGA_E_THREADS / 32) - 1) / (32 * (TGA_E_THREADS / 32))), (unsigned)count);
    if (s->w % 4 == 0 && s->w >= 128)
        tga_ebits_kernel<true><<<egrid, TGA_E_THREADS, 0, ctx->stream>>>(s->color, npix, first, s->w, s->h, t.ebits, L);
    else
        tga_ebits_kernel<false><<<egrid, TGA_E_THREADS, 0, ctx->stream>>>(s->color, npix, first, s->w, s->h, t.ebits, L);
    tga_structure_kernel<<<(unsigned)count, TGA_B_THREADS, 0, ctx->stream>>>(t.ebits, t.recs, L);
    const dim3 cgrid((unsigned)((L.nbatch + TGA_C_THREADS / 32 - 1) / (TGA_C_THREADS / 32)), (unsigned)count);
    tga_count_kernel<<<cgrid, TGA_C_THREADS, 0, ctx->stream>>>(t.ebits, t.recs, t.counts, t.batch_sums, L);
    tga_scan_kernel<<<(unsigned)count, TGA_S_THREADS, 0, ctx->stream>>>(t.batch_sums, t.batch_offs, t.sizes, L.nbatch);
    tga_offsets_kernel<<<1, 32, 0, ctx->stream>>>(t.sizes, t.offsets, count);
    tga_write_kernel<<<cgrid, TGA_C_THREADS, 0, ctx->stream>>>(s->color, npix, first, s->w, s->h, t.ebits, t.recs, t.counts, t.batch_offs, t.sizes,
                                                               t.offsets, t.packed, L);
    prof_end(ctx, PROF_OTHER, a, b);
    ctx->launches += 6;
    CU_TRY(cudaGetLastError());
    HANA_TRY(post_to_host(ctx, t.meta_pin, t.offsets, sizeof(unsigned long long) * (count + 1), ctx->stream));
    HANA_TRY(post_to_host(ctx, t.meta_pin + s->max_frames + 1, t.sizes, sizeof(unsigned long long) * count, ctx->stream));
    CU_TRY(cudaEventRecord(t.ev, ctx->stream));
    t.first = first;
    t.count = count;
    t.rerender_seen = s->rerender_count;
    t.valid = true;
    return HANA_OK;
}

/* Queues the encoding of frames [first, first+count) of the sweep's last render as RLE-compressed 24-bit TGA files
 * (TGAImage::write_tga_file(rle = true), tgaimage.cpp:145-246: byte-identical). Asynchronous, ordered after the render on
 * the context's stream; nothing is read back. */
extern "C" int hana_sweep_encode_tga(hana_sweep* s, int first, int count) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    if (first < 0 || count < 1 || first + count > s->max_frames) return fail(HANA_E_INVALID, "frame range out of bounds");
    if (s->w > 32767 || s->h > 32767) return fail(HANA_E_INVALID, "TGA: image larger than 32767 pixels on a side");
    HANA_TRY(use_device(s->ctx));
    return tga_encode_launch(s, first, count);
}

/* Delivers the files of the last hana_sweep_encode_tga: waits for the encoder (and for the render it followed: a batch
 * that ran out of scratch is rendered and encoded again first), writes offsets[0..count] (file f occupies
 * dst_host[offsets[f], offsets[f] + sizes[f]); starts are 16-byte aligned; offsets[count] = bytes used) and sizes[0..count)
 * (may be NULL), and copies the bytes to dst_host (pinned memory overlaps the next batch) on the copy stream: complete
 * after hana_sync() / hana_timer_stop(). HANA_E_OVERFLOW if dst_capacity is too small (hana_last_error names the need). */
extern "C" int hana_sweep_fetch_tga(hana_sweep* s, uint8_t* dst_host, size_t dst_capacity, uint64_t* offsets, uint64_t* sizes) {
    if (!s || !dst_host || !offsets) return fail(HANA_E_INVALID, "NULL argument");
    hana_ctx* ctx = s->ctx;
    hana_sweep::Tga& t = s->tga;
    if (!t.valid) return fail(HANA_E_INVALID, "no hana_sweep_encode_tga to fetch");
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s));
    if (t.rerender_seen != s->rerender_count) HANA_TRY(tga_encode_launch(s, t.first, t.count)); /* the frames were rendered again */
    CU_TRY(cudaEventSynchronize(t.ev));
    const unsigned long long total = t.meta_pin[t.count];
    if (total > dst_capacity)
        return fail(HANA_E_OVERFLOW, "TGA files need " + std::to_string(total) + " bytes, destination holds " + std::to_string(dst_capacity));
    for (int f = 0; f <= t.count; f++) offsets[f] = t.meta_pin[f];
    if (sizes)
        for (int f = 0; f < t.count; f++) sizes[f] = t.meta_pin[s->max_frames + 1 + f];
    CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, t.ev, 0));
    CU_TRY(cudaMemcpyAsync(dst_host, t.packed, total, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU_TRY(cudaEventRecord(s->ev_copy, ctx->copy_stream));
    s->copy_in_flight = true;
    return HANA_OK;
}

extern "C" int hana_sweep_device_ptrs(hana_sweep* s, void** color_dev, void** depth_dev, size_t* frame_stride_pixels) {
    if (!s) return fail(HANA_E_INVALID, "sweep is NULL");
    HANA_TRY(use_device(s->ctx));
    HANA_TRY(sweep_verify(s));
    if (color_dev) *color_dev = s->color;
    if (depth_dev) *depth_dev = s->depth;
    if (frame_stride_pixels) *frame_stride_pixels = (size_t)s->w * s->h;
    return HANA_OK;
}
extern "C" int hana_sweep_checksums(hana_sweep* s, int n_frames, uint64_t* out_host) {
    if (!s || !out_host) return fail(HANA_E_INVALID, "NULL argument");
    if (n_frames < 1 || n_frames > s->max_frames) return fail(HANA_E_INVALID, "n_frames out of range");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s));
    CU_TRY(cudaMemsetAsync(s->checksums, 0, sizeof(unsigned long long) * n_frames, ctx->stream));
    size_t n = (size_t)s->w * s->h;
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 256), n_frames);
    cudaEvent_t a, b;
    prof_begin(ctx, PROF_OTHER, &a, &b);
    checksum_kernel<<<grid, 256, 0, ctx->stream>>>(s->color, s->depth, n, n, s->checksums);
    prof_end(ctx, PROF_OTHER, a, b);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out_host, s->checksums, sizeof(uint64_t) * n_frames, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return HANA_OK;
}
extern "C" int hana_sweep_stats(hana_sweep* s, int frame, HanaStats* out) {
    if (!s || !out) return fail(HANA_E_INVALID, "NULL argument");
    if (frame < 0 || frame >= s->last_frames) return fail(HANA_E_INVALID, "frame outside the last batch");
    hana_ctx* ctx = s->ctx;
    HANA_TRY(use_device(ctx));
    HANA_TRY(sweep_verify(s));
    CU_TRY(cudaMemsetAsync(s->pix_counts, 0, sizeof(uint32_t) * s->last_frames, ctx->stream));
    size_t n = (size_t)s->w * s->h;
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 256), s->last_frames);
    count_written_kernel<<<grid, 256, 0, ctx->stream>>>(s->depth, n, n, s->last_clear_depth, s->pix_counts);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    uint32_t px = 0;
    CU_TRY(cudaMemcpyAsync(&px, s->pix_counts + frame, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *out = s->last_stats;
    out->pixels_covered = px;
    if (s->stats_lazy) { /* the main pass's counters are still where it left them */
        const Scratch& sc = scratch_of(ctx, s->stats_scratch);
        PassCounters c;
        uint32_t ways[TRI_COUNT_WAYS];
        CU_TRY(cudaMemcpy(&c, sc.counters, sizeof(c), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(ways, sc.tri_count + (size_t)frame * TRI_COUNT_WAYS, sizeof(ways), cudaMemcpyDeviceToHost));
        out->tile_refs = c.pool_used;
        out->tiles_touched = c.tiles_touched;
        out->tris_out = 0;
        for (int w = 0; w < TRI_COUNT_WAYS; w++) out->tris_out += ways[w];
    } else if (frame < (int)s->last_tri_counts[1].size()) {
        out->tris_out = s->last_tri_counts[1][frame];
    }
    return HANA_OK;
}

/* ---- host-buffer entry point (what the drop-in graphics shim and the e2e bench call) ---- */
extern "C" int hana_draw_model_host(hana_ctx* ctx, int width, int height, uint8_t* frame_color_rgba, float* frame_depth,
                                    const hana_model* model, int shader_id, const HanaUniforms* uniforms,
                                    const hana_texture* diffuse, const hana_texture* normal, int assume_cleared,
                                    const uint8_t clear_rgba[4], float clear_depth) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!frame_color_rgba || !frame_depth) return fail(HANA_E_INVALID, "frame buffers are NULL");
    HANA_TRY(use_device(ctx));
    if (assume_cleared) {
        if (!clear_rgba) return fail(HANA_E_INVALID, "clear_rgba is NULL");
        hana_sweep*& sw = ctx->host_sweep;
        if (sw && (sw->w != width || sw->h != height)) {
            hana_sweep_destroy(sw);
            sw = nullptr;
        }
        if (!sw) HANA_TRY(hana_sweep_create(ctx, width, height, 1, &sw));
        HANA_TRY(hana_sweep_render(sw, model, shader_id, uniforms, 1, diffuse, normal, clear_rgba, clear_depth));
        return hana_sweep_download(sw, 0, frame_color_rgba, frame_depth);
    }
    hana_rb*& fr = ctx->host_frame;
    hana_rb*& sh = ctx->host_shadow;
    if (fr && (fr->w != width || fr->h != height)) {
        hana_rb_destroy(fr);
        fr = nullptr;
        hana_rb_destroy(sh);
        sh = nullptr;
    }
    if (!fr) HANA_TRY(hana_rb_create(ctx, width, height, &fr));
    if (uniforms->enable_shadow && !sh) {
        HANA_TRY(hana_rb_create(ctx, width, height, &sh));
        HANA_TRY(hana_rb_clear_color(sh, 0, 0, 0, 1));
        HANA_TRY(hana_rb_clear_depth(sh, 3.402823466e+38f));
    }
    HANA_TRY(hana_rb_upload(fr, frame_color_rgba, frame_depth));
    HANA_TRY(hana_draw_model(ctx, fr, sh, model, shader_id, uniforms, diffuse, normal));
    return hana_rb_download(fr, frame_color_rgba, frame_depth);
}

extern "C" int hana_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(HANA_E_INVALID, "out is NULL");
    CU_TRY(cudaMallocHost(out, bytes ? bytes : 1));
    return HANA_OK;
}
extern "C" int hana_host_free(void* p) {
    if (p) CU_TRY(cudaFreeHost(p));
    return HANA_OK;
}

/* ---- stage-level entry points ------------------------------------------------------- */
extern "C" int hana_stage_vertex(hana_ctx* ctx, const hana_model* model, int shader_id, const HanaUniforms* uniforms,
                                 float* out_v2f_host) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!out_v2f_host) return fail(HANA_E_INVALID, "out is NULL");
    HANA_TRY(use_device(ctx));
    HANA_TRY(upload_uniforms(ctx, uniforms, 1, ctx->u_raw, ctx->u_dev));
    int n = model->ncorners;
    if (n == 0) return HANA_OK;
    float* out = nullptr;
    CU_TRY(cudaMalloc(&out, (size_t)n * V2F_N * 4));
    vertex_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(model->posu, model->nrmv, n, shader_id, ctx->u_dev, out);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_v2f_host, out, (size_t)n * V2F_N * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(out);
    if (e != cudaSuccess) return fail(HANA_E_CUDA, cudaGetErrorString(e));
    return HANA_OK;
}

extern "C" int hana_stage_setup(hana_ctx* ctx, const hana_model* model, int shader_id, const HanaUniforms* uniforms, int width,
                                int height, int capacity, uint32_t* out_order, float* out_v2f, int* out_count) {
    HANA_TRY(check_draw_args(ctx, model, uniforms, shader_id));
    if (!out_order || !out_v2f || !out_count || capacity < 1) return fail(HANA_E_INVALID, "bad output arguments");
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return fail(HANA_E_INVALID, "bad target size");
    HANA_TRY(use_device(ctx));
    HANA_TRY(upload_uniforms(ctx, uniforms, 1, ctx->u_raw, ctx->u_dev));
    float* dbg = nullptr;
    CU_TRY(cudaMalloc(&dbg, (size_t)capacity * 39 * 4));
    PassDesc d;
    d.shader = shader_id;
    d.mode = MODE_RMW;
    d.n_frames = 1;
    d.W = width;
    d.H = height;
    d.model = model;
    d.uniforms = ctx->u_dev;
    d.dbg_v2f = dbg;
    d.dbg_cap = (uint32_t)capacity;
    d.setup_only = true;
    std::vector<uint32_t> tri_counts, slots;
    d.tri_counts_out = &tri_counts;
    d.slots_out = &slots;
    uint32_t saved_hint = ctx->tri_cap_hint;
    ctx->tri_cap_hint = 0;
    int r = run_pass(ctx, d);
    ctx->tri_cap_hint = saved_hint;
    if (r != HANA_OK) {
        cudaFree(dbg);
        return r;
    }
    const uint32_t nslots = slots.empty() ? 0 : std::min<uint32_t>(slots[0], (uint32_t)capacity); /* slot = face index, or beyond for clipped fans */
    std::vector<float> recs((size_t)nslots * 16);
    std::vector<float> v2f((size_t)nslots * 39);
    cudaError_t e = cudaSuccess;
    if (nslots) {
        e = cudaMemcpy(recs.data(), ctx->sc.tri_rec, sizeof(float) * 16 * nslots, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(v2f.data(), dbg, (size_t)nslots * 39 * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dbg);
    if (e != cudaSuccess) return fail(HANA_E_CUDA, cudaGetErrorString(e));
    std::vector<uint2> bbs(nslots);
    if (nslots && cudaMemcpy(bbs.data(), ctx->sc.tri_bbox, sizeof(uint2) * nslots, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(HANA_E_CUDA, "cudaMemcpy of the slot table failed");
    std::vector<uint32_t> idx;
    for (uint32_t i = 0; i < nslots; i++) /* slots of faces that emitted nothing carry the empty pixel range */
        if (bbs[i].y != DEAD_BBY) idx.push_back(i);
    const uint32_t n = (uint32_t)idx.size();
    auto key_of = [&](uint32_t i) {
        uint32_t k;
        memcpy(&k, &recs[(size_t)i * 16 + 11], 4);
        return k;
    };
    std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key_of(a) < key_of(b); });
    for (uint32_t i = 0; i < n; i++) {
        out_order[i] = key_of(idx[i]);
        memcpy(out_v2f + (size_t)i * 39, v2f.data() + (size_t)idx[i] * 39, 39 * 4);
    }
    *out_count = (int)n;
    return HANA_OK;
}
