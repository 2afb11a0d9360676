/*
 * hana_b200.h — C ABI of the B200-native rasterisation path behind
 * Hana-SoftwareRenderer's draw API.
 *
 * Every entry point replaces one piece of the reference's hot path
 * (all citations relative to /root/reference/Hana-SoftwareRenderer/):
 *
 *   hana_draw            <- void graphics_draw_triangle(DrawData*)     graphics.h:15, graphics.cpp:378-407
 *   hana_draw_model      <- DrawModel::draw(Camera*,RenderBuffer*,bool) scene.h:53-99 (shadow pass, main pass, shadow clear)
 *   hana_sweep_*         <- the main loop calling Scene::tick per frame main.cpp:87-156 (batched: N frames per submission)
 *   hana_rb_*            <- class RenderBuffer                          renderbuffer.h:5-22, renderbuffer.cpp:3-76
 *   hana_model_upload    <- Model::vert/normal/uv(iface,nth) gathers    model.cpp:70,99,108 as read by graphics.cpp:383-385
 *   hana_texture_upload  <- TGAImage storage read by TGAImage::get      tgaimage.cpp:248-253
 *   HanaUniforms         <- struct ShaderData + struct Material         IShader.h:7-32
 *   HANA_SHADER_*        <- the IShader subclasses                      IShader.h:132-165, IShader.cpp
 *
 * Plain pointers and sizes only; no C++ or torch types. All functions return
 * 0 on success, a negative HANA_E_* code otherwise; hana_last_error() gives
 * the message of the last failure on the calling thread. There is no CPU
 * fallback: without a CUDA device every compute entry point fails.
 *
 * Threading: one context per GPU; calls on one context must be serialised by
 * the caller (the reference is single-threaded: SURVEY.md §8b).
 */
#ifndef HANA_B200_H
#define HANA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HANA_API __attribute__((visibility("default")))
#else
#define HANA_API
#endif

#define HANA_OK 0
#define HANA_E_INVALID (-1)   /* bad argument */
#define HANA_E_CUDA (-2)      /* CUDA runtime/driver error (message has the cudaError string) */
#define HANA_E_NODEVICE (-3)  /* no usable CUDA device: there is no CPU fallback */
#define HANA_E_UNSUPPORTED (-4) /* e.g. an IShader subclass outside the closed device set */
#define HANA_E_OVERFLOW (-5)  /* internal capacity exceeded even after the automatic retry */

/* The closed set of device shaders (IShader.cpp). User-defined IShader
 * subclasses cannot run on the device and are rejected by the shim. */
#define HANA_SHADER_SHADOW 0        /* ShadowShader            IShader.cpp:170-180 */
#define HANA_SHADER_BLINN 1         /* BlinnShader             IShader.cpp:85-109  */
#define HANA_SHADER_NORMALMAP 2     /* NormalMapShader         IShader.cpp:117-162 */
#define HANA_SHADER_GROUND 3        /* GroundShader            IShader.cpp:5-15    */
#define HANA_SHADER_TOON 4          /* ToonShader              IShader.cpp:23-39   */
#define HANA_SHADER_TEXTURE 5       /* TextureShader           IShader.cpp:47-57   */
#define HANA_SHADER_TEXTURE_LIGHT 6 /* TextureWithLightShader  IShader.cpp:65-77   */
#define HANA_SHADER_COUNT 7

/* POD image of ShaderData (IShader.h:17-32) + Material scalars (IShader.h:7-15).
 * Matrices are row-major float[16] exactly as Matrix4x4 rows[4] lie in memory
 * (matrix.h:31). The library forms camera_vp*model and light_vp*model on the
 * host with the reference's own evaluation order (matrix.h:118-123,
 * vector.h:69-73), which gives the bits IShader.h:56,60 produce per vertex. */
typedef struct HanaUniforms {
    float model[16];        /* ShaderData::model_matrix     */
    float model_I[16];      /* ShaderData::model_matrix_I   */
    float camera_vp[16];    /* ShaderData::camera_vp_matrix */
    float light_vp[16];     /* ShaderData::light_vp_matrix  */
    float view_pos[3];      /* ShaderData::view_Pos         */
    float gloss;            /* Material::gloss              */
    float light_dir[3];     /* ShaderData::light_dir        */
    float bump_scale;       /* Material::bump_scale         */
    float light_color[4];   /* ShaderData::light_color (r,g,b,a) */
    float ambient[4];       /* ShaderData::ambient          */
    float mat_color[4];     /* Material::color              */
    float mat_specular[4];  /* Material::specular           */
    int32_t enable_shadow;  /* ShaderData::enable_shadow    */
    int32_t reserved[3];
} HanaUniforms;

/* Per-draw counters (device-side, read back on request). */
typedef struct HanaStats {
    uint32_t faces_in;        /* faces submitted                              */
    uint32_t tris_clipped;    /* faces that went through Sutherland-Hodgman   */
    uint32_t tris_out;        /* triangles after clip+cull+degenerate reject  */
    uint32_t tile_refs;       /* (tile, triangle) pairs binned                */
    uint32_t tiles_touched;   /* 16x16 tiles with at least one triangle       */
    uint32_t pixels_covered;  /* pixels whose final value was written by this draw */
    uint32_t overflow;        /* non-zero if a capacity retry happened        */
    uint32_t reserved;
} HanaStats;

typedef struct hana_ctx hana_ctx;
typedef struct hana_model hana_model;
typedef struct hana_texture hana_texture;
typedef struct hana_rb hana_rb;
typedef struct hana_sweep hana_sweep;

HANA_API const char* hana_last_error(void);
HANA_API int hana_version(void);
HANA_API int hana_device_count(void);

/* --- context ------------------------------------------------------------ */
HANA_API int hana_ctx_create(int device, hana_ctx** out);
HANA_API int hana_ctx_destroy(hana_ctx* ctx);
/* Launch on an existing cudaStream_t (e.g. torch's current stream) instead of
 * the context's own non-blocking stream; pass NULL to go back. */
HANA_API int hana_ctx_set_stream(hana_ctx* ctx, void* cuda_stream);
HANA_API int hana_sync(hana_ctx* ctx);
/* Number of kernels this context has launched so far (bench "gpu_launches"). */
HANA_API int hana_ctx_launch_count(hana_ctx* ctx, uint64_t* out);
/* Tile flush / clear through TMA bulk tensor stores (default when the driver exposes
 * cuTensorMapEncodeTiled and the target's row stride is a multiple of 16 bytes) or
 * plain global stores. HANA_NO_TMA=1 in the environment selects the latter at creation. */
HANA_API int hana_ctx_uses_tma(hana_ctx* ctx);
HANA_API int hana_ctx_set_tma(hana_ctx* ctx, int enable);
/* Pipelined sweep submissions (see hana_sweep_create) on / off; waits for the work queued so far. On by default unless
 * HANA_NO_PIPELINE=1 is in the environment at creation. Off = one submission after the other in the context's stream,
 * which is also what gives per-kernel profile times that add up to the step. */
HANA_API int hana_ctx_set_pipeline(hana_ctx* ctx, int enable);
HANA_API int hana_ctx_sm_count(hana_ctx* ctx);
/* Shadow-map passes that took the wide-slot rasteriser (a pass whose per-frame triangle capacity exceeds what the
 * 24-bit slot field of the packed {shadow byte, triangle slot} state addresses; HANA_R8_SLOT_LIMIT in the environment
 * lowers the threshold so that tests can force the path). */
HANA_API int hana_ctx_wide_r8_launches(hana_ctx* ctx, uint64_t* out);

/* --- measurement ------------------------------------------------------------ */
/* CUDA-event timing on the context's stream. hana_timer_stop synchronises. */
HANA_API int hana_timer_start(hana_ctx* ctx);
HANA_API int hana_timer_stop(hana_ctx* ctx, float* ms);
/* Per-kernel-class device time: when enabled, every launch is bracketed by a
 * pair of events on the launching stream; _get sums them (after they completed). */
#define HANA_PROF_BEGIN 0
#define HANA_PROF_SETUP 1
#define HANA_PROF_SCAN 2
#define HANA_PROF_FILL 3
#define HANA_PROF_RASTER_SHADOW 4
#define HANA_PROF_RASTER_MAIN 5
#define HANA_PROF_OTHER 6
#define HANA_PROF_KINDS 7
HANA_API int hana_ctx_profile(hana_ctx* ctx, int enable);
HANA_API int hana_ctx_profile_reset(hana_ctx* ctx);
HANA_API int hana_ctx_profile_get(hana_ctx* ctx, int kind, double* total_ms, uint64_t* launches);

/* --- inputs --------------------------------------------------------------- */
/* a2v: ncorners records of shader_struct_a2v (IShader.h:35-39): 8 floats
 * {obj_pos.xyz, obj_normal.xyz, uv.xy}, 3 consecutive corners per face in the
 * order graphics.cpp:380-386 visits them. ncorners must be a multiple of 3. */
HANA_API int hana_model_upload(hana_ctx* ctx, const float* a2v, int ncorners, hana_model** out);
/* Replace the stream of an uploaded model in place (same corner count): the reference's
 * Model::normal() re-normalises its normals on every access (model.cpp:108-111), so a faithful
 * caller re-gathers the stream per pass. */
HANA_API int hana_model_update(hana_model* m, const float* a2v, int ncorners);
HANA_API int hana_model_destroy(hana_model* m);
HANA_API int hana_model_ncorners(const hana_model* m);

/* data: TGAImage::buffer() layout (tgaimage.cpp:248-253): w*h texels of
 * bytespp (1, 3 or 4) bytes in B,G,R,A order, row y at data + y*w*bytespp. */
HANA_API int hana_texture_upload(hana_ctx* ctx, const uint8_t* data, int w, int h, int bytespp,
                        hana_texture** out);
HANA_API int hana_texture_destroy(hana_texture* t);

/* --- render target (RenderBuffer, renderbuffer.h:5-22) -------------------- */
/* Device-resident colour (RGBA8, index (y*w+x)*4, y up) + depth (f32, y*w+x).
 * A fresh buffer holds what the reference ctor leaves: colour (0,0,0,255),
 * depth 1.0 (renderbuffer.cpp:6-7,16-17). */
HANA_API int hana_rb_create(hana_ctx* ctx, int width, int height, hana_rb** out);
HANA_API int hana_rb_destroy(hana_rb* rb);
HANA_API int hana_rb_size(const hana_rb* rb, int* width, int* height);
/* renderbuffer_clear_color writes all four bytes (renderbuffer.cpp:59-68);
 * the bytes are given directly because Color{...,a=255}*255 is out of range
 * for unsigned char in the reference. */
HANA_API int hana_rb_clear_color(hana_rb* rb, uint8_t r, uint8_t g, uint8_t b, uint8_t a);
HANA_API int hana_rb_clear_depth(hana_rb* rb, float depth);
/* Either pointer may be NULL to skip that plane. Host pointers; synchronous. */
HANA_API int hana_rb_upload(hana_rb* rb, const uint8_t* color_rgba, const float* depth);
HANA_API int hana_rb_download(hana_rb* rb, uint8_t* color_rgba, float* depth);
/* Raw device pointers (for zero-copy consumers such as the bench harness). */
HANA_API int hana_rb_device_ptrs(hana_rb* rb, void** color_dev, void** depth_dev);

/* --- the draw entry points -------------------------------------------------- */
/* One pass over all faces of `model` with device shader `shader_id`:
 * vertex -> clip -> cull -> setup -> bin -> tile raster + depth test +
 * fragment shading -> tile flush. Semantics are those of
 * graphics_draw_triangle (graphics.cpp:378-407): existing contents of `rb`
 * take part in the depth test, only covered pixels change (R,G,B and depth;
 * alpha is never written). diffuse/normal may be NULL (fetches return 0 as
 * TGAImage::get does without data); shadow_map may be NULL
 * (IShader.h:109 -> lit). Asynchronous on the context's stream. */
HANA_API int hana_draw(hana_ctx* ctx, hana_rb* rb, const hana_model* model, int shader_id,
              const HanaUniforms* uniforms, const hana_texture* diffuse,
              const hana_texture* normal, const hana_rb* shadow_map);

/* DrawModel::draw (scene.h:53-99): if u->enable_shadow, ShadowShader pass
 * into `shadow_map`, then `shader_id` pass into `frame` reading it, then
 * shadow_map cleared to (0,0,0,[255*255 -> 1]) / FLT_MAX (scene.h:94-98). */
HANA_API int hana_draw_model(hana_ctx* ctx, hana_rb* frame, hana_rb* shadow_map, const hana_model* model,
                    int shader_id, const HanaUniforms* uniforms, const hana_texture* diffuse,
                    const hana_texture* normal);

/* Same as hana_draw_model but with HOST render buffers, exactly the memory
 * the reference's DrawModel::draw reads and writes: frame colour/depth are
 * uploaded, both passes run, results are downloaded back (the shadow map's
 * final state is the cleared one, so it is not transferred). This is the
 * call the drop-in graphics_draw_triangle shim and the bench's e2e leg use.
 * If assume_cleared != 0 the frame is taken to hold (clear_rgba, clear_depth)
 * everywhere and the upload is skipped. */
HANA_API int hana_draw_model_host(hana_ctx* ctx, int width, int height, uint8_t* frame_color_rgba,
                         float* frame_depth, const hana_model* model, int shader_id,
                         const HanaUniforms* uniforms, const hana_texture* diffuse,
                         const hana_texture* normal, int assume_cleared,
                         const uint8_t clear_rgba[4], float clear_depth);

HANA_API int hana_last_stats(hana_ctx* ctx, HanaStats* out); /* synchronises */

/* --- batched frame sweep (BASELINE.json configs[2]) ------ */
/* Renders n_frames independent frames of one model per submission: for each
 * frame clear (clear_rgba, clear_depth), optional shadow pass, main pass —
 * i.e. main.cpp:152-153 + DrawModel::draw — with the frame index as a grid
 * dimension, so launch latency is paid once per batch. Frames land in a
 * device ring of `n_frames` colour+depth targets.
 * Renders are ASYNCHRONOUS: the call returns once the batch is queued, without
 * reading anything back. The sweep's next synchronising call (download*,
 * checksums, stats, device_ptrs, hana_sync, hana_timer_stop) waits for it and,
 * if the batch ran out of internal scratch, transparently renders it again
 * with more. Model, textures and a device uniform buffer must stay alive and
 * unchanged until then.
 * Consecutive submissions of a context are PIPELINED: the uniform upload and the binning of both passes of a
 * submission run on internal streams beside the rasterisers of the submission before it (their scratch and uniform
 * blocks alternate), the rasterisers follow in submission order on the context's stream. Nothing changes for the
 * caller except that inputs handed to a render (model, textures, uniform buffers) must be COMPLETE when the call is
 * made, not merely queued on some stream. A context that runs on a caller's stream (hana_ctx_set_stream), or
 * HANA_NO_PIPELINE=1 in the environment, keeps everything in that one stream's order. */
HANA_API int hana_sweep_create(hana_ctx* ctx, int width, int height, int max_frames, hana_sweep** out);
HANA_API int hana_sweep_destroy(hana_sweep* s);
/* uniforms: host array of n_frames HanaUniforms (copied H2D inside). */
HANA_API int hana_sweep_render(hana_sweep* s, const hana_model* model, int shader_id,
                      const HanaUniforms* uniforms, int n_frames, const hana_texture* diffuse,
                      const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth);
/* Device buffer (max_frames * sizeof(HanaUniforms)) a caller may fill itself and hand to
 * hana_sweep_render_dev to keep the whole submission free of host->device copies. */
HANA_API int hana_sweep_uniforms_dev(hana_sweep* s, void** out);
/* Same, uniforms already resident on the device (n_frames * sizeof(HanaUniforms)); enable_shadow is what the
 * caller wrote into them (scene.h:73), passed separately so that nothing has to be read back. */
HANA_API int hana_sweep_render_dev(hana_sweep* s, const hana_model* model, int shader_id,
                          const void* uniforms_dev, int enable_shadow, int n_frames,
                          const hana_texture* diffuse, const hana_texture* normal,
                          const uint8_t clear_rgba[4], float clear_depth);
/* Copy frame `i` of the last batch to host (either may be NULL). Synchronous. */
HANA_API int hana_sweep_download(hana_sweep* s, int frame, uint8_t* color_rgba, float* depth);
/* Async copy of frames [first, first+count) into caller-provided PINNED host
 * memory (count*W*H*4 bytes each plane; either may be NULL). The host waits for
 * the sweep's render, then the copies run on the context's copy stream so that
 * the next batch (rendered into ANOTHER sweep) overlaps them; they are complete
 * after hana_sync(). A later render into this sweep waits for them on the device. */
HANA_API int hana_sweep_download_async(hana_sweep* s, int first, int count, uint8_t* color_rgba_pinned,
                              float* depth_pinned);
HANA_API int hana_sweep_device_ptrs(hana_sweep* s, void** color_dev, void** depth_dev,
                           size_t* frame_stride_pixels);
/* Per-frame 64-bit checksum (FNV-style over RGB bytes and depth bits) of the
 * last batch, computed on the device; used by the multi-GPU sharding tests. */
HANA_API int hana_sweep_checksums(hana_sweep* s, int n_frames, uint64_t* out_host);
HANA_API int hana_sweep_stats(hana_sweep* s, int frame, HanaStats* out); /* synchronises */
/* Batches of this sweep that ran out of internal scratch AND were overwritten by a later submission before a
 * synchronising call could render them again (only possible when renders are queued back to back, as a throughput
 * loop does). Waits for everything queued so far. 0 = no frame handed out so far was incomplete. */
HANA_API int hana_sweep_overflow_count(hana_sweep* s, uint64_t* out);

/* Optional, OFF by default ("shadow-map reuse when the light is static — algorithmic change"): the
 * ShadowShader pass of DrawModel::draw (scene.h:73-88) reads nothing but light_vp_matrix * model_matrix (IShader.cpp:170),
 * which an orbiting camera does not change (scene.h:69: the light looks at the camera's fixed TARGET). With reuse enabled,
 * a hana_sweep_render batch whose frames all carry bit-identical `light_vp` and `model` renders ONE shadow map and every
 * frame's BlinnShader/NormalMapShader pass reads it; the frames are byte-identical to those of the per-frame passes
 * (the reference renders the same map again every frame, then clears it, scene.h:94-98). A batch whose frames differ in
 * either matrix is rendered pass by pass as always. hana_sweep_render_dev (uniforms already on the device, nothing to
 * compare on the host) ignores the setting. bench.py reports it as an extra key, never as the headline. */
HANA_API int hana_sweep_set_shadow_reuse(hana_sweep* s, int enable);

/* --- one frame split by screen tiles over several GPUs (BASELINE.json north_star, optional mode) --
 * The frame's 16x16 tiles are dealt to the GPUs in bands of tile rows. Geometry is replicated; each GPU rasterises
 * only the tiles of its band, per pass. Replaces nothing in the reference (it has one thread); the per-GPU work is
 * still graphics_draw_triangle (graphics.cpp:378-407) restricted to a pixel range.
 *   hana_sweep_set_bands   tile rows [first, first+count) of the shadow pass / of the main pass this sweep renders
 *                          from now on; count 0 = every row (the default).
 *   hana_sweep_render_pass ONE pass of DrawModel::draw (scene.h:86 or :91) for n_frames frames: HANA_PASS_SHADOW
 *                          leaves this GPU's band of the 1-byte shadow maps in HBM, HANA_PASS_MAIN shades its band of
 *                          the frames from whatever the maps hold. Between the two the caller makes the maps complete
 *                          on every GPU (NCCL all-gather of the bands: sharding.py) — pass 2 looks up arbitrary
 *                          light-space texels (IShader.h:107-129), so that exchange is a real data dependency.
 *   hana_sweep_shadow_ptrs the maps' device memory: texel (x,y) of frame f at r8[f*frame_stride + y*pitch + x].
 * A band of tile rows is a contiguous byte range of every plane (rows are row-major, y up), so bands gather in place. */
#define HANA_PASS_SHADOW 1
#define HANA_PASS_MAIN 2
HANA_API int hana_sweep_set_bands(hana_sweep* s, int shadow_row_first, int shadow_row_count, int main_row_first,
                         int main_row_count);
HANA_API int hana_sweep_render_pass(hana_sweep* s, int pass, const hana_model* model, int shader_id,
                           const HanaUniforms* uniforms, int n_frames, const hana_texture* diffuse,
                           const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth);
HANA_API int hana_sweep_shadow_ptrs(hana_sweep* s, void** r8_dev, int* pitch_bytes, size_t* frame_stride_bytes);
/* hana_sweep_render_pass without host synchronisation: the pass is queued on the context's stream with the scratch
 * capacities the context has (hana_sweep_render_pass reads the needs back and retries by itself). With the context on
 * the caller's stream (hana_ctx_set_stream) the exchanges between the passes are ordered by the stream alone.
 * hana_sweep_passes_ok waits for the queued passes and reports whether both had room (*ok = 1); if not, the scratch has
 * been grown and the frame must be queued again — on every rank of a split frame, or none. */
HANA_API int hana_sweep_render_pass_async(hana_sweep* s, int pass, const hana_model* model, int shader_id,
                                 const HanaUniforms* uniforms, int n_frames, const hana_texture* diffuse,
                                 const hana_texture* normal, const uint8_t clear_rgba[4], float clear_depth);
HANA_API int hana_sweep_passes_ok(hana_sweep* s, int* ok);

/* --- present / output ------------------------------------ */
/* Replaces window_draw_buffer's conversion loop (win32.cpp:348-370): frames [first, first+count) of the last batch
 * as top-down B,G,R,255 (BGRA8) or B,G,R (BGR8, a TGA payload) surfaces, converted on the device; copied to dst_host
 * (pinned memory overlaps the next batch; may be NULL) on the copy stream, complete after hana_sync(). */
#define HANA_PRESENT_BGRA8 0
#define HANA_PRESENT_BGR8 1
HANA_API int hana_sweep_present(hana_sweep* s, int first, int count, int format, uint8_t* dst_host, void** dst_dev_out);
/* Replaces TGAImage::write_tga_file(rle = true) (tgaimage.cpp:145-246, packets of unload_rle_data :206-246) for frames
 * [first, first+count) of the sweep's last render: complete 24-bit RLE TGA files (header, packets, footer), byte-identical
 * to the reference's writer and to hana_tga_write, produced on the device so that a frame leaves the GPU compressed.
 *   hana_sweep_encode_tga  queues the encoder behind the render on the context's stream; reads nothing back.
 *   hana_sweep_fetch_tga   waits for it, writes offsets[0..count] and sizes[0..count) (file f = dst_host[offsets[f],
 *                          offsets[f] + sizes[f]); starts are 16-byte aligned, offsets[count] = bytes used) and copies the
 *                          bytes to dst_host (pinned memory overlaps the next batch) on the copy stream: complete after
 *                          hana_sync(). HANA_E_OVERFLOW if dst_capacity is too small. */
HANA_API int hana_sweep_encode_tga(hana_sweep* s, int first, int count);
HANA_API int hana_sweep_fetch_tga(hana_sweep* s, uint8_t* dst_host, size_t dst_capacity, uint64_t* offsets, uint64_t* sizes);
/* Replaces TGAImage::write_tga_file (tgaimage.cpp:145-246): `data` = w*h*bytespp bytes in file order (top-left
 * origin is flagged); raw or RLE; byte-identical files. */
HANA_API int hana_tga_write(const char* path, const uint8_t* data, int w, int h, int bytespp, int rle);

/* --- asset ingestion -------------------------------------- */
/* Replaces Model::Model's OBJ parse + the per-corner gather of graphics.cpp:380-386 (model.cpp:6-48, 66-111):
 * out_a2v = malloc'd ncorners*8 floats (obj_pos, obj_normal, uv) ready for hana_model_upload; release with hana_free.
 * normal_pass >= 1: which walk of the reference's draw loop the normals correspond to (Model::normal re-normalises
 * in place on every access, model.cpp:108-111). */
HANA_API int hana_obj_load(const char* path, int normal_pass, float** out_a2v, int* out_ncorners);
/* Replaces TGAImage::read_tga_file (tgaimage.cpp:40-143) and, with model_flip != 0, Model::load_texture's extra
 * flip (model.cpp:74-84): malloc'd w*h*bytespp bytes ready for hana_texture_upload; release with hana_free. */
HANA_API int hana_tga_load(const char* path, int model_flip, uint8_t** out_data, int* out_w, int* out_h, int* out_bytespp);
HANA_API void hana_free(void* p);

/* Pinned host memory helpers for the e2e path. */
HANA_API int hana_host_alloc(size_t bytes, void** out);
HANA_API int hana_host_free(void* p);

/* --- host-side mirror of the path's caller ---------------- */
/* Camera (camera.h:13-33) and the per-draw constants of SingleModelScene /
 * DrawModel (scene.cpp:7,77-84; scene.h:8-9; gameobject.cpp:12-17), as PODs. */
typedef struct HanaCamera {
    float position[3];
    float target[3];
    float aspect;
} HanaCamera;
typedef struct HanaSceneDesc {
    float light_pos[3];
    float model_pos[3];
    float model_rot_deg[3];
    float model_scale[3];
    float light_color[4];
    float ambient[4];
    float mat_color[4];
    float mat_specular[4];
    float gloss;
    float bump_scale;
} HanaSceneDesc;
HANA_API int hana_camera_init(HanaCamera* cam, const float position[3], const float target[3], float aspect);
/* Camera::update_transform(Motion) camera.cpp:63-70 */
HANA_API int hana_camera_update(HanaCamera* cam, float orbit_x, float orbit_y, float pan_x, float pan_y, float dolly);
/* light (2,2,2), identity model transform, LightColor/AMBIENT, white material, gloss 50, bump 1 */
HANA_API int hana_scene_defaults(HanaSceneDesc* s);
/* The ShaderData block DrawModel::draw builds before its passes (scene.h:55-71),
 * bit-identical to the reference's for the same camera (host float arithmetic in
 * the reference's evaluation order). */
HANA_API int hana_scene_uniforms(const HanaCamera* cam, const HanaSceneDesc* s, int width, int height, int enable_shadow,
                                 HanaUniforms* out);
/* BASELINE.json configs[2]: frame k of the orbit sweep = Camera(CAMERA_POSITION,
 * CAMERA_TARGET, W/H) advanced k times by orbit (1/frames_per_turn, 0). */
HANA_API int hana_orbit_sweep_uniforms(const HanaSceneDesc* s, int width, int height, int enable_shadow, int first, int count,
                                       int frames_per_turn, HanaUniforms* out);

/* --- stage-level entry points (parity tests, SURVEY.md §4 tier 1) --------- */
/* Runs only the vertex kernel; out_v2f receives ncorners records of
 * shader_struct_v2f (IShader.h:41-47): 13 floats each. Fields the shader
 * does not set are written as 0. */
HANA_API int hana_stage_vertex(hana_ctx* ctx, const hana_model* model, int shader_id,
                      const HanaUniforms* uniforms, float* out_v2f_host);
/* Runs vertex + clip + cull + setup for a `width` x `height` target and
 * returns the surviving triangles sorted by order key: per triangle
 * 1 order key (face*8 + fan index) in out_order, and 3x13 floats of the
 * post-clip v2f records in out_v2f (capacity in triangles). */
HANA_API int hana_stage_setup(hana_ctx* ctx, const hana_model* model, int shader_id,
                     const HanaUniforms* uniforms, int width, int height, int capacity,
                     uint32_t* out_order, float* out_v2f, int* out_count);
/* Primitive-ID buffer of the last hana_draw on `rb` is not kept by the
 * reference; this variant of hana_draw also writes, per pixel, the order key
 * of the primitive that owns it (0xFFFFFFFF where the draw wrote nothing). */
HANA_API int hana_draw_primid(hana_ctx* ctx, hana_rb* rb, const hana_model* model, int shader_id,
                     const HanaUniforms* uniforms, const hana_texture* diffuse,
                     const hana_texture* normal, const hana_rb* shadow_map,
                     uint32_t* out_primid_host);

#ifdef __cplusplus
}
#endif
#endif /* HANA_B200_H */
