/* portal_b200 -- C ABI of the B200-native renderer for optozorax/portal's per-pixel ray loop.
 *
 * This is the drop-in boundary.  In the reference the renderer
 * (`SceneRenderer`, /root/reference/src/main.rs:732-1976) talks to the GPU through a
 * macroquad `Material`:
 *
 *   load_material(fragment GLSL, MaterialParams{uniforms, textures})   src/gui/scene.rs:1132-1143
 *   material.set_uniform(name, mat4 | f32 | i32 | (f32,f32))           src/gui/scene.rs:587-656,
 *                                                                      src/main.rs:1266-1359
 *   material.set_texture(name, Texture2D)                              src/main.rs:1066-1098
 *   gl_use_material + draw_rectangle(0,0,w,h) into a render target     src/main.rs:1411-1428
 *   render_target.texture.get_texture_data() -> RGBA8                  src/main.rs:2939-2943
 *
 * The functions below replace exactly those five operations.  The scene is handed over in
 * the same decomposition the reference's shader generator walks (objects, materials,
 * intersection materials, library: src/gui/scene.rs:693-1110) with the user's GLSL
 * snippets as text; this library generates the sm_100a program from them (NVRTC) instead
 * of a GLSL fragment shader.  Plain pointers and sizes only; one context per host thread
 * and GPU; all functions return 0 on success, non-zero on failure (see pe_last_error).
 */
#ifndef PORTAL_B200_H
#define PORTAL_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PE_API __attribute__((visibility("default")))
#else
#define PE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pe_ctx pe_ctx;

/* ---- context ------------------------------------------------------------------------- */

/* device >= 0: CUDA device ordinal.  device < 0: compile-only context (program generation and
 * NVRTC work without a GPU; pe_render* fail loudly).  Returns NULL on failure. */
PE_API pe_ctx* pe_create(int device);
PE_API void pe_destroy(pe_ctx* ctx);
/* Message of the last failed call on this context (never NULL). With ctx == NULL: why the last
 * pe_create failed. */
PE_API const char* pe_last_error(pe_ctx* ctx);
/* ABI version of this header: major * 100 + minor. */
PE_API int pe_abi_version(void);

/* ---- scene program  (replaces load_material, src/gui/scene.rs:1112-1176) ----------------- */

enum { PE_SUBSPACE_NORMAL = 0, PE_SUBSPACE_SUBSPACE = 1, PE_SUBSPACE_BOTH = 2 }; /* src/gui/object.rs:39-45 */
enum { PE_UNIFORM_MAT4 = 0, PE_UNIFORM_FLOAT = 1, PE_UNIFORM_INT = 2 };          /* src/gui/scene.rs:673-677 */

/* Start describing a new scene program; discards the previous description. */
PE_API int pe_scene_begin(pe_ctx* ctx);
/* User library code, in scene order (src/gui/scene.rs:1037-1044). */
PE_API int pe_scene_add_library(pe_ctx* ctx, const char* name, const char* glsl);
/* Materials in scene order; ids are USER_MATERIAL_OFFSET + k (src/gui/scene.rs:720-780). */
PE_API int pe_scene_add_material_simple(pe_ctx* ctx, const char* name, const double color[3], double normal_coef, int grid,
                                 double grid_scale, double grid_coef, int grid2, int grid3);
PE_API int pe_scene_add_material_reflect(pe_ctx* ctx, const char* name, const double add_to_color[3]);
PE_API int pe_scene_add_material_refract(pe_ctx* ctx, const char* name, const double add_to_color[3], double refractive_index);
PE_API int pe_scene_add_material_complex(pe_ctx* ctx, const char* name, const char* glsl);
/* Objects in scene order (src/gui/scene.rs:847-1009).  matrix_b == NULL: ObjectType::Simple(matrix_a),
 * otherwise ObjectType::Portal(matrix_a, matrix_b).  Matrix names are the uniform stems
 * (`<name>_mat`, `<name>_mat_inv`, `<a>_to_<b>_mat_teleport`: src/gui/object.rs:18-30). */
PE_API int pe_scene_add_object_flat(pe_ctx* ctx, const char* name, int subspace, const char* matrix_a, const char* matrix_b,
                             const char* is_inside_glsl);
PE_API int pe_scene_add_object_complex(pe_ctx* ctx, const char* name, int subspace, const char* matrix_a, const char* matrix_b,
                                const char* intersect_glsl);
PE_API int pe_scene_add_object_debug_matrix(pe_ctx* ctx, const char* name, const char* matrix);
/* Intersection materials in scene order (src/gui/scene.rs:1011-1035). */
PE_API int pe_scene_add_intersection_material(pe_ctx* ctx, const char* name, const char* glsl);
/* Uniform table (src/gui/scene.rs:424-495): full GLSL names, e.g. "a_mat", "a_mat_inv", "progress_u".
 * The renderer's `_`-prefixed uniforms (src/gui/scene.rs:497-535) are built in and need no declaration. */
PE_API int pe_scene_declare_uniform(pe_ctx* ctx, const char* name, int type);
/* Texture samplers, name without the `_tex` suffix (src/gui/texture.rs:12-16). */
PE_API int pe_scene_declare_texture(pe_ctx* ctx, const char* name);
/* Skybox (src/gui/scene.rs:1052-1063): rays that hit nothing sample this texture by the primary
 * ray's direction (through `_camera_mul_inv`) instead of the constant grey. */
PE_API int pe_scene_set_skybox(pe_ctx* ctx, const char* texture_name);
/* Generate + compile the program for sm_100a.  Integer uniforms are specialisation constants:
 * the program is compiled for their current values and transparently re-specialised (cached)
 * when pe_render* sees different ones.  On failure pe_last_error holds the compiler log with
 * each message attributed to the scene element that owns the snippet and its local line
 * (the role of src/shader_error_parser.rs + src/code_generation.rs:36-43). */
PE_API int pe_scene_compile(pe_ctx* ctx);
/* Generated CUDA source / compiled cubin of the current specialisation (valid until the next
 * compile); for inspection, caching and cuobjdump. */
PE_API const char* pe_scene_source(pe_ctx* ctx);
PE_API int pe_scene_cubin(pe_ctx* ctx, const void** data, size_t* size);
/* Host image of the constant uniform block (`PeConstBlock` of pe_scene_source) exactly as a render of a
 * width x height frame would upload it now: uploaded uniforms, `_resolution`, and the values the host derives
 * per upload (tan(view_angle/2), per-plane unit normals).  Valid until the next call on ctx; for inspection and
 * for running the generated program outside the GPU in tests.  Texture slots hold device pointers. */
PE_API int pe_scene_uniform_block(pe_ctx* ctx, int width, int height, const void** data, size_t* size);
/* Options: "persistent" (0/1, default 0), "specialize_ints" (0/1, default 1), "specialize_matrices"
 * (0/1, default 1: the exact-0 / exact-1 structure of every uploaded matrix is baked in), "block_threads" (a multiple of 64 in [64, 1024], default 512),
 * "min_blocks", "hoist_planes" (0/1, default 1: per-plane normal algebra evaluated once per upload on the
 * host), "lazy_planes" (0/1, default 1: a plane test stops as soon as its result is certain to be
 * rejected, same pixels), "with_probe" (0/1, default 0: also generate the camera-teleportation probe kernel; pe_probe_ray
 * switches it on by itself), "adaptive" (0/1, default 1: an int uniform or a matrix structure that differed between
 * renders more than 4 times stops being a specialisation constant, which bounds recompilation when an animation
 * drives it), "uniforms_in_smem" (0 = constant bank (default), 1 = every block copies the uniform block into
 * shared memory with a cooperative loop, 2 = with one TMA bulk copy + mbarrier; same pixels, measured in DESIGN.md section 3),
 * "tile_w" (8 / 16 / 32: a warp covers 8x4, 16x2 or 32x1 pixels -- longer contiguous stores for peer-mapped targets), "lineinfo" (0/1, default 1), "unroll_loops" (0/1, default 1; 0 keeps the loops of
 * user snippets rolled).  Set before pe_scene_compile. */
PE_API int pe_set_option(pe_ctx* ctx, const char* key, int value);

/* ---- uniforms and textures  (replace material.set_uniform / set_texture) ------------------ */

PE_API int pe_set_uniform_mat4(pe_ctx* ctx, const char* name, const float column_major[16]);
PE_API int pe_set_uniform_f32(pe_ctx* ctx, const char* name, float value);
PE_API int pe_set_uniform_i32(pe_ctx* ctx, const char* name, int32_t value);
/* Bulk form of the above for the per-frame upload: n names with one value record each. */
PE_API int pe_set_uniforms_mat4(pe_ctx* ctx, int n, const char* const* names, const float* column_major_16n);
/* RGBA8 texels, row 0 first (what Texture2D::from_file_with_format holds, src/main.rs:1075). */
PE_API int pe_set_texture(pe_ctx* ctx, const char* name, const uint8_t* rgba8, int width, int height);

/* ---- render  (replaces gl_use_material + draw_rectangle + get_texture_data) --------------- */

/* Which rows of the frame this call renders: `n_strips` strips of `strip_rows` rows; local
 * strip k is global strip strip_first + k * strip_step.  Whole frame on one GPU:
 * {w, h, h, 0, 1, 1, 1}.  Cyclic sharding over G GPUs, rank r: strip_first = r, strip_step = G. */
typedef struct pe_target {
    int32_t width, height;       /* full frame, `_resolution` */
    int32_t strip_rows;          /* rows per strip */
    int32_t strip_first;         /* first global strip of this call */
    int32_t strip_step;          /* distance between consecutive strips of this call */
    int32_t n_strips;            /* strips rendered by this call */
    int32_t full_frame_layout;   /* 1: `out` holds the full frame; 0: only this call's rows, compacted */
} pe_target;

/* Number of pixels a target's compact output holds. */
PE_API size_t pe_target_pixels(const pe_target* t);
/* Asynchronous render into DEVICE memory (float RGBA, 16 B/pixel: the value the reference's
 * shader writes to FragColor, src/frag.glsl:526,551, before the render target quantises it).
 * `stream` is a cudaStream_t or NULL for the context's own stream.  `out_device` may be a
 * peer-mapped pointer (see pe_ipc_*). `bounces_device` (int32 per pixel) may be NULL. */
/* A context has ONE uniform block and ONE tile-queue counter on the device: renders of one context must be issued on one
 * stream at a time (different contexts are independent).  A target with n_strips == 0 renders nothing and succeeds. */
PE_API int pe_render(pe_ctx* ctx, const pe_target* target, void* out_device, void* bounces_device, void* stream);
/* Synchronous render into HOST memory, device->host copy included (float RGBA). */
PE_API int pe_render_host(pe_ctx* ctx, const pe_target* target, float* out_host);
/* Asynchronous render into DEVICE memory as RGBA8 (4 B/pixel): the kernel itself applies the
 * render target's float -> unorm8 conversion (src/main.rs:2939-2943 draws into an RGBA8 target),
 * bit-identical to pe_render followed by pe_quantize_rgba8. */
PE_API int pe_render_rgba8(pe_ctx* ctx, const pe_target* target, void* out_device_rgba8, void* stream);
/* Same, quantised to RGBA8 as the reference's render target + get_texture_data deliver it
 * (src/main.rs:2945-2947). Synchronous. */
PE_API int pe_render_host_rgba8(pe_ctx* ctx, const pe_target* target, uint8_t* out_host);
/* Pipelined form of pe_render_host_rgba8 for frame sequences (the offline `render` loop,
 * src/main.rs:2925-2968, renders and reads back one frame after another): queue render + device->host
 * copy of one frame and return at once.  Up to PE_PIPELINE_DEPTH frames are in flight; frame i's copy
 * overlaps frame i+1's kernel.  Uniforms set between submits apply to later frames only.  `out_host`
 * should be page-locked (pe_host_malloc) or the copy is not asynchronous; it must stay untouched until
 * pe_wait_host(ticket) returns. */
#define PE_PIPELINE_DEPTH 4
PE_API int pe_submit_host_rgba8(pe_ctx* ctx, const pe_target* target, uint8_t* out_host, uint64_t* ticket);
PE_API int pe_wait_host(pe_ctx* ctx, uint64_t ticket);
/* Multi-GPU host delivery without a device-side gather: `target` selects this rank's row strips, `host_frame`
 * is the WHOLE width x height RGBA8 frame in host memory shared by all ranks (e.g. POSIX shared memory,
 * page-locked in each process with pe_host_register).  The strips are rendered as RGBA8 and each one is copied
 * over this GPU's own PCIe link to its rows of the frame; same ticket / pipeline rules as pe_submit_host_rgba8.
 * When every rank's ticket has completed the frame is whole. */
PE_API int pe_submit_host_strips_rgba8(pe_ctx* ctx, const pe_target* target, uint8_t* host_frame, uint64_t* ticket);
/* Page-lock / unlock caller-owned host memory (cudaHostRegister, portable). */
PE_API int pe_host_register(pe_ctx* ctx, void* p, size_t bytes);
PE_API int pe_host_unregister(pe_ctx* ctx, void* p);
/* Page-locked host memory for the readback calls. */
PE_API int pe_host_malloc(pe_ctx* ctx, size_t bytes, void** out);
PE_API int pe_host_free(pe_ctx* ctx, void* p);
/* Camera-teleportation probe (replaces teleport_external_ray, src/main.rs:1361-1409 +
 * src/frag.glsl:166-257, 527-547): follow the segment a -> b through the scene's portals (at most
 * 10, `teleport_light_u` forced on) and return where its end point lands -- directly, instead of
 * floats encoded into a 2x3 RGBA8 target.  have_result == 0: no portal was crossed (pos = 0). */
PE_API int pe_probe_ray(pe_ctx* ctx, const float a[3], const float b[3], float pos_out[3], int32_t* have_result,
                        int32_t* encounter_object, int32_t* change_subspace);
/* Per-scene autotuning: render `target` with each of the launch / specialisation alternatives that keep the pixels (512- or
 * 1024-thread blocks, canonical rays and w-aware products on or off), `reps` times each into context-owned scratch, timed with CUDA events, and
 * keep the fastest as the context's setting.  A candidate whose frame is not bit-identical to the first candidate's (the
 * default variant's) is reported as REJECTED and never chosen.  Synchronous; the programs are compiled (or taken from the
 * cache) as needed.  `report` (optional) receives one text line per candidate and the choice. */
PE_API int pe_autotune(pe_ctx* ctx, const pe_target* target, int reps, char* report, size_t report_len);
PE_API int pe_sync(pe_ctx* ctx);
/* Number of kernel launches this context has issued (render + helper kernels). */
PE_API uint64_t pe_launch_count(pe_ctx* ctx);

/* ---- multi-GPU helpers -------------------------------------------------------------------- */

/* De-interleave strips gathered rank-major ([rank][local strip][row][x], every rank padded to
 * `strips_per_rank` strips) into a row-major frame.  Device pointers; async on `stream`. */
PE_API int pe_deinterleave_strips(pe_ctx* ctx, const void* gathered_device, void* frame_device, int width, int height,
                           int strip_rows, int n_ranks, int strips_per_rank, void* stream);
/* Device memory owned by the context (cudaMalloc, so its base address can be exported over IPC). */
PE_API int pe_device_malloc(pe_ctx* ctx, size_t bytes, void** device_ptr_out);
PE_API int pe_device_free(pe_ctx* ctx, void* device_ptr);
PE_API int pe_memcpy_d2h(pe_ctx* ctx, void* host_dst, const void* device_src, size_t bytes, void* stream);
/* Context-owned device scratch, slot 0..127: allocated on first use, grown (never shrunk) when a larger size is asked
 * for, freed by pe_destroy.  For per-frame intermediates of a frame loop (motion-blur sub-frames): no cudaMalloc /
 * cudaFree -- each a device-wide synchronisation -- inside the loop. */
PE_API int pe_scratch_buffer(pe_ctx* ctx, int slot, size_t bytes, void** device_ptr_out);
/* Stream-ordered cross-GPU signalling without a collective: pe_signal_u32 publishes `value` to up to 8
 * flag words (device pointers, local or peer-mapped) after everything earlier in the stream -- remote
 * stores of a render kernel included; pe_stream_wait_geq_u32 makes the stream wait until a LOCAL flag
 * word is >= value (wrap-around safe).  pe_memset_u32 initialises flag words (synchronous). */
PE_API int pe_signal_u32(pe_ctx* ctx, void* const* device_ptrs, int n, uint32_t value, void* stream);
PE_API int pe_stream_wait_geq_u32(pe_ctx* ctx, void* local_device_ptr, uint32_t value, void* stream);
PE_API int pe_memset_u32(pe_ctx* ctx, void* device_ptr, uint32_t value, size_t count, void* stream);
/* CUDA IPC: let another process's render kernel store its pixels straight into this GPU's
 * frame over NVLink.  handle_out/handle_in are 64-byte cudaIpcMemHandle_t blobs. */
PE_API int pe_ipc_export(pe_ctx* ctx, void* device_ptr, uint8_t handle_out[64]);
PE_API int pe_ipc_open(pe_ctx* ctx, const uint8_t handle_in[64], void** device_ptr_out);
PE_API int pe_ipc_close(pe_ctx* ctx, void* device_ptr);

/* ---- sharded frames behind the C ABI --------------------------------------
 * One process per GPU, all on one box; the ranks rendezvous through the POSIX shared-memory segment /dev/shm/<name>
 * (no collective library).  Every rank calls pe_sharder_create with the same name / geometry / mode / format and its
 * own rank; rank 0 creates the segment (and unlinks it in pe_sharder_destroy).  `name` must be FRESH for every sharder: a
 * file of the same name left behind by a run that died could be opened by a late rank before rank 0 has replaced it, so
 * derive it from something the launch agrees on anew each time (a launcher-supplied random token; the Python wrapper has
 * rank 0 broadcast one).  Uniforms are set on the context as usual before each render / submit.
 *   PE_SHARD_OWNER  each rank keeps the strips it renders in its own HBM (compact rows: pe_sharder_target); nothing moves.
 *   PE_SHARD_P2P    every rank's render kernel stores its pixels straight into rank 0's whole frame over NVLink (CUDA
 *                   IPC mapping); frame completion / buffer reuse are stream-ordered flag words.  After
 *                   pe_sharder_render on rank 0, work enqueued on `stream` sees the assembled frame (*frame_out); once
 *                   the consumer is enqueued, pe_sharder_release lets the other ranks reuse the buffer 2 frames later.
 *   PE_SHARD_HOST   RGBA8 strips go over each GPU's own PCIe link into one shared page-locked host frame (ring of
 *                   PE_HOST_RING_DEPTH): submit -> complete on every rank; wait_frame / release_frame on rank 0. */
typedef struct pe_sharder pe_sharder;
enum { PE_SHARD_OWNER = 0, PE_SHARD_P2P = 1, PE_SHARD_HOST = 2 };
enum { PE_FRAME_F32 = 0, PE_FRAME_RGBA8 = 1 };
#define PE_HOST_RING_DEPTH 6
/* Cyclic strip layout: rank r of `world` owns global strips r, r + world, ... */
PE_API int pe_shard_target(int width, int height, int rank, int world, int strip_rows, int full_frame_layout, pe_target* out);
/* On failure *out may still hold an object: read pe_sharder_last_error, then pe_sharder_destroy it. */
PE_API int pe_sharder_create(pe_ctx* ctx, const char* name, int width, int height, int rank, int world, int strip_rows,
                             int mode, int format, pe_sharder** out);
PE_API void pe_sharder_destroy(pe_sharder* s);
PE_API const char* pe_sharder_last_error(pe_sharder* s);
PE_API int pe_sharder_target(pe_sharder* s, pe_target* out);
/* OWNER / P2P: enqueue this rank's strips of the next frame on `stream` (asynchronous).  *frame_out: OWNER -- this rank's
 * compact strips (device); P2P -- on rank 0 the assembled frame (device), NULL elsewhere. */
PE_API int pe_sharder_render(pe_sharder* s, void* stream, void** frame_out);
PE_API int pe_sharder_release(pe_sharder* s, void* stream);
/* OWNER mode with two frames in flight: consecutive frames alternate between two internal streams and two loaded instances
 * of the program (each with its own uniform block), so the partly filled last wave of frame f runs under the first waves of
 * frame f + 1 -- at 8 GPUs a frame is ~7 waves of blocks per GPU and that tail is ~15 % of it.  The frame is enqueued after
 * everything already on `stream`; `stream` is made to wait for the frame BEFORE it, returned in *prev_frame_out (NULL for the
 * first): the caller consumes frames one call late.  pe_sharder_flush orders `stream` after the last frame as well. */
PE_API int pe_sharder_render_overlapped(pe_sharder* s, void* stream, void** prev_frame_out);
PE_API int pe_sharder_flush(pe_sharder* s, void* stream, void** last_frame_out);
/* HOST: queue this rank's strips of the next frame (returns its number); block until they are in host memory and publish;
 * rank 0: block until every rank's strips of a frame have landed (*frame = the whole RGBA8 frame); give its slot back. */
PE_API int pe_sharder_submit(pe_sharder* s, uint64_t* frame_no);
PE_API int pe_sharder_complete(pe_sharder* s, uint64_t frame_no);
PE_API int pe_sharder_wait_frame(pe_sharder* s, uint64_t frame_no, const uint8_t** frame);
PE_API int pe_sharder_release_frame(pe_sharder* s, uint64_t frame_no);

/* ---- frame post-processing (the offline `render` caller, src/main.rs:640-722) -------------- */

/* Motion-blur average of n RGBA8 frames in gamma-2 space (average_images). Device pointers. */
PE_API int pe_average_frames_rgba8(pe_ctx* ctx, const void* const* frames_device, int n_frames, void* out_device,
                            size_t n_pixels, void* stream);
/* float RGBA -> RGBA8, the render target's quantisation. Device pointers. */
PE_API int pe_quantize_rgba8(pe_ctx* ctx, const void* rgba_f32_device, void* rgba8_device, size_t n_pixels, void* stream);
/* Bit comparison of two device buffers (16-byte aligned, `bytes` a multiple of 16): *words_out = how many 16-byte words
 * (= float RGBA pixels) differ.  Synchronous.  pe_autotune's guard uses it -- a candidate variant is only eligible when its
 * frame is the first candidate's frame bit for bit -- and a caller can use it to compare frames without a read-back (the
 * reference has no counterpart: its frames are compared by eye). */
PE_API int pe_frames_differ(pe_ctx* ctx, const void* a_device, const void* b_device, size_t bytes, void* stream, uint32_t* words_out);

#ifdef __cplusplus
}
#endif
#endif /* PORTAL_B200_H */
