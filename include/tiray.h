/* tiray.h — C-ABI of libtiray.so, the B200 (sm_100a) replacement for the Taichi kernels on the
 * hot path of lyd405121/ti-raytrace.
 *
 * The reference has no FFI layer: its device work is Taichi `@ti.kernel`s reached from the Python
 * classes Scene / Camera / LBvh.Bvh / PT_RGB.PathTrace.  Each entry point below replaces one of
 * those kernel groups; the Python classes in ti-raytrace_b200/ (same names and methods as the
 * reference's) call them through ctypes.  Plain pointers and sizes only; host arrays are borrowed
 * for the duration of the call; all device memory is owned by the context.
 *
 * Every function returns TR_OK (0) or a negative tr_status; tr_last_error() gives the text.
 * A context is bound to one CUDA device and one non-default stream and is not thread-safe.
 * Citations are file:line in the reference tree (commit 70ccd57).
 */
#ifndef TIRAY_H
#define TIRAY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tr_ctx tr_ctx;

typedef enum tr_status {
    TR_OK = 0,
    TR_ERR_CUDA = -1,          /* CUDA runtime error (text in tr_last_error) */
    TR_ERR_INVALID = -2,       /* bad argument / call order */
    TR_ERR_AABB = -3,          /* "aabb gen error": refit did not reach n-1 nodes (accel/LBvh.py:215-216) */
    TR_ERR_NO_DEVICE = -4,     /* no CUDA device: there is no CPU fallback */
    TR_ERR_STACK = -5,         /* "overflow, need larger stack" (Scene.py:741-742): the tree is deeper than the traversal stack */
    TR_ERR_COMM = -6           /* NCCL error / libnccl.so.2 not loadable (text in tr_last_error) */
} tr_status;

/* Counters and timings of the last tr_render_pt_rgb batch (device counters, CUDA events). */
typedef struct tr_stats {
    uint64_t rays_closest;     /* closest-hit traversals  (integrator/PT_RGB.py:65)  */
    uint64_t rays_shadow;      /* shadow traversals       (integrator/PT_RGB.py:104) */
    uint64_t node_visits;      /* internal-node visits (only when built with -DTR_COUNTERS) */
    uint64_t leaf_tests;       /* leaf tests           (only when built with -DTR_COUNTERS) */
    uint64_t kernel_launches;  /* kernels launched by the last render call */
    float    ms_total;         /* device time of the last render call (CUDA events on ctx stream) */
    float    ms_trace;         /* sum over stages, only when stage timing is enabled */
    float    ms_shade;
    float    ms_shadow;
    float    ms_build;         /* device time of the last tr_bvh_build */
    int32_t  frames;           /* frames rendered by the last call */
    int32_t  paths_in_flight;  /* path slots per batch */
    uint64_t node_visits_shadow; /* shadow-query visits (only with -DTR_COUNTERS) */
    uint64_t leaf_tests_shadow;
    int32_t  chains;           /* independent wavefront chains per batch */
    int32_t  pad_;
    uint64_t shade_terminal;   /* shaded vertices that were terminal (miss or emitter hit): the rest of rays_closest were
                                  Disney / glass vertices (roofline accounting of the shade kernel) */
} tr_stats;

/* ---- context ---------------------------------------------------------------------------- */
/* replaces ti.init(arch=ti.gpu) (example/cornell_box.py:15) */
int  tr_ctx_create(int device, tr_ctx** out);
void tr_ctx_destroy(tr_ctx* ctx);
const char* tr_last_error(tr_ctx* ctx);          /* ctx may be NULL: last global error */
int  tr_device_count(void);
int  tr_synchronize(tr_ctx* ctx);
/* run all subsequent work of this context on the caller's CUDA stream (cudaStream_t as void*), e.g.
 * torch's current stream, so that host-side events and NCCL calls order naturally; NULL restores the
 * context's own stream. */
int  tr_stream_set(tr_ctx* ctx, void* cuda_stream);

/* Page-lock a host array the caller keeps handing to the upload functions (e.g. the packed vertex table of a scene that is
 * re-uploaded every frame): uploads of page-locked arrays are DMA'd straight from the caller's memory instead of being copied
 * into the library's staging buffer first.  Purely an optimisation; the arrays are still only borrowed for the call.  No context
 * needed; unregister before freeing the memory. */
int tr_host_register(void* p, size_t bytes);
int tr_host_unregister(void* p);

/* ---- scene upload: replaces the from_numpy calls of Scene.setup_data_gpu (Scene.py:299-309) ---
 * vertex nv x 9 f32, prim np x 3 i32, material nm x 10 f32, shape ns x 10 f32 (may be NULL, ns=0),
 * light nl i32 (may be NULL, nl=0): exactly the tables packed at Scene.py:225-273. */
int tr_scene_upload(tr_ctx* ctx, const float* vertex, int nv, const int32_t* prim, int np,
                    const float* material, int nm, const float* shape, int ns,
                    const int32_t* light, int nl, const float bmin[3], const float bmax[3]);
/* material table may be re-uploaded alone (values frozen at first launch in Taichi) */
int tr_material_upload(tr_ctx* ctx, const float* material, int nm);
/* replaces Texture.setup_data_gpu (texture/Texture.py:41-42): buf[x][y] packed RGB i32 */
int tr_env_upload(tr_ctx* ctx, const int32_t* rgb, int w, int h, float power);

/* ---- OBJ / MTL ingest (host code): replaces pywavefront.Wavefront(filename).parse() and the per-vertex loops of Scene.add_obj
 * (Scene.py:59-141).  tr_obj_open parses the file (and the MTL it names); materials come back in PyWavefront's order with
 * props = Kd[3], Ke[3], d (transparency), Ns (shininess), Ni (optical density) and n_vertices = 3 x triangles;
 * tr_obj_material_vertices fills n_vertices x 9 f64 rows (pos3, normal3, tex3; zeros where the file has no vn / vt). */
typedef struct tr_obj tr_obj;
int  tr_obj_open(const char* path, tr_obj** out);
void tr_obj_close(tr_obj* obj);
const char* tr_obj_last_error(void);
int  tr_obj_material_count(const tr_obj* obj);
int  tr_obj_material(const tr_obj* obj, int k, char* name, int name_cap, double props[9], int64_t* n_vertices, int* has_vt, int* has_vn);
int  tr_obj_material_vertices(const tr_obj* obj, int k, double* rows);

/* ---- LBVH: replaces Bvh.setup_data_gpu (accel/LBvh.py:192-226): build_morton_3d, radix_sort_host,
 * build_lbvh, gen_aabb loop and the host-side flatten_tree, all on the device. */
int tr_bvh_build(tr_ctx* ctx);
/* reference-layout views (any pointer may be NULL): morton_code_s n x 2 i32 (code, prim) after the
 * sort; bvh_node (2n-1) x 11 f32 (UtilsFunc.py:24-26); compact_node (2n-1) x 9 f32 (:28-30) */
int tr_bvh_download(tr_ctx* ctx, int32_t* morton_sorted, float* bvh_node, float* compact_node);
/* unsorted Morton codes as written by build_morton_3d (accel/LBvh.py:318-336): n x 2 i32 */
int tr_morton_download(tr_ctx* ctx, int32_t* morton_unsorted);

/* replaces Scene.process_normal (Scene.py:754-798) and Scene.total_area (:747-750).  tr_process_normal is asynchronous (enqueued on
 * the context stream behind the build); tr_vertex_download / tr_total_area synchronise. */
int tr_process_normal(tr_ctx* ctx);
int tr_vertex_download(tr_ctx* ctx, float* vertex /* nv x 9 */);
int tr_total_area(tr_ctx* ctx, float* area);

/* ---- camera: replaces the from_numpy calls of Camera.update (Camera.py:87-93) */
int tr_camera_set(tr_ctx* ctx, const float view[16], const float view_inv[16], const float eye[3],
                  float fx, float fy, float cx, float cy);

/* ---- film: replaces the hdr / rgb_film fields (integrator/PT_RGB.py:27-37); index [x*H+y][3] */
int tr_film_create(tr_ctx* ctx, int W, int H);
int tr_film_clear(tr_ctx* ctx);
int tr_film_download(tr_ctx* ctx, float* hdr /* W*H*3 or NULL */, float* rgb /* W*H*3 or NULL */);
int tr_film_upload(tr_ctx* ctx, const float* hdr /* W*H*3 */);
/* the same download without the last host copy: the film is DMA'd into pinned host buffers owned by the context and their
 * addresses are returned (valid until the next download / tr_film_create; NULL for a film that was not requested).  This is what
 * `.to_numpy()` of the Python classes hands out when asked for a view (example/Example.py:44 reads the film every frame). */
int tr_film_download_pinned(tr_ctx* ctx, int want_hdr, int want_rgb, const float** hdr, const float** rgb);
/* device pointer of hdr (W*H*3 f32) for zero-copy interop */
int tr_film_device_ptr(tr_ctx* ctx, void** hdr_dev, void** rgb_dev);

/* pixel-tile sharding: this context renders the 32x32 tiles t with (tx + 3*ty) % nranks == rank;
 * pixels of other ranks stay 0 in hdr so that a SUM over ranks gives the image. Default (0,1). */
int tr_set_shard(tr_ctx* ctx, int rank, int nranks);

/* ---- multi-GPU (the reference is single-device) ----------------------------------------------------------
 * One context (= one process or thread) per GPU.  Rank 0 obtains an id with tr_comm_unique_id and hands the 128 bytes to the
 * other ranks by any host channel (torch.distributed / MPI broadcast, a file, a socket); every rank then calls tr_comm_init,
 * which creates the NCCL communicator and shards the film (tr_set_shard(rank, nranks)).  tr_film_reduce enqueues ONE
 * ncclReduce (sum, f32, W*H*3) of the per-rank partial films on the context's stream, right behind whatever was rendered:
 * rank 0 (all_ranks != 0: every rank, ncclAllReduce) then presents the full image through tr_tonemap / tr_film_download*.
 * The partial films themselves are left untouched, so rendering more frames and reducing again never double-counts; any
 * render, clear or upload of the film makes the context present its own partial film again.  nranks == 1 is a no-op.
 * libnccl.so.2 is loaded at run time on the first call (TR_ERR_COMM if absent). */
#define TR_COMM_ID_BYTES 128
int tr_comm_unique_id(void* id_out /* TR_COMM_ID_BYTES */);
int tr_comm_init(tr_ctx* ctx, int rank, int nranks, const void* unique_id /* TR_COMM_ID_BYTES, may be NULL when nranks == 1 */);
int tr_film_reduce(tr_ctx* ctx, int all_ranks);
int tr_comm_destroy(tr_ctx* ctx);

/* ---- integrators ------------------------------------------------------------------------ */
/* replaces PT_RGB.PathTrace.render (integrator/PT_RGB.py:44-136) for frames
 * [frame_begin, frame_begin+n_frames): wavefront generate -> trace -> shade -> shadow -> accumulate.
 * Asynchronous on the context stream. stack_size is accepted for API parity (the traversal is
 * stackless); max_depth is the reference's MAX_DEPTH (15). */
int tr_render_pt_rgb(tr_ctx* ctx, int frame_begin, int n_frames, int max_depth, uint64_t seed);
/* replaces PT_Spec.PathTrace.render (integrator/PT_Spec.py:189-280): the same wavefront with four hero wavelengths per
 * path (spectrum/HeroSample.py), spectral reflectances through the rgb2spec model or measured tables (MAT_SPECTRAL),
 * the Sellmeier glass of Glass.sample_lambda (brdf/Glass.py:39-65), the sky dome on a miss, and AddSplat
 * (CIE XYZ -> linear sRGB) as accumulation.  max_depth is the reference's MAX_DEPTH (10).  Requires every
 * tr_spec_*_upload below. */
int tr_render_pt_spec(tr_ctx* ctx, int frame_begin, int n_frames, int max_depth, uint64_t seed);
/* replaces BDPT.render (integrator/BDPT_RGB.py:600-641) for frames [frame_begin, frame_begin+n_frames): eye and light
 * sub-paths (eye_path :104-187, light_path :189-250, Scene.sample_light Scene.py:430-474), every (e, l) connection with
 * 0 <= e+l-2 <= MAX_DEPTH = 5 (connect_path :436-580, Camera.get_image_point Camera.py:144-158) weighted by mis_weight
 * (:258-434), light-tracing (e == 1) contributions splatted across pixels, running mean into hdr.  Needs the view matrix of
 * tr_camera_set and at least one emitter.  With tile sharding every rank's hdr also receives that rank's splats on foreign
 * pixels, so the film reduce must be a SUM.  Asynchronous like tr_render_pt_rgb (tr_stats_get waits).  The e == 1 splats
 * are accumulated with float atomics (as the reference does): their summation order, hence the last bits of the film, may differ
 * from run to run; the e >= 2 strategies are summed in a fixed order. */
int tr_render_bdpt_rgb(tr_ctx* ctx, int frame_begin, int n_frames, uint64_t seed);
/* with "stage_timing" on: device time of the last tr_render_bdpt_rgb spent in the closest-hit kernels of the sub-path stages and
 * in the connection shadow-query kernel (CUDA events around every launch) */
int tr_bdpt_kernel_ms(tr_ctx* ctx, float* ms_trace_kernels, float* ms_shadow_kernel);
/* replaces Debug.render (integrator/Debug.py:44-66); also fills the first-hit buffers */
int tr_render_debug(tr_ctx* ctx);
/* frame-0 primary rays and first hits, index [x*H+y]; any pointer may be NULL */
int tr_first_hit_download(tr_ctx* ctx, float* t, int32_t* prim, float* uv /*2*/, float* pos /*3*/,
                          float* gnormal /*3*/, float* normal /*3*/, float* dir /*3*/);
/* replaces UF.tone_map (UtilsFunc.py:583-586): rgb = srgb(ACES(hdr * exposure)) */
int tr_tonemap(tr_ctx* ctx, float exposure);
int tr_stats_get(tr_ctx* ctx, tr_stats* out);
/* tuning: "batch_frames" (0 = auto: as many frames per batch as "max_paths" allows), "max_paths" (path slots per batch, 252 B each;
 * default 64 Mi = a 64-spp step of a 1024^2 film in ONE batch; BDPT takes twice the byte budget for its 2.9 KB samples),
 * "chains" (parallel wavefront chains per batch), "chain_skew" (percent of a batch's frames given to chain 0), "stage_timing", "graph"
 * (CUDA-graph replay), "smem_bvh" (TMA staging of small trees into shared memory), "replicas" (their bank-conflict-free
 * 8-way replicated image), "top_nodes" (large trees: breadth-first top nodes staged per CTA, 0 = off), "tail_max", "tail_chunk",
 * "shadow_overlap", "pdl", "bdpt_wavefront".  Out-of-range values are rejected with TR_ERR_INVALID. */
int tr_set_option(tr_ctx* ctx, const char* name, int value);

/* ---- spectral tables: replace the from_numpy calls of PT_Spec.setup_data_gpu (integrator/PT_Spec.py:89-98) ---------
 * sensor: CIE 1931 colour matching functions, n x 3 f32 on a regular grid lmin..lmax nm (:57-79,90) */
int tr_spec_sensor_upload(tr_ctx* ctx, const float* xyz, int n, float lmin, float lmax);
/* one tabulated spectrum (spectrum/Spectrum.py:18-40): which = 0 D65 illuminant, 1 white, 2 red, 3 green reflectance
 * (the MAT_SPECTRAL tables selected by the material's alebdoTex, integrator/PT_Spec.py:119-135) */
int tr_spec_spectrum_upload(tr_ctx* ctx, int which, const float* data, int n, float lmin, float lmax);
int tr_spec_spectrum_download(tr_ctx* ctx, int which, float* data /* n */);
/* Jakob-Hanika coefficient table (spectrum/Rgb2Spec.py:15-42): scale res f32, data 3*res^3*3 f32 */
int tr_spec_rgb2spec_upload(tr_ctx* ctx, const float* scale, const float* data, int res);
/* Hosek-Wilkie sky state after Sky.update (sky/Sky.py:84-96,107-159): configs 11 x 9, radiances 11, sun direction */
int tr_spec_sky_upload(tr_ctx* ctx, const float* configs, const float* radiances, const float sun_dir[3]);
/* replaces PT_Spec.normalize_spec (integrator/PT_Spec.py:101-107) = cal_white_point (:174-187) + Spectrum.scale
 * (spectrum/Spectrum.py:53-56): scales table `which` so that its CIE Y is 1; returns the white point before scaling */
int tr_spec_normalize(tr_ctx* ctx, int which, float white_point[3]);

/* ---- unit hooks: the device functions of the shading/traversal kernels run on arrays, for parity
 * tests against the oracle (brdf/Disney.py:17-108, brdf/Glass.py:9-34, UtilsFunc.py:440-461,
 * Scene.py:671-744). Host pointers in, host pointers out. */
int tr_test_disney_evaluate_pdf(tr_ctx* ctx, int n, const float* N, const float* V, const float* L,
                                float metal, float rough, float* out /* n x 2 */);
int tr_test_disney_sample(tr_ctx* ctx, int n, const float* dir, const float* N, float metal,
                          float rough, const float* u /* n x 3 */, float* out /* n x 3 */);
int tr_test_glass_sample(tr_ctx* ctx, int n, const float* dir, const float* N, float ior,
                         const float* u /* n */, float* out /* n x 4 */);
int tr_test_offset_ray(tr_ctx* ctx, int n, const float* p, const float* nrm, float* out /* n x 3 */);
int tr_test_rng(tr_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t frame, uint32_t block, float* out4);
/* include/trmath.h on the device: fn 0 sin, 1 cos, 2 exp, 3 acos (of a), 4 atan2(a, b), 5 pow(a, b); the oracle compiles the same
 * header, so the results must agree bit for bit */
int tr_test_math(tr_ctx* ctx, int fn, int n, const float* a, const float* b /* fn >= 4 */, float* out);
/* arbitrary rays through the simple one-lane-per-ray walk (the Debug integrator's); shadow != 0 adds the shadow-query cross-check */
int tr_test_trace(tr_ctx* ctx, int n, const float* o, const float* d, int shadow,
                  float* t, int32_t* prim, float* uv /* n x 2 or NULL */);
/* the same rays through the PRODUCTION kernels, as queue records, in the tree mode the renderer would use (options "smem_bvh",
 * "replicas", "top_nodes" apply): kernel 0 = tr_test_trace; 1 = persistent closest-hit kernel k_trace; 2 = persistent shadow kernel
 * k_shadow (target[k] = primitive ray k must see: prim[k] = target[k] and t[k] = 1 when the target is the nearest hit, else -2 / 0);
 * 3 = the tail megakernel k_tail (first walk of each path). */
int tr_test_trace_kernel(tr_ctx* ctx, int kernel, int n, const float* o, const float* d, const int32_t* target /* kernel 2 */, int shadow,
                         float* t, int32_t* prim, float* uv /* n x 2 or NULL */);

/* BDPT internals of the last tr_render_bdpt_rgb batch (its first frame) for n pixels: verts n x 13 x 20 f32 (eye 0..6, light
 * 0..5; pos3 normal3 snormal3 beta3 wo3 fpdf rpdf type+16*delta prim mat, zero beyond the sub-path depth), depths n x 2 i32,
 * contrib n x 7 x 7 x 4 f32 indexed [e-1][l] (MIS-weighted radiance of each e >= 2 strategy; e == 1 rows are splats, not kept) */
int tr_test_bdpt_dump(tr_ctx* ctx, int n, const int32_t* px, const int32_t* py, float* verts, int32_t* depths, float* contrib);

/* spectral device functions: Hero.srgb_to_spec (spectrum/HeroSample.py:46-57) for n (srgb, hero wavelength) pairs;
 * Sky.get_solar_radiance (sky/Sky.py:258-265); Spectrum.sample (spectrum/Spectrum.py:43-51) of table `which`, or
 * PT_Spec.sample of the sensor (which = -1, out n x 3) */
int tr_test_srgb_to_spec(tr_ctx* ctx, int n, const float* rgb /* n x 3 */, const float* lambda0 /* n */, float* out4 /* n x 4 */);
int tr_test_sky_radiance(tr_ctx* ctx, int n, const float* theta, const float* gamma, const float* wavelength, float* out);
int tr_test_spectrum_sample(tr_ctx* ctx, int which, int n, const float* lambda, float* out);

#ifdef __cplusplus
}
#endif
#endif /* TIRAY_H */
