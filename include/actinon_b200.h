/* actinon_b200.h — C ABI of the B200-native Actinon sample tracer.
 *
 * The drop-in boundary replaces ONE internal call of the reference renderer:
 *
 *     void lum_machine_s_run( const scene_s* scene, lum_arr_s* lum_arr );
 *                                        (reference src/scene.c:1017-1028, call site src/scene.c:1141)
 *
 * i.e. "given a read-only scene and an array of sample positions in pixel units, write the
 * gamma-applied, clamped RGB of every sample".  Everything below that call (camera rays,
 * scene traversal, every shape / CSG / envelope type, shading, direct and path sampling)
 * runs in hand-written sm_100a CUDA kernels; everything above it stays on the host.
 *
 * Plain C: pointers + sizes only, no C++/torch types.  All functions return 0 on success
 * and a negative acn_status on failure; nothing aborts or throws across this boundary.
 * A handle is thread-compatible (use one handle per host thread / per GPU).
 */
#ifndef ACTINON_B200_H
#define ACTINON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * status codes
 * ------------------------------------------------------------------------------------------- */
typedef enum acn_status
{
    ACN_OK                 =   0,
    ACN_ERR_INVALID_ARG    =  -1,  /* null pointer, bad size, bad handle                              */
    ACN_ERR_BAD_SCENE      =  -2,  /* flat scene fails validation (unknown kind, index out of range)  */
    ACN_ERR_NO_DEVICE      =  -3,  /* no CUDA device / driver: the tracer has NO CPU fallback         */
    ACN_ERR_CUDA           =  -4,  /* a CUDA runtime call failed; see acn_last_error()                */
    ACN_ERR_UNSUPPORTED    =  -5,  /* e.g. a light whose shape has no fov function (objects.c:254-259) */
    ACN_ERR_CANCELLED      =  -6,  /* *cancel became non-zero; outputs are incomplete (scene.c:978)    */
    ACN_ERR_OUT_OF_MEMORY  =  -7,
    ACN_ERR_PARSE          =  -8,  /* .acn front-end: syntax / evaluation error                        */
    ACN_ERR_IO             =  -9
} acn_status;

/* ---------------------------------------------------------------------------------------------
 * Flat scene: what crosses the boundary instead of `const scene_s*`.
 * Indices instead of pointers, doubles as in the reference (f3_t = double, vectors.h:53,106);
 * the device converts to its own packed FP32 layout at upload.
 * ------------------------------------------------------------------------------------------- */

/* node kinds = the reference's object types (quicktypes.h:38-47) */
enum
{
    ACN_KIND_COMPOUND      = 0,  /* compound_s            compound.c:36-50   */
    ACN_KIND_PLANE         = 1,  /* obj_plane_s           objects.c:481-551  */
    ACN_KIND_SPHERE        = 2,  /* obj_sphere_s          objects.c:556-661  */
    ACN_KIND_SQUAROID      = 3,  /* obj_squaroid_s        objects.c:669-831  */
    ACN_KIND_DIST_SPHERE   = 4,  /* obj_distance_s + distance_sphere_s  objects.c:836-970, distance.c:39-42 */
    ACN_KIND_DIST_TORUS    = 5,  /* obj_distance_s + distance_torus_s   distance.c:83-92  */
    ACN_KIND_PAIR_INSIDE   = 6,  /* obj_pair_inside_s     objects.c:975-1120 */
    ACN_KIND_PAIR_OUTSIDE  = 7,  /* obj_pair_outside_s    objects.c:1125-1277 */
    ACN_KIND_NEG           = 8,  /* obj_neg_s             objects.c:1282-1348 */
    ACN_KIND_SCALE         = 9,  /* obj_scale_s           objects.c:1353-1459 */
    ACN_KIND_COUNT         = 10
};

enum { ACN_TEX_NONE = 0, ACN_TEX_PLAIN = 1, ACN_TEX_CHESS = 2 };   /* textures.c:77-148 */

/* One node = one compound_s or one obj_*_s (properties_s + shape tail), objects.h:51-97. */
typedef struct acn_flat_node
{
    int32_t kind;            /* ACN_KIND_*                                                          */
    int32_t child0;          /* CSG: o1.  compound: first index into acn_flat_scene.children        */
    int32_t child1;          /* pair: o2. compound: number of children. otherwise -1                */
    int32_t material;        /* index into materials (objects); -1 for compounds                    */
    int32_t has_envelope;    /* properties_s.envelope / compound_s.envelope present                 */
    int32_t reserved;
    double  env_pos[3];      /* envelope_s { v3d_s pos; f3_t radius; }  objects.c:35-40             */
    double  env_radius;
    double  pos[3];          /* properties_s.pos                                                    */
    double  rax[9];          /* properties_s.rax, rows x,y,z                                        */
    double  surface_roughness;
    double  tail[4];         /* sphere: radius | squaroid: a,b,c,r | distance: inv_scale, ex_radius,
                                cycles | scale: inv_scale.xyz                                       */
} acn_flat_node;

/* Surface part of properties_s (objects.h:51-78) + the texture field if any. */
typedef struct acn_flat_material
{
    double  color[3];
    double  radiance;
    double  refractive_index;
    double  fresnel_reflectivity;     /* raw value; binarised by the shader like scene.c:451 */
    double  chromatic_reflectivity;
    double  diffuse_reflectivity;
    double  sigma;
    double  transparency[3];
    int32_t texture_kind;             /* ACN_TEX_* */
    int32_t reserved;
    double  tex_color1[3];            /* plain: colour.  chess: color1 */
    double  tex_color2[3];
    double  tex_scale;
} acn_flat_material;

/* scene_s render parameters (scene.c:153-213) + what lum_machine_s_func derives from them. */
typedef struct acn_flat_params
{
    int32_t image_width;
    int32_t image_height;
    double  gamma;
    double  background_color[3];
    double  camera_position[3];
    double  camera_view_direction[3];
    double  camera_top_direction[3];
    double  camera_focal_length;
    int32_t trace_depth;
    int32_t direct_samples;
    int32_t path_samples;
    int32_t gradient_samples;
    int32_t gradient_cycles;
    int32_t threads;                  /* informational (CPU reference thread count) */
    double  trace_min_intensity;
    double  max_path_length;
    double  gradient_threshold;
} acn_flat_params;

typedef struct acn_flat_scene
{
    acn_flat_params           params;
    int32_t                   n_nodes;
    int32_t                   n_children;
    int32_t                   n_materials;
    int32_t                   light_root;     /* node index of the `light` compound  (scene.c:178) */
    int32_t                   matter_root;    /* node index of the `matter` compound (scene.c:179) */
    int32_t                   reserved;
    const acn_flat_node*      nodes;
    const int32_t*            children;       /* child node indices of all compounds, concatenated */
    const acn_flat_material*  materials;
} acn_flat_scene;

/* ---------------------------------------------------------------------------------------------
 * Tracer options that have no counterpart in scene_s.
 * ------------------------------------------------------------------------------------------- */
enum
{
    ACN_SEED_POSITION_HASH = 0,  /* reference behaviour: seed from the mantissa bits of hit position
                                    and normal (scene.c:537, vectors.h:177-190, objects.c:269)       */
    ACN_SEED_INDEX_KEYED   = 1   /* implementation-independent: seed = hash(sample index, path in
                                    the ray tree); shared with the oracle for the 1e-3 parity check  */
};

enum { ACN_PRECISION_F32 = 0, ACN_PRECISION_F64 = 1 /* validation only */ };

enum
{
    ACN_CSG_AUTO      = 0,  /* intervals in f32, reference march in f64 (bit-for-bit validation)      */
    ACN_CSG_INTERVALS = 1,  /* classify the ray against every leaf once, combine interval lists       */
    ACN_CSG_MARCH     = 2   /* the reference's alternating march (objects.c:1052-1094,1209-1251)      */
};

enum
{
    ACN_SPECIALIZE_AUTO = 0,  /* on when the environment has ACN_SPECIALIZE=1, else off                       */
    ACN_SPECIALIZE_ON   = 1,  /* compile kernels for this scene's structure at acn_tracer_create (NVRTC, cached
                                 in memory and under ACN_CACHE_DIR); fails with ACN_ERR_UNSUPPORTED if it cannot */
    ACN_SPECIALIZE_OFF  = 2   /* the generic kernels, which interpret the scene tables                          */
};

typedef struct acn_options
{
    int32_t seed_mode;        /* ACN_SEED_*                                                         */
    int32_t precision;        /* ACN_PRECISION_*                                                    */
    double  eps;              /* shell thickness (f3_eps, vectors.h:33). <=0: automatic — 1e-6 in
                                 f64, scene-scale aware in f32                      */
    int64_t wave_budget;      /* path children / explicit rays traced per wavefront iteration; the queues take
                                 ~2.9 KB of device memory per unit.  <=0: follows the call — 64 x its samples,
                                 rounded up to a power of two, between 2^16 (190 MB) and 2^23 (24 GB); the queues
                                 only grow.  Samples do not depend on it (fixed-point sums)                */
    int32_t device;           /* CUDA device ordinal; <0: current device                            */
    int32_t csg_mode;         /* ACN_CSG_*: how composite objects are intersected                   */
    int32_t specialize;       /* ACN_SPECIALIZE_*: scene-specialised kernels (same results, faster)  */
    int32_t reserved;
} acn_options;

/* counters filled per render call: rays by class */
typedef struct acn_stats
{
    uint64_t samples;
    uint64_t rays_primary;
    uint64_t rays_reflection;
    uint64_t rays_chromatic;
    uint64_t rays_refraction;
    uint64_t rays_path;
    uint64_t rays_shadow;        /* shadow test + light self-hit count as one each (scene.c:564,569) */
    uint64_t rays_light;
    uint64_t diffuse_hits;
    uint64_t kernel_launches;    /* CUDA kernels launched by this call                              */
    uint64_t waves;              /* wavefront iterations                                            */
    double   device_ms;          /* CUDA-event time of the device work of this call                 */
    /* per-kernel CUDA-event times, filled when the environment has ACN_PROFILE_KERNELS=1
       (index 0 primary, 1 explicit rays, 2 path children, 3 direct lighting, 4 surface response,
       5 list building, 6 scheduler + ray pop, 7 unused) */
    double   kernel_ms[8];
    uint64_t kernel_launches_by_class[8];
} acn_stats;

typedef struct acn_tracer acn_tracer;   /* opaque: device copy of one scene + queues */

/* ---------------------------------------------------------------------------------------------
 * Tracer (device side)
 * ------------------------------------------------------------------------------------------- */

/* fills defaults: position-hash seeding, f32, automatic eps */
void acn_options_default( acn_options* opt );

/* Validates and uploads a flat scene.  Replaces the implicit `const scene_s*` argument of
 * lum_machine_s_run (scene.c:1017).  The flat scene may be freed afterwards. */
int acn_tracer_create( const acn_flat_scene* scene, const acn_options* opt, acn_tracer** out );
void acn_tracer_destroy( acn_tracer* t );

/* lum_machine_s_run( scene, lum_arr ) (scene.c:1017-1028):
 *   xy      n sample positions in pixel units, interleaved x,y (lum_s.pos, scene.c:682-687);
 *           x right, y down, pixel centre = +0.5
 *   rgb     n output triples: cl_s_sat( colour, gamma ) per sample (scene.c:1010), same order
 *   index_base  global index of xy[0] in the caller's sample stream (used by ACN_SEED_INDEX_KEYED)
 *   cancel  optional; polled between wavefront iterations like signal_received_g (scene.c:978)
 * Host pointers; the copies to and from the device are part of the call. */
int acn_render_samples( acn_tracer* t, const double* xy, uint64_t n, uint64_t index_base,
                        float* rgb, const volatile int* cancel, acn_stats* stats );

/* Same, with xy and rgb resident in device memory (used by the benchmark's device-resident leg
 * and by the multi-GPU accumulation path).  stream is a cudaStream_t cast to void*; 0 is CUDA's legacy
 * default stream (also PyTorch's default stream), so the work is ordered against whatever else the caller
 * enqueued there.  The call returns when the samples are traced (the wavefront scheduler is polled). */
int acn_render_samples_device( acn_tracer* t, const double* d_xy, uint64_t n, uint64_t index_base,
                               float* d_rgb, void* stream, const volatile int* cancel, acn_stats* stats );

/* lum_image_s_push_arr (scene.c:804-820) on the device: adds n samples to a per-pixel
 * accumulator float4[width*height] = (sum r, sum g, sum b, sum weight); pixel = truncation of
 * the sample position.  d_accum is a device pointer (reduced across GPUs by the caller, NCCL). */
int acn_accumulate_device( acn_tracer* t, const double* d_xy, const float* d_rgb, uint64_t n,
                           float* d_accum, void* stream );

/* Diagnostics of the run-time specialisation (needs no GPU: NVRTC cross-compiles for sm_100a): writes the C++ source
 * generated for the scene's structure to src (cap bytes, NUL-terminated; *len = its full length, 0 when the scene does
 * not qualify) and, if compile != 0, compiles it (ACN_ERR_UNSUPPORTED + acn_last_error() on failure). */
int acn_spec_probe( const acn_flat_scene* scene, const acn_options* opt, int compile, char* src, uint64_t cap,
                    uint64_t* len, double* seconds );

/* The tracer's private non-blocking stream (cudaStream_t as void*), the one acn_render_samples uses. */
void* acn_tracer_stream( acn_tracer* t );

const char* acn_last_error( void );
const char* acn_version( void );

/* Number of CUDA devices visible, or a negative acn_status (ACN_ERR_NO_DEVICE). */
int acn_device_count( void );

/* Measures an FP32 FMA peak on the current device (dependent-free FFMA chains on every SM) and
 * returns TFLOP/s; used as the live roofline denominator.  <0: error. */
double acn_measure_fp32_peak_tflops( int device );

/* ---------------------------------------------------------------------------------------------
 * Host side: scene-description API (the part of objects.c:1463-1716, compound.c:140-207,380-455,
 * container.c:376-421 and scene.c:238-279 that the flattener must understand), the .acn
 * front-end, the pass controller and the pnm writer.  Pure host code — works without a GPU.
 * ------------------------------------------------------------------------------------------- */
typedef struct acn_scene acn_scene;     /* opaque: scene_s equivalent (params + light + matter) */
typedef int32_t acn_obj;                /* handle of a host-side value (object / compound / list) */

int  acn_scene_create( acn_scene** out );
void acn_scene_destroy( acn_scene* s );
acn_flat_params* acn_scene_params( acn_scene* s );                    /* mutable, scene_s fields */

/* shape constructors (closures.c:417-591) — return a handle >= 0 or a negative status */
acn_obj acn_create_plane( acn_scene* s );
acn_obj acn_create_sphere( acn_scene* s, double radius );
acn_obj acn_create_squaroid( acn_scene* s, double a, double b, double c, double r );
acn_obj acn_create_ellipsoid( acn_scene* s, double rx, double ry, double rz );
acn_obj acn_create_cylinder( acn_scene* s, double rx, double ry );
acn_obj acn_create_cone( acn_scene* s, double rx, double ry, double rz );
acn_obj acn_create_hyperboloid1( acn_scene* s, double rx, double ry, double rz );
acn_obj acn_create_hyperboloid2( acn_scene* s, double rx, double ry, double rz );
acn_obj acn_create_torus( acn_scene* s, double r1, double r2 );
acn_obj acn_create_distance_sphere( acn_scene* s );

/* value semantics: every combinator clones its operands (objects.c:1011-1018 etc.) */
acn_obj acn_clone( acn_scene* s, acn_obj o );
acn_obj acn_pair_inside( acn_scene* s, acn_obj o1, acn_obj o2 );      /* a & b */
acn_obj acn_pair_outside( acn_scene* s, acn_obj o1, acn_obj o2 );     /* a | b */
acn_obj acn_neg( acn_scene* s, acn_obj o1 );                          /* !a    */
acn_obj acn_scale_object( acn_scene* s, acn_obj o1, const double scale[3] ); /* obj * vec */
acn_obj acn_list_create( acn_scene* s );
int     acn_list_push( acn_scene* s, acn_obj list, acn_obj item );    /* clones item */
acn_obj acn_list_inside_composite( acn_scene* s, acn_obj list );      /* container.c:376-392 */
acn_obj acn_list_outside_composite( acn_scene* s, acn_obj list );     /* container.c:394-410 */
acn_obj acn_list_create_compound( acn_scene* s, acn_obj list );       /* container.c:412-421 */

/* in-place edits; apply recursively to lists / compounds / CSG children like the reference */
int acn_move( acn_scene* s, acn_obj o, const double v[3] );
int acn_rotate( acn_scene* s, acn_obj o, const double m[9] );         /* rows x,y,z */
int acn_scale( acn_scene* s, acn_obj o, double f );
int acn_set_color( acn_scene* s, acn_obj o, const double rgb[3] );
int acn_set_transparency( acn_scene* s, acn_obj o, const double rgb[3] );
int acn_set_refractive_index( acn_scene* s, acn_obj o, double n );    /* rewrites fresnel, objects.c:436-448 */
int acn_set_radiance( acn_scene* s, acn_obj o, double v );
int acn_set_fresnel_reflectivity( acn_scene* s, acn_obj o, double v );
int acn_set_chromatic_reflectivity( acn_scene* s, acn_obj o, double v );
int acn_set_diffuse_reflectivity( acn_scene* s, acn_obj o, double v );
int acn_set_sigma( acn_scene* s, acn_obj o, double v );
int acn_set_surface_roughness( acn_scene* s, acn_obj o, double v );
int acn_set_material( acn_scene* s, acn_obj o, const char* name );    /* objects.c:1589-1682 */
int acn_set_envelope( acn_scene* s, acn_obj o, const double pos[3], double radius );
int acn_set_auto_envelope( acn_scene* s, acn_obj o );                 /* objects.c:470-476, compound.c:73-107 */
/* A conservative bounding sphere from the shape's parameters instead of the reference's Monte-Carlo estimate (1000 random
 * rays, x 1.1): sphere, ellipsoid, torus, A&B, A|B of bounded shapes.  ACN_ERR_UNSUPPORTED for unbounded shapes.  An
 * envelope is part of a shape's semantics (obj_side reports "outside" beyond it, objects.c:365-370), so a scene built
 * with this call renders like the reference renders the same scene with the same envelopes set by set_envelope. */
int acn_set_bounding_envelope( acn_scene* s, acn_obj o );
int acn_set_texture_plain( acn_scene* s, acn_obj o, const double rgb[3] );
int acn_set_texture_chess( acn_scene* s, acn_obj o, const double rgb1[3], const double rgb2[3], double scale );

int acn_scene_clear( acn_scene* s );                                  /* scene.c:671-675 */
int acn_scene_push( acn_scene* s, acn_obj o );                        /* scene.c:238-279 */

/* Evaluates an .acn script (interpreter.c) up to and excluding image creation.  Every
 * `scene.create_image(file)` call is recorded instead of rendered: *n_images receives the count;
 * use acn_scene_image_name / acn_scene_select_image to get each frame's scene state. */
int acn_scene_load_acn( acn_scene* s, const char* path, int argc, const char* const* argv, int* n_images );
const char* acn_scene_image_name( acn_scene* s, int image_index );
int acn_scene_select_image( acn_scene* s, int image_index );

/* Flatten (the "thin C-ABI layer" of the north star).  The returned flat scene is owned by the
 * acn_scene and stays valid until the next flatten / destroy. */
int acn_scene_flatten( acn_scene* s, const acn_flat_scene** out );

/* ---------------------------------------------------------------------------------------------
 * Pass controller (scene.c:1032-1165): pass 0 = pixel centres, passes 1..gradient_cycles =
 * jittered re-sampling of high-gradient pixels; accumulation, .pnm output, resume.
 * ------------------------------------------------------------------------------------------- */
typedef struct acn_image acn_image;    /* lum_image_s equivalent (scene.c:760-768) */

int  acn_image_create( int32_t width, int32_t height, acn_image** out );   /* lum_image_s_reset, rval = 21943294 */
void acn_image_destroy( acn_image* im );
int  acn_image_size( const acn_image* im, int32_t* width, int32_t* height );
int32_t  acn_image_cycle( const acn_image* im );
uint64_t acn_image_rval( const acn_image* im );
/* builds the sample list of the next pass (scene.c:1108-1139); *xy is owned by the image */
int  acn_image_next_pass( acn_image* im, const acn_flat_params* prm, const double** xy, uint64_t* n );
/* lum_image_s_push_arr for the pass just built (scene.c:1156), advances the cycle */
int  acn_image_push( acn_image* im, const double* xy, const float* rgb, uint64_t n );
/* per-pixel averages (lum_image_s_get_avg, scene.c:824-835): rgb float[w*h*3] */
int  acn_image_average( const acn_image* im, float* rgb );
/* raw sums: double[w*h*6] = pos.x,pos.y,clr.r,clr.g,clr.b,weight */
int  acn_image_sums( const acn_image* im, double* sums );
int  acn_image_add_sums( acn_image* im, const double* sums );
/* P6 writer + FNV-1a fold hash of the packed pixels (scene.c:76-82,122-146,866-885) */
int  acn_image_write_pnm( const acn_image* im, const char* path, uint64_t* hash );
/* checkpoint / resume (scene.c:1068-1106,1143-1153): own format, same fields */
int  acn_image_save( const acn_image* im, const char* path );
int  acn_image_load( const char* path, acn_image** out );
/* replaces the whole state: sums as acn_image_sums returns them, pass counter, jitter stream */
int  acn_image_set_state( acn_image* im, int32_t cycle, uint64_t rval, const double* sums );

/* ---------------------------------------------------------------------------------------------
 * The same pass controller with the image RESIDENT ON THE DEVICE (scene.c:804-862,1103-1159 as kernels): gradient
 * selection, sample list (prefix sum + LCG skip-ahead: bit-identical to acn_image_next_pass), accumulation.  Sums are
 * 64-bit fixed point, so an image is bit-identical however its samples are spread over GPUs.
 *
 *   one GPU:      while( acn_dimage_render_pass( d, t, prm, base, &n, 0, cancel, &st ) == 0 && n ) base += n;
 *   several GPUs: acn_dimage_set_shard( d, n_ranks, rank, 4 ) on every rank's image, then per pass
 *                 acn_dimage_render_pass (traces and accumulates the rank's own pixel tiles into the pass delta),
 *                 sum acn_dimage_delta over the ranks (ncclAllReduce, ncclUint64, ncclSum), acn_dimage_end_pass.
 *                 acn_group_* below does exactly that inside one process.
 * ------------------------------------------------------------------------------------------- */
typedef struct acn_dimage acn_dimage;

int  acn_dimage_create( int device, int32_t width, int32_t height, acn_dimage** out );     /* device < 0: current */
void acn_dimage_destroy( acn_dimage* d );
/* pixels are dealt to the ranks in tile x tile squares along a Morton curve; default: one rank */
int  acn_dimage_set_shard( acn_dimage* d, int32_t n_ranks, int32_t rank, int32_t tile );
/* the rank that owns pixel (x, y) under that dealing (pure host function; -1 on bad arguments) */
int32_t acn_pixel_owner( int32_t x, int32_t y, int32_t tile, int32_t n_ranks );
int  acn_dimage_reset( acn_dimage* d );                                                    /* lum_image_s_reset, scene.c:790-800 */
int32_t  acn_dimage_cycle( const acn_dimage* d );
uint64_t acn_dimage_rval( const acn_dimage* d );
void*    acn_dimage_stream( acn_dimage* d );                                                /* cudaStream_t of its kernels */
/* builds this rank's sample list of the next pass on the device (scene.c:1108-1139).  *n_local = 0 and *n_total = 0
 * when all gradient_cycles + 1 passes are done; *d_xy (device, n_local pairs) stays valid until the next begin_pass */
int  acn_dimage_begin_pass( acn_dimage* d, const acn_flat_params* prm, const double** d_xy, uint64_t* n_local, uint64_t* n_total );
/* lum_image_s_push_arr into the pass delta (device pointers) */
int  acn_dimage_accumulate( acn_dimage* d, const double* d_xy, const float* d_rgb, uint64_t n, void* stream );
/* the pass delta: uint64[ width*height*6 ] = pos.x, pos.y, r, g, b (Q20.44), weight — device pointer */
uint64_t* acn_dimage_delta( acn_dimage* d, uint64_t* n_words );
/* delta -> / <- a caller-owned device buffer of n_words uint64 (e.g. a torch tensor that is all-reduced over NCCL) */
int  acn_dimage_copy_delta( acn_dimage* d, uint64_t* d_dst, void* stream );
int  acn_dimage_set_delta( acn_dimage* d, const uint64_t* d_src, void* stream );
/* this rank's sample list of the current / last pass, copied to the host (n_local pairs; tests, diagnostics) */
int  acn_dimage_read_pass_xy( acn_dimage* d, double* xy );
/* totals += delta, delta = 0, advances the pass counter and the jitter stream */
int  acn_dimage_end_pass( acn_dimage* d, void* stream );
/* begin_pass + lum_machine_s_run on the device + accumulate (+ end_pass when the image has one rank) */
int  acn_dimage_render_pass( acn_dimage* d, acn_tracer* t, const acn_flat_params* prm, uint64_t index_base, uint64_t* n_local,
                             uint64_t* n_total, const volatile int* cancel, acn_stats* stats );
/* to / from the host image (pnm output, checkpoint, resume) */
int  acn_dimage_download( acn_dimage* d, acn_image* im );
int  acn_dimage_upload( acn_dimage* d, const acn_image* im );

/* ---------------------------------------------------------------------------------------------
 * One image on several GPUs of one box, inside one process: what replaces the single call site
 * lum_machine_s_run( scene, lum_arr ) + lum_image_s_push_arr (scene.c:1141,1156) when more than one device is to work
 * on an image.  One worker thread, one tracer and one device image per GPU; pixel tiles are dealt to the GPUs, each
 * pass's per-pixel sums are exchanged straight through peer memory (NVLink / NVSwitch).  devices: n CUDA ordinals
 * (NULL: 0..n-1; an ordinal may repeat).  The images are bit-identical for every n.
 *
 *   acn_group_create( flat, &opt, NULL, 8, &g );
 *   while( acn_group_render_pass( g, base, &n, cancel, &st ) == 0 && n ) { base += n; acn_group_download( g, image ); ... }
 * ------------------------------------------------------------------------------------------- */
typedef struct acn_group acn_group;

int  acn_group_create( const acn_flat_scene* scene, const acn_options* opt, const int32_t* devices, int32_t n, acn_group** out );
void acn_group_destroy( acn_group* g );
int  acn_group_size( const acn_group* g );
int  acn_group_uses_peer_access( const acn_group* g );
/* one pass of the controller on all GPUs; *n_samples = samples of the pass (0: all passes done); stats summed over the GPUs */
int  acn_group_render_pass( acn_group* g, uint64_t index_base, uint64_t* n_samples, const volatile int* cancel, acn_stats* stats );
int  acn_group_download( acn_group* g, acn_image* im );
int  acn_group_upload( acn_group* g, const acn_image* im );
acn_dimage* acn_group_image( acn_group* g, int32_t rank );

#ifdef __cplusplus
}
#endif

#endif /* ACTINON_B200_H */
