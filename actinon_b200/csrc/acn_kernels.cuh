// acn_kernels.cuh — device side of the wavefront sample tracer for sm_100a (templated on the real type).
//
// This header is compiled twice: ahead of time by nvcc into libactinon_b200.so (the generic kernels, which interpret the
// scene tables), and at run time by NVRTC (acn_spec.h) together with a generated header that restates ONE scene's
// structure as straight-line code (ACN_SPEC).  It therefore includes no host headers.
//
//
// Replaces lum_machine_s_func + scene_s_lum (reference src/scene.c:420-667,956-1013): the
// reference's recursive, branching ray tree becomes three kinds of device work items
//
//   explicit rays   primary (generated from the camera in-kernel), Fresnel-reflection, chromatic
//                   and refraction rays — SoA queue, 64 B/ray in f32
//   tasks           diffuse hits (position, normal, Oren-Nayar terms, RGB throughput, RNG state,
//                   sample counts) — 104 B/task; their direct_samples*lights shadow rays and
//                   path_samples indirect rays are never materialised: lane i of a warp
//                   regenerates child i from (task, i) with an O(1) LCG skip-ahead table
//   contributions   atomically added to a per-sample RGB accumulator; gamma + clamp at the end
//
// All of scene_s_lum is linear in its children, so colour products and exit absorption fold into
// a per-ray RGB throughput; the scalar intensity is carried separately because it drives sample
// counts and termination (scene.c:428,553,593).
//
// Scheduling is depth-first in bulk: a wave pops at most `budget` work items from the top of the
// ray stack (or tasks worth at most `budget` path children), and everything it spawns lands on
// top, so the memory in flight stays bounded however large ds*ps*ps gets.
#pragma once
#pragma once

#include "acn_geom.h"
#include "acn_isect.cuh"

namespace acn {

enum { TEX_NONE = 0, TEX_PLAIN = 1, TEX_CHESS = 2 };      // = ACN_TEX_* of the C ABI (checked in acn_tracer.cu)

// ---------------------------------------------------------------------------------------------
// device-side scene tables
// ---------------------------------------------------------------------------------------------
template <typename R> struct DMat
{
    R   color[ 3 ];
    R   radiance;
    R   refr;
    R   fresnel01;      // (fresnel_reflectivity != 0 && n != 1) ? 1 : 0      scene.c:451
    R   chroma;
    R   diffuse;
    R   on_a, on_b;     // scene.c:455-461
    R   transp[ 3 ];
    R   tex1[ 3 ], tex2[ 3 ], tex_scale;
    int transparent;    // |transparency|^2 > 0                               scene.c:454
    int tex_kind;
};

template <typename R> struct DLight
{
    int node;
    R   pos[ 3 ];       // prp.pos of the light
    R   color[ 3 ];     // obj_color( light, light.pos )                       scene.c:552
    R   radiance;
};

template <typename R> struct DParams
{
    SceneView<R> sv;
    const DMat<R>*   mats;
    const DLight<R>* lights;
    int   n_lights, n_materials;
    int   n_nodes, n_children, n_prog;
    int   width, height;
    R     gamma;
    V3<R> background;
    V3<R> cam_pos, cam_rx, cam_ry, cam_rz;   // camera_rotation columns (scene.c:963-973)
    R     focal;
    int   trace_depth;
    int   direct_samples, path_samples;
    R     min_intensity;
    R     max_path_length;
    R     eps_rel;                            // per-ray shell thickness = max( sv.eps, eps_rel * |origin|_inf ); 0: constant
    int   stage_bytes;                        // node table bytes staged into shared memory (0: none)
    unsigned int off_geo, off_link, off_pref, off_crec, off_par, off_prog;   // byte offsets of the staged tables (env at 0)
    int   n_heavy;                            // envelopes of the expensive top-level objects (CSG, distance fields); -1: no split
    int   heavy[ 8 ];
    const u64* skipA;                         // LCG skip table: state after 2k steps = s*A[k] + C[k]
    const u64* skipC;
    int   skip_n;
};

// ---------------------------------------------------------------------------------------------
// work-item storage
// ---------------------------------------------------------------------------------------------
enum { RC_PRIMARY = 0, RC_REFLECT = 1, RC_CHROMATIC = 2, RC_REFRACT = 3, RC_PATH = 4 };
enum { RAYF_PROBE = 1 << 16 };     // child whose shading would return 0: only "hit anything?" matters

template <typename R> struct RayBuf     // SoA
{
    R4<R>* o_i;     // origin.xyz, intensity
    R4<R>* d_;      // dir.xyz, -
    R4<R>* tp;      // throughput rgb, -
    I4*    meta;    // depth | class<<8 | flags, sample, key lo, key hi
};

template <typename R> struct HitBuf     // SoA: hits waiting for scene_s_lum (k_shade)
{
    R4<R>* o_a;       // ray origin, hit distance
    R4<R>* d_i;       // ray direction, intensity
    R4<R>* n_e;       // trans.exit_nor, hit_eps
    R4<R>* tp;        // throughput rgb, -
    I4*    meta;      // depth, sample, exit_obj, enter_obj
    u64*   key;       // ray-tree key
};

template <typename R> struct TaskBuf    // SoA
{
    R4<R>* pos_id;    // pos.xyz, diffuse_intensity
    R4<R>* nrm_ci;    // surface.d (= -exit_nor), cos(theta_i)
    R4<R>* prj_a;     // ray projection, on_a
    R4<R>* tpc_b;     // throughput * surface colour, on_b
    I4*    meta;      // sample, depth, n_direct, n_path
    u64*   rv0;       // RNG state at the hit (scene.c:537)
    u64*   key;       // ray-tree key (index-keyed seeding)
    u64*   cum;       // inclusive running sum of n_path on the task stack
};


enum
{
    ST_PRIMARY = 0, ST_REFLECT, ST_CHROMATIC, ST_REFRACT, ST_PATH, ST_SHADOW, ST_LIGHT, ST_DIFFUSE, ST_COUNT
};

#define ACN_TASK_SHIFT 38
#define ACN_TASK_MASK  ( ( 1ull << ACN_TASK_SHIFT ) - 1 )
#define ACN_NONE64     ( ~0ull )

// Device-resident scheduler state.  The host never sizes a launch from it: every kernel runs on a
// fixed persistent grid and takes its work range from the plan k_sched wrote, so a whole wavefront
// iteration is enqueued without a host round trip (the host only polls `done` every few iterations).
struct Sched
{
    // stacks
    unsigned long long nr_a, nr_b;      // rays on the two ends of the ray stack: A = reflection / chromatic (grows up from slot 0),
                                        // B = refraction (grows down from the last slot)
    unsigned long long nt, nt_cum;      // tasks on the task stack, their outstanding path children
    // plan of the current iteration
    unsigned long long base_a, base_b;  // new rays are appended at A[ base_a + out_a++ ] / B[ base_b + out_b++ ]
    unsigned long long take_a, take_b;  // rays popped from the top of each end for k_rays (read in place)
    unsigned long long path_blk_lo, path_blk_hi;   // 32-child blocks of the task stack traced by k_path
    unsigned long long path_c_hi;       // first child index beyond the stack top
    unsigned long long path_nt;         // stack height seen by k_path (window bound)
    unsigned long long fix_slot, fix_cum;   // partially consumed task: cum[ fix_slot ] = fix_cum after k_path
    unsigned long long prim_first, prim_count;
    // work cursors of the persistent kernels (units: chunks)
    unsigned long long cur_rays, cur_path, cur_index, cur_direct, cur_primary, cur_shade;
    // per-iteration outputs
    unsigned long long out_ab;          // rays appended in this iteration: end A in the low, end B in the high 32 bits (one atomic reserves both)
    unsigned long long hits;            // hits appended to the hit queue by the tracing kernels
    unsigned long long tasks_new;       // diffuse hits appended to the new-task scratch
    unsigned long long dl_packed;       // direct list: entries << 38 | shadow children
    unsigned long long task_stack;      // task stack: height << 38 | outstanding path children
    // status
    unsigned long long waves;
    unsigned long long stats[ ST_COUNT ];
    int done;
    int overflow;
};

template <typename R> struct Acc;
template <> struct Acc<float>  { typedef unsigned long long T; };
template <> struct Acc<double> { typedef double T; };

template <typename R> struct Wave       // everything a kernel needs
{
    DParams<R>  prm;
    RayBuf<R>   rays_out;     // ray stack
    TaskBuf<R>  tasks_out;    // new-task scratch
    HitBuf<R>   hits_out;     // hit queue
    Sched*      sc;
    typename Acc<R>::T* accum; // per-sample sums: r, g, b, saturation flags (4 words per sample)
    unsigned long long rays_cap;
    unsigned long long tasks_cap;
    unsigned long long hits_cap;
    u64         index_base;   // global index of sample 0 (index-keyed seeding)
};

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
#define ACN_FULL 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------
// Per-sample accumulation.  A sample's colour is the sum of thousands of contributions that arrive in an order the
// wavefront scheduler decides anew in every run.  The f32 product path therefore sums them as 64-bit FIXED-POINT
// integers (Q28.36): integer addition is associative, so a sample's value is bit-identical from run to run, whatever
// the wave budget, the chunking of the work lists or the number of GPUs the image is spread over — the property the
// multi-GPU accumulation relies on.  Resolution 1.5e-11; a contribution of 1024 or more (the sample
// saturates to 1 after cl_s_sat anyway) only sets the channel's saturation flag, so the integer cannot wrap.  The FP64
// validation mode keeps plain double atomics (it is compared with the oracle to 1e-6, not run to run).
// ---------------------------------------------------------------------------------------------
#define ACN_ACC_SCALE   68719476736.0f          // 2^36
#define ACN_ACC_INV     1.4551915228366852e-11  // 2^-36
#define ACN_ACC_SAT     1024.0f
__device__ __forceinline__ unsigned long long to_acc( float c, unsigned int* sat, int ch )
{
    if( !( c > 0.0f ) ) return 0ull;
    if( c >= ACN_ACC_SAT ) { *sat |= 1u << ch; return 0ull; }
    return ( unsigned long long )__float2ll_rn( c * ACN_ACC_SCALE );
}
__device__ __forceinline__ double to_acc( double c, unsigned int*, int ) { return c; }
template <typename R> struct AccV { typename Acc<R>::T x, y, z; unsigned int sat; };
template <typename R> __device__ __forceinline__ AccV<R> acc_of( V3<R> c )
{
    AccV<R> a; a.sat = 0;
    a.x = to_acc( c.x, &a.sat, 0 ); a.y = to_acc( c.y, &a.sat, 1 ); a.z = to_acc( c.z, &a.sat, 2 );
    return a;
}
template <typename R> __device__ __forceinline__ AccV<R> acc_zero() { AccV<R> a; a.x = a.y = a.z = 0; a.sat = 0; return a; }
template <typename R> __device__ __forceinline__ bool acc_any( const AccV<R>& a ) { return a.x != 0 || a.y != 0 || a.z != 0 || a.sat != 0; }
__device__ __forceinline__ void atomic_add_acc( unsigned long long* p, unsigned long long v ) { if( v ) atomicAdd( p, v ); }
__device__ __forceinline__ void atomic_add_acc( double* p, double v ) { if( v != 0.0 ) atomicAdd( p, v ); }
__device__ __forceinline__ void atomic_or_acc( unsigned long long* p, unsigned int v ) { atomicOr( p, ( unsigned long long )v ); }
__device__ __forceinline__ void atomic_or_acc( double*, unsigned int ) {}

// warp-aggregated counter increment: one atomic per converged group
__device__ __forceinline__ unsigned long long agg_inc( unsigned long long* ctr )
{
    const unsigned int m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs( m ) - 1;
    unsigned long long base = 0;
    if( lane == leader ) base = atomicAdd( ctr, ( unsigned long long )__popc( m ) );
    return __shfl_sync( m, base, leader ) + __popc( m & ( ( 1u << lane ) - 1u ) );
}

__device__ __forceinline__ void agg_count( unsigned long long* ctr )
{
    const unsigned int m = __activemask();
    if( ( int )( threadIdx.x & 31 ) == __ffs( m ) - 1 ) atomicAdd( ctr, ( unsigned long long )__popc( m ) );
}

// whole-warp sum of a per-lane count, added to a global counter by lane 0
__device__ __forceinline__ void warp_count( unsigned long long* ctr, unsigned long long v, int lane )
{
    #pragma unroll
    for( int o = 16; o > 0; o >>= 1 ) v += __shfl_down_sync( ACN_FULL, v, o );
    if( lane == 0 && v ) atomicAdd( ctr, v );
}

// persistent-warp work fetch: `per` consecutive units per atomic
__device__ __forceinline__ unsigned long long warp_fetch( unsigned long long* cursor, unsigned long long per, int lane )
{
    unsigned long long b = 0;
    if( lane == 0 ) b = atomicAdd( cursor, per );
    return __shfl_sync( ACN_FULL, b, 0 );
}

// lanes hold the inclusive running counts of 32 consecutive list entries (ACN_NONE64 beyond the list):
// number of entries whose count is <= idx, i.e. the entry that owns child idx
__device__ __forceinline__ int window_find( unsigned long long incl, unsigned long long idx )
{
    int j = 0;
    #pragma unroll
    for( int s = 16; s > 0; s >>= 1 )
    {
        const unsigned long long v = __shfl_sync( ACN_FULL, incl, j + s - 1 );
        if( v <= idx ) j += s;
    }
    return j;
}

// sum over runs of equal keys (keys are non-decreasing over the lanes); valid at the first lane of a run
template <typename T> __device__ __forceinline__ T seg_sum( T v, int key, int lane )
{
    #pragma unroll
    for( int o = 1; o < 32; o <<= 1 )
    {
        const T   v2 = __shfl_down_sync( ACN_FULL, v, o );
        const int k2 = __shfl_down_sync( ACN_FULL, key, o );
        if( lane + o < 32 && k2 == key ) v += v2;
    }
    return v;
}

// segmented sum of per-lane contributions over runs of equal keys (non-decreasing over the lanes); valid at the first
// lane of a run.  Integer (or double) adds: the grouping of the children into warps does not change the total.
// whole-warp sum of a fixed-point value < 2^46 per lane with the warp-reduce unit: two 23-bit halves, each sum < 2^28
__device__ __forceinline__ unsigned long long warp_sum_acc( unsigned long long v )
{
    const unsigned int lo = __reduce_add_sync( ACN_FULL, ( unsigned int )( v & 0x7FFFFFull ) );
    const unsigned int hi = __reduce_add_sync( ACN_FULL, ( unsigned int )( v >> 23 ) );
    return ( unsigned long long )lo + ( ( unsigned long long )hi << 23 );
}
__device__ __forceinline__ double warp_sum_acc( double v )
{
    #pragma unroll
    for( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync( ACN_FULL, v, o );
    return v;
}

template <typename R> __device__ __forceinline__ AccV<R> seg_sum_acc( AccV<R> v, int key, int lane )
{
    // one segment over the whole warp (the usual case: a task has hundreds of children): three warp reductions
    // (several tasks in the block: the shuffle ladder below.  __match_any_sync + __reduce_add_sync over each lane's own group was
    // measured slower: wine_glass 25.8 -> 27.0 ms, diamond 10.6 -> 11.3)
    if( __all_sync( ACN_FULL, key == __shfl_sync( ACN_FULL, key, 0 ) ) )
    {
        v.x = warp_sum_acc( v.x ); v.y = warp_sum_acc( v.y ); v.z = warp_sum_acc( v.z );
        v.sat = __reduce_or_sync( ACN_FULL, v.sat );
        return v;
    }
    #pragma unroll
    for( int o = 1; o < 32; o <<= 1 )
    {
        const int k2 = __shfl_down_sync( ACN_FULL, key, o );
        const typename Acc<R>::T x2 = __shfl_down_sync( ACN_FULL, v.x, o ), y2 = __shfl_down_sync( ACN_FULL, v.y, o ), z2 = __shfl_down_sync( ACN_FULL, v.z, o );
        const unsigned int s2 = __shfl_down_sync( ACN_FULL, v.sat, o );
        if( lane + o < 32 && k2 == key ) { v.x += x2; v.y += y2; v.z += z2; v.sat |= s2; }
    }
    return v;
}

template <typename R> __device__ __forceinline__ void add_sample_acc( const Wave<R>& w, int sample, const AccV<R>& c )
{
    typename Acc<R>::T* a = w.accum + 4ull * ( unsigned long long )sample;
    atomic_add_acc( a + 0, c.x );
    atomic_add_acc( a + 1, c.y );
    atomic_add_acc( a + 2, c.z );
    if( c.sat ) atomic_or_acc( a + 3, c.sat );
}
template <typename R> __device__ __forceinline__ void add_sample( const Wave<R>& w, int sample, V3<R> c )
{
    add_sample_acc( w, sample, acc_of( c ) );
}

// the node table of the kernel instantiation: SH = true copies it into shared memory (C1/C2/C4: a few KB; the host
// launches these instantiations only when it fits), SH = false reads it where it lies (many_spheres: 1.3 MB in L2)
template <typename R, bool SH> __device__ __forceinline__ SceneView<R, SH> stage_scene( const DParams<R>& prm, unsigned char* smem )
{
    if constexpr( !SH ) return prm.sv;
    else
    {
        const int n = prm.n_nodes;
        // table offsets come from the host through the parameter bank: a table address is then one constant-bank add
        // away from the element index (computed here from n they were re-derived, ~40 instructions, before every access)
        R4<R>* s_env  = reinterpret_cast<R4<R>*>( smem );
        R4<R>* s_geo  = reinterpret_cast<R4<R>*>( smem + prm.off_geo );
        I4*    s_link = reinterpret_cast<I4*>( smem + prm.off_link );
        I4*    s_pref = reinterpret_cast<I4*>( smem + prm.off_pref );
        CRec<R>* s_crec = reinterpret_cast<CRec<R>*>( smem + prm.off_crec );
        int*   s_par  = reinterpret_cast<int*>( smem + prm.off_par );
        int*   s_prog = reinterpret_cast<int*>( smem + prm.off_prog );
        for( int i = threadIdx.x; i < n; i += blockDim.x )
        {
            s_env[ i ] = prm.sv.env[ i ]; s_link[ i ] = prm.sv.link[ i ];
            s_pref[ i ] = prm.sv.prog_ref[ i ]; s_par[ i ] = prm.sv.parent[ i ];
        }
        for( int i = threadIdx.x; i < n * GEO_STRIDE; i += blockDim.x ) s_geo[ i ] = prm.sv.geo[ i ];
        for( int i = threadIdx.x; i < prm.n_children; i += blockDim.x ) s_crec[ i ] = prm.sv.crec[ i ];
        for( int i = threadIdx.x; i < prm.n_prog; i += blockDim.x ) s_prog[ i ] = prm.sv.prog[ i ];
        __syncthreads();
        SceneView<R, true> sv;
#ifndef ACN_NO_OPAQUE_BASE
        unsigned int base;      // opaque to the optimiser: held in one register instead of being re-derived from SR_CgaCtaId at every use
        asm volatile( "{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"( base ) : "l"( smem ) );
#else
        const unsigned int base = ( unsigned int )__cvta_generic_to_shared( smem );
#endif
        sv.env.a      = base;
        sv.geo.a      = base + prm.off_geo;
        sv.link.a     = base + prm.off_link;
        sv.prog_ref.a = base + prm.off_pref;
        sv.crec.a     = base + prm.off_crec;
        sv.parent.a   = base + prm.off_par;
        sv.prog.a     = base + prm.off_prog;
        sv.children.a = 0;                       // march / host only
        sv.eps = prm.sv.eps; sv.light_root = prm.sv.light_root; sv.matter_root = prm.sv.matter_root; sv.seed_mode = prm.sv.seed_mode;
        sv.rec_light = prm.sv.rec_light; sv.rec_matter = prm.sv.rec_matter;
        for( int k = 0; k < 8; k++ ) sv.rec_matter_oct[ k ] = prm.sv.rec_matter_oct[ k ];
        return sv;
    }
}

// obj_color (objects.c:411-422) with txm_plain / txm_chess (textures.c:99-102,142-148)
template <typename R, bool SH> __device__ __forceinline__ V3<R> obj_color( const DMat<R>* mats, const SceneView<R, SH>& sv0, int node, V3<R> pos )
{
    const DMat<R>& m = mats[ sv0.link[ node ].w ];
    if( m.tex_kind == TEX_NONE )  return v3<R>( m.color[ 0 ], m.color[ 1 ], m.color[ 2 ] );
    if( m.tex_kind == TEX_PLAIN ) return v3<R>( m.tex1[ 0 ], m.tex1[ 1 ], m.tex1[ 2 ] );
    R u, v;
    obj_projection( sv0, node, pos, &u, &v );
    long long x = llrint( ( double )( u * m.tex_scale ) );
    long long y = llrint( ( double )( v * m.tex_scale ) );
    return ( ( x ^ y ) & 1 ) ? v3<R>( m.tex1[ 0 ], m.tex1[ 1 ], m.tex1[ 2 ] ) : v3<R>( m.tex2[ 0 ], m.tex2[ 1 ], m.tex2[ 2 ] );
}

// The reference's shell thickness is an absolute 1e-6 in FP64.  In FP32 a hit distance carries an error
// of a few ulp of the ray origin's magnitude, so the product path scales the shell with the origin
// (16 ulp) and never goes below the reference's 1e-6; see DESIGN.md "eps".
template <typename R, bool SH> __device__ __forceinline__ SceneView<R, SH> ray_view( const DParams<R>& prm, const SceneView<R, SH>& sv0, V3<R> o )
{
    SceneView<R, SH> sv = sv0;
    if( prm.eps_rel > R( 0 ) ) sv.eps = r_max( sv.eps, prm.eps_rel * r_max( r_max( r_abs( o.x ), r_abs( o.y ) ), r_abs( o.z ) ) );
    return sv;
}

template <typename R> __device__ __forceinline__ u64 skip2( const DParams<R>& prm, u64 s, unsigned long long k )
{
    if( k < ( unsigned long long )prm.skip_n ) return s * prm.skipA[ k ] + prm.skipC[ k ];
    return lcg00_skip( s, 2ull * k );
}

template <typename R> __device__ __forceinline__ void write_ray( const Wave<R>& w, unsigned long long slot, V3<R> p, V3<R> d, R intensity, int depth,
                                                               V3<R> tp, int cls, int sample, u64 key )
{
    int flags = 0;
    if( depth == 0 || intensity < w.prm.min_intensity ) flags |= RAYF_PROBE;
    R4<R> a; a.x = p.x; a.y = p.y; a.z = p.z; a.w = intensity;
    R4<R> b; b.x = d.x; b.y = d.y; b.z = d.z; b.w = R( 0 );
    R4<R> c; c.x = tp.x; c.y = tp.y; c.z = tp.z; c.w = R( 0 );
    I4 m; m.x = depth | ( cls << 8 ) | flags; m.y = sample; m.z = ( int )( unsigned )( key & 0xFFFFFFFFull ); m.w = ( int )( unsigned )( key >> 32 );
    w.rays_out.o_i[ slot ] = a; w.rays_out.d_[ slot ] = b; w.rays_out.tp[ slot ] = c; w.rays_out.meta[ slot ] = m;
}

#define ACN_BLOCK 128
#define ACN_SHADE_MATS 24
struct ShadeShared
{
    unsigned long long n_ab[ ACN_BLOCK / 32 ], base_ab, base_t, c0;
    unsigned int n_t[ ACN_BLOCK / 32 ];
};

// ---------------------------------------------------------------------------------------------
// scene_s_lum (scene.c:420-667) for one hit per lane: emits child rays and at most one diffuse task.
// BLOCK-COOPERATIVE: all threads of the block call it (live = this thread holds a hit).  The surface response is planned first —
// which of the reflection / chromatic / refraction rays and the diffuse task this hit spawns, with the intensity
// each stage passes on (scene.c updates `intensity` between the stages) — then the warp reserves the queue slots of
// ALL its emissions with one atomic for the rays and one for the tasks, then the lanes write.  With an atomic per
// emission site (up to four dependent round trips to one contended address per 32 hits) k_shade ran at the
// latency of those atomics: 2 TB/s, and slower with more resident warps.
// ---------------------------------------------------------------------------------------------
template <typename R, bool SH> __device__ __forceinline__ void shade_hits( const Wave<R>& w, const SceneView<R, SH>& sv0, bool live, const Ray<R>& ray, R a, R hit_eps,
                                                                          const Trans<R>& tr, int depth, R I, V3<R> tp, int sample, u64 key, int lane,
                                                                          ShadeShared* sh, const DMat<R>* mats, unsigned long long base_a, unsigned long long base_b )
{
    const DParams<R>& prm = w.prm;
    live = live && !( depth == 0 || I < prm.min_intensity );                             // scene.c:428
    const V3<R> pos = madd( ray.p, ray.d, a );
    const DMat<R>* me = nullptr;
    if( live && tr.enter_obj >= 0 ) me = &mats[ sv0.link[ tr.enter_obj ].w ];
    if( live && me && me->radiance > R( 0 ) )                                            // scene.c:432-437
    {
        R d2 = sqr( pos - xyz( sv0.geo[ tr.enter_obj * GEO_STRIDE ] ) );
        R li = d2 > R( 0 ) ? me->radiance / d2 : Num<R>::mag();
        add_sample( w, sample, mul( obj_color( mats, sv0, tr.enter_obj, pos ), tp ) * ( li * I ) );
        agg_count( &w.sc->stats[ ST_LIGHT ] );
        live = false;
    }

    // ---- plan
    R nrel = R( 1 ), F = R( 0 ), C = R( 0 ), Dff = R( 0 ), on_a = R( 1 ), on_b = R( 0 );
    bool T = false;
    R I_refl = R( 0 ), I_chro = R( 0 ), I_diff = R( 0 ), I_refr = R( 0 );
    bool do_refl = false, do_chro = false, do_diff = false, do_refr = false;
    if( live )
    {
        if( me )
        {
            nrel = me->refr; F = me->fresnel01; C = me->chroma; Dff = me->diffuse;
            T = me->transparent != 0; on_a = me->on_a; on_b = me->on_b;
        }
        if( tr.exit_obj >= 0 )                                                           // scene.c:464-470, 656-664
        {
            const DMat<R>& mx = mats[ sv0.link[ tr.exit_obj ].w ];
            nrel /= mx.refr; F = R( 1 ); C = R( 0 ); Dff = R( 0 ); T = true;
            if( a > R( 0 ) )
            {
                tp.x *= r_pow( mx.transp[ 0 ], a );
                tp.y *= r_pow( mx.transp[ 1 ], a );
                tp.z *= r_pow( mx.transp[ 2 ], a );
            }
        }
        if( F > R( 0 ) && I >= prm.min_intensity )                                       // scene.c:473-495
        {
            const R refl = fresnel_reflectance( ray.d, tr.exit_nor, nrel ) * F;
            do_refl = true; I_refl = refl * I; I *= ( R( 1 ) - refl );
        }
        if( C > R( 0 ) && I >= prm.min_intensity )                                       // scene.c:498-523
        {
            do_chro = true; I_chro = C * I; I *= ( R( 1 ) - C );
        }
        if( tr.enter_obj >= 0 && I * Dff >= prm.min_intensity )                          // scene.c:526-630
        {
            do_diff = true; I_diff = I * Dff; I *= ( R( 1 ) - Dff );
        }
        if( T && I >= prm.min_intensity ) { do_refr = true; I_refr = I; }                // scene.c:633-653
    }

    // ---- reserve: ONE atomic per block for the rays of both ends of the stack (A in the low, B in the high 32 bits of one
    // word) and one for the tasks.  The L2 atomic unit serialises operations on one address (about one per clock): with
    // an atomic per warp and 32 hits, k_shade ran at the rate of that one address (4 M atomics, 4 ms per wine_glass
    // pass) whatever else was done to it.  Reflection and chromatic rays go to end A of the stack, refraction rays to end
    // B: a k_rays warp then traces 32 rays of one kind (reflections leave the solid they were born on, refractions cross
    // it: different envelope gates, different numbers of crossings) without a separate partitioning pass over the wave.
    const unsigned int lt = ( 1u << lane ) - 1u;
    const int wid = threadIdx.x >> 5;
    const unsigned int m_refl = __ballot_sync( ACN_FULL, do_refl ), m_chro = __ballot_sync( ACN_FULL, do_chro );
    const unsigned int m_refr = __ballot_sync( ACN_FULL, do_refr ), tmask = __ballot_sync( ACN_FULL, do_diff );
    if( lane == 0 )
    {
        sh->n_ab[ wid ] = ( unsigned long long )( __popc( m_refl ) + __popc( m_chro ) ) | ( ( unsigned long long )__popc( m_refr ) << 32 );
        sh->n_t[ wid ] = ( unsigned int )__popc( tmask );
    }
    __syncthreads();
    if( threadIdx.x == 0 )
    {
        unsigned long long tot_ab = 0; unsigned int tot_t = 0;
        #pragma unroll
        for( int k = 0; k < ACN_BLOCK / 32; k++ )
        {
            const unsigned long long v = sh->n_ab[ k ]; const unsigned int t = sh->n_t[ k ];
            sh->n_ab[ k ] = tot_ab; sh->n_t[ k ] = tot_t;            // exclusive prefix over the warps
            tot_ab += v; tot_t += t;
        }
        sh->base_ab = tot_ab ? atomicAdd( &w.sc->out_ab, tot_ab ) : 0ull;
        sh->base_t = tot_t ? atomicAdd( &w.sc->tasks_new, ( unsigned long long )tot_t ) : 0ull;
    }
    __syncthreads();
    const unsigned long long abbase = sh->base_ab + sh->n_ab[ wid ], tbase = sh->base_t + sh->n_t[ wid ];
    __syncthreads();                                                    // the scratch is rewritten by the next group
    const unsigned long long abase = abbase & 0xFFFFFFFFull, bbase = abbase >> 32;
    if( !live ) return;
    // slots: end A counts up from slot 0, end B down from the last slot; the two ends must not meet (checked again, for
    // the whole iteration, by k_sched: a thread only sees its own slots)
    unsigned long long aslot = base_a + abase + __popc( m_refl & lt ) + __popc( m_chro & lt );
    const unsigned long long bidx = base_b + bbase + __popc( m_refr & lt );
    if( ( ( do_refl || do_chro ) && aslot + 2 > w.rays_cap ) || ( do_refr && bidx >= w.rays_cap ) ) { w.sc->overflow = 1; return; }
    const unsigned long long bslot = w.rays_cap - 1ull - bidx;

    // ---- emit
    if( do_refl )
        write_ray( w, aslot++, pos, reflect( ray.d, tr.exit_nor ), I_refl, depth - 1, tp, RC_REFLECT, sample, mix64( key, KEY_REFLECT ) );
    if( do_chro )
    {
        V3<R> col = obj_color( mats, sv0, tr.enter_obj, pos );
        write_ray( w, aslot++, pos, reflect( ray.d, tr.exit_nor ), I_chro, depth - 1, mul( tp, col ), RC_CHROMATIC, sample, mix64( key, KEY_CHROMATIC ) );
    }
    if( do_diff )
    {
        const R Id = I_diff;
        const V3<R> nrm = -tr.exit_nor;
        const R cos_i = dot( ray.d, tr.exit_nor );
        const V3<R> prj = unit( ray.d - nrm * dot( ray.d, nrm ) );
        u64 rv0 = sv0.seed_mode == SEED_POSITION_HASH
                      ? random_seed( pos, ( u64 )3294479285ull ) + random_seed( nrm, ( u64 )3247146734ull )
                      : mix64( key, KEY_DIFFUSE );
        V3<R> col = obj_color( mats, sv0, tr.enter_obj, pos );
        unsigned long long nd = ( unsigned long long )( ( double )prm.direct_samples * ( double )Id );   // scene.c:553
        if( nd == 0 ) nd = 1;
        unsigned long long np = 0;
        if( prm.path_samples && depth > 10 )                                             // scene.c:584,593
        {
            np = ( unsigned long long )( ( double )prm.path_samples * ( double )Id );
            if( np == 0 ) np = 1;
        }
        const unsigned long long slot = tbase + __popc( tmask & lt );
        if( slot < w.tasks_cap )
        {
            V3<R> tpc = mul( tp, col );
            R4<R> q;
            q.x = pos.x; q.y = pos.y; q.z = pos.z; q.w = Id;      w.tasks_out.pos_id[ slot ] = q;
            q.x = nrm.x; q.y = nrm.y; q.z = nrm.z; q.w = cos_i;   w.tasks_out.nrm_ci[ slot ] = q;
            q.x = prj.x; q.y = prj.y; q.z = prj.z; q.w = on_a;    w.tasks_out.prj_a[ slot ] = q;
            q.x = tpc.x; q.y = tpc.y; q.z = tpc.z; q.w = on_b;    w.tasks_out.tpc_b[ slot ] = q;
            I4 m; m.x = sample; m.y = depth; m.z = ( int )nd; m.w = ( int )np;
            w.tasks_out.meta[ slot ] = m;
            w.tasks_out.rv0[ slot ] = rv0;
            w.tasks_out.key[ slot ] = key;
        }
        else w.sc->overflow = 2;
    }
    if( do_refr )
        write_ray( w, bslot, madd( ray.p, ray.d, a + R( 2 ) * hit_eps ), refract( ray.d, tr.exit_nor, nrel ), I_refr, depth - 1, tp,
                   RC_REFRACT, sample, mix64( key, KEY_REFRACT ) );
}

// a hit of a ray of the tree goes to the hit queue: the surface response (scene_s_lum) runs in its own kernel (k_shade)
// over the compacted hits.  eps: the shell thickness the ray was traced with.
template <typename R> __device__ __forceinline__ void emit_hit( const Wave<R>& w, const Ray<R>& ray, R a, R eps, const Trans<R>& tr, R I, int depth, V3<R> tp, int sample, u64 key )
{
    // the hit distance itself is only good to a few ulp of its magnitude: keep the shading point that far in front
    const R hit_eps = r_max( eps, w.prm.eps_rel * a );
    a -= hit_eps - eps;
    const unsigned long long slot = agg_inc( &w.sc->hits );
    if( slot >= w.hits_cap ) { w.sc->overflow = 5; return; }
    R4<R> q;
    q.x = ray.p.x; q.y = ray.p.y; q.z = ray.p.z; q.w = a;                          w.hits_out.o_a[ slot ] = q;
    q.x = ray.d.x; q.y = ray.d.y; q.z = ray.d.z; q.w = I;                          w.hits_out.d_i[ slot ] = q;
    q.x = tr.exit_nor.x; q.y = tr.exit_nor.y; q.z = tr.exit_nor.z; q.w = hit_eps;  w.hits_out.n_e[ slot ] = q;
    q.x = tp.x; q.y = tp.y; q.z = tp.z; q.w = R( 0 );                              w.hits_out.tp[ slot ] = q;
    I4 m; m.x = depth; m.y = sample; m.z = tr.exit_obj; m.w = tr.enter_obj;        w.hits_out.meta[ slot ] = m;
    w.hits_out.key[ slot ] = key;
}

// trace one ray of the tree and shade its hit; returns true when the ray leaves the scene, in which
// case the caller owes the sample  background * tp * I  (scene.c:484-491, 613-616)
//   cls != RC_PATH: scene_s_trans_hit (lights + matter)
//   cls == RC_PATH: matter only; "leaves" = nothing closer than max_path_length        (scene.c:606-616)
//   probe: the hit's shading would return 0 (depth 0 or I < Imin) — only "anything hit?" matters
template <typename R, int MARCH, bool SH> __device__ __forceinline__ bool trace_ray( const Wave<R>& w, const SceneView<R, SH>& sv0, const CsgMem<R>& cm, const Ray<R>& ray, R I, int depth, V3<R> tp, int cls,
                                                                bool probe, int sample, u64 key )
{
    const DParams<R>& prm = w.prm;
    const R inf = Num<R>::inf();
    HitCtx ctx; ctx.key = key;
    const SceneView<R, SH> sv = ray_view( prm, sv0, ray.p );
    const bool path = cls == RC_PATH;
    // probe: "anything at all?" (path children: "anything closer than max_path_length?")
    const int flags = ( path ? Q_MATTER : ( Q_LIGHT | Q_MATTER ) ) | ( probe ? 0 : Q_TRANS );
    const R t_lim = path ? prm.max_path_length : inf;
    Trans<R> tr;
    tr.exit_obj = tr.enter_obj = -1; tr.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    R a = query<R, MARCH>( sv, ray, flags, t_lim, &tr, ctx, cm );     // nothing at or beyond t_lim matters
    if( !( a < t_lim ) ) return true;
    if( probe ) return false;
    emit_hit( w, ray, a, sv.eps, tr, I, depth, tp, sample, key );
    return false;
}

// Warp-level ray compaction.  Most rays of a wave touch only planes and spheres; a minority enters the
// envelope of an expensive object (a CSG solid, a distance field) and then costs 10-50x more.  Traced
// in the order they come, the expensive rays run a few lanes at a time while the rest of the warp
// idles.  Each persistent warp therefore DEFERS the rays that pass the envelope of an expensive object
// into a small ring in shared memory and traces them only when a full group of 32 has collected (or the
// input is exhausted): cheap groups run without them, expensive groups run dense.
#define ACN_PEND 64
template <typename R, bool SH> __device__ __forceinline__ bool ray_is_heavy( const DParams<R>& prm, const SceneView<R, SH>& sv0, const Ray<R>& ray )
{
    bool h = false;
    for( int k = 0; k < prm.n_heavy; k++ ) h = h || envelope_hits( sv0.env[ prm.heavy[ k ] ], ray );
    return h;
}

// appends the items of the lanes with `take` to the warp's ring (order of lanes preserved)
__device__ __forceinline__ void pend_push( unsigned long long* ring, int head, int& count, bool take, unsigned long long item, int lane )
{
    const unsigned int mask = __ballot_sync( ACN_FULL, take );
    if( take ) ring[ ( head + count + __popc( mask & ( ( 1u << lane ) - 1u ) ) ) & ( ACN_PEND - 1 ) ] = item;
    count += __popc( mask );
    __syncwarp();
}

// obj_ray_hit of a light for a direct sample (scene.c:564): spheres in line, any other shape out of line
template <typename R, int MARCH, bool SH> __device__ __noinline__ R light_hit_cold( SceneView<R, SH> sv, int node, Ray<R> ray, HitCtx ctx, CsgMem<R> cm )
{
    const I4 lk = sv.link[ node ];
    if( ( node_flags( lk ) & F_ENV ) && !envelope_hits( sv.env[ node ], ray ) ) return Num<R>::inf();
    return elem_hit<R, MARCH>( sv, lk, node, ray, ( V3<R>* )nullptr, ctx, cm, Num<R>::inf() );
}

template <typename R, int MARCH, bool SH> __device__ __forceinline__ R light_hit( const SceneView<R, SH>& sv, int node, const Ray<R>& ray, HitCtx ctx, const CsgMem<R>& cm )
{
    const I4 lk = sv.link[ node ];
    if( node_kind( lk ) == K_SPHERE )
    {
        if( ( node_flags( lk ) & F_ENV ) && !envelope_hits( sv.env[ node ], ray ) ) return Num<R>::inf();
        const R4<R> g0 = sv.geo[ node * GEO_STRIDE ];
        return sphere_hit<R>( xyz( g0 ), g0.w, ray, sv.eps, nullptr );
    }
    return light_hit_cold<R, MARCH>( sv, node, ray, ctx, cm );
}

// ---------------------------------------------------------------------------------------------
// kernels — all persistent: fixed grid, warps fetch chunks of work through a cursor in Sched
// ---------------------------------------------------------------------------------------------
#ifndef ACN_CHUNK
#define ACN_CHUNK 2        // 32-item groups per cursor fetch when a launch has plenty of work
#endif
#ifndef ACN_CHUNK_MIN_GROUPS
#define ACN_CHUNK_MIN_GROUPS 16
#endif
// Groups per cursor fetch of this launch.  Big launches (wine_glass: ~30 groups per warp) amortise the cursor atomic
// over ACN_CHUNK groups; small ones (diamond: ~6 per warp) fetch single groups, because the last chunk a warp takes is
// the tail the rest of the machine waits for (diamond 15.2 -> 14.0 ms/step).
__device__ __forceinline__ int launch_chunk( unsigned long long groups )
{
    return groups >= ( unsigned long long )gridDim.x * ( ACN_BLOCK / 32 ) * ACN_CHUNK_MIN_GROUPS ? ACN_CHUNK : 1;
}
// minimum resident blocks per SM the compiler must fit the registers of each tracing kernel into.  Staged scenes
// (tables in shared memory, latency ~30 cycles) run best at 5 blocks x 96 registers; scenes whose tables stay in
// L2 (many_spheres) are latency-bound and want more resident warps at the price of fewer registers.
#ifndef ACN_MINB_RAYS
#define ACN_MINB_RAYS 5
#endif
#ifndef ACN_MINB_PATH
#define ACN_MINB_PATH 5
#endif
#ifndef ACN_MINB_DIRECT
#define ACN_MINB_DIRECT 5
#endif
#ifndef ACN_MINB_RAYS_G
#define ACN_MINB_RAYS_G 6
#endif
#ifndef ACN_MINB_PATH_G
#define ACN_MINB_PATH_G 6
#endif
#ifndef ACN_MINB_DIRECT_G
#define ACN_MINB_DIRECT_G 6     // lamps, after the packed evaluation programs: 8 blocks x 64 registers 1.81 s, 7: 1.67, 6 x 80: 1.64, 5: 1.76
#endif
#ifndef ACN_MINB_DIRECT_R
#define ACN_MINB_DIRECT_R 6     // refill body: 8 x 64 registers kept the walk's ray in local memory (19 LDL/STL per step); 6 x 80: 3, and no slower
#endif

enum { SCHED_PRIMARY = 0, SCHED_WAVE = 1 };

// one thread: closes the books of the previous iteration and plans the next one
__global__ void k_sched( Sched* s, const u64* cum, const unsigned int* pdir, unsigned long long budget,
                         unsigned long long ray_min, unsigned long long ray_cap, int mode, unsigned long long prim_first, unsigned long long prim_count )
{
    if( threadIdx.x != 0 || blockIdx.x != 0 ) return;
    // ---- previous iteration
    const unsigned long long out_a = s->out_ab & 0xFFFFFFFFull, out_b = s->out_ab >> 32;
    const unsigned long long nr_a = s->base_a + out_a, nr_b = s->base_b + out_b, nr = nr_a + nr_b;
    if( nr > ray_cap && !s->overflow ) s->overflow = 1;            // the two ends of the ray stack met
    unsigned long long nt = s->task_stack >> ACN_TASK_SHIFT, nt_cum = s->task_stack & ACN_TASK_MASK;
    s->stats[ ST_DIFFUSE ] += s->tasks_new;
    if( s->out_ab | s->tasks_new | s->take_a | s->take_b | ( s->path_blk_hi - s->path_blk_lo ) | s->prim_count ) s->waves++;
    // dl_packed and cur_direct belong to k_direct, which may still be running on the second stream: k_shade resets them
    s->out_ab = 0; s->tasks_new = 0; s->hits = 0;
    s->cur_rays = s->cur_path = s->cur_index = s->cur_primary = s->cur_shade = 0;
    // ---- plan
    unsigned long long take_a = 0, take_b = 0, blk_lo = 0, blk_hi = 0, fix_slot = ACN_NONE64, fix_cum = 0;
    s->path_nt = nt; s->path_c_hi = nt_cum;
    s->prim_first = prim_first; s->prim_count = 0;
    if( mode == SCHED_PRIMARY )
    {
        s->prim_count = prim_count; s->done = 0;
    }
    else
    {
        const bool path_avail = nt_cum > 0;
        if( nr > 0 && ( nr >= ray_min || !path_avail ) )
        {   // the newest rays of both ends; when they exceed the budget each end gets at least half of it
            take_b = nr_b < budget ? nr_b : budget;
            const unsigned long long room = budget - take_b, half = budget >> 1;
            take_a = nr_a < ( room > half ? room : half ) ? nr_a : ( room > half ? room : half );
            if( take_a + take_b > budget ) take_b = budget - take_a;
        }
        if( path_avail )
        {
            const unsigned long long c_lo = nt_cum > budget ? ( ( nt_cum - budget ) & ~31ull ) : 0ull;
            blk_lo = c_lo >> 5; blk_hi = ( nt_cum + 31 ) >> 5;
            const unsigned long long t0 = pdir[ blk_lo ];                // the task that owns child c_lo
            const unsigned long long excl = t0 ? cum[ t0 - 1 ] : 0ull;
            const bool partial = excl < c_lo;                            // its low children stay on the stack
            if( partial ) { fix_slot = t0; fix_cum = c_lo; }
            nt = t0 + ( partial ? 1 : 0 ); nt_cum = c_lo;
            s->task_stack = ( nt << ACN_TASK_SHIFT ) | nt_cum;
        }
        if( take_a + take_b == 0 && !path_avail ) s->done = 1;
    }
    if( s->overflow ) { s->done = 1; take_a = take_b = 0; blk_lo = blk_hi = 0; fix_slot = ACN_NONE64; s->prim_count = 0; }
    s->nr_a = nr_a - take_a; s->nr_b = nr_b - take_b; s->nt = nt; s->nt_cum = nt_cum;
    s->base_a = nr_a - take_a; s->base_b = nr_b - take_b; s->take_a = take_a; s->take_b = take_b;
    s->path_blk_lo = blk_lo; s->path_blk_hi = blk_hi;
    s->fix_slot = fix_slot; s->fix_cum = fix_cum;
}

// camera rays (scene.c:976-990) fused with their first trace + shade
template <typename R, int MARCH, bool SH> __global__ void __launch_bounds__( ACN_BLOCK )
k_primary( Wave<R> w, const double* __restrict__ xy )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    const unsigned long long first = w.sc->prim_first, count = w.sc->prim_count;
    if( count == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( count >> 5 );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * 32ull * chunk >= count ) return;     // more warps than chunks: no need to stage the scene
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const int lane = threadIdx.x & 31;
    unsigned long long n_rays = 0;
    for( ;; )
    {
        const unsigned long long c0 = warp_fetch( &w.sc->cur_primary, 32ull * chunk, lane );
        if( c0 >= count ) break;
        for( int g = 0; g < chunk; g++ )
        {
            const unsigned long long i = c0 + 32ull * g + lane;
            if( i >= count ) continue;
            const unsigned long long s = first + i;
            const double mx = xy[ 2 * s ], my = xy[ 2 * s + 1 ];
            const int unit_sz = prm.height >> 1;
            const double unit_f = 1.0 / ( double )unit_sz;
            const R z = ( R )( unit_f * ( ( double )unit_sz - my ) );
            const R x = ( R )( unit_f * ( mx - ( double )( prm.width >> 1 ) ) );
            V3<R> d = unit( v3<R>( x, prm.focal, z ) );
            Ray<R> ray;
            ray.p = prm.cam_pos;
            ray.d = prm.cam_rx * d.x + prm.cam_ry * d.y + prm.cam_rz * d.z;
            n_rays++;
            const V3<R> one = v3<R>( R( 1 ), R( 1 ), R( 1 ) );
            if( trace_ray<R, MARCH>( w, sv0, cm, ray, R( 1 ), prm.trace_depth, one, RC_PRIMARY, false, ( int )s, mix64( w.index_base + s, 0x5EEDull ) ) )
                add_sample( w, ( int )s, prm.background );
        }
    }
    warp_count( &w.sc->stats[ ST_PRIMARY ], n_rays, lane );
}

// explicit rays, read in place from the tops of the two ends of the ray stack: items [0, pad_a) are the take_a newest
// rays of end A (reflection / chromatic), padded to a whole number of 32-ray groups so that no warp mixes the kinds,
// items [pad_a, pad_a + take_b) the take_b newest of end B (refraction).  What k_shade spawns afterwards overwrites them.
template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_rays_groups( const Wave<R>& w, const RayBuf<R>& in )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    __shared__ unsigned long long ring_all[ ACN_BLOCK / 32 ][ ACN_PEND ];
    const unsigned long long take_a = w.sc->take_a, take_b = w.sc->take_b, pad_a = ( take_a + 31ull ) & ~31ull;
    const unsigned long long base_a = w.sc->base_a, top_b = w.rays_cap - 1ull - w.sc->base_b;      // slot of B's item j: top_b - j
    const unsigned long long count = pad_a + take_b;
    if( count == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( count >> 5 );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * 32ull * chunk >= count ) return;     // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const int lane = threadIdx.x & 31;
    unsigned long long* ring = ring_all[ threadIdx.x >> 5 ];
    int pend_head = 0, pend_n = 0;
    bool input_done = false;
    unsigned long long c_cur = 0, c_end = 0;
    const bool split = w.prm.n_heavy > 0;
    unsigned int n_refl = 0, n_chro = 0, n_refr = 0;
    for( ;; )
    {
        unsigned long long i = ACN_NONE64;
        if( pend_n >= 32 || ( input_done && pend_n > 0 ) )
        {
            const int k = pend_n < 32 ? pend_n : 32;
            if( lane < k ) i = ring[ ( pend_head + lane ) & ( ACN_PEND - 1 ) ];
            pend_head += k; pend_n -= k;
            __syncwarp();
        }
        else if( !input_done )
        {
            if( c_cur >= c_end )
            {
                c_cur = warp_fetch( &w.sc->cur_rays, 32ull * chunk, lane );
                c_end = c_cur + 32ull * chunk;
                if( c_cur >= count ) { input_done = true; continue; }
            }
            i = c_cur + lane; c_cur += 32;
            if( i < pad_a ) i = i < take_a ? base_a + i : ACN_NONE64;                 // the ring and the trace below hold SLOTS
            else            i = i < count ? top_b - ( i - pad_a ) : ACN_NONE64;
            if( split )
            {
                bool heavy = false;
                if( i != ACN_NONE64 )
                {
                    Ray<R> ray; ray.p = xyz( in.o_i[ i ] ); ray.d = xyz( in.d_[ i ] );
                    heavy = ray_is_heavy( w.prm, sv0, ray );
                }
                pend_push( ring, pend_head, pend_n, heavy, i, lane );
                if( heavy ) i = ACN_NONE64;
            }
        }
        else break;
        if( i != ACN_NONE64 )
        {
            const R4<R> a = in.o_i[ i ], b = in.d_[ i ], c = in.tp[ i ];
            const I4 m = in.meta[ i ];
            Ray<R> ray; ray.p = xyz( a ); ray.d = xyz( b );
            const int depth = m.x & 0xFF, cls = ( m.x >> 8 ) & 0xFF;
            const u64 key = ( u64 )( unsigned )m.z | ( ( u64 )( unsigned )m.w << 32 );
            n_refl += cls == RC_REFLECT; n_chro += cls == RC_CHROMATIC; n_refr += cls == RC_REFRACT;
            if( trace_ray<R, MARCH>( w, sv0, cm, ray, a.w, depth, xyz( c ), cls, ( m.x & RAYF_PROBE ) != 0, m.y, key ) )
                add_sample( w, m.y, mul( w.prm.background, xyz( c ) ) * a.w );
        }
        __syncwarp();
    }
    warp_count( &w.sc->stats[ ST_REFLECT ], n_refl, lane );
    warp_count( &w.sc->stats[ ST_CHROMATIC ], n_chro, lane );
    warp_count( &w.sc->stats[ ST_REFRACT ], n_refr, lane );
}

// ---------------------------------------------------------------------------------------------
// The kernels that WALK the scene (no specialised code for its structure): refill loops.
// A ray's walk through the threaded traversal records takes anything between one record (it misses the first bound)
// and hundreds (many_spheres: a ray grazing five levels of clusters), and the children of a diffuse hit point in
// all directions.  Traced a group of 32 at a time, a warp waited for its longest walk with a third of its lanes
// busy (ncu, many_spheres: 11.7 of the 21.7 lanes that had a ray at all were inside the walk on average; the other
// ten had drawn a direction below the horizon).  Here a warp keeps 32 walks going: a lane whose ray is through takes
// the next LIVE ray at once.  Rays that have to be generated first (shadow and path children: window search,
// sampling, light test — warp-collective work at 32 lanes) go through a ring in shared memory; the walk of a lane is
// one record index (acn_isect.cuh: walk_*_step) and survives the refill in registers.
// Results are added per ray with integer (fixed-point) atomics: the sums do not depend on the order.
// ---------------------------------------------------------------------------------------------
#ifndef ACN_REFILL_MIN
#define ACN_REFILL_MIN 8          // idle lanes that trigger a refill (all 32 at the latest)
#endif

// per-ray shell thickness (ray_view) without copying the scene view
template <typename R> __device__ __forceinline__ R ray_eps( const DParams<R>& prm, R eps0, V3<R> o )
{
    return prm.eps_rel > R( 0 ) ? r_max( eps0, prm.eps_rel * r_max( r_max( r_abs( o.x ), r_abs( o.y ) ), r_abs( o.z ) ) ) : eps0;
}

// lights of a ray of the tree (first pass of scene_s_trans_hit, scene.c:362-382): the few records of the light root, walked
// at once by the lanes that have just been given a ray.  Returns the closest light hit; its transition goes to *tl.
template <typename R, int MARCH, bool SH> __device__ __forceinline__ R walk_lights( const SceneView<R, SH>& sv, const Ray<R>& ray, const bool want_trans, Trans<R>* tl,
                                                                               HitCtx ctx, const CsgMem<R>& cm )
{
    const R inf = Num<R>::inf();
    WalkT<R> s; s.reset();
    int c = walk_root( sv, ray, true, inf );
    while( c >= 0 ) c = walk_step<R, MARCH>( sv, ray, inf, want_trans, c, s, ctx, cm );
    if( !want_trans ) return c == WALK_FOUND ? R( 0 ) : inf;            // a probe only asks "anything?"
    walk_trans_finish( sv, ray, s );
    *tl = s.tl;
    return s.min_a;
}

template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_rays_refill( const Wave<R>& w, const RayBuf<R>& in )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    __shared__ R   s_ln[ 3 ][ ACN_BLOCK ];          // transition of the closest light hit of the lane's ray
    __shared__ int s_lo[ 2 ][ ACN_BLOCK ];
    const unsigned long long take_a = w.sc->take_a, take_b = w.sc->take_b, pad_a = ( take_a + 31ull ) & ~31ull;
    const unsigned long long base_a = w.sc->base_a, top_b = w.rays_cap - 1ull - w.sc->base_b;      // slot of B's item j: top_b - j
    const unsigned long long count = pad_a + take_b;
    if( count == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( count >> 5 );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * 32ull * chunk >= count ) return;     // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const R inf = Num<R>::inf();
    const int lane = threadIdx.x & 31;
    const unsigned int lt = ( 1u << lane ) - 1u;
    bool input_done = false;
    unsigned long long c_cur = 0, c_end = 0;
    unsigned int n_refl = 0, n_chro = 0, n_refr = 0;
    // the lane's walk
    int cur = WALK_END;
    unsigned long long slot = 0;
    Ray<R> ray; ray.p = ray.d = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    R eps = sv0.eps, best_l = inf;
    bool probe = false, pend = false;
    HitCtx ctx; ctx.key = 0;
    WalkT<R> ws; ws.reset();
    for( ;; )
    {
        const unsigned int idle = __ballot_sync( ACN_FULL, cur < 0 );
        if( idle == ACN_FULL || ( !input_done && __popc( idle ) >= ACN_REFILL_MIN ) )
        {
            if( pend )
            {   // ---- the rays that came to their end since the last refill, together: hit record or background
                pend = false;
                SceneView<R, SH> sv = sv0; sv.eps = eps;
                const R4<R> a = in.o_i[ slot ], c = in.tp[ slot ];
                const I4 m = in.meta[ slot ];
                if( !probe ) walk_trans_finish( sv, ray, ws );
                if( !probe && ws.min_a < best_l ) emit_hit( w, ray, ws.min_a, eps, ws.tl, a.w, m.x & 0xFF, xyz( c ), m.y, ctx.key );
                else if( !probe && best_l < inf )
                {
                    Trans<R> tl;
                    tl.exit_nor = v3<R>( s_ln[ 0 ][ threadIdx.x ], s_ln[ 1 ][ threadIdx.x ], s_ln[ 2 ][ threadIdx.x ] );
                    tl.exit_obj = s_lo[ 0 ][ threadIdx.x ]; tl.enter_obj = s_lo[ 1 ][ threadIdx.x ];
                    emit_hit( w, ray, best_l, eps, tl, a.w, m.x & 0xFF, xyz( c ), m.y, ctx.key );
                }
                else add_sample( w, m.y, mul( prm.background, xyz( c ) ) * a.w );          // scene.c:484-491: the ray leaves the scene
            }
            if( c_cur >= c_end && !input_done )
            {
                c_cur = warp_fetch( &w.sc->cur_rays, 32ull * chunk, lane );
                c_end = c_cur + 32ull * chunk;
                if( c_cur >= count ) input_done = true;
            }
            if( input_done ) { if( idle == ACN_FULL ) break; }
            else
            {   // the r-th idle lane takes item c_cur + r
                const int n_idle = __popc( idle );
                const unsigned long long avail = c_end - c_cur;
                const int n_take = ( unsigned long long )n_idle < avail ? n_idle : ( int )avail;
                const int r = __popc( idle & lt );
                if( cur < 0 && r < n_take )
                {
                    unsigned long long i = c_cur + r;
                    if( i < pad_a ) i = i < take_a ? base_a + i : ACN_NONE64;
                    else            i = i < count ? top_b - ( i - pad_a ) : ACN_NONE64;
                    if( i != ACN_NONE64 )
                    {
                        const R4<R> a = in.o_i[ i ], b = in.d_[ i ];
                        const I4 m = in.meta[ i ];
                        slot = i;
                        ray.p = xyz( a ); ray.d = xyz( b );
                        const int cls = ( m.x >> 8 ) & 0xFF;
                        n_refl += cls == RC_REFLECT; n_chro += cls == RC_CHROMATIC; n_refr += cls == RC_REFRACT;
                        probe = ( m.x & RAYF_PROBE ) != 0;
                        ctx.key = ( u64 )( unsigned )m.z | ( ( u64 )( unsigned )m.w << 32 );
                        eps = ray_eps( prm, sv0.eps, ray.p );
                        SceneView<R, SH> sv = sv0; sv.eps = eps;
                        Trans<R> tl; tl.exit_obj = tl.enter_obj = -1; tl.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
                        best_l = walk_lights<R, MARCH>( sv, ray, !probe, &tl, ctx, cm );
                        if( !probe && best_l < inf )
                        {
                            s_ln[ 0 ][ threadIdx.x ] = tl.exit_nor.x; s_ln[ 1 ][ threadIdx.x ] = tl.exit_nor.y; s_ln[ 2 ][ threadIdx.x ] = tl.exit_nor.z;
                            s_lo[ 0 ][ threadIdx.x ] = tl.exit_obj; s_lo[ 1 ][ threadIdx.x ] = tl.enter_obj;
                        }
                        ws.reset();
                        cur = ( probe && best_l < inf ) ? WALK_FOUND : walk_root( sv, ray, false, best_l );
                        if( cur == WALK_END )
                        {   // no matter in the way: a light hit, or the ray leaves the scene
                            if( !probe && best_l < inf ) emit_hit( w, ray, best_l, eps, tl, a.w, m.x & 0xFF, xyz( in.tp[ i ] ), m.y, ctx.key );
                            else add_sample( w, m.y, mul( prm.background, xyz( in.tp[ i ] ) ) * a.w );
                        }
                    }
                }
                c_cur += n_take;
            }
        }
        if( cur >= 0 )
        {
            SceneView<R, SH> sv = sv0; sv.eps = eps;
            const int nx = walk_step<R, MARCH>( sv, ray, best_l, !probe, cur, ws, ctx, cm );
            // what a finished ray owes (a hit record, the background) is a hundred instructions of its own: done one ray at
            // a time they cost more issue slots than the whole walk.  The lane waits with it for the next refill, where the
            // lanes that finished since the last one do it together.
            pend = nx == WALK_END;
            cur = nx;
        }
    }
    warp_count( &w.sc->stats[ ST_REFLECT ], n_refl, lane );
    warp_count( &w.sc->stats[ ST_CHROMATIC ], n_chro, lane );
    warp_count( &w.sc->stats[ ST_REFRACT ], n_refr, lane );
}
template <typename R, int MARCH, bool SH> __global__ void __launch_bounds__( ACN_BLOCK, sizeof( R ) == 8 ? 4 : SH ? ACN_MINB_RAYS : ACN_MINB_RAYS_G )
k_rays( Wave<R> w, RayBuf<R> in )
{
#if defined(ACN_SPEC_SCENE)
    k_rays_groups<R, MARCH, SH>( w, in );
#else
    // scenes of planes, spheres and quadrics walk with refill; where composite objects are met the rays of a group stay together
    // (they reach the same object in the same iteration and run ONE event sweep program side by side)
    if constexpr( MARCH == 2 ) k_rays_refill<R, MARCH, SH>( w, in );
    else k_rays_groups<R, MARCH, SH>( w, in );
#endif
}

// scene_s_lum (scene.c:420-667) over the hits of the iteration: emits child rays and diffuse tasks
template <typename R, bool SH> __global__ void __launch_bounds__( ACN_BLOCK )
k_shade( Wave<R> w, HitBuf<R> in )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    // the previous iteration's k_direct is through (the host orders k_shade behind it), k_index of this iteration
    // rebuilds the direct list after this kernel: the one place where its counters can be reset
    if( blockIdx.x == 0 && threadIdx.x == 0 ) { w.sc->dl_packed = 0; w.sc->cur_direct = 0; }
    unsigned long long count = w.sc->hits;
    if( count > w.hits_cap ) count = w.hits_cap;
    if( count == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( count >> 5 );
    if( ( unsigned long long )blockIdx.x * ACN_BLOCK * chunk >= count ) return;                      // more blocks than chunks
    __shared__ ShadeShared sh;
    __shared__ DMat<R> s_mats[ ACN_SHADE_MATS ];
    // the materials next to the staged node table: a hit's surface response starts with two material records
    const DMat<R>* mats = w.prm.mats;
    if( w.prm.n_materials <= ACN_SHADE_MATS )
    {
        for( int i = threadIdx.x; i < w.prm.n_materials; i += blockDim.x ) s_mats[ i ] = w.prm.mats[ i ];
        mats = s_mats;
    }
    const unsigned long long base_a = w.sc->base_a, base_b = w.sc->base_b;      // fixed for the iteration
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // the BLOCK fetches ACN_BLOCK * chunk hits per cursor atomic and reserves the queue slots of ACN_BLOCK hits per atomic
    for( ;; )
    {
        if( threadIdx.x == 0 ) sh.c0 = atomicAdd( &w.sc->cur_shade, ( unsigned long long )ACN_BLOCK * chunk );
        __syncthreads();
        const unsigned long long c0 = sh.c0;
        if( c0 >= count ) break;
        for( int g = 0; g < chunk; g++ )
        {
            const unsigned long long i = c0 + ( unsigned long long )ACN_BLOCK * g + threadIdx.x;
            if( c0 + ( unsigned long long )ACN_BLOCK * g >= count ) break;                           // block-uniform
            const bool live = i < count;
            const unsigned long long il = live ? i : count - 1;          // dead threads of the last group read a valid record and ignore it
            const R4<R> oa = in.o_a[ il ], di = in.d_i[ il ], ne = in.n_e[ il ], tp = in.tp[ il ];
            const I4 m = in.meta[ il ];
            Ray<R> ray; ray.p = xyz( oa ); ray.d = xyz( di );
            Trans<R> tr; tr.exit_nor = xyz( ne ); tr.exit_obj = m.z; tr.enter_obj = m.w;
            shade_hits( w, sv0, live, ray, oa.w, ne.w, tr, m.x, di.w, xyz( tp ), m.y, in.key[ il ], lane, &sh, mats, base_a, base_b );
        }
    }
}

// Lists of work entries with implicit children.  An entry owns the children [ excl, incl ) of a global
// child index space; dir[ b ] names the entry that owns child 32*b, so a warp that takes block b finds
// the owners of its 32 children inside a window of 32 consecutive entries (every entry has >= 1 child).
struct ListWindow { unsigned long long incl, excl0; unsigned int e0; };

__device__ __forceinline__ ListWindow list_window( const u64* __restrict__ cum, const unsigned int* __restrict__ dir,
                                                   unsigned long long blk, unsigned long long n_entries, int lane )
{
    ListWindow lw;
    lw.e0 = dir[ blk ];
    const unsigned long long e = ( unsigned long long )lw.e0 + lane;
    lw.incl = e < n_entries ? cum[ e ] : ACN_NONE64;
    unsigned long long x = 0;
    if( lane == 0 && lw.e0 > 0 ) x = cum[ lw.e0 - 1 ];
    lw.excl0 = __shfl_sync( ACN_FULL, x, 0 );
    return lw;
}

// direct lighting (scene.c:542-581): one lane per (task, light, sample); the shadow rays exist only
// as (entry, child index) and are regenerated from the task with an O(1) LCG skip-ahead
template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_direct_groups( const Wave<R>& w, const TaskBuf<R>& in, const u64* __restrict__ dl_cum, const unsigned int* __restrict__ dl_slot,
          const unsigned int* __restrict__ dl_dir )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    const unsigned long long n_entries = w.sc->dl_packed >> ACN_TASK_SHIFT, total = w.sc->dl_packed & ACN_TASK_MASK;
    if( total == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( total >> 5 );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * 32ull * chunk >= total ) return;     // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_blocks = ( total + 31 ) >> 5;
    unsigned long long n_shadow = 0;
    for( ;; )
    {
        const unsigned long long b0 = warp_fetch( &w.sc->cur_direct, chunk, lane );
        if( b0 >= n_blocks ) break;
        for( unsigned long long blk = b0; blk < b0 + chunk && blk < n_blocks; blk++ )
        {
            const ListWindow lw = list_window( dl_cum, dl_dir, blk, n_entries, lane );
            const unsigned long long idx = ( blk << 5 ) + lane;
            const bool live = idx < total;
            const int j = window_find( lw.incl, live ? idx : ( blk << 5 ) );
            const unsigned long long prev = __shfl_sync( ACN_FULL, lw.incl, ( j + 31 ) & 31 );
            AccV<R> sum = acc_zero<R>();
            int sample = -1;
            if( live )
            {
                const unsigned long long r = idx - ( j ? prev : lw.excl0 );
                const unsigned int t = dl_slot[ lw.e0 + j ];
                const I4 m = in.meta[ t ];
                sample = m.x;
                const unsigned int nd = ( unsigned int )m.z;
                const unsigned int li = ( unsigned int )( r / nd ), jj = ( unsigned int )( r - ( unsigned long long )li * nd );
                const R4<R> pi = in.pos_id[ t ], nc = in.nrm_ci[ t ], pa = in.prj_a[ t ], tb = in.tpc_b[ t ];
                const V3<R> pos = xyz( pi ), nrm = xyz( nc ), prj = xyz( pa );
                const DLight<R>& lg = prm.lights[ li ];

                const SceneView<R, SH> sv = ray_view( prm, sv0, pos );
                V3<R> axis; R cos_rs;
                obj_fov( sv, lg.node, pos, &axis, &cos_rs );
                const Basis<R> bs = basis_con_z( axis );
                const R h = R( 1 ) - cos_rs;                                                 // areal_coverage, vectors.h:362
                u64 rv = skip2( prm, in.rv0[ t ], ( unsigned long long )li * nd + jj );
                Ray<R> out; out.p = pos;
                out.d = from_basis( bs, sphere_cap<R>( &rv, h ) );
                R wgt = dot( out.d, nrm );
                if( wgt > R( 0 ) )
                {
                    HitCtx ctx; ctx.key = 0;
                    n_shadow++;
                    R a = light_hit<R, MARCH>( sv, lg.node, out, ctx, cm );                                // scene.c:564
                    if( a < Num<R>::inf() )
                    {
                        if( tb.w > R( 0 ) ) wgt = oren_nayar( wgt, nc.w, pa.w, tb.w, out.d, nrm, prj );
                        n_shadow++;
                        R sh = query<R, MARCH>( sv, out, Q_MATTER, a, nullptr, ctx, cm );          // scene.c:569
                        if( sh > a )
                        {
                            V3<R> hp = madd( out.p, out.d, a );
                            R d2 = sqr( hp - v3<R>( lg.pos[ 0 ], lg.pos[ 1 ], lg.pos[ 2 ] ) );
                            R lint = d2 > R( 0 ) ? lg.radiance / d2 : Num<R>::mag();
                            R f = lint * wgt * pi.w * ( R( 2 ) * h / ( R )nd );               // scene.c:574,579
                            sum = acc_of( v3<R>( lg.color[ 0 ] * f * tb.x, lg.color[ 1 ] * f * tb.y, lg.color[ 2 ] * f * tb.z ) );
                        }
                    }
                }
            }
            __syncwarp();
            // one atomic triple per task segment of the block instead of one per shadow ray
            const int key = live ? j : -1;
            if( __any_sync( ACN_FULL, acc_any( sum ) ) )
            {
                sum = seg_sum_acc( sum, key, lane );
                const int kprev = __shfl_up_sync( ACN_FULL, key, 1 );
                if( live && ( lane == 0 || kprev != key ) && acc_any( sum ) ) add_sample_acc( w, sample, sum );
            }
        }
    }
    warp_count( &w.sc->stats[ ST_SHADOW ], n_shadow, lane );
}

// refill loop (see k_rays): blocks of 32 children are GENERATED by the whole warp (owner search, cone sampling, light test,
// contribution) and the live ones — direction above the horizon, light met — queue in a ring; idle lanes take them
// from there and walk the matter root for an occluder.  An unoccluded ray adds its contribution at once.
#define ACN_RING 64
template <typename R> struct ShadowRing { R ox[ ACN_RING ], oy[ ACN_RING ], oz[ ACN_RING ], dx[ ACN_RING ], dy[ ACN_RING ], dz[ ACN_RING ], tf[ ACN_RING ], cr[ ACN_RING ], cg[ ACN_RING ], cb[ ACN_RING ]; int smp[ ACN_RING ]; };

template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_direct_refill( const Wave<R>& w, const TaskBuf<R>& in, const u64* __restrict__ dl_cum, const unsigned int* __restrict__ dl_slot,
          const unsigned int* __restrict__ dl_dir )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    __shared__ ShadowRing<R> rings[ ACN_BLOCK / 32 ];
    __shared__ R   s_con[ 3 ][ ACN_BLOCK ];         // what the lane's ray adds to its sample when it reaches the light: kept out of the
    __shared__ int s_smp[ ACN_BLOCK ];              // registers of the walk (written when the ray is taken, read when it is through)
    const unsigned long long n_entries = w.sc->dl_packed >> ACN_TASK_SHIFT, total = w.sc->dl_packed & ACN_TASK_MASK;
    if( total == 0 || w.sc->overflow ) return;
    const int chunk = launch_chunk( total >> 5 );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * 32ull * chunk >= total ) return;     // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const int lane = threadIdx.x & 31;
    const unsigned int lt = ( 1u << lane ) - 1u;
    const unsigned long long n_blocks = ( total + 31 ) >> 5;
    ShadowRing<R>& rg = rings[ threadIdx.x >> 5 ];
    int head = 0, ring_n = 0;
    bool input_done = false;
    unsigned int b_cur = 0, b_end = 0;          // blocks of the list: < 2^32 (the list holds < 2^38 children)
    unsigned int n_shadow = 0;
    HitCtx ctx; ctx.key = 0;
    // the lane's walk
    int cur = WALK_END;
    bool pend = false;
    Ray<R> ray; ray.p = ray.d = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    R eps = sv0.eps, tfar = R( 0 );
    for( ;; )
    {
        const unsigned int idle = __ballot_sync( ACN_FULL, cur < 0 );
        if( idle == ACN_FULL || ( __popc( idle ) >= ACN_REFILL_MIN && !( input_done && ring_n == 0 ) ) )
        {
            const int n_idle = __popc( idle );
            if( pend )
            {   // the unoccluded rays since the last refill, together (see k_rays)
                pend = false;
                add_sample( w, s_smp[ threadIdx.x ], v3<R>( s_con[ 0 ][ threadIdx.x ], s_con[ 1 ][ threadIdx.x ], s_con[ 2 ][ threadIdx.x ] ) );
            }
            while( !input_done && ring_n < n_idle )
            {   // ---- generate the next block of 32 children
                if( b_cur >= b_end )
                {
                    const unsigned long long b0 = warp_fetch( &w.sc->cur_direct, chunk, lane );
                    if( b0 >= n_blocks ) { input_done = true; break; }
                    b_cur = ( unsigned int )b0;
                    b_end = b0 + chunk < n_blocks ? ( unsigned int )( b0 + chunk ) : ( unsigned int )n_blocks;
                }
                const unsigned long long blk = b_cur++;
                const ListWindow lw = list_window( dl_cum, dl_dir, blk, n_entries, lane );
                const unsigned long long idx = ( blk << 5 ) + lane;
                const bool live = idx < total;
                const int j = window_find( lw.incl, live ? idx : ( blk << 5 ) );
                const unsigned long long prev = __shfl_sync( ACN_FULL, lw.incl, ( j + 31 ) & 31 );
                bool want = false;
                Ray<R> out; out.p = out.d = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
                R a = R( 0 ); V3<R> c3 = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); int sample = 0;
                if( live )
                {
                    const unsigned long long r = idx - ( j ? prev : lw.excl0 );
                    const unsigned int t = dl_slot[ lw.e0 + j ];
                    const I4 m = in.meta[ t ];
                    sample = m.x;
                    const unsigned int nd = ( unsigned int )m.z;
                    const unsigned int li = ( unsigned int )( r / nd ), jj = ( unsigned int )( r - ( unsigned long long )li * nd );
                    const R4<R> pi = in.pos_id[ t ], nc = in.nrm_ci[ t ], pa = in.prj_a[ t ], tb = in.tpc_b[ t ];
                    const V3<R> pos = xyz( pi ), nrm = xyz( nc ), prj = xyz( pa );
                    const DLight<R>& lg = prm.lights[ li ];
                    const SceneView<R, SH> sv = ray_view( prm, sv0, pos );
                    V3<R> axis; R cos_rs;
                    obj_fov( sv, lg.node, pos, &axis, &cos_rs );
                    const Basis<R> bs = basis_con_z( axis );
                    const R h = R( 1 ) - cos_rs;                                                 // areal_coverage, vectors.h:362
                    u64 rv = skip2( prm, in.rv0[ t ], ( unsigned long long )li * nd + jj );
                    out.p = pos;
                    out.d = from_basis( bs, sphere_cap<R>( &rv, h ) );
                    R wgt = dot( out.d, nrm );
                    if( wgt > R( 0 ) )
                    {
                        n_shadow++;
                        a = light_hit<R, MARCH>( sv, lg.node, out, ctx, cm );                  // scene.c:564
                        if( a < Num<R>::inf() )
                        {
                            if( tb.w > R( 0 ) ) wgt = oren_nayar( wgt, nc.w, pa.w, tb.w, out.d, nrm, prj );
                            n_shadow++;
                            const V3<R> hp = madd( out.p, out.d, a );
                            const R d2 = sqr( hp - v3<R>( lg.pos[ 0 ], lg.pos[ 1 ], lg.pos[ 2 ] ) );
                            const R lint = d2 > R( 0 ) ? lg.radiance / d2 : Num<R>::mag();
                            const R f = lint * wgt * pi.w * ( R( 2 ) * h / ( R )nd );           // scene.c:574,579
                            c3 = v3<R>( lg.color[ 0 ] * f * tb.x, lg.color[ 1 ] * f * tb.y, lg.color[ 2 ] * f * tb.z );
                            want = true;
                        }
                    }
                }
                const unsigned int wm = __ballot_sync( ACN_FULL, want );
                if( want )
                {
                    const int q = ( head + ring_n + __popc( wm & lt ) ) & ( ACN_RING - 1 );
                    rg.ox[ q ] = out.p.x; rg.oy[ q ] = out.p.y; rg.oz[ q ] = out.p.z;
                    rg.dx[ q ] = out.d.x; rg.dy[ q ] = out.d.y; rg.dz[ q ] = out.d.z;
                    rg.tf[ q ] = a; rg.cr[ q ] = c3.x; rg.cg[ q ] = c3.y; rg.cb[ q ] = c3.z; rg.smp[ q ] = sample;
                }
                ring_n += __popc( wm );
                __syncwarp();
            }
            // ---- hand out: the r-th idle lane takes the r-th ray of the ring
            const int n_take = n_idle < ring_n ? n_idle : ring_n;
            if( n_take > 0 )
            {
                const int r = __popc( idle & lt );
                if( cur < 0 && r < n_take )
                {
                    const int q = ( head + r ) & ( ACN_RING - 1 );
                    ray.p = v3<R>( rg.ox[ q ], rg.oy[ q ], rg.oz[ q ] ); ray.d = v3<R>( rg.dx[ q ], rg.dy[ q ], rg.dz[ q ] );
                    tfar = rg.tf[ q ];
                    s_con[ 0 ][ threadIdx.x ] = rg.cr[ q ]; s_con[ 1 ][ threadIdx.x ] = rg.cg[ q ]; s_con[ 2 ][ threadIdx.x ] = rg.cb[ q ]; s_smp[ threadIdx.x ] = rg.smp[ q ];
                    eps = ray_eps( prm, sv0.eps, ray.p );
                    SceneView<R, SH> sv = sv0; sv.eps = eps;
                    cur = walk_root( sv, ray, false, tfar, false );
                    pend = cur == WALK_END;                  // no matter at all: settled at the next refill
                }
                head = ( head + n_take ) & ( ACN_RING - 1 ); ring_n -= n_take;
                __syncwarp();
            }
            else if( idle == ACN_FULL ) break;           // nothing left to generate, nothing queued, nothing in flight
        }
        if( cur >= 0 )
        {   // ---- one record of the shadow test (scene.c:569: the light is seen when nothing lies at or before it)
            SceneView<R, SH> sv = sv0; sv.eps = eps;
            const int nx = walk_any_step<R, MARCH, SH, false>( sv, ray, tfar, cur, ctx, cm );
            pend = nx == WALK_END;
            cur = nx;
        }
    }
    warp_count( &w.sc->stats[ ST_SHADOW ], ( unsigned long long )n_shadow, lane );
}
template <typename R, int MARCH, bool SH> __global__ void __launch_bounds__( ACN_BLOCK, sizeof( R ) == 8 ? 4 : SH ? ACN_MINB_DIRECT : MARCH == 2 ? ACN_MINB_DIRECT_R : ACN_MINB_DIRECT_G )
k_direct( Wave<R> w, TaskBuf<R> in, const u64* __restrict__ dl_cum, const unsigned int* __restrict__ dl_slot,
          const unsigned int* __restrict__ dl_dir )
{
#if defined(ACN_SPEC_SCENE)
    k_direct_groups<R, MARCH, SH>( w, in, dl_cum, dl_slot, dl_dir );
#else
    // scenes of planes, spheres and quadrics walk with refill; where composite objects are met the rays of a group stay together
    // (they reach the same object in the same iteration and run ONE event sweep program side by side)
    if constexpr( MARCH == 2 ) k_direct_refill<R, MARCH, SH>( w, in, dl_cum, dl_slot, dl_dir );
    else k_direct_groups<R, MARCH, SH>( w, in, dl_cum, dl_slot, dl_dir );
#endif
}

// indirect rays (scene.c:584-621): one lane per (task, path sample); the child ray is generated,
// traced and shaded in place, never stored.
template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_path_groups( const Wave<R>& w, const TaskBuf<R>& in, const unsigned int* __restrict__ pdir )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    __shared__ unsigned long long ring_all[ ACN_BLOCK / 32 ][ ACN_PEND ];
    const unsigned long long blk_lo = w.sc->path_blk_lo, blk_hi = w.sc->path_blk_hi;
    if( blk_hi <= blk_lo || w.sc->overflow ) return;
    const unsigned long long c_hi = w.sc->path_c_hi, n_entries = w.sc->path_nt;
    const int chunk = launch_chunk( blk_hi - blk_lo );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * chunk >= blk_hi - blk_lo ) return;   // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const int lane = threadIdx.x & 31;
    const int L = prm.n_lights;
    const unsigned long long n_blocks = blk_hi - blk_lo;
    unsigned long long* ring = ring_all[ threadIdx.x >> 5 ];
    int pend_head = 0, pend_n = 0;
    bool input_done = false;
    unsigned long long b_cur = 0, b_end = 0;
    const bool split = prm.n_heavy > 0;
    unsigned long long n_path = 0;
    for( ;; )
    {
        // one item per lane: ( task entry t, child i ) packed as t << 32 | i
        unsigned long long item = ACN_NONE64;
        bool fresh = false;
        if( pend_n >= 32 || ( input_done && pend_n > 0 ) )
        {
            const int k = pend_n < 32 ? pend_n : 32;
            if( lane < k ) item = ring[ ( pend_head + lane ) & ( ACN_PEND - 1 ) ];
            pend_head += k; pend_n -= k;
            __syncwarp();
        }
        else if( !input_done )
        {
            if( b_cur >= b_end )
            {
                b_cur = warp_fetch( &w.sc->cur_path, chunk, lane );
                b_end = b_cur + chunk < n_blocks ? b_cur + chunk : n_blocks;
                if( b_cur >= n_blocks ) { input_done = true; continue; }
            }
            const unsigned long long blk = blk_lo + b_cur; b_cur++;
            const ListWindow lw = list_window( in.cum, pdir, blk, n_entries, lane );
            const unsigned long long idx = ( blk << 5 ) + lane;
            const bool live = idx < c_hi;
            const int j = window_find( lw.incl, live ? idx : ( blk << 5 ) );
            const unsigned long long prev = __shfl_sync( ACN_FULL, lw.incl, ( j + 31 ) & 31 );
            if( live ) item = ( ( unsigned long long )( lw.e0 + j ) << 32 ) | ( unsigned int )( idx - ( j ? prev : lw.excl0 ) );
            fresh = true;
        }
        else break;

        R miss = R( 0 );
        V3<R> tpm = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
        int sample = -1, key = -1;
        bool defer = false;
        if( item != ACN_NONE64 )
        {
            const unsigned long long t = item >> 32;
            const unsigned int i = ( unsigned int )item;
            const I4 m = in.meta[ t ];
            const R4<R> pi = in.pos_id[ t ], nc = in.nrm_ci[ t ], pa = in.prj_a[ t ], tb = in.tpc_b[ t ];
            const V3<R> nrm = xyz( nc );
            const Basis<R> bs = basis_con_z( nrm );
            u64 rv = skip2( prm, in.rv0[ t ], ( unsigned long long )L * ( unsigned int )m.z + i );
            Ray<R> out; out.p = xyz( pi );
            out.d = from_basis( bs, sphere_cap<R>( &rv, R( 1 ) ) );
            R wgt = dot( out.d, nrm );
            // every lane of a task carries the task's key, sample and throughput: the lane that ends up
            // adding the run's sum may itself be a deferred or back-facing child
            sample = m.x; key = ( int )( t & 0x7FFFFFFFull );
            tpm = xyz( tb ) * ( R( 2 ) / ( R )( unsigned int )m.w );                         // scene.c:620
            if( wgt > R( 0 ) )                                                           // scene.c:600
            {
                if( fresh && split && ray_is_heavy( prm, sv0, out ) ) defer = true;
                else
                {
                    if( tb.w > R( 0 ) ) wgt = oren_nayar( wgt, nc.w, pa.w, tb.w, out.d, nrm, xyz( pa ) );
                    n_path++;
                    const R ci = wgt * pi.w;
                    const bool probe = ( m.y - 10 ) == 0 || ci < prm.min_intensity;
                    if( trace_ray<R, MARCH>( w, sv0, cm, out, ci, m.y - 10, tpm, RC_PATH, probe, m.x, mix64( in.key[ t ], KEY_PATH0 + i ) ) ) miss = ci;
                }
            }
        }
        __syncwarp();
        if( fresh && split ) pend_push( ring, pend_head, pend_n, defer, item, lane );
        // children that left the scene: background * throughput * sum of their intensities, one atomic
        // triple per run of lanes of the same task (items are in non-decreasing task order in both kinds of group)
        const unsigned int has = __ballot_sync( ACN_FULL, miss != R( 0 ) );
        if( has )
        {
            AccV<R> ms = miss != R( 0 ) ? acc_of( mul( prm.background, tpm ) * miss ) : acc_zero<R>();
            {
                ms = seg_sum_acc( ms, key, lane );
                const int kprev = __shfl_up_sync( ACN_FULL, key, 1 );
                if( key >= 0 && ( lane == 0 || kprev != key ) && acc_any( ms ) ) add_sample_acc( w, sample, ms );
            }
        }
    }
    warp_count( &w.sc->stats[ ST_PATH ], n_path, lane );
}

// refill loop (see k_rays, k_direct): generated children above the horizon queue in a ring as ( ray, intensity, task, child );
// what a finished walk needs beyond that (throughput, sample, depth, key of the task) is read again from the task.
template <typename R> struct PathRing { R ox[ ACN_RING ], oy[ ACN_RING ], oz[ ACN_RING ], dx[ ACN_RING ], dy[ ACN_RING ], dz[ ACN_RING ], ci[ ACN_RING ]; unsigned int t[ ACN_RING ], i[ ACN_RING ]; };

template <typename R, int MARCH, bool SH> __device__ __forceinline__ void k_path_refill( const Wave<R>& w, const TaskBuf<R>& in, const unsigned int* __restrict__ pdir )
{
    extern __shared__ __align__( 32 ) unsigned char smem[];
    __shared__ PathRing<R> rings[ ACN_BLOCK / 32 ];
    const unsigned long long blk_lo = w.sc->path_blk_lo, blk_hi = w.sc->path_blk_hi;
    if( blk_hi <= blk_lo || w.sc->overflow ) return;
    const unsigned long long c_hi = w.sc->path_c_hi, n_entries = w.sc->path_nt;
    const int chunk = launch_chunk( blk_hi - blk_lo );
    if( ( unsigned long long )blockIdx.x * ( ACN_BLOCK / 32 ) * chunk >= blk_hi - blk_lo ) return;   // more warps than chunks
    const SceneView<R, SH> sv0 = stage_scene<R, SH>( w.prm, smem );
    const CsgMem<R> cm = csg_mem<R>( smem + ( SH ? w.prm.stage_bytes : 0 ), ACN_BLOCK, threadIdx.x );
    const DParams<R>& prm = w.prm;
    const int lane = threadIdx.x & 31;
    const unsigned int lt = ( 1u << lane ) - 1u;
    const int L = prm.n_lights;
    const unsigned long long n_blocks = blk_hi - blk_lo;
    const R t_lim = prm.max_path_length;
    PathRing<R>& rg = rings[ threadIdx.x >> 5 ];
    int head = 0, ring_n = 0;
    bool input_done = false;
    unsigned long long b_cur = 0, b_end = 0;
    unsigned long long n_path = 0;
    // the lane's walk
    int cur = WALK_END;
    unsigned int tsk = 0, chi = 0;               // task entry, child (bit 31: probe)
    Ray<R> ray; ray.p = ray.d = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    R eps = sv0.eps, ci = R( 0 );
    bool pend = false;
    HitCtx ctx; ctx.key = 0;
    WalkT<R> ws; ws.reset();
    for( ;; )
    {
        const unsigned int idle = __ballot_sync( ACN_FULL, cur < 0 );
        if( idle == ACN_FULL || ( __popc( idle ) >= ACN_REFILL_MIN && !( input_done && ring_n == 0 ) ) )
        {
            const int n_idle = __popc( idle );
            if( pend )
            {   // ---- the rays that came to their end since the last refill, together (see k_rays)
                pend = false;
                const bool probe = ( chi & 0x80000000u ) != 0;
                SceneView<R, SH> sv = sv0; sv.eps = eps;
                const I4 m = in.meta[ tsk ];
                const V3<R> tpm = xyz( in.tpc_b[ tsk ] ) * ( R( 2 ) / ( R )( unsigned int )m.w );   // scene.c:620
                if( !probe ) walk_trans_finish( sv, ray, ws );
                if( !probe && ws.min_a < t_lim ) emit_hit( w, ray, ws.min_a, eps, ws.tl, ci, m.y - 10, tpm, m.x, ctx.key );
                else add_sample( w, m.x, mul( prm.background, tpm ) * ci );                       // scene.c:613-616
            }
            while( !input_done && ring_n < n_idle )
            {   // ---- generate the next block of 32 children
                if( b_cur >= b_end )
                {
                    b_cur = warp_fetch( &w.sc->cur_path, chunk, lane );
                    b_end = b_cur + chunk < n_blocks ? b_cur + chunk : n_blocks;
                    if( b_cur >= n_blocks ) { input_done = true; break; }
                }
                const unsigned long long blk = blk_lo + b_cur; b_cur++;
                const ListWindow lw = list_window( in.cum, pdir, blk, n_entries, lane );
                const unsigned long long idx = ( blk << 5 ) + lane;
                const bool live = idx < c_hi;
                const int j = window_find( lw.incl, live ? idx : ( blk << 5 ) );
                const unsigned long long prev = __shfl_sync( ACN_FULL, lw.incl, ( j + 31 ) & 31 );
                bool want = false;
                Ray<R> out; out.p = out.d = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
                R cint = R( 0 ); unsigned int t = 0, i = 0;
                if( live )
                {
                    t = lw.e0 + j;
                    i = ( unsigned int )( idx - ( j ? prev : lw.excl0 ) );
                    const I4 m = in.meta[ t ];
                    const R4<R> pi = in.pos_id[ t ], nc = in.nrm_ci[ t ], pa = in.prj_a[ t ], tb = in.tpc_b[ t ];
                    const V3<R> nrm = xyz( nc );
                    const Basis<R> bs = basis_con_z( nrm );
                    u64 rv = skip2( prm, in.rv0[ t ], ( unsigned long long )L * ( unsigned int )m.z + i );
                    out.p = xyz( pi );
                    out.d = from_basis( bs, sphere_cap<R>( &rv, R( 1 ) ) );
                    R wgt = dot( out.d, nrm );
                    if( wgt > R( 0 ) )                                                           // scene.c:600
                    {
                        if( tb.w > R( 0 ) ) wgt = oren_nayar( wgt, nc.w, pa.w, tb.w, out.d, nrm, xyz( pa ) );
                        n_path++;
                        cint = wgt * pi.w;
                        if( ( m.y - 10 ) == 0 || cint < prm.min_intensity ) i |= 0x80000000u;    // probe: its shading would return 0
                        want = true;
                    }
                }
                const unsigned int wm = __ballot_sync( ACN_FULL, want );
                if( want )
                {
                    const int q = ( head + ring_n + __popc( wm & lt ) ) & ( ACN_RING - 1 );
                    rg.ox[ q ] = out.p.x; rg.oy[ q ] = out.p.y; rg.oz[ q ] = out.p.z;
                    rg.dx[ q ] = out.d.x; rg.dy[ q ] = out.d.y; rg.dz[ q ] = out.d.z;
                    rg.ci[ q ] = cint; rg.t[ q ] = t; rg.i[ q ] = i;
                }
                ring_n += __popc( wm );
                __syncwarp();
            }
            const int n_take = n_idle < ring_n ? n_idle : ring_n;
            if( n_take > 0 )
            {
                const int r = __popc( idle & lt );
                if( cur < 0 && r < n_take )
                {
                    const int q = ( head + r ) & ( ACN_RING - 1 );
                    ray.p = v3<R>( rg.ox[ q ], rg.oy[ q ], rg.oz[ q ] ); ray.d = v3<R>( rg.dx[ q ], rg.dy[ q ], rg.dz[ q ] );
                    ci = rg.ci[ q ]; tsk = rg.t[ q ]; chi = rg.i[ q ];
                    eps = ray_eps( prm, sv0.eps, ray.p );
                    ctx.key = mix64( in.key[ tsk ], KEY_PATH0 + ( chi & 0x7FFFFFFFu ) );
                    ws.reset();
                    SceneView<R, SH> sv = sv0; sv.eps = eps;
                    cur = walk_root( sv, ray, false, t_lim );
                    if( cur == WALK_END )
                    {   // scene.c:613-616: nothing within max_path_length: background
                        const I4 m = in.meta[ tsk ];
                        add_sample( w, m.x, mul( prm.background, xyz( in.tpc_b[ tsk ] ) * ( R( 2 ) / ( R )( unsigned int )m.w ) ) * ci );
                    }
                }
                head = ( head + n_take ) & ( ACN_RING - 1 ); ring_n -= n_take;
                __syncwarp();
            }
            else if( idle == ACN_FULL ) break;
        }
        if( cur >= 0 )
        {
            const bool probe = ( chi & 0x80000000u ) != 0;
            SceneView<R, SH> sv = sv0; sv.eps = eps;
            const int nx = walk_step<R, MARCH>( sv, ray, t_lim, !probe, cur, ws, ctx, cm );
            pend = nx == WALK_END;
            cur = nx;
        }
    }
    warp_count( &w.sc->stats[ ST_PATH ], n_path, lane );
}
template <typename R, int MARCH, bool SH> __global__ void __launch_bounds__( ACN_BLOCK, sizeof( R ) == 8 ? 4 : SH ? ACN_MINB_PATH : ACN_MINB_PATH_G )
k_path( Wave<R> w, TaskBuf<R> in, const unsigned int* __restrict__ pdir )
{
#if defined(ACN_SPEC_SCENE)
    k_path_groups<R, MARCH, SH>( w, in, pdir );
#else
    // scenes of planes, spheres and quadrics walk with refill; where composite objects are met the rays of a group stay together
    // (they reach the same object in the same iteration and run ONE event sweep program side by side)
    if constexpr( MARCH == 2 ) k_path_refill<R, MARCH, SH>( w, in, pdir );
    else k_path_groups<R, MARCH, SH>( w, in, pdir );
#endif
}

// New tasks of the iteration -> (a) the direct list: one entry per task with >= 1 shadow child,
// (b) the task stack: tasks that still have path children.  A warp allocates its entries and their
// child ranges with ONE packed atomic (entries << 38 | children), so entry order and child order agree.
template <typename R> __global__ void __launch_bounds__( 256 )
k_index( Sched* s, TaskBuf<R> in, unsigned long long in_cap, int n_lights,
         u64* __restrict__ dl_cum, unsigned int* __restrict__ dl_slot, unsigned int* __restrict__ dl_dir,
         unsigned long long dl_cap, unsigned long long dl_dir_cap,
         TaskBuf<R> stack, unsigned int* __restrict__ pdir, unsigned long long stack_cap, unsigned long long pdir_cap )
{
    if( blockIdx.x == 0 && threadIdx.x == 0 && s->fix_slot != ACN_NONE64 ) { stack.cum[ s->fix_slot ] = s->fix_cum; s->fix_slot = ACN_NONE64; }
    unsigned long long count = s->tasks_new;
    if( count > in_cap ) count = in_cap;
    if( count == 0 || s->overflow ) return;
    const int lane = threadIdx.x & 31;
    const unsigned int lt = ( 1u << lane ) - 1u;
    const unsigned long long warps = ( ( unsigned long long )gridDim.x * blockDim.x ) >> 5;
    for( unsigned long long g = ( ( unsigned long long )blockIdx.x * blockDim.x + threadIdx.x ) >> 5; ( g << 5 ) < count; g += warps )
    {
        const unsigned long long i = ( g << 5 ) + lane;
        I4 m; m.x = m.y = m.z = m.w = 0;
        if( i < count ) m = in.meta[ i ];
        // ---- direct list
        {
            const unsigned int cd = ( unsigned int )m.z * ( unsigned int )n_lights;
            const unsigned int mask = __ballot_sync( ACN_FULL, cd > 0 );
            if( mask )
            {
                unsigned int inc = cd;
                #pragma unroll
                for( int o = 1; o < 32; o <<= 1 ) { unsigned int v = __shfl_up_sync( ACN_FULL, inc, o ); if( lane >= o ) inc += v; }
                const unsigned int total = __shfl_sync( ACN_FULL, inc, 31 );
                unsigned long long base = 0;
                if( lane == 0 ) base = atomicAdd( &s->dl_packed, ( ( unsigned long long )__popc( mask ) << ACN_TASK_SHIFT ) | total );
                base = __shfl_sync( ACN_FULL, base, 0 );
                if( cd > 0 )
                {
                    const unsigned long long e = ( base >> ACN_TASK_SHIFT ) + __popc( mask & lt );
                    const unsigned long long incl = ( base & ACN_TASK_MASK ) + inc, excl = incl - cd;
                    if( e < dl_cap && ( ( incl + 31 ) >> 5 ) <= dl_dir_cap )
                    {
                        dl_cum[ e ] = incl; dl_slot[ e ] = ( unsigned int )i;
                        for( unsigned long long b = ( excl + 31 ) >> 5; ( b << 5 ) < incl; b++ ) dl_dir[ b ] = ( unsigned int )e;
                    }
                    else s->overflow = 4;
                }
            }
        }
        // ---- task stack
        {
            const unsigned int np = ( unsigned int )m.w;
            const unsigned int mask = __ballot_sync( ACN_FULL, np > 0 );
            if( mask )
            {
                unsigned int inc = np;
                #pragma unroll
                for( int o = 1; o < 32; o <<= 1 ) { unsigned int v = __shfl_up_sync( ACN_FULL, inc, o ); if( lane >= o ) inc += v; }
                const unsigned int total = __shfl_sync( ACN_FULL, inc, 31 );
                unsigned long long base = 0;
                if( lane == 0 ) base = atomicAdd( &s->task_stack, ( ( unsigned long long )__popc( mask ) << ACN_TASK_SHIFT ) | total );
                base = __shfl_sync( ACN_FULL, base, 0 );
                if( np > 0 )
                {
                    const unsigned long long e = ( base >> ACN_TASK_SHIFT ) + __popc( mask & lt );
                    const unsigned long long incl = ( base & ACN_TASK_MASK ) + inc, excl = incl - np;
                    if( e < stack_cap && ( ( incl + 31 ) >> 5 ) <= pdir_cap )
                    {
                        stack.pos_id[ e ] = in.pos_id[ i ]; stack.nrm_ci[ e ] = in.nrm_ci[ i ];
                        stack.prj_a[ e ]  = in.prj_a[ i ];  stack.tpc_b[ e ]  = in.tpc_b[ i ];
                        stack.meta[ e ] = m; stack.rv0[ e ] = in.rv0[ i ]; stack.key[ e ] = in.key[ i ];
                        stack.cum[ e ] = incl;
                        for( unsigned long long b = ( excl + 31 ) >> 5; ( b << 5 ) < incl; b++ ) pdir[ b ] = ( unsigned int )e;
                    }
                    else s->overflow = 3;
                }
            }
        }
    }
}

// cl_s_sat (vectors.h:372-384, scene.c:1010): pow(c, gamma) then clamp, per sample
__device__ __forceinline__ float  acc_value( unsigned long long a, unsigned long long flags, int ch, float )
{
    return ( ( flags >> ch ) & 1ull ) ? ACN_ACC_SAT : ( float )( ( double )a * ACN_ACC_INV );
}
__device__ __forceinline__ double acc_value( double a, double, int, double ) { return a; }

template <typename R> __global__ void k_finish( const typename Acc<R>::T* accum, unsigned long long n, R gamma, float* rgb )
{
    unsigned long long i = ( unsigned long long )blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n * 3 ) return;
    const unsigned long long smp = i / 3; const int ch = ( int )( i - smp * 3 );
    R v = r_pow( acc_value( accum[ 4 * smp + ch ], accum[ 4 * smp + 3 ], ch, R( 0 ) ), gamma );
    v = v > R( 0 ) ? ( v < R( 1 ) ? v : R( 1 ) ) : R( 0 );
    rgb[ i ] = ( float )v;
}

// lum_image_s_push (scene.c:804-813) on the device
__global__ void k_accumulate( const double* xy, const float* rgb, unsigned long long n, int width, int height, float* accum )
{
    unsigned long long i = ( unsigned long long )blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    int x = ( int )xy[ 2 * i ], y = ( int )xy[ 2 * i + 1 ];
    if( x < 0 || x >= width || y < 0 || y >= height ) return;
    float* a = accum + 4ull * ( ( unsigned long long )y * width + x );
    atomicAdd( a + 0, rgb[ 3 * i + 0 ] );
    atomicAdd( a + 1, rgb[ 3 * i + 1 ] );
    atomicAdd( a + 2, rgb[ 3 * i + 2 ] );
    atomicAdd( a + 3, 1.0f );
}

} // namespace acn
