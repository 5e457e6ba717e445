// acn_tracer.cuh — host side of the wavefront sample tracer: scene upload, CSG program builder, queues, the launch loop.
// The kernels are in acn_kernels.cuh; the run-time specialisation of the kernels to one scene in acn_spec.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <string>
#include <algorithm>
#include <functional>

#include <stdarg.h>

#include "acn_kernels.cuh"
#include "../../include/actinon_b200.h"
#include "acn_rtc.h"
#include "acn_spec.h"

namespace acn {

// ---------------------------------------------------------------------------------------------
// host side of the tracer
// ---------------------------------------------------------------------------------------------
struct TracerBase
{
    virtual ~TracerBase() {}
    virtual int render( const double* d_xy, uint64_t n, uint64_t index_base, float* d_rgb, cudaStream_t st,
                        const volatile int* cancel, acn_stats* stats ) = 0;
    int width = 0, height = 0, device = 0;
    cudaStream_t own_stream = nullptr;
    // second stream: k_path runs beside k_rays and k_direct beside the next iteration's k_sched/k_rays, so the
    // tail of one persistent kernel (last warps still tracing) is filled by the blocks of the next
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_sched = nullptr, ev_path = nullptr, ev_index = nullptr, ev_direct = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;      // device time of a render call
    // staging for the host-pointer entry point
    double* d_xy_stage = nullptr; float* d_rgb_stage = nullptr; uint64_t stage_cap = 0;
};

void set_error( const char* fmt, ... );

#define ACN_CUDA( call ) do { cudaError_t e__ = ( call ); if( e__ != cudaSuccess ) { \
    set_error( "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString( e__ ) ); return ACN_ERR_CUDA; } } while( 0 )

template <typename T> static int dev_alloc( T** p, size_t n )
{
    cudaError_t e = cudaMalloc( ( void** )p, n * sizeof( T ) );
    if( e != cudaSuccess ) { set_error( "cudaMalloc(%zu bytes) failed: %s", n * sizeof( T ), cudaGetErrorString( e ) ); return ACN_ERR_OUT_OF_MEMORY; }
    return ACN_OK;
}

template <typename R> static int alloc_rays( RayBuf<R>& b, size_t n )
{
    int rc;
    if( ( rc = dev_alloc( &b.o_i, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.d_, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.tp, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.meta, n ) ) ) return rc;
    return ACN_OK;
}
template <typename R> static void free_rays( RayBuf<R>& b ) { cudaFree( b.o_i ); cudaFree( b.d_ ); cudaFree( b.tp ); cudaFree( b.meta ); }

template <typename R> static int alloc_hits( HitBuf<R>& b, size_t n )
{
    int rc;
    if( ( rc = dev_alloc( &b.o_a, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.d_i, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.n_e, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.tp, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.meta, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.key, n ) ) ) return rc;
    return ACN_OK;
}
template <typename R> static void free_hits( HitBuf<R>& b )
{
    cudaFree( b.o_a ); cudaFree( b.d_i ); cudaFree( b.n_e ); cudaFree( b.tp ); cudaFree( b.meta ); cudaFree( b.key );
}

template <typename R> static int alloc_tasks( TaskBuf<R>& b, size_t n )
{
    int rc;
    if( ( rc = dev_alloc( &b.pos_id, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.nrm_ci, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.prj_a, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.tpc_b, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.meta, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.rv0, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.key, n ) ) ) return rc;
    if( ( rc = dev_alloc( &b.cum, n ) ) ) return rc;
    return ACN_OK;
}
template <typename R> static void free_tasks( TaskBuf<R>& b )
{
    cudaFree( b.pos_id ); cudaFree( b.nrm_ci ); cudaFree( b.prj_a ); cudaFree( b.tpc_b );
    cudaFree( b.meta ); cudaFree( b.rv0 ); cudaFree( b.key ); cudaFree( b.cum );
}

template <typename R> struct Tracer : TracerBase
{
    DParams<R> prm;
    // device copies of the scene tables
    R4<R>* d_env = nullptr; I4* d_link = nullptr; R4<R>* d_geo = nullptr; int* d_children = nullptr; CRec<R>* d_crec = nullptr;
    int* d_prog = nullptr; I4* d_prog_ref = nullptr; int* d_parent = nullptr; int n_prog = 0;
    size_t n_rec = 0; int rec_light = -1, rec_matter = -1;     // traversal records: count, first record of each root list (-1: empty)
    int rec_matter_oct[ 8 ] = { -1, -1, -1, -1, -1, -1, -1, -1 };   // the matter list laid out front to back per octant of ray directions (= rec_matter when not built)
    DMat<R>* d_mats = nullptr; DLight<R>* d_lights = nullptr;
    u64* d_skipA = nullptr; u64* d_skipC = nullptr;
    // queues
    RayBuf<R>  ray_stack;
    TaskBuf<R> task_stack, task_new;
    HitBuf<R>  hit_q;
    u64* d_dl_cum = nullptr; unsigned int* d_dl_slot = nullptr; unsigned int* d_dl_dir = nullptr; unsigned int* d_pdir = nullptr;
    uint64_t   budget_opt = 0;      // acn_options.wave_budget (0: follow the size of the call)
    uint64_t   budget_floor = 0;    // raised when a call overflowed the queues of an automatic budget
    uint64_t   budget = 0, ray_min = 0, ray_cap = 0, task_stack_cap = 0, task_new_cap = 0, dl_dir_cap = 0, pdir_cap = 0, prim_chunk = 0;
    Sched*     d_sc = nullptr;
    Sched*     h_sc = nullptr;        // pinned
    typename Acc<R>::T* d_accum = nullptr; uint64_t accum_cap = 0;
    int        smem_bytes = 0;
    int        max_csg_depth = 0;
    int        grid_trace[ 5 ] = { 0, 0, 0, 0, 0 };   // persistent grids: primary, rays, path, direct, shade
    int        grid_util = 0;
    std::shared_ptr<SpecModule> spec;   // kernels specialised to this scene's structure (NVRTC); null: the generic kernels
    bool       march = false;     // kernels instantiated with the reference's recursive CSG march (scale nodes, CSG over distance fields, f64 validation)
    void ( *kp_primary )( Wave<R>, const double* ) = nullptr;
    void ( *kp_rays )( Wave<R>, RayBuf<R> ) = nullptr;
    void ( *kp_path )( Wave<R>, TaskBuf<R>, const unsigned int* ) = nullptr;
    void ( *kp_direct )( Wave<R>, TaskBuf<R>, const u64*, const unsigned int*, const unsigned int* ) = nullptr;
    void ( *kp_shade )( Wave<R>, HitBuf<R> ) = nullptr;
    template <int MARCH, bool SH> void select_kernels()
    {
        kp_primary = k_primary<R, MARCH, SH>; kp_rays = k_rays<R, MARCH, SH>; kp_path = k_path<R, MARCH, SH>;
        kp_direct = k_direct<R, MARCH, SH>; kp_shade = k_shade<R, SH>;
    }

    ~Tracer() override
    {
        cudaSetDevice( device );
        cudaFree( d_env ); cudaFree( d_link ); cudaFree( d_geo ); cudaFree( d_children ); cudaFree( d_crec );
        cudaFree( d_prog ); cudaFree( d_prog_ref ); cudaFree( d_parent );
        cudaFree( d_mats ); cudaFree( d_lights ); cudaFree( d_skipA ); cudaFree( d_skipC );
        free_rays( ray_stack );
        free_tasks( task_stack ); free_tasks( task_new ); free_hits( hit_q );
        cudaFree( d_dl_cum ); cudaFree( d_dl_slot ); cudaFree( d_dl_dir ); cudaFree( d_pdir );
        cudaFree( d_sc ); if( h_sc ) cudaFreeHost( h_sc );
        cudaFree( d_accum ); cudaFree( d_xy_stage ); cudaFree( d_rgb_stage );
        if( own_stream ) cudaStreamDestroy( own_stream );
        if( side_stream ) cudaStreamDestroy( side_stream );
        if( ev_sched ) cudaEventDestroy( ev_sched ); if( ev_path ) cudaEventDestroy( ev_path );
        if( ev_index ) cudaEventDestroy( ev_index ); if( ev_direct ) cudaEventDestroy( ev_direct );
        if( ev_t0 ) cudaEventDestroy( ev_t0 ); if( ev_t1 ) cudaEventDestroy( ev_t1 );
    }

    int init( const acn_flat_scene* fs, const acn_options* opt );
    int ensure_queues( uint64_t n );
    int render( const double* d_xy, uint64_t n, uint64_t index_base, float* d_rgb, cudaStream_t st,
                const volatile int* cancel, acn_stats* stats ) override;

    Wave<R> make_wave( uint64_t index_base )
    {
        Wave<R> w;
        w.prm = prm; w.rays_out = ray_stack; w.tasks_out = task_new; w.hits_out = hit_q; w.sc = d_sc; w.accum = d_accum;
        w.rays_cap = ray_cap; w.tasks_cap = task_new_cap; w.hits_cap = task_new_cap; w.index_base = index_base;
        return w;
    }
};

// depth of the CSG recursion below node n (for the device stack size)
static int csg_depth( const acn_flat_scene* fs, int n, int guard )
{
    if( guard > 64 ) return 64;
    const acn_flat_node& nd = fs->nodes[ n ];
    switch( nd.kind )
    {
        case ACN_KIND_PAIR_INSIDE: case ACN_KIND_PAIR_OUTSIDE:
        {
            int a = csg_depth( fs, nd.child0, guard + 1 ), b = csg_depth( fs, nd.child1, guard + 1 );
            return 1 + ( a > b ? a : b );
        }
        case ACN_KIND_NEG: case ACN_KIND_SCALE: return 1 + csg_depth( fs, nd.child0, guard + 1 );
        case ACN_KIND_COMPOUND:
        {
            int m = 0;
            for( int i = 0; i < nd.child1; i++ ) { int d = csg_depth( fs, fs->children[ nd.child0 + i ], guard + 1 ); if( d > m ) m = d; }
            return m;
        }
        default: return 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Cull bounds of the traversal records.  The reference tests an element's envelope before the element
// (objects.c:261-266, compound.c:215-244): a hit needs "ray meets the envelope AND ray meets the contents".
// The envelopes of the shipped scenes are Monte-Carlo estimates (objects.c:312-363): loose where they matter (a
// cluster of eight spheres of many_spheres.acn: radius 0.092 where 0.068 encloses it) and off-centre, so that they
// do not even contain their contents everywhere.  With B a ball that contains the contents, "ray meets B AND ray
// meets the envelope" is necessary for a hit; when B lies inside the envelope the first implies the second.  The
// traversal records therefore carry B, the tightest ball the host can construct (plus a safety margin for FP32):
//   sphere       B = the sphere itself: the record holds the shape (F_SELF), one test instead of two, no
//                second load of the geometry table
//   compound     B = enclosing ball of the children's balls (Badoiu-Clarkson iteration), also for compounds the
//                script gave no envelope at all
//   anything else (planes, quadrics, CSG, distance fields) keeps the reference's envelope.
// Where B does not fit inside the element's envelope the record asks for the envelope test as well (F_ENV2, read
// from env[]): the rays that pass B — about half of those that pass the envelope — pay a second test.  That second
// test is a plain gate (objects.c:264), NOT culled against the horizon: contents that stick out of their envelope can be
// hit in front of it.  (Culling an envelope against the best hit so far assumes that the contents lie inside it.  The
// walk relies on that for the elements whose extent the host cannot compute — composite objects, quadrics — as the
// scripts of the reference do when they set bounding envelopes by hand; for spheres and compounds of spheres it is
// checked here.)  Same rays accepted and rejected as by the reference's order of tests; the node tables (env[], link[])
// are untouched: they serve the CSG programs, the light tests and the FP64 march of the validation mode.
// ---------------------------------------------------------------------------------------------
struct CullBounds
{
    enum { KEEP = 0, TIGHT = 1, SELF = 2 };
    struct Rec { int mode = KEEP; bool env_too = false; double c[ 3 ] = { 0, 0, 0 }, r = 0; };
    struct Ball { bool ok = false; double c[ 3 ] = { 0, 0, 0 }, r = 0; };
    std::vector<Rec> rec;
    std::vector<Ball> ball;
    int n_tight = 0, n_self = 0, n_both = 0;

    static double dist( const double* a, const double* b ) { return sqrt( ( a[ 0 ] - b[ 0 ] ) * ( a[ 0 ] - b[ 0 ] ) + ( a[ 1 ] - b[ 1 ] ) * ( a[ 1 ] - b[ 1 ] ) + ( a[ 2 ] - b[ 2 ] ) * ( a[ 2 ] - b[ 2 ] ) ); }
    static double margin( const double* c, double r ) { return 2E-5 * fmax( 1.0, fmax( fmax( fabs( c[ 0 ] ), fabs( c[ 1 ] ) ), fabs( c[ 2 ] ) ) + r ); }

    // enclosing ball of balls
    static Ball enclose( const std::vector<Ball>& in )
    {
        Ball b; b.ok = true;
        for( const Ball& k : in ) for( int j = 0; j < 3; j++ ) b.c[ j ] += k.c[ j ] / ( double )in.size();
        auto radius = [ & ]( const double* c, int* far ) { double r = 0; for( size_t i = 0; i < in.size(); i++ ) { const double d = dist( c, in[ i ].c ) + in[ i ].r; if( d > r ) { r = d; if( far ) *far = ( int )i; } } return r; };
        double c[ 3 ] = { b.c[ 0 ], b.c[ 1 ], b.c[ 2 ] };
        int far = 0;
        b.r = radius( c, &far );
        for( int it = 1; it <= 256 && in.size() > 1; it++ )
        {   // step towards the farthest point of the farthest ball
            const Ball& f = in[ far ];
            const double d = dist( c, f.c );
            double fp[ 3 ];
            for( int j = 0; j < 3; j++ ) fp[ j ] = f.c[ j ] + ( d > 0 ? ( f.c[ j ] - c[ j ] ) / d * f.r : 0.0 );
            for( int j = 0; j < 3; j++ ) c[ j ] += ( fp[ j ] - c[ j ] ) / ( double )( it + 1 );
            const double r = radius( c, &far );
            if( r < b.r ) { b.r = r; for( int j = 0; j < 3; j++ ) b.c[ j ] = c[ j ]; }
        }
        return b;
    }

    CullBounds( const acn_flat_scene* fs, bool enable )
    {
        const int n = fs->n_nodes;
        rec.resize( n ); ball.resize( n );
        if( !enable ) return;
        // children have larger or smaller indices than their compound depending on the flattener: resolve recursively
        std::vector<char> done( n, 0 );
        std::function<void( int )> solve = [ & ]( int i )
        {
            if( done[ i ] ) return;
            done[ i ] = 1;
            const acn_flat_node& nd = fs->nodes[ i ];
            Ball b;
            if( nd.kind == ACN_KIND_SPHERE && nd.tail[ 0 ] > 0 ) { b.ok = true; for( int j = 0; j < 3; j++ ) b.c[ j ] = nd.pos[ j ]; b.r = nd.tail[ 0 ]; }
            else if( nd.kind == ACN_KIND_COMPOUND && nd.child1 > 0 )
            {
                std::vector<Ball> ch;
                bool all = true;
                for( int k = 0; k < nd.child1; k++ )
                {
                    const int c = fs->children[ nd.child0 + k ];
                    solve( c );
                    if( !ball[ c ].ok ) { all = false; break; }
                    ch.push_back( ball[ c ] );
                }
                if( all ) b = enclose( ch );
            }
            ball[ i ] = b;
            if( !b.ok ) return;
            const double m = margin( b.c, b.r );
            Rec& r = rec[ i ];
            r.env_too = nd.has_envelope && !( dist( b.c, nd.env_pos ) + b.r + 2 * m <= nd.env_radius );
            if( r.env_too ) n_both++;
            for( int j = 0; j < 3; j++ ) r.c[ j ] = b.c[ j ];
            if( nd.kind == ACN_KIND_SPHERE ) { r.mode = SELF; r.r = b.r; n_self++; }
            else                             { r.mode = TIGHT; r.r = b.r + m; n_tight++; }
        };
        for( int i = 0; i < n; i++ ) solve( i );
    }
};

static int compound_depth( const acn_flat_scene* fs, int n, int guard )
{
    if( guard > 64 ) return 64;
    const acn_flat_node& nd = fs->nodes[ n ];
    if( nd.kind != ACN_KIND_COMPOUND ) return 0;
    int m = 0;
    for( int i = 0; i < nd.child1; i++ ) { int d = compound_depth( fs, fs->children[ nd.child0 + i ], guard + 1 ); if( d > m ) m = d; }
    return 1 + m;
}


// ---------------------------------------------------------------------------------------------
// postfix programs for the interval CSG evaluator (acn_geom.h: csg_fast_hit)
// ---------------------------------------------------------------------------------------------
struct CsgBuilder
{
    const acn_flat_scene* fs;
    std::vector<int> prog, parent;
    std::vector<I4>  prog_ref;

    static bool is_dist( int k ) { return k == ACN_KIND_DIST_SPHERE || k == ACN_KIND_DIST_TORUS; }
    static bool is_leaf( int k ) { return k == ACN_KIND_PLANE || k == ACN_KIND_SPHERE || k == ACN_KIND_SQUAROID || is_dist( k ); }
    static bool is_pair( int k ) { return k == ACN_KIND_PAIR_INSIDE || k == ACN_KIND_PAIR_OUTSIDE; }

    bool eligible( int n, int guard, bool in_scale = false ) const
    {
        if( guard > 64 ) return false;
        const acn_flat_node& nd = fs->nodes[ n ];
        if( is_dist( nd.kind ) && getenv( "ACN_NO_SWEEP_DIST" ) ) return false;      // diagnostics
        if( is_leaf( nd.kind ) ) return true;
        if( is_pair( nd.kind ) ) return eligible( nd.child0, guard + 1, in_scale ) && eligible( nd.child1, guard + 1, in_scale );
        if( nd.kind == ACN_KIND_NEG ) return eligible( nd.child0, guard + 1, in_scale );
        // obj_scale_s: its child is classified against the ray carried into the scaled frame (XFORM .. XEND); one level
        if( nd.kind == ACN_KIND_SCALE ) return !in_scale && !getenv( "ACN_NO_SWEEP_SCALE" ) && eligible( nd.child0, guard + 1, true );
        return false;
    }

    // operands of the maximal chain of one operator below n: children of the same pair kind are opened up
    // unless they carry an envelope (which clips THEIR set, so they stay a unit).  A&(B&C) = (A&B)&C as point sets.
    void operands( int n, int kind, bool root, std::vector<int>& out ) const
    {
        const acn_flat_node& nd = fs->nodes[ n ];
        if( nd.kind == kind && ( root || !nd.has_envelope ) ) { operands( nd.child0, kind, false, out ); operands( nd.child1, kind, false, out ); }
        else out.push_back( n );
    }

    // convex along every ray, so that its inside-set is one interval: half-space, ball, and the squaroids
    // a x^2 + b y^2 + c z^2 + r <= 0 with a,b,c >= 0 > r (ellipsoid, elliptic cylinder, slab)
    bool convex_leaf( int n ) const
    {
        const acn_flat_node& nd = fs->nodes[ n ];
        if( nd.has_envelope ) return false;
        if( nd.kind == ACN_KIND_PLANE || nd.kind == ACN_KIND_SPHERE ) return true;
        if( nd.kind == ACN_KIND_SQUAROID ) return nd.tail[ 0 ] >= 0 && nd.tail[ 1 ] >= 0 && nd.tail[ 2 ] >= 0 && nd.tail[ 3 ] < 0;
        return false;
    }
    // member word of a RUN for operand n, or -1: a convex leaf, or the negation of a plane (the other half-space)
    int run_member( int n ) const
    {
        const acn_flat_node& nd = fs->nodes[ n ];
        if( convex_leaf( n ) ) return CSG_MEMBER | ( n << 4 );
        if( nd.kind == ACN_KIND_NEG && !nd.has_envelope )
        {
            const acn_flat_node& c = fs->nodes[ nd.child0 ];
            if( c.kind == ACN_KIND_PLANE && !c.has_envelope ) return CSG_MEMBER_NEG | ( nd.child0 << 4 );
        }
        return -1;
    }

    int n_vars = 0;
    bool has_dist_leaf = false;       // some program has a distance-field leaf: the kernels with that support (MARCH instantiation) are needed
    bool has_scale = false;           // some program carries the ray into a scaled frame: same kernels

    void emit( int n, bool root )
    {
        const acn_flat_node& nd = fs->nodes[ n ];
        const bool clip = nd.has_envelope && !root;
        size_t skip_slot = 0;
        const int vars_before = n_vars;
        if( clip ) { prog.push_back( CSG_ENV | ( n << 4 ) ); skip_slot = prog.size(); prog.push_back( 0 ); }
        if( is_leaf( nd.kind ) )
        {
            prog.push_back( CSG_LEAF | ( n << 4 ) ); n_vars++;
            if( nd.kind == ACN_KIND_DIST_TORUS ) prog.push_back( CSG_MORE | ( n << 4 ) );      // a torus has up to four crossings
            if( is_dist( nd.kind ) ) has_dist_leaf = true;
        }
        else if( nd.kind == ACN_KIND_NEG ) { emit( nd.child0, false ); prog.push_back( CSG_NEG | ( n << 4 ) ); }
        else if( nd.kind == ACN_KIND_SCALE )
        {
            prog.push_back( CSG_XFORM | ( n << 4 ) ); emit( nd.child0, false ); prog.push_back( CSG_XEND | ( n << 4 ) );
            has_scale = true;
        }
        else
        {
            std::vector<int> ops; operands( n, nd.kind, true, ops );
            const int opc = ( nd.kind == ACN_KIND_PAIR_INSIDE ? CSG_AND : CSG_OR ) | ( n << 4 );
            std::vector<int> members, rest;
            if( nd.kind == ACN_KIND_PAIR_INSIDE )
            {
                for( int o : ops ) { if( run_member( o ) >= 0 ) members.push_back( o ); else rest.push_back( o ); }
                if( members.size() < 2 ) { rest = ops; members.clear(); }
            }
            else rest = ops;
            int emitted = 0;
            if( !members.empty() )
            {
                prog.push_back( CSG_RUN | ( ( int )members.size() << 4 ) );
                for( int o : members ) prog.push_back( run_member( o ) );
                n_vars++; emitted++;
            }
            for( int o : rest ) { emit( o, false ); if( emitted++ > 0 ) prog.push_back( opc ); }
        }
        if( clip )
        {
            prog.push_back( CSG_CLIP | ( n << 4 ) ); n_vars++;
            prog[ skip_slot ] = ( int )( prog.size() - 1 - skip_slot ) | ( ( n_vars - vars_before ) << 16 );
        }
    }

    // the boolean function of a program for one variable assignment (host mirror of csg_state)
    int eval( size_t start, size_t len, unsigned long long vars ) const
    {
        unsigned long long stk = 0;
        int v = 0;
        for( size_t pc = start; pc < start + len; pc++ )
        {
            const int ins = prog[ pc ], op = ins & 15;
            if( op == CSG_LEAF )      { stk = ( stk << 1 ) | ( ( vars >> v ) & 1ull ); v++; }
            else if( op == CSG_RUN )  { stk = ( stk << 1 ) | ( ( vars >> v ) & 1ull ); v++; pc += ( size_t )( ins >> 4 ); }
            else if( op == CSG_CLIP ) { stk &= ~1ull | ( ( vars >> v ) & 1ull ); v++; }
            else if( op == CSG_NEG )  stk ^= 1ull;
            else if( op == CSG_AND )  { const unsigned long long t = stk & 1ull; stk >>= 1; stk &= t | ~1ull; }
            else if( op == CSG_OR )   { const unsigned long long t = stk & 1ull; stk >>= 1; stk |= t; }
            else if( op == CSG_ENV ) pc++;
        }
        return ( int )( stk & 1ull );
    }

    // Do two leaves below n lie on one surface (the same plane, the same sphere or quadric)?  Scripts butt the pieces of
    // a solid together on common cut planes; a ray through such a seam has two crossings at one t, and those must be
    // judged the way the reference's march judges them (csg_eval, "group of crossings").
    void leaves_of( int n, int guard, std::vector<int>& out ) const
    {
        if( guard > 64 ) return;
        const acn_flat_node& nd = fs->nodes[ n ];
        if( is_pair( nd.kind ) ) { leaves_of( nd.child0, guard + 1, out ); leaves_of( nd.child1, guard + 1, out ); }
        else if( nd.kind == ACN_KIND_NEG || nd.kind == ACN_KIND_SCALE ) leaves_of( nd.child0, guard + 1, out );
        else out.push_back( n );
    }
    bool same_surface( int a, int b ) const
    {
        const acn_flat_node& x = fs->nodes[ a ]; const acn_flat_node& y = fs->nodes[ b ];
        if( x.kind != y.kind ) return false;
        const double tol = 4e-6;
        auto near = [ & ]( double p, double q ) { return fabs( p - q ) <= tol * ( 1.0 + fabs( p ) ); };
        if( x.kind == ACN_KIND_PLANE )
        {   // same plane: parallel normals (rax row z) and one offset
            double d = 0, ox = 0, oy = 0;
            for( int k = 0; k < 3; k++ ) { d += x.rax[ 6 + k ] * y.rax[ 6 + k ]; ox += x.pos[ k ] * x.rax[ 6 + k ]; oy += y.pos[ k ] * x.rax[ 6 + k ]; }
            return fabs( fabs( d ) - 1.0 ) < 1e-9 && fabs( ox - oy ) <= tol;
        }
        for( int k = 0; k < 3; k++ ) if( !near( x.pos[ k ], y.pos[ k ] ) ) return false;
        for( int k = 0; k < 4; k++ ) if( !near( x.tail[ k ], y.tail[ k ] ) ) return false;
        if( x.kind == ACN_KIND_SPHERE ) return true;
        for( int k = 0; k < 9; k++ ) if( !near( x.rax[ k ], y.rax[ k ] ) ) return false;
        return true;
    }
    bool coincident_leaves( int n ) const
    {
        std::vector<int> lv; leaves_of( n, 0, lv );
        for( size_t i = 0; i < lv.size(); i++ ) for( size_t j = i + 1; j < lv.size(); j++ ) if( same_surface( lv[ i ], lv[ j ] ) ) return true;
        return false;
    }
    bool has_coincident = false;      // some swept object has two leaves on one surface: crossings must be judged in groups

    int depth( int n, int guard ) const
    {
        if( guard > 64 ) return 64;
        const acn_flat_node& nd = fs->nodes[ n ];
        if( is_pair( nd.kind ) ) { int a = depth( nd.child0, guard + 1 ), b = depth( nd.child1, guard + 1 ); return 1 + ( a > b ? a : b ); }
        if( nd.kind == ACN_KIND_NEG || nd.kind == ACN_KIND_SCALE ) return 1 + depth( nd.child0, guard + 1 );
        return 1;
    }

    void set_parents( int n, int guard )
    {
        if( guard > 64 ) return;
        const acn_flat_node& nd = fs->nodes[ n ];
        if( is_pair( nd.kind ) ) { parent[ nd.child0 ] = n; parent[ nd.child1 ] = n; set_parents( nd.child0, guard + 1 ); set_parents( nd.child1, guard + 1 ); }
        else if( nd.kind == ACN_KIND_NEG || nd.kind == ACN_KIND_SCALE ) { parent[ nd.child0 ] = n; set_parents( nd.child0, guard + 1 ); }
    }

    // ---- evaluation programs for functions of more than CSG_TABLE_VARS variables.
    // The interpreter of the full program costs ~12 instructions per word and is called for every crossing of a ray: the five
    // big objects of a hanging lamp (17 - 35 variables, 42 - 108 words) made it 55 % of k_direct's instructions.  Their
    // functions are chains — unions of butted pieces, a body minus twenty drill holes — and a chain of one operator can be
    // cut anywhere: every stretch of operands with <= CSG_TABLE_VARS variables in all becomes ONE truth table over its
    // (consecutive) variables, and what is left of the program is a handful of table lookups joined by the chain's operator.
    // Words: E_TAB | n << 4 | v0 << 8 | ( table offset behind the program's length word ) << 14;  E_VAR | v << 4;  E_CLIP | v << 4;  E_NEG;  E_AND;  E_OR.
    struct TNode { int kind = 0, var = 0, v0 = 0, nv = 0; std::vector<int> ch; };     // kind: 0 variable, 1 not, 2 and, 3 or
    std::vector<TNode> tn;
    int tn_var( int v ) { TNode t; t.kind = 0; t.var = v; t.v0 = v; t.nv = 1; tn.push_back( t ); return ( int )tn.size() - 1; }
    int tn_op( int kind, int a, int b )
    {
        if( tn[ a ].kind == kind ) { tn[ a ].ch.push_back( b ); tn[ a ].nv += tn[ b ].nv; return a; }      // a op b op c: one chain
        TNode t; t.kind = kind; t.ch = { a, b }; t.v0 = tn[ a ].v0; t.nv = tn[ a ].nv + tn[ b ].nv; tn.push_back( t ); return ( int )tn.size() - 1;
    }
    int tn_eval( int n, unsigned long long vars ) const
    {
        const TNode& t = tn[ n ];
        if( t.kind == 0 ) return ( int )( ( vars >> t.var ) & 1ull );
        if( t.kind == 1 ) return !tn_eval( t.ch[ 0 ], vars );
        for( int c : t.ch ) { const int v = tn_eval( c, vars ); if( t.kind == 2 ? !v : v ) return t.kind == 2 ? 0 : 1; }
        return t.kind == 2 ? 1 : 0;
    }
    // the expression tree of a program (same stack discipline as eval)
    int parse_tree( size_t start, size_t len )
    {
        tn.clear();
        std::vector<int> stk;
        int v = 0;
        for( size_t pc = start; pc < start + len; pc++ )
        {
            const int ins = prog[ pc ], op = ins & 15;
            if( op == CSG_LEAF ) stk.push_back( tn_var( v++ ) );
            else if( op == CSG_RUN ) { stk.push_back( tn_var( v++ ) ); pc += ( size_t )( ins >> 4 ); }
            else if( op == CSG_CLIP ) { const int a = stk.back(); stk.pop_back(); TNode t; t.kind = 2; t.ch = { a, tn_var( v++ ) }; t.v0 = tn[ a ].v0; t.nv = tn[ a ].nv + 1; tn.push_back( t ); stk.push_back( ( int )tn.size() - 1 ); }
            else if( op == CSG_NEG ) { const int a = stk.back(); stk.pop_back(); TNode t; t.kind = 1; t.ch = { a }; t.v0 = tn[ a ].v0; t.nv = tn[ a ].nv; tn.push_back( t ); stk.push_back( ( int )tn.size() - 1 ); }
            else if( op == CSG_AND || op == CSG_OR ) { const int b = stk.back(); stk.pop_back(); const int a = stk.back(); stk.pop_back(); stk.push_back( tn_op( op == CSG_AND ? 2 : 3, a, b ) ); }
            else if( op == CSG_ENV ) pc++;
        }
        return stk.size() == 1 ? stk[ 0 ] : -1;
    }
    void pack_table( int kind, const std::vector<int>& members, std::vector<int>& code, std::vector<std::pair<size_t, std::vector<int>>>& tables )
    {   // one truth table over the consecutive variables of `members`, joined by `kind`
        const int v0 = tn[ members.front() ].v0;
        int nv = 0; for( int m : members ) nv += tn[ m ].nv;
        if( members.size() == 1 && tn[ members[ 0 ] ].kind == 0 ) { code.push_back( E_VAR | ( v0 << 4 ) ); return; }
        const size_t rows = ( size_t )1 << nv;
        std::vector<int> tab( ( rows + 31 ) / 32, 0 );
        for( size_t a = 0; a < rows; a++ )
        {
            const unsigned long long vars = ( unsigned long long )a << v0;
            int r = kind == 2 ? 1 : 0;
            for( int m : members ) { const int x = tn_eval( m, vars ); if( kind == 2 ) r &= x; else r |= x; }
            if( r ) tab[ a >> 5 ] |= ( int )( 1u << ( a & 31 ) );
        }
        tables.push_back( { code.size(), tab } );
        code.push_back( E_TAB | ( nv << 4 ) | ( v0 << 8 ) );     // | table offset relative to the program's first word << 14, patched when the tables are appended
    }
    void pack( int n, std::vector<int>& code, std::vector<std::pair<size_t, std::vector<int>>>& tables )
    {
        const TNode& t = tn[ n ];
        if( t.nv <= CSG_TABLE_VARS ) { pack_table( 3, { n }, code, tables ); return; }
        if( t.kind == 1 ) { pack( t.ch[ 0 ], code, tables ); code.push_back( E_NEG ); return; }
        std::vector<int> bin; int bin_nv = 0, items = 0;
        auto flush = [ & ]() { if( bin.empty() ) return; pack_table( t.kind, bin, code, tables ); if( items++ > 0 ) code.push_back( t.kind == 2 ? E_AND : E_OR ); bin.clear(); bin_nv = 0; };
        for( int c : t.ch )
        {
            if( tn[ c ].nv > CSG_TABLE_VARS ) { flush(); pack( c, code, tables ); if( items++ > 0 ) code.push_back( t.kind == 2 ? E_AND : E_OR ); continue; }
            if( bin_nv + tn[ c ].nv > CSG_TABLE_VARS ) flush();
            bin.push_back( c ); bin_nv += tn[ c ].nv;
        }
        flush();
    }
    // appends the evaluation program of [ start, start + len ) to prog; returns its offset (first word: number of words) or -1
    int build_eval_program( size_t start, size_t len )
    {
        const int root = parse_tree( start, len );
        if( root < 0 ) return -1;
        std::vector<int> code;
        std::vector<std::pair<size_t, std::vector<int>>> tables;
        pack( root, code, tables );
        const int at = ( int )prog.size();
        prog.push_back( ( int )code.size() );
        prog.insert( prog.end(), code.begin(), code.end() );
        for( auto& tb : tables )
        {
            const size_t rel = prog.size() - ( size_t )at;
            if( rel >= ( ( size_t )1 << 17 ) ) { prog.resize( at ); return -1; }      // 18 bits of the word, sign bit kept clear
            prog[ ( size_t )at + 1 + tb.first ] |= ( int )( rel << 14 );
            prog.insert( prog.end(), tb.second.begin(), tb.second.end() );
        }
        // the packed function must be the function of the program: checked on random assignments (and on all of a small one)
        const int nv = tn[ root ].nv;
        unsigned long long x = 0x9E3779B97F4A7C15ull;
        for( int k = 0; k < 4096; k++ )
        {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            const unsigned long long vars = nv >= 64 ? x : ( x & ( ( 1ull << nv ) - 1ull ) );
            if( eval_packed( at, vars ) != eval( start, len, vars ) ) { prog.resize( at ); return -1; }
        }
        return at;
    }
    int eval_packed( int at, unsigned long long vars ) const           // host mirror of the device interpreter (acn_isect.cuh: csg_state)
    {
        unsigned int stk = 0;
        const int n = prog[ at ];
        for( int pc = at + 1; pc < at + 1 + n; pc++ )
        {
            const int ins = prog[ pc ], op = ins & 15;
            if( op == E_TAB )
            {
                const int nv = ( ins >> 4 ) & 15, v0 = ( ins >> 8 ) & 63;
                const unsigned int idx = ( unsigned int )( ( vars >> v0 ) & ( ( 1ull << nv ) - 1ull ) );
                const int off = at + ( int )( ( unsigned int )ins >> 14 );
                stk = ( stk << 1 ) | ( ( ( unsigned int )prog[ off + ( int )( idx >> 5 ) ] >> ( idx & 31u ) ) & 1u );
            }
            else if( op == E_VAR )  stk = ( stk << 1 ) | ( unsigned int )( ( vars >> ( ins >> 4 ) ) & 1ull );
            else if( op == E_CLIP ) stk &= ~1u | ( unsigned int )( ( vars >> ( ins >> 4 ) ) & 1ull );
            else if( op == E_NEG )  stk ^= 1u;
            else if( op == E_AND )  { const unsigned int t = stk & 1u; stk >>= 1; stk &= t | ~1u; }
            else if( op == E_OR )   { const unsigned int t = stk & 1u; stk >>= 1; stk |= t; }
        }
        return ( int )( stk & 1u );
    }

    void visit_compound( int c, int guard, bool enable )
    {
        if( guard > 64 ) return;
        const acn_flat_node& cn = fs->nodes[ c ];
        for( int i = 0; i < cn.child1; i++ )
        {
            const int n = fs->children[ cn.child0 + i ];
            const acn_flat_node& nd = fs->nodes[ n ];
            if( nd.kind == ACN_KIND_COMPOUND ) { visit_compound( n, guard + 1, enable ); continue; }
            set_parents( n, 0 );
            // the bit stack of the interpreter holds 32 levels; a left-deep chain needs 2 however long it is
            const char* skip = getenv( "ACN_NO_SWEEP_NODE" );       // diagnostics: keep one object on the reference march
            if( skip && atoi( skip ) == n ) continue;
            if( enable && ( is_pair( nd.kind ) || nd.kind == ACN_KIND_NEG || nd.kind == ACN_KIND_SCALE ) && eligible( n, 0 ) && depth( n, 0 ) < 30 && prog_ref[ n ].y == 0 )
            {
                const size_t mark = prog.size();
                n_vars = 0;
                emit( n, true );
                const size_t len = prog.size() - mark;
                if( len < ( size_t )CSG_VIRTUAL && n_vars <= CSG_MAX_VARS )     // crossing ids are one byte: program-relative leaf offsets
                {
                    I4 r; r.x = ( int )mark; r.y = ( int )len; r.z = -1; r.w = n_vars;
                    if( n_vars <= CSG_TABLE_VARS )
                    {
                        r.z = ( int )prog.size();
                        const size_t rows = ( size_t )1 << n_vars;
                        std::vector<int> tab( ( rows + 31 ) / 32, 0 );
                        for( size_t a = 0; a < rows; a++ ) if( eval( mark, len, a ) ) tab[ a >> 5 ] |= ( int )( 1u << ( a & 31 ) );
                        prog.insert( prog.end(), tab.begin(), tab.end() );
                    }
                    else if( !getenv( "ACN_NO_PACKED_EVAL" ) )
                    {
                        const int at = build_eval_program( mark, len );
                        if( at >= 0 ) r.z = -2 - at;             // z <= -2: evaluation program at prog[ -z - 2 ]
                    }
                    prog_ref[ n ] = r;
                    if( coincident_leaves( n ) ) has_coincident = true;
                }
                else
                {
                    if( getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: object %d: program of %d words, %d variables exceeds the sweep's limits\n", n, ( int )len, n_vars );
                    prog.resize( mark );
                }
            }
        }
    }

    void build( const acn_flat_scene* scene, bool enable )
    {
        fs = scene;
        prog.clear();
        I4 none; none.x = 0; none.y = 0; none.z = -1; none.w = 0;
        prog_ref.assign( ( size_t )fs->n_nodes, none );
        parent.assign( fs->n_nodes, -1 );
        visit_compound( fs->light_root, 0, enable );
        visit_compound( fs->matter_root, 0, enable );
        if( prog.empty() ) prog.push_back( 0 );
    }
};

// diagnostics (no device needed): variables / words of every event-sweep program of a scene
static void dump_programs( const acn_flat_scene* fs )
{
    CsgBuilder d; d.build( fs, true );
    for( int i = 0; i < fs->n_nodes; i++ )
        if( d.prog_ref[ i ].y > 0 )
        {
            fprintf( stderr, "acn: program of node %d: %d variables, %d words, %s", i, d.prog_ref[ i ].w, d.prog_ref[ i ].y, d.prog_ref[ i ].z >= 0 ? "table\n" : d.prog_ref[ i ].z == -1 ? "interpreted\n" : "packed: " );
            if( d.prog_ref[ i ].z <= -2 ) fprintf( stderr, "%d evaluation words\n", d.prog[ -2 - d.prog_ref[ i ].z ] );
            if( d.prog_ref[ i ].z == -1 && atoi( getenv( "ACN_DUMP_PROGRAMS" ) ) > 1 )
            {
                static const char* nm[] = { "LEAF", "NEG", "AND", "OR", "CLIP", "ENV", "RUN", "MEMBER", "MEMBER_NEG", "MORE", "XFORM", "XEND" };
                for( int pc = d.prog_ref[ i ].x; pc < d.prog_ref[ i ].x + d.prog_ref[ i ].y; pc++ )
                {
                    const int ins = d.prog[ pc ], op = ins & 15;
                    fprintf( stderr, " %s:%d", op < 12 ? nm[ op ] : "?", ins >> 4 );
                    if( op == CSG_ENV ) { pc++; fprintf( stderr, "(skip %d vars %d)", d.prog[ pc ] & 0xFFFF, d.prog[ pc ] >> 16 ); }
                }
                fprintf( stderr, "\n" );
            }
        }
}

int validate_flat_scene( const acn_flat_scene* fs );   // acn_tracer.cu

// does the scene need the full-featured (MARCH) kernel instantiation?  A top-level object the event sweep does not cover
// (scale nodes, oversized programs) runs the reference's recursive march; distance-field leaves and groups of coincident
// crossings are compiled into that instantiation only.
static bool needs_march( const acn_flat_scene* fs, const CsgBuilder& cb )
{
    bool march = false;
    std::function<void( int )> walk = [ & ]( int c )
    {
        const acn_flat_node& cn = fs->nodes[ c ];
        for( int i = 0; i < cn.child1; i++ )
        {
            const int e = fs->children[ cn.child0 + i ];
            const acn_flat_node& nd = fs->nodes[ e ];
            if( nd.kind == ACN_KIND_COMPOUND ) { walk( e ); continue; }
            if( nd.kind >= ACN_KIND_PAIR_INSIDE && cb.prog_ref[ e ].y == 0 ) march = true;
        }
    };
    walk( fs->light_root ); walk( fs->matter_root );
    return march || cb.has_dist_leaf || cb.has_coincident || cb.has_scale;
}

// composite objects the event sweep does not cover (programs of more than CSG_MAX_VARS variables or 254 words, nested
// scale nodes): the FP64 validation mode runs the reference's recursive march on them, the FP32 tracer refuses the scene
static int first_unswept( const acn_flat_scene* fs, const CsgBuilder& cb )
{
    int bad = -1;
    std::function<void( int )> walk = [ & ]( int c )
    {
        const acn_flat_node& cn = fs->nodes[ c ];
        for( int i = 0; i < cn.child1 && bad < 0; i++ )
        {
            const int e = fs->children[ cn.child0 + i ];
            const acn_flat_node& nd = fs->nodes[ e ];
            if( nd.kind == ACN_KIND_COMPOUND ) { walk( e ); continue; }
            if( nd.kind >= ACN_KIND_PAIR_INSIDE && cb.prog_ref[ e ].y == 0 ) bad = e;
        }
    };
    walk( fs->light_root ); if( bad < 0 ) walk( fs->matter_root );
    return bad;
}

// number of traversal records: one per element of every child list reached from the two roots (a compound that two
// lists share is laid out once per reference — the records form a TREE, see ThreadedRecords); capped
static size_t threaded_record_count( const acn_flat_scene* fs, size_t cap = ( size_t )1 << 26 )
{
    size_t n = 0;
    std::function<void( int, int )> walk = [ & ]( int c, int guard )
    {
        const acn_flat_node& cn = fs->nodes[ c ];
        n += ( size_t )cn.child1;
        if( n > cap || guard > 64 ) return;
        for( int i = 0; i < cn.child1; i++ )
        {
            const int e = fs->children[ cn.child0 + i ];
            if( fs->nodes[ e ].kind == ACN_KIND_COMPOUND ) walk( e, guard + 1 );
        }
    };
    walk( fs->light_root, 0 ); walk( fs->matter_root, 0 );
    return n;
}

// bytes of the node table as staged into shared memory (layout in Tracer::init)
template <typename R> static size_t staged_table_bytes( const acn_flat_scene* fs, size_t n_prog )
{
    const size_t n = ( size_t )fs->n_nodes;
    return n * ( sizeof( R4<R> ) * ( 1 + GEO_STRIDE ) + 2 * sizeof( I4 ) + sizeof( int ) ) + threaded_record_count( fs ) * sizeof( CRec<R> ) + n_prog * sizeof( int );
}
template <typename R> static bool stages_table( const acn_flat_scene* fs, size_t n_prog )
{
    return sizeof( R ) == 4 && staged_table_bytes<R>( fs, n_prog ) <= 96 * 1024 && !getenv( "ACN_NO_STAGING" );
}

// the source SpecGen writes for a scene ("" when the scene does not qualify) and the instantiation it belongs to
struct SpecPlan { std::string src; bool march = false, sh = false; };
template <typename R> static SpecPlan plan_spec( const acn_flat_scene* fs, const CsgBuilder& cb )
{
    SpecPlan pl;
    pl.march = needs_march( fs, cb );
    pl.sh = stages_table<R>( fs, cb.prog.size() );
    SpecGen g; g.fs = fs; g.prog = &cb.prog; g.prog_ref = &cb.prog_ref;
    const CullBounds cbnd( fs, true );
    std::vector<char> gate( ( size_t )fs->n_nodes, 0 );
    for( int i = 0; i < fs->n_nodes; i++ ) gate[ i ] = cbnd.rec[ i ].env_too ? 1 : 0;
    g.env_gate_only = &gate;
    if( g.generate() ) pl.src = g.out;
    return pl;
}

template <typename R> int Tracer<R>::init( const acn_flat_scene* fs, const acn_options* opt )
{
    int rc = validate_flat_scene( fs );
    if( rc ) return rc;
    const acn_flat_params& p = fs->params;
    width = p.image_width; height = p.image_height;
    const int n = fs->n_nodes;
    // ---- pack the node table
    std::vector<R4<R>> env( n ), geo( ( size_t )n * GEO_STRIDE );
    std::vector<I4> link( n );
    double scale = 0;
    for( int k = 0; k < 3; k++ ) scale = fmax( scale, fabs( p.camera_position[ k ] ) );
    for( int i = 0; i < n; i++ )
    {
        const acn_flat_node& nd = fs->nodes[ i ];
        int flags = 0;
        if( nd.has_envelope ) flags |= F_ENV;
        if( nd.surface_roughness > 0 && nd.kind != ACN_KIND_COMPOUND ) flags |= F_ROUGH;
        env[ i ].x = ( R )nd.env_pos[ 0 ]; env[ i ].y = ( R )nd.env_pos[ 1 ]; env[ i ].z = ( R )nd.env_pos[ 2 ];
        env[ i ].w = nd.has_envelope ? ( R )nd.env_radius : ( R )-1;
        link[ i ].x = nd.kind | ( flags << 8 ); link[ i ].y = nd.child0; link[ i ].z = nd.child1; link[ i ].w = nd.material;
        R4<R>* g = &geo[ ( size_t )i * GEO_STRIDE ];
        g[ 0 ].x = ( R )nd.pos[ 0 ]; g[ 0 ].y = ( R )nd.pos[ 1 ]; g[ 0 ].z = ( R )nd.pos[ 2 ]; g[ 0 ].w = ( R )nd.tail[ 0 ];
        for( int r = 0; r < 3; r++ )
        {
            g[ 1 + r ].x = ( R )nd.rax[ 3 * r + 0 ]; g[ 1 + r ].y = ( R )nd.rax[ 3 * r + 1 ]; g[ 1 + r ].z = ( R )nd.rax[ 3 * r + 2 ];
            g[ 1 + r ].w = ( R )nd.tail[ 1 + r ];
        }
        g[ 4 ].x = ( R )nd.surface_roughness; g[ 4 ].y = g[ 4 ].z = g[ 4 ].w = ( R )0;
        if( nd.kind != ACN_KIND_COMPOUND && nd.kind != ACN_KIND_PLANE )
            for( int k = 0; k < 3; k++ ) scale = fmax( scale, fabs( nd.pos[ k ] ) );
    }

    // ---- shell thickness: the reference's absolute 1e-6 (f64 and as the f32 floor); the f32 path adds a
    // per-ray term of 16 ulp of the ray origin (ray_view) and of the hit distance (trace_ray)
    double eps = opt->eps, eps_rel = 0;
    if( !( eps > 0 ) )
    {
        eps = 1E-6;
        if( sizeof( R ) == 4 ) eps_rel = 16.0 * 1.1920929E-7;
        if( sizeof( R ) == 4 && getenv( "ACN_EPS_ULPS" ) ) eps_rel = atof( getenv( "ACN_EPS_ULPS" ) ) * 1.1920929E-7;     // experiments (tools/c1_probe.py)
    }
    ( void )scale;

    cudaError_t ce = cudaSetDevice( device );
    if( ce != cudaSuccess ) { set_error( "cudaSetDevice(%d): %s", device, cudaGetErrorString( ce ) ); return ACN_ERR_NO_DEVICE; }
    ACN_CUDA( cudaStreamCreateWithFlags( &own_stream, cudaStreamNonBlocking ) );
    ACN_CUDA( cudaStreamCreateWithFlags( &side_stream, cudaStreamNonBlocking ) );
    ACN_CUDA( cudaEventCreateWithFlags( &ev_sched, cudaEventDisableTiming ) );
    ACN_CUDA( cudaEventCreateWithFlags( &ev_path, cudaEventDisableTiming ) );
    ACN_CUDA( cudaEventCreateWithFlags( &ev_index, cudaEventDisableTiming ) );
    ACN_CUDA( cudaEventCreateWithFlags( &ev_direct, cudaEventDisableTiming ) );
    ACN_CUDA( cudaEventCreate( &ev_t0 ) ); ACN_CUDA( cudaEventCreate( &ev_t1 ) );

    if( ( rc = dev_alloc( &d_env, n ) ) ) return rc;
    if( ( rc = dev_alloc( &d_link, n ) ) ) return rc;
    if( ( rc = dev_alloc( &d_geo, ( size_t )n * GEO_STRIDE ) ) ) return rc;
    if( ( rc = dev_alloc( &d_children, ( size_t )( fs->n_children > 0 ? fs->n_children : 1 ) ) ) ) return rc;
    ACN_CUDA( cudaMemcpy( d_env, env.data(), n * sizeof( R4<R> ), cudaMemcpyHostToDevice ) );
    ACN_CUDA( cudaMemcpy( d_link, link.data(), n * sizeof( I4 ), cudaMemcpyHostToDevice ) );
    ACN_CUDA( cudaMemcpy( d_geo, geo.data(), ( size_t )n * GEO_STRIDE * sizeof( R4<R> ), cudaMemcpyHostToDevice ) );
    if( fs->n_children > 0 ) ACN_CUDA( cudaMemcpy( d_children, fs->children, fs->n_children * sizeof( int ), cudaMemcpyHostToDevice ) );
    {   // ---- traversal records.  The child lists of the two root compounds as ONE array of packed records
        // ( cull bound | kind + flags, first record of the child list, SKIP record, node ), threaded: a ray that fails an
        // element's bound (or is through with a leaf) goes to the element's skip record — its next sibling, or, behind
        // the last element of a list, whatever follows the list's compound — and a ray that passes a compound's bound goes
        // to the first record of its list.  The walk is a single index per ray: no stack, no "back to the parent list"
        // steps, and a traversal can be parked and resumed with four bytes of state (acn_kernels.cuh: refill loops).
        // A compound referenced from two lists gets two copies of its list.
        n_rec = threaded_record_count( fs );
        if( n_rec > ( ( size_t )1 << 26 ) ) { set_error( "flat scene: more than 2^26 traversal records (shared compounds expand into a tree)" ); return ACN_ERR_UNSUPPORTED; }
        std::vector<CRec<R>> crec( n_rec > 0 ? n_rec : 1 );
        const CullBounds cbnd( fs, !getenv( "ACN_NO_TIGHT_BOUNDS" ) );
        // PRE-ORDER: an element's record is followed by the records of everything it contains, so first-child and skip
        // links both point FORWARD.  The lockstep walk of a warp (acn_isect.cuh: scene_query) relies on that: it always
        // advances the lanes at the lowest record, which keeps lanes that skipped a subtree waiting for the others at the
        // next common record instead of running ahead through different objects.
        size_t next_free = 0;
        const int SAME_AS_SKIP = -4;
        // FRONT TO BACK.  For a scene of planes, spheres and quadrics whose tables stay in global memory (many_spheres) the
        // matter records exist eight more times, once per octant of ray directions, the elements of every nested list sorted
        // along the octant's diagonal: a ray walks the copy of its octant and meets near elements first — a closest-hit search
        // tightens its horizon early, an any-hit search ends early.  Inside a nested compound the order of the tests does not
        // matter (plain minimum; two hits at EXACTLY the same distance, which only coincident geometry produces, may then
        // name the other object); the order of the root's own elements, whose coincident surfaces are merged in sequence
        // (compound.c:246-299), is never changed.
        const bool grouping = staged_table_bytes<R>( fs, 0 ) > 96 * 1024 && !getenv( "ACN_NO_GROUP_RECORDS" );
        int n_groups = 0;
        auto alloc_rec = [ & ]() -> int
        {
            if( next_free >= crec.size() ) crec.resize( crec.size() * 2 + 16 );
            return ( int )next_free++;
        };
        // the bound a record tests first (with the horizon): the tight ball of CullBounds or the reference's envelope
        auto first_bound = [ & ]( int c, CullBounds::Ball* out ) -> bool
        {
            const acn_flat_node& nd = fs->nodes[ c ];
            const CullBounds::Rec& b = cbnd.rec[ c ];
            out->ok = true;
            if( b.mode != CullBounds::KEEP ) { for( int j = 0; j < 3; j++ ) out->c[ j ] = b.c[ j ]; out->r = b.r; return true; }
            if( nd.has_envelope ) { for( int j = 0; j < 3; j++ ) out->c[ j ] = nd.env_pos[ j ]; out->r = nd.env_radius; return true; }
            out->ok = false;
            return false;
        };
        int oct = -1;
        auto front_key = [ & ]( int c ) -> double
        {
            const acn_flat_node& nd = fs->nodes[ c ];
            const double* p = cbnd.ball[ c ].ok ? cbnd.ball[ c ].c : nd.has_envelope ? nd.env_pos : nd.pos;
            if( nd.kind == ACN_KIND_PLANE ) return -1e300;
            return ( ( oct & 1 ) ? -p[ 0 ] : p[ 0 ] ) + ( ( oct & 2 ) ? -p[ 1 ] : p[ 1 ] ) + ( ( oct & 4 ) ? -p[ 2 ] : p[ 2 ] );
        };
        std::function<int( int, bool, std::vector<int>& )> emit = [ & ]( int compound, bool top, std::vector<int>& open ) -> int
        {   // returns the first record of the list (-1: empty); `open` collects the records whose skip is the escape of this list
            const acn_flat_node& cn = fs->nodes[ compound ];
            int first = -1;
            std::vector<int> wait;              // records whose skip is the next record of this list
            std::vector<int> kids( fs->children + cn.child0, fs->children + cn.child0 + cn.child1 );
            if( oct >= 0 && !top )
                std::stable_sort( kids.begin(), kids.end(), [ & ]( int a, int b ) { return front_key( a ) < front_key( b ); } );
            // GROUP records over runs of consecutive elements of a long list (a lamp holds 61 objects): a pure bound, the ball round
            // the members' own first bounds.  A ray that misses it (or enters it beyond its horizon) would fail every member's test
            // one by one; it skips them in one step.  The members and the order of their tests are unchanged.
            std::vector<int> group_len( ( size_t )cn.child1, 0 );        // at the first member of a group: its length
            std::vector<CullBounds::Ball> group_ball( ( size_t )cn.child1 );
            if( grouping && cn.child1 > 12 )
            {
                for( int i = 0; i < cn.child1; )
                {
                    CullBounds::Ball b0;
                    if( !first_bound( kids[ i ], &b0 ) ) { i++; continue; }
                    std::vector<CullBounds::Ball> mem = { b0 };
                    CullBounds::Ball gb = b0;
                    double maxr = b0.r;
                    int j = i + 1;
                    for( ; j < cn.child1 && ( int )mem.size() < 8; j++ )
                    {
                        CullBounds::Ball bj;
                        if( !first_bound( kids[ j ], &bj ) ) break;
                        mem.push_back( bj );
                        const CullBounds::Ball nb = CullBounds::enclose( mem );
                        if( nb.r > 3.0 * fmax( maxr, bj.r ) ) { mem.pop_back(); break; }
                        gb = nb; maxr = fmax( maxr, bj.r );
                    }
                    if( mem.size() >= 3 ) { group_len[ i ] = ( int )mem.size(); group_ball[ i ] = gb; n_groups++; i += ( int )mem.size(); }
                    else i++;
                }
            }
            int group_rec = -1, group_left = 0;                           // the open group: its record, members still to come
            for( int i = 0; i < cn.child1; i++ )
            {
                const int c = kids[ i ];
                if( group_len[ i ] > 0 )
                {
                    const int g = alloc_rec();
                    for( int r : wait ) crec[ r ].link.z = g;
                    wait.clear();
                    if( first < 0 ) first = g;
                    const CullBounds::Ball& gb = group_ball[ i ];
                    const double m = CullBounds::margin( gb.c, gb.r );
                    CRec<R>& gr = crec[ g ];
                    gr.env.x = ( R )gb.c[ 0 ]; gr.env.y = ( R )gb.c[ 1 ]; gr.env.z = ( R )gb.c[ 2 ]; gr.env.w = ( R )( gb.r + m );
                    gr.link.x = K_GROUP | ( ( F_ENV | ( top ? F_TOP : 0 ) ) << 8 );
                    gr.link.y = g + 1; gr.link.z = -1; gr.link.w = -1;
                    group_rec = g; group_left = group_len[ i ];
                }
                const int idx = alloc_rec();
                for( int r : wait ) crec[ r ].link.z = idx;
                wait.clear();
                if( first < 0 ) first = idx;
                CRec<R>& r = crec[ idx ];
                r.env = env[ c ];
                int flags = node_flags( link[ c ] );
                const CullBounds::Rec& b = cbnd.rec[ c ];
                if( b.mode != CullBounds::KEEP )
                {
                    r.env.x = ( R )b.c[ 0 ]; r.env.y = ( R )b.c[ 1 ]; r.env.z = ( R )b.c[ 2 ]; r.env.w = ( R )b.r;
                    flags = ( flags & ~F_ENV ) | ( b.mode == CullBounds::SELF ? F_SELF : F_ENV );
                    if( b.env_too ) flags |= F_ENV2;
                }
                if( top ) flags |= F_TOP;
                r.link.x = node_kind( link[ c ] ) | ( flags << 8 );
                r.link.y = -1; r.link.z = -1; r.link.w = c;
                wait.push_back( idx );
                if( fs->nodes[ c ].kind == ACN_KIND_COMPOUND )
                {
                    std::vector<int> sub;
                    const int f = emit( c, false, sub );
                    crec[ idx ].link.y = f >= 0 ? f : SAME_AS_SKIP;
                    wait.insert( wait.end(), sub.begin(), sub.end() );
                }
                if( group_rec >= 0 && --group_left == 0 ) { wait.push_back( group_rec ); group_rec = -1; }     // the group's skip: whatever follows its last member
            }
            open.insert( open.end(), wait.begin(), wait.end() );
            return first;
        };
        {
            std::vector<int> open;
            rec_light = emit( fs->light_root, true, open );
            for( int r : open ) crec[ r ].link.z = -1;
            open.clear();
            rec_matter = emit( fs->matter_root, true, open );
            for( int r : open ) crec[ r ].link.z = -1;
            for( int k = 0; k < 8; k++ ) rec_matter_oct[ k ] = rec_matter;
            bool prims = true;
            for( int i = 0; i < n && prims; i++ ) prims = fs->nodes[ i ].kind <= ACN_KIND_SQUAROID;
            const size_t n_matter = next_free - ( size_t )( rec_matter >= 0 ? rec_matter : ( int )next_free );
            if( prims && n_matter > 64 && n_matter * 8 * sizeof( CRec<R> ) <= ( ( size_t )32 << 20 ) && staged_table_bytes<R>( fs, 0 ) > 96 * 1024 && !getenv( "ACN_NO_OCTANT_ORDER" ) )
            {
                crec.resize( next_free + 8 * n_matter );
                for( oct = 0; oct < 8; oct++ )
                {
                    open.clear();
                    rec_matter_oct[ oct ] = emit( fs->matter_root, true, open );
                    for( int r : open ) crec[ r ].link.z = -1;
                }
                oct = -1;
                n_rec = next_free;
            }
            for( size_t i = 0; i < next_free; i++ ) if( crec[ i ].link.y == SAME_AS_SKIP ) crec[ i ].link.y = crec[ i ].link.z;
            crec.resize( next_free > 0 ? next_free : 1 );
            n_rec = next_free;
        }
        if( getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: %zu traversal records%s, %d group records: %d compounds with a tight cull bound, %d spheres held in their record, %d of both test the reference's envelope as well\n", n_rec, rec_matter_oct[ 0 ] != rec_matter ? " (matter list eight more times, front to back per octant)" : "", n_groups, cbnd.n_tight, cbnd.n_self, cbnd.n_both );
        if( ( rc = dev_alloc( &d_crec, crec.size() ) ) ) return rc;
        ACN_CUDA( cudaMemcpy( d_crec, crec.data(), crec.size() * sizeof( CRec<R> ), cudaMemcpyHostToDevice ) );
    }

    // ---- CSG interval programs
    CsgBuilder cb;
    {
        const bool fast = opt->csg_mode == ACN_CSG_INTERVALS || ( opt->csg_mode == ACN_CSG_AUTO && sizeof( R ) == 4 );
        cb.build( fs, fast );
        n_prog = ( int )cb.prog.size();
        if( ( rc = dev_alloc( &d_prog, cb.prog.size() ) ) ) return rc;
        if( ( rc = dev_alloc( &d_prog_ref, cb.prog_ref.size() ) ) ) return rc;
        if( ( rc = dev_alloc( &d_parent, cb.parent.size() ) ) ) return rc;
        ACN_CUDA( cudaMemcpy( d_prog, cb.prog.data(), cb.prog.size() * sizeof( int ), cudaMemcpyHostToDevice ) );
        ACN_CUDA( cudaMemcpy( d_prog_ref, cb.prog_ref.data(), cb.prog_ref.size() * sizeof( I4 ), cudaMemcpyHostToDevice ) );
        ACN_CUDA( cudaMemcpy( d_parent, cb.parent.data(), cb.parent.size() * sizeof( int ), cudaMemcpyHostToDevice ) );
        march = needs_march( fs, cb );
        if( sizeof( R ) == 4 )
        {   // the FP32 product path has no recursive march: every composite object is swept, or the scene is refused
            const int bad = first_unswept( fs, cb );
            if( bad >= 0 )
            {
                set_error( fast ? "object %d (kind %d) exceeds what the CSG event sweep covers (%d variables, 254 program words, one level of scale nodes); "
                                  "the reference's recursive march runs in the FP64 validation mode only"
                                : "csg_mode MARCH (object %d, kind %d): the reference's recursive march runs in the FP64 validation mode only (limit %d)",
                           bad, fs->nodes[ bad ].kind, ( int )CSG_MAX_VARS );
                return ACN_ERR_UNSUPPORTED;
            }
        }
        if( getenv( "ACN_VERBOSE" ) )
        {
            int swept = 0, marched = 0;
            std::function<void( int )> cnt = [ & ]( int c )
            {
                const acn_flat_node& cn = fs->nodes[ c ];
                for( int i = 0; i < cn.child1; i++ )
                {
                    const int e = fs->children[ cn.child0 + i ];
                    const acn_flat_node& nd = fs->nodes[ e ];
                    if( nd.kind == ACN_KIND_COMPOUND ) { cnt( e ); continue; }
                    if( nd.kind < ACN_KIND_PAIR_INSIDE ) continue;
                    if( cb.prog_ref[ e ].y > 0 ) swept++;
                    else { marched++; fprintf( stderr, "acn: object %d (kind %d) runs the reference march\n", e, nd.kind ); }
                }
            };
            cnt( fs->light_root ); cnt( fs->matter_root );
            fprintf( stderr, "acn: %d composite objects swept, %d marched; program words %d; dist leaves %d, coincident surfaces %d, scale nodes %d -> %s kernels\n",
                     swept, marched, n_prog, ( int )cb.has_dist_leaf, ( int )cb.has_coincident, ( int )cb.has_scale, march ? "full-featured" : "lean" );
        }
    }

    // ---- materials
    std::vector<DMat<R>> mats( fs->n_materials > 0 ? fs->n_materials : 1 );
    for( int i = 0; i < fs->n_materials; i++ )
    {
        const acn_flat_material& fm = fs->materials[ i ];
        DMat<R>& m = mats[ i ];
        for( int k = 0; k < 3; k++ ) { m.color[ k ] = ( R )fm.color[ k ]; m.transp[ k ] = ( R )fm.transparency[ k ]; m.tex1[ k ] = ( R )fm.tex_color1[ k ]; m.tex2[ k ] = ( R )fm.tex_color2[ k ]; }
        m.radiance = ( R )fm.radiance; m.refr = ( R )fm.refractive_index;
        m.fresnel01 = ( fm.fresnel_reflectivity != 0 && fm.refractive_index != 1.0 ) ? ( R )1 : ( R )0;
        m.chroma = ( R )fm.chromatic_reflectivity; m.diffuse = ( R )fm.diffuse_reflectivity;
        double oa = 1, ob = 0;
        if( fm.sigma > 0 ) { double s2 = fm.sigma * fm.sigma; oa = 1.0 - 0.5 * s2 / ( s2 + 0.33 ); ob = 0.45 * s2 / ( s2 + 0.09 ); }
        m.on_a = ( R )oa; m.on_b = ( R )ob;
        m.transparent = ( fm.transparency[ 0 ] * fm.transparency[ 0 ] + fm.transparency[ 1 ] * fm.transparency[ 1 ] + fm.transparency[ 2 ] * fm.transparency[ 2 ] ) > 0;
        m.tex_kind = fm.texture_kind; m.tex_scale = ( R )fm.tex_scale;
    }
    if( ( rc = dev_alloc( &d_mats, mats.size() ) ) ) return rc;
    ACN_CUDA( cudaMemcpy( d_mats, mats.data(), mats.size() * sizeof( DMat<R> ), cudaMemcpyHostToDevice ) );

    // ---- lights = elements of the light compound (scene.c:542-552)
    const acn_flat_node& lroot = fs->nodes[ fs->light_root ];
    std::vector<DLight<R>> lights( lroot.child1 > 0 ? lroot.child1 : 1 );
    {
        // host view (double) for evaluating obj_color( light, light.pos ) once
        std::vector<R4<double>> henv( n ), hgeo( ( size_t )n * GEO_STRIDE );
        for( int i = 0; i < n; i++ )
        {
            henv[ i ].x = env[ i ].x; henv[ i ].y = env[ i ].y; henv[ i ].z = env[ i ].z; henv[ i ].w = env[ i ].w;
            for( int k = 0; k < GEO_STRIDE; k++ )
            {
                const R4<R>& s = geo[ ( size_t )i * GEO_STRIDE + k ]; R4<double>& d = hgeo[ ( size_t )i * GEO_STRIDE + k ];
                d.x = s.x; d.y = s.y; d.z = s.z; d.w = s.w;
            }
        }
        SceneView<double> hv; hv.env = henv.data(); hv.geo = hgeo.data(); hv.link = link.data(); hv.children = fs->children;
        hv.prog = nullptr; hv.prog_ref = nullptr; hv.parent = nullptr;
        hv.eps = eps; hv.light_root = fs->light_root; hv.matter_root = fs->matter_root; hv.seed_mode = 0; hv.rec_light = hv.rec_matter = -1; for( int k = 0; k < 8; k++ ) hv.rec_matter_oct[ k ] = -1;
        for( int i = 0; i < lroot.child1; i++ )
        {
            const int ln = fs->children[ lroot.child0 + i ];
            const acn_flat_node& nd = fs->nodes[ ln ];
            if( !( nd.kind == ACN_KIND_SPHERE || nd.kind == ACN_KIND_PLANE || nd.kind == ACN_KIND_PAIR_INSIDE || nd.kind == ACN_KIND_PAIR_OUTSIDE ) )
            {
                set_error( "light %d (node %d, kind %d) has no fov function (reference objects.c:254-259 would abort)", i, ln, nd.kind );
                return ACN_ERR_UNSUPPORTED;
            }
            const acn_flat_material& fm = fs->materials[ nd.material ];
            DLight<R>& lg = lights[ i ];
            lg.node = ln; lg.radiance = ( R )fm.radiance;
            double col[ 3 ] = { fm.color[ 0 ], fm.color[ 1 ], fm.color[ 2 ] };
            if( fm.texture_kind == ACN_TEX_PLAIN ) { for( int k = 0; k < 3; k++ ) col[ k ] = fm.tex_color1[ k ]; }
            else if( fm.texture_kind == ACN_TEX_CHESS )
            {
                double u, v;
                obj_projection<double>( hv, ln, v3<double>( nd.pos[ 0 ], nd.pos[ 1 ], nd.pos[ 2 ] ), &u, &v );
                long long x = llrint( u * fm.tex_scale ), y = llrint( v * fm.tex_scale );
                for( int k = 0; k < 3; k++ ) col[ k ] = ( ( x ^ y ) & 1 ) ? fm.tex_color1[ k ] : fm.tex_color2[ k ];
            }
            for( int k = 0; k < 3; k++ ) { lg.pos[ k ] = ( R )nd.pos[ k ]; lg.color[ k ] = ( R )col[ k ]; }
        }
    }
    if( ( rc = dev_alloc( &d_lights, lights.size() ) ) ) return rc;
    ACN_CUDA( cudaMemcpy( d_lights, lights.data(), lights.size() * sizeof( DLight<R> ), cudaMemcpyHostToDevice ) );

    // ---- LCG skip table: state after 2k steps
    const int skip_n = 1 << 14;
    {
        std::vector<u64> A( skip_n ), C( skip_n );
        u64 a2 = ACN_LCG00_A * ACN_LCG00_A, c2 = ACN_LCG00_C * ( ACN_LCG00_A + 1 );
        u64 a = 1, c = 0;
        for( int k = 0; k < skip_n; k++ ) { A[ k ] = a; C[ k ] = c; c = c * a2 + c2; a *= a2; }
        if( ( rc = dev_alloc( &d_skipA, skip_n ) ) ) return rc;
        if( ( rc = dev_alloc( &d_skipC, skip_n ) ) ) return rc;
        ACN_CUDA( cudaMemcpy( d_skipA, A.data(), skip_n * sizeof( u64 ), cudaMemcpyHostToDevice ) );
        ACN_CUDA( cudaMemcpy( d_skipC, C.data(), skip_n * sizeof( u64 ), cudaMemcpyHostToDevice ) );
    }

    // ---- params
    prm.sv.env = d_env; prm.sv.link = d_link; prm.sv.geo = d_geo; prm.sv.children = d_children; prm.sv.crec = d_crec;
    prm.sv.prog = d_prog; prm.sv.prog_ref = d_prog_ref; prm.sv.parent = d_parent; prm.n_prog = n_prog;
    prm.sv.eps = ( R )eps; prm.sv.light_root = fs->light_root; prm.sv.matter_root = fs->matter_root;
    prm.sv.rec_light = rec_light; prm.sv.rec_matter = rec_matter;
    for( int k = 0; k < 8; k++ ) prm.sv.rec_matter_oct[ k ] = rec_matter_oct[ k ];
    prm.sv.seed_mode = opt->seed_mode;
    prm.mats = d_mats; prm.lights = d_lights; prm.n_lights = lroot.child1; prm.n_materials = fs->n_materials;
    prm.n_nodes = n; prm.n_children = ( int )n_rec;
    prm.width = p.image_width; prm.height = p.image_height;
    prm.gamma = ( R )p.gamma;
    prm.background = v3<R>( ( R )p.background_color[ 0 ], ( R )p.background_color[ 1 ], ( R )p.background_color[ 2 ] );
    {   // camera_rotation (scene.c:963-973), in double on the host
        V3<double> view = v3<double>( p.camera_view_direction[ 0 ], p.camera_view_direction[ 1 ], p.camera_view_direction[ 2 ] );
        V3<double> top  = v3<double>( p.camera_top_direction[ 0 ], p.camera_top_direction[ 1 ], p.camera_top_direction[ 2 ] );
        V3<double> ry = unit( view );
        V3<double> rz = von( ry, unit( top ) );
        V3<double> rx = cross( ry, rz );
        prm.cam_rx = v3<R>( ( R )rx.x, ( R )rx.y, ( R )rx.z );
        prm.cam_ry = v3<R>( ( R )ry.x, ( R )ry.y, ( R )ry.z );
        prm.cam_rz = v3<R>( ( R )rz.x, ( R )rz.y, ( R )rz.z );
        prm.cam_pos = v3<R>( ( R )p.camera_position[ 0 ], ( R )p.camera_position[ 1 ], ( R )p.camera_position[ 2 ] );
    }
    prm.focal = ( R )p.camera_focal_length;
    prm.trace_depth = p.trace_depth > 255 ? 255 : p.trace_depth;
    prm.direct_samples = p.direct_samples; prm.path_samples = p.path_samples;
    prm.min_intensity = ( R )p.trace_min_intensity;
    prm.max_path_length = ( R )p.max_path_length;
    prm.eps_rel = ( R )eps_rel;
    prm.skipA = d_skipA; prm.skipC = d_skipC; prm.skip_n = skip_n;

    // expensive top-level objects for the warp-level compaction: everything that is not a plane/sphere/squaroid
    {
        prm.n_heavy = 0;
        std::vector<int> hv;
        bool open = false;      // an expensive object without envelope: every ray is expensive, nothing to split
        std::function<void( int )> walk = [ & ]( int c )
        {
            const acn_flat_node& cn = fs->nodes[ c ];
            for( int i = 0; i < cn.child1; i++ )
            {
                const int e = fs->children[ cn.child0 + i ];
                const acn_flat_node& nd = fs->nodes[ e ];
                if( nd.kind == ACN_KIND_COMPOUND ) { walk( e ); continue; }
                if( nd.kind == ACN_KIND_PLANE || nd.kind == ACN_KIND_SPHERE || nd.kind == ACN_KIND_SQUAROID ) continue;
                if( nd.has_envelope ) hv.push_back( e ); else open = true;
            }
        };
        walk( fs->light_root ); walk( fs->matter_root );
        if( !open && !hv.empty() && hv.size() <= 8 && !getenv( "ACN_NO_COMPACTION" ) )
        {
            prm.n_heavy = ( int )hv.size();
            for( size_t i = 0; i < hv.size(); i++ ) prm.heavy[ i ] = hv[ i ];
        }
    }

    // shared-memory staging of the node table
    const size_t table = staged_table_bytes<R>( fs, ( size_t )n_prog );
    {   // layout: env | geo | link | prog_ref | crec | parent | prog — every table starts 16-byte aligned (32 for the R4<double> ones by construction)
        size_t o = ( size_t )n * sizeof( R4<R> );
        prm.off_geo  = ( unsigned int )o; o += ( size_t )n * GEO_STRIDE * sizeof( R4<R> );
        prm.off_link = ( unsigned int )o; o += ( size_t )n * sizeof( I4 );
        prm.off_pref = ( unsigned int )o; o += ( size_t )n * sizeof( I4 );
        prm.off_crec = ( unsigned int )o; o += n_rec * sizeof( CRec<R> );
        prm.off_par  = ( unsigned int )o; o += ( size_t )n * sizeof( int );
        prm.off_prog = ( unsigned int )o;
    }
    // (f32 only: the FP64 validation instantiations read the tables from global memory)
    prm.stage_bytes = stages_table<R>( fs, ( size_t )n_prog ) ? ( int )( ( table + 31 ) & ~( size_t )31 ) : 0;
    // kernel instantiation (acn_isect.cuh: elem_hit): 2 = planes, spheres and quadrics only, 1 = full-featured, 0 = lean sweep
    bool prims_only = !march;
    for( int i = 0; i < n && prims_only; i++ ) prims_only = fs->nodes[ i ].kind <= ACN_KIND_SQUAROID;
    if constexpr( sizeof( R ) == 4 )
    {
        if( prm.stage_bytes > 0 ) { if( march ) select_kernels<1, true>(); else if( prims_only ) select_kernels<2, true>(); else select_kernels<0, true>(); }
        else                      { if( march ) select_kernels<1, false>(); else if( prims_only ) select_kernels<2, false>(); else select_kernels<0, false>(); }
    }
    else { if( march ) select_kernels<1, false>(); else if( prims_only ) select_kernels<2, false>(); else select_kernels<0, false>(); }
    {   // staged tables, then the per-thread scratch of the CSG event sweep (none for a scene without composite objects)
        bool any_prog = false;
        for( size_t i = 0; i < cb.prog_ref.size(); i++ ) any_prog = any_prog || cb.prog_ref[ i ].y > 0;
        smem_bytes = prm.stage_bytes + ( any_prog ? ( int )csg_mem_bytes<R>( ACN_BLOCK ) : 0 );
    }
    if( smem_bytes > 30 * 1024 )
    {
        ACN_CUDA( cudaFuncSetAttribute( ( const void* )kp_primary, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes ) );
        ACN_CUDA( cudaFuncSetAttribute( ( const void* )kp_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes ) );
        ACN_CUDA( cudaFuncSetAttribute( ( const void* )kp_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes ) );
        ACN_CUDA( cudaFuncSetAttribute( ( const void* )kp_path, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes ) );
        ACN_CUDA( cudaFuncSetAttribute( ( const void* )kp_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes ) );
    }

    // ---- kernels specialised to this scene's structure (acn_spec.h, acn_rtc.h)
    {
        const char* env = getenv( "ACN_SPECIALIZE" );
        const bool want = opt->specialize == ACN_SPECIALIZE_ON || ( opt->specialize == ACN_SPECIALIZE_AUTO && env && env[ 0 ] == '1' );
        if( want )
        {
            const SpecPlan pl = plan_spec<R>( fs, cb );
            bool ok = false;
            if( !pl.src.empty() )
            {
                cudaDeviceProp pr;
                ACN_CUDA( cudaGetDeviceProperties( &pr, device ) );
                char arch[ 32 ]; snprintf( arch, sizeof( arch ), "sm_%d%d%s", pr.major, pr.minor, pr.major >= 9 ? "a" : "" );
                const char* xo = getenv( "ACN_SPEC_OPTS" );
                std::shared_ptr<SpecBinary> bin = spec_compile( pl.src, sizeof( R ) == 8, march, prm.stage_bytes > 0, arch, xo ? xo : "" );
                if( bin ) spec = spec_load( *bin );
                if( spec )
                {
                    ok = true;
                    DriverApi& drv = DriverApi::get();
                    if( smem_bytes > 30 * 1024 )
                        for( int k = 0; k < SPEC_K_COUNT; k++ )
                            if( drv.FuncSetAttribute( spec->fn[ k ], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_bytes ) != CUDA_SUCCESS ) ok = false;
                    if( !ok ) { spec.reset(); set_error( "scene specialisation: cannot set the shared-memory size of the compiled kernels" ); }
                    else if( getenv( "ACN_VERBOSE" ) )
                        fprintf( stderr, "acn: scene-specialised kernels (%s, %.1f s): registers primary %d rays %d path %d direct %d shade %d\n",
                                 bin->from_disk ? "disk cache" : "compiled", bin->compile_seconds, spec->regs[ 0 ], spec->regs[ 1 ], spec->regs[ 2 ], spec->regs[ 3 ], spec->regs[ 4 ] );
                }
            }
            else set_error( "scene specialisation: the scene has more than %d top-level elements", ( int )SpecGen::MAX_ELEMENTS );
            // an explicit request must not fall back silently; ACN_SPECIALIZE=1 in the environment is a preference
            if( !ok && opt->specialize == ACN_SPECIALIZE_ON ) return ACN_ERR_UNSUPPORTED;
            if( !ok && getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: generic kernels (%s)\n", acn_last_error() );
        }
    }

    // device stack for the CSG recursion (obj_ray_hit <-> pair_hit <-> obj_side)
    max_csg_depth = 0;
    { int a = csg_depth( fs, fs->light_root, 0 ), b = csg_depth( fs, fs->matter_root, 0 ); max_csg_depth = a > b ? a : b; }
    {
        size_t want = 8192 + ( size_t )max_csg_depth * ( sizeof( R ) == 4 ? 1024 : 1792 );
        size_t cur = 0;
        cudaDeviceGetLimit( &cur, cudaLimitStackSize );
        if( cur < want ) ACN_CUDA( cudaDeviceSetLimit( cudaLimitStackSize, want ) );
    }


    // ---- queues: sized at the first render, from the wave budget of that call (ensure_queues)
    budget_opt = opt->wave_budget > 0 ? ( uint64_t )opt->wave_budget : 0;
    if( ( rc = dev_alloc( &d_sc, 1 ) ) ) return rc;
    ACN_CUDA( cudaMallocHost( ( void** )&h_sc, sizeof( Sched ) ) );

    // ---- persistent grids: as many blocks as are resident at once
    {
        cudaDeviceProp pr;
        ACN_CUDA( cudaGetDeviceProperties( &pr, device ) );
        const int sms = pr.multiProcessorCount;
        const void* aot[ SPEC_K_COUNT ] = { ( const void* )kp_primary, ( const void* )kp_rays, ( const void* )kp_path, ( const void* )kp_direct, ( const void* )kp_shade };
        for( int k = 0; k < SPEC_K_COUNT; k++ )
        {
            int b = 0;
            if( spec ) { if( DriverApi::get().OccupancyMaxActiveBlocksPerMultiprocessor( &b, spec->fn[ k ], ACN_BLOCK, ( size_t )smem_bytes ) != CUDA_SUCCESS ) b = 0; }
            else ACN_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &b, aot[ k ], ACN_BLOCK, smem_bytes ) );
            grid_trace[ k ] = sms * ( b > 0 ? b : 1 );
        }
        grid_util = sms * 4;
    }
    return ACN_OK;
}

static inline unsigned grid_for( uint64_t n, unsigned block ) { return ( unsigned )( ( n + block - 1 ) / block ); }

// The wavefront queues.  budget = path children (and explicit rays) traced per wavefront iteration; everything an iteration
// can spawn must fit: 24 ray slots, 6 + 2 task slots, 2 hits per unit of budget, ~2.9 KB in all — 24 GB at the default
// ceiling of 2^23, which is what a 400x400 pass of wine_glass wants and absurd for a 64x48 test render.
// Without an explicit acn_options.wave_budget the budget follows the call: 64 x the samples of the call rounded up to a power
// of two, between 2^16 (190 MB) and 2^23.  Queues only grow; samples do not depend on the budget (fixed-point sums).
template <typename R> int Tracer<R>::ensure_queues( uint64_t n )
{
    uint64_t want = budget_opt;
    if( want == 0 )
    {
        want = 1ull << 16;
        while( want < ( 1ull << 23 ) && want < n * 64 ) want <<= 1;
        if( want < budget_floor ) want = budget_floor;
    }
    if( want < 64 ) want = 64;
    if( want <= budget ) return ACN_OK;
    int rc;
    ACN_CUDA( cudaDeviceSynchronize() );
    free_rays( ray_stack ); free_tasks( task_stack ); free_tasks( task_new ); free_hits( hit_q );
    cudaFree( d_dl_cum ); cudaFree( d_dl_slot ); cudaFree( d_dl_dir ); cudaFree( d_pdir );
    ray_stack = RayBuf<R>(); task_stack = TaskBuf<R>(); task_new = TaskBuf<R>(); hit_q = HitBuf<R>();
    d_dl_cum = nullptr; d_dl_slot = nullptr; d_dl_dir = nullptr; d_pdir = nullptr;
    budget = 0;
    const uint64_t bd = want;
    ray_min = bd / 4 < 65536 ? bd / 4 : 65536;      // smaller ray waves wait for company while path work is pending
    ray_cap = bd * 24;
    task_stack_cap = bd * 6;
    task_new_cap = bd * 2 + 64;
    dl_dir_cap = task_new_cap * 16 > ( 1ull << 20 ) ? task_new_cap * 16 : ( 1ull << 20 );
    pdir_cap = bd * 32 > ( 1ull << 20 ) ? bd * 32 : ( 1ull << 20 );
    if( pdir_cap > ( 1ull << 27 ) ) pdir_cap = 1ull << 27;
    prim_chunk = bd;
    {   // first-generation tasks must fit the task-stack directory: chunk * path_samples children
        const uint64_t ps = prm.path_samples > 0 ? ( uint64_t )prm.path_samples : 1;
        const uint64_t lim = pdir_cap * 32 / ps / 2;
        if( prim_chunk > lim ) prim_chunk = lim > 32 ? lim : 32;
    }
    if( ( rc = alloc_rays( ray_stack, ray_cap ) ) ) return rc;
    if( ( rc = alloc_tasks( task_stack, task_stack_cap ) ) ) return rc;
    if( ( rc = alloc_tasks( task_new, task_new_cap ) ) ) return rc;
    if( ( rc = alloc_hits( hit_q, task_new_cap ) ) ) return rc;
    if( ( rc = dev_alloc( &d_dl_cum, task_new_cap ) ) ) return rc;
    if( ( rc = dev_alloc( &d_dl_slot, task_new_cap ) ) ) return rc;
    if( ( rc = dev_alloc( &d_dl_dir, dl_dir_cap ) ) ) return rc;
    if( ( rc = dev_alloc( &d_pdir, pdir_cap ) ) ) return rc;
    budget = bd;
    if( getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: wavefront queues for a budget of %llu: %.2f GB\n", ( unsigned long long )bd,
                                          ( double )( ray_cap * 4 * sizeof( R4<R> ) + ( task_stack_cap + task_new_cap ) * ( 4 * sizeof( R4<R> ) + 40 ) + task_new_cap * ( 4 * sizeof( R4<R> ) + 24 + 12 ) + ( dl_dir_cap + pdir_cap ) * 4 ) / 1e9 );
    return ACN_OK;
}

template <typename R> int Tracer<R>::render( const double* d_xy, uint64_t n, uint64_t index_base, float* d_rgb, cudaStream_t st,
                                             const volatile int* cancel, acn_stats* stats )
{
    ACN_CUDA( cudaSetDevice( device ) );
    if( stats ) memset( stats, 0, sizeof( *stats ) );
    if( n == 0 ) return ACN_OK;
    if( n > 0x7FFFFFFFull ) { set_error( "at most 2^31-1 samples per call" ); return ACN_ERR_INVALID_ARG; }
    { const int rq = ensure_queues( n ); if( rq ) return rq; }
    if( accum_cap < n )
    {
        cudaFree( d_accum ); d_accum = nullptr; accum_cap = 0;
        int rc = dev_alloc( &d_accum, ( size_t )n * 4 );
        if( rc ) return rc;
        accum_cap = n;
    }
    const cudaEvent_t ev0 = ev_t0, ev1 = ev_t1;
    ACN_CUDA( cudaEventRecord( ev0, st ) );
    ACN_CUDA( cudaMemsetAsync( d_accum, 0, ( size_t )n * 4 * sizeof( typename Acc<R>::T ), st ) );
    ACN_CUDA( cudaMemsetAsync( d_sc, 0, sizeof( Sched ), st ) );

    uint64_t launches = 0;
    // optional per-kernel timing (ACN_PROFILE_KERNELS=1): an event pair around every launch of the four tracing kernels
    struct KProf { bool on = false; std::vector<cudaEvent_t> a[ 8 ], b[ 8 ]; } kp;
    { const char* e = getenv( "ACN_PROFILE_KERNELS" ); kp.on = e && e[ 0 ] == '1'; }
    auto kp_begin = [ & ]( int c ) { if( kp.on ) { cudaEvent_t e; cudaEventCreate( &e ); cudaEventRecord( e, st ); kp.a[ c ].push_back( e ); } };
    auto kp_end = [ & ]( int c ) { if( kp.on ) { cudaEvent_t e; cudaEventCreate( &e ); cudaEventRecord( e, st ); kp.b[ c ].push_back( e ); } };

    const Wave<R> w = make_wave( index_base );
    int result = ACN_OK;

    // one k_sched + the kernels it planned; nothing here depends on device-side counts.  Two streams (unless the
    // per-kernel timing is on, which needs the kernels one after the other):
    //   st    k_sched  k_rays .................... | k_shade  k_index
    //   s2    [k_direct of the previous iteration]  k_path   |                  k_direct
    // k_path needs k_sched's plan; k_shade needs the hits of k_rays and k_path and — because it clears the direct list
    // and k_index rebuilds it — the previous k_direct, which precedes k_path on s2; k_direct needs k_index.
    const bool two = !kp.on && !getenv( "ACN_ONE_STREAM" );
    cudaStream_t s2 = two ? side_stream : st;
    bool direct_pending = false;
    // a tracing kernel on its persistent grid: the scene-specialised module when there is one, else the generic kernels
    const void* aot[ SPEC_K_COUNT ] = { ( const void* )kp_primary, ( const void* )kp_rays, ( const void* )kp_path, ( const void* )kp_direct, ( const void* )kp_shade };
    auto launch_trace = [ & ]( int k, cudaStream_t s, void** args )
    {
        if( spec ) DriverApi::get().LaunchKernel( spec->fn[ k ], ( unsigned )grid_trace[ k ], 1, 1, ACN_BLOCK, 1, 1, ( unsigned )smem_bytes, ( CUstream )s, args, nullptr );
        else cudaLaunchKernel( aot[ k ], dim3( ( unsigned )grid_trace[ k ] ), dim3( ACN_BLOCK ), args, ( size_t )smem_bytes, s );
    };
    auto enqueue = [ & ]( int mode, uint64_t first, uint64_t cnt )
    {
        kp_begin( 6 );
        k_sched<<< 1, 32, 0, st >>>( d_sc, task_stack.cum, d_pdir, budget, ray_min, ray_cap, mode, first, cnt );
        launches++;
        if( mode == SCHED_PRIMARY )
        {
            kp_end( 6 );
            kp_begin( 0 );
            { void* a[] = { ( void* )&w, ( void* )&d_xy }; launch_trace( SPEC_K_PRIMARY, st, a ); }
            kp_end( 0 );
            launches++;
            if( two && direct_pending ) cudaStreamWaitEvent( st, ev_direct, 0 );
        }
        else
        {
            if( two ) { cudaEventRecord( ev_sched, st ); cudaStreamWaitEvent( s2, ev_sched, 0 ); }
            kp_end( 6 );
            kp_begin( 1 );
            { void* a[] = { ( void* )&w, ( void* )&ray_stack }; launch_trace( SPEC_K_RAYS, st, a ); }
            kp_end( 1 );
            kp_begin( 2 );
            { void* a[] = { ( void* )&w, ( void* )&task_stack, ( void* )&d_pdir }; launch_trace( SPEC_K_PATH, s2, a ); }
            kp_end( 2 );
            if( two ) { cudaEventRecord( ev_path, s2 ); cudaStreamWaitEvent( st, ev_path, 0 ); }
            launches += 2;
        }
        kp_begin( 4 );
        { void* a[] = { ( void* )&w, ( void* )&hit_q }; launch_trace( SPEC_K_SHADE, st, a ); }
        kp_end( 4 );
        kp_begin( 5 );
        k_index<R><<< grid_util, 256, 0, st >>>( d_sc, task_new, task_new_cap, prm.n_lights, d_dl_cum, d_dl_slot, d_dl_dir, task_new_cap, dl_dir_cap,
                                                 task_stack, d_pdir, task_stack_cap, pdir_cap );
        kp_end( 5 );
        if( two ) { cudaEventRecord( ev_index, st ); cudaStreamWaitEvent( s2, ev_index, 0 ); }
        kp_begin( 3 );
        { void* a[] = { ( void* )&w, ( void* )&task_new, ( void* )&d_dl_cum, ( void* )&d_dl_slot, ( void* )&d_dl_dir }; launch_trace( SPEC_K_DIRECT, s2, a ); }
        kp_end( 3 );
        if( two ) { cudaEventRecord( ev_direct, s2 ); direct_pending = true; }
        launches += 3;
    };
    // everything on the side stream joins st: before the host reads the scheduler state for good, and before k_finish
    auto join = [ & ]() { if( two && direct_pending ) { cudaStreamWaitEvent( st, ev_direct, 0 ); direct_pending = false; } };

    const int iters_per_poll = 4;
    for( uint64_t first = 0; first < n && result == ACN_OK; first += prim_chunk )
    {
        const uint64_t cnt = ( n - first < prim_chunk ) ? n - first : prim_chunk;
        enqueue( SCHED_PRIMARY, first, cnt );
        for( uint64_t polls = 0;; polls++ )
        {
            if( polls > ( 1ull << 22 ) ) { set_error( "wavefront scheduler did not terminate" ); result = ACN_ERR_CUDA; break; }   // cannot happen; never spin forever
            for( int k = 0; k < iters_per_poll; k++ ) enqueue( SCHED_WAVE, 0, 0 );
            cudaError_t ce = cudaMemcpyAsync( h_sc, d_sc, sizeof( Sched ), cudaMemcpyDeviceToHost, st );
            if( ce == cudaSuccess ) ce = cudaStreamSynchronize( st );
            if( ce != cudaSuccess ) { set_error( "wavefront iteration failed: %s", cudaGetErrorString( ce ) ); result = ACN_ERR_CUDA; break; }
            if( h_sc->overflow )
            {
                set_error( "wavefront queue overflow (code %d): change acn_options.wave_budget", h_sc->overflow );
                result = ACN_ERR_OUT_OF_MEMORY; break;
            }
            if( h_sc->done ) break;
            if( cancel && *cancel ) { result = ACN_ERR_CANCELLED; break; }
        }
    }

    join();
    if( result == ACN_OK )
    {
        k_finish<R><<< grid_for( n * 3, 256 ), 256, 0, st >>>( d_accum, n, prm.gamma, d_rgb );
        launches++;
    }
    // single exit: whatever happened above, both streams are drained before the call returns (a k_direct still pending
    // on the side stream must not meet the next call's memsets)
    cudaMemcpyAsync( h_sc, d_sc, sizeof( Sched ), cudaMemcpyDeviceToHost, st );
    cudaEventRecord( ev1, st );
    cudaError_t se = cudaStreamSynchronize( st );
    if( two ) { const cudaError_t s2e = cudaStreamSynchronize( side_stream ); if( se == cudaSuccess ) se = s2e; }
    if( se != cudaSuccess && result == ACN_OK ) { set_error( "render: %s", cudaGetErrorString( se ) ); result = ACN_ERR_CUDA; }
    cudaError_t le = cudaGetLastError();
    if( le != cudaSuccess ) { set_error( "kernel failure: %s", cudaGetErrorString( le ) ); result = ACN_ERR_CUDA; }
    float ms = 0; cudaEventElapsedTime( &ms, ev0, ev1 );
    if( stats )
    {
        const unsigned long long* ss = h_sc->stats;
        stats->samples = n;
        stats->rays_primary = ss[ ST_PRIMARY ]; stats->rays_reflection = ss[ ST_REFLECT ];
        stats->rays_chromatic = ss[ ST_CHROMATIC ]; stats->rays_refraction = ss[ ST_REFRACT ];
        stats->rays_path = ss[ ST_PATH ]; stats->rays_shadow = ss[ ST_SHADOW ];
        stats->rays_light = ss[ ST_LIGHT ]; stats->diffuse_hits = ss[ ST_DIFFUSE ];
        stats->kernel_launches = launches; stats->waves = h_sc->waves; stats->device_ms = ms;
    }
    if( result == ACN_ERR_OUT_OF_MEMORY && budget_opt == 0 && budget < ( 1ull << 23 ) && !kp.on )
    {   // an automatic budget proved too small for this scene's fan-out: twice the queues, once more from the start
        budget_floor = budget * 2;
        if( getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: queue overflow at a budget of %llu, retrying with %llu\n", ( unsigned long long )budget, ( unsigned long long )budget_floor );
        return render( d_xy, n, index_base, d_rgb, st, cancel, stats );
    }
    if( kp.on )
    {
        for( int c = 0; c < 8; c++ )
        {
            double tot = 0;
            for( size_t i = 0; i < kp.a[ c ].size() && i < kp.b[ c ].size(); i++ )
            {
                float m = 0;
                if( cudaEventElapsedTime( &m, kp.a[ c ][ i ], kp.b[ c ][ i ] ) == cudaSuccess ) tot += m;
                cudaEventDestroy( kp.a[ c ][ i ] ); cudaEventDestroy( kp.b[ c ][ i ] );
            }
            if( stats ) { stats->kernel_ms[ c ] = tot; stats->kernel_launches_by_class[ c ] = kp.a[ c ].size(); }
        }
        cudaGetLastError();
    }
    return result;
}

} // namespace acn
