// acn_isect.cuh — device-only scene queries of the wavefront tracer.
//
//   csg_eval      composite objects (obj_pair_inside_s / obj_pair_outside_s / obj_neg_s over plane, sphere
//                 and squaroid leaves) by an event sweep over a boolean function
//   scene_query   ONE traversal routine for every kind of ray: scene_s_trans_hit (scene.c:362-382),
//                 compound_s_ray_trans_hit / compound_s_ray_hit (compound.c:215-299) and the shadow /
//                 probe "anything closer than t?" tests, selected by flags, so that each kernel holds a
//                 single copy of the intersection code (instruction-cache footprint) and none of it
//                 needs a call stack.
#pragma once

#include "acn_geom.h"

namespace acn {

// ---------------------------------------------------------------------------------------------
// CSG by event sweep.  The reference finds the first boundary of A&B / A|B by an alternating
// march over the children (objects.c:1052-1094,1209-1251), re-tracing whole subtrees for every
// rejected candidate: O(n^2) ray tests for an n-leaf solid and hopelessly divergent on a GPU.
// The same first boundary comes out of classifying the ray against every leaf ONCE:
//   * a leaf is a boolean VARIABLE: its state at the ray origin plus its (<= 2) crossings for t > 0;
//   * the solid is a boolean FUNCTION of the variables ('!' complement, '&' and, '|' or);
//   * a node's envelope clips its inside-set (obj_side reports "outside" beyond the own envelope,
//     objects.c:365-370, also for negations): one more variable whose crossings are VIRTUAL — they
//     change the state but are never reported as hits (obj_ray_hit only returns shape crossings) —
//     and a ray that misses the envelope skips the whole subtree (objects.c:264);
//   * the hit is the first crossing, in order of t, at which the function of the toggled variables
//     flips — the reference's acceptance test ("the boundary point of one child lies on the wanted
//     side of the other") applied once per crossing instead of once per re-traced subtree.
// Two reductions keep the variable count small: chains of one operator are opened up (A&(B&C) is the
// same point set as (A&B)&C), and the CONVEX operands of an '&' chain (half-spaces, negated
// half-spaces, balls, ellipsoids, elliptic cylinders) are fused into ONE variable, a RUN, whose
// inside-set along the ray is the single interval [max entry, min exit] (slab clipping) — the
// 58-facet brilliant of diamond.acn is one variable.  With <= CSG_TABLE_VARS variables the function
// is a truth table of 2^n bits stored behind the program (64 bytes for the wine glass), so one state
// evaluation is one shared-memory load; larger functions are interpreted with a bit stack.
//
// Per-thread storage: CSG_E events (t, leaf id, variable) in shared memory, element k of thread tid at
// [k*nthreads + tid] — conflict-free, no local-memory frame.  If a ray produces more than CSG_E
// crossings the smallest CSG_E are kept and the sweep is exact up to the first dropped one; a sweep
// that gets that far without a hit falls back to the reference march.
//
// Program words (op | arg << 4):  LEAF node | RUN count, then count MEMBER/MEMBER_NEG node words |
// NEG | AND | OR | ENV node, then (words to skip | variables skipped << 16) | CLIP node |
// XFORM node ... XEND node: the words in between belong to the child of an obj_scale_s (objects.c:1418-1443) and are
// classified against the ray carried into its frame; their crossings come back as ray parameters of the outer ray.
// ---------------------------------------------------------------------------------------------
#ifndef ACN_CSG_INLINE
#define ACN_CSG_INLINE __forceinline__
#endif

// per-thread scratch of a scene query in shared memory, [slot][thread]: the CSG events
// element i of an array in shared memory, addressed by its 32-bit shared-window byte address (LDS/STS, no 64-bit pointer registers)
template <typename T> struct SPtr
{
    unsigned int a;
    __device__ __forceinline__ T& operator[]( int i ) const
    {
        return *reinterpret_cast<T*>( __cvta_shared_to_generic( a + ( unsigned int )i * ( unsigned int )sizeof( T ) ) );
    }
};

template <typename R> struct CsgMem
{
    SPtr<R> t; SPtr<unsigned short> iv;         // CSG_E events: crossing, leaf id | variable << 8
    int stride;
};

template <typename R> __host__ __device__ inline size_t csg_mem_bytes( int nthreads )
{
    return ( size_t )( sizeof( R ) + 2 ) * CSG_E * nthreads;
}

template <typename R> __device__ __forceinline__ CsgMem<R> csg_mem( unsigned char* base, int nthreads, int tid )
{
    CsgMem<R> m;
    const unsigned int b = ( unsigned int )__cvta_generic_to_shared( base );
    m.t.a  = b + ( unsigned int )( tid * sizeof( R ) );
    m.iv.a = b + ( unsigned int )( sizeof( R ) * CSG_E * nthreads + tid * 2 );
    m.stride = nthreads;
    return m;
}

// crossings of a sphere for t > 0 and the state at the origin
template <typename R> __device__ __forceinline__ int sphere_events( V3<R> c, R r, const Ray<R>& ray, int* s0, R* t0, R* t1 )
{
    V3<R> p = ray.p - c;
    R s = dot( p, ray.d );
    R q = sqr( p ) - r * r;
    V3<R> l = p - ray.d * s;
    R disc = r * r - sqr( l );
    *s0 = q > R( 0 ) ? 0 : 1;
    if( disc < R( 0 ) ) return 0;
    R sq = r_sqrt( disc );
    if( q > R( 0 ) )
    {
        if( !( s < R( 0 ) ) ) return 0;
        *t0 = -s - sq; *t1 = -s + sq;
        return 2;
    }
    if( s < R( 0 ) || q < R( 0 ) ) { *t0 = -s + sq; return 1; }
    return 0;
}

// Crossings of a distance-field object (sphere-traced, objects.c:903-959) beyond ray parameter `from`, at most two per
// call, and the state at `from`.  A torus has up to four crossings: the program carries a CSG_MORE word behind the
// leaf that asks for the next two, starting behind the second one.  Each crossing is one march of the reference's
// own hit function from where the previous one ended (+ eps), which is how its alternating march advances too
// (objects.c:1087,1244: offs += a + 2 eps).
template <typename R, bool SH> __device__ __forceinline__ int dist_events( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, R from, int* s0, R* t0, R* t1 )
{
    const R inf = Num<R>::inf();
    Ray<R> r2; r2.d = ray.d; r2.p = madd( ray.p, ray.d, from );
    *s0 = prim_side( sv, kind, n, r2.p ) < 0 ? 1 : 0;
    const R a0 = dist_hit( sv, kind, n, r2, ( V3<R>* )nullptr );
    if( !( a0 < inf ) ) return 0;
    *t0 = from + a0 + sv.eps;
    r2.p = madd( ray.p, ray.d, *t0 + sv.eps );
    const R a1 = dist_hit( sv, kind, n, r2, ( V3<R>* )nullptr );
    if( !( a1 < inf ) ) return 1;
    *t1 = *t0 + sv.eps + a1 + sv.eps;
    return 2;
}

// DIST: distance-field leaves compiled in (only the MARCH instantiations of the kernels: their register appetite must
// not touch the lean kernels)
template <bool DIST, typename R, bool SH> __device__ __forceinline__ int leaf_events( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, int* s0, R* t0, R* t1 )
{
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );
    if( kind == K_SPHERE ) return sphere_events( pos, g0.w, ray, s0, t0, t1 );
    if( kind == K_PLANE )
    {
        V3<R> nz = xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
        R g = dot( ray.p - pos, nz );
        R dn = dot( nz, ray.d );
        *s0 = g > R( 0 ) ? 0 : 1;
        if( dn == R( 0 ) ) return 0;
        R tt = -g / dn;
        if( tt > R( 0 ) ) { *t0 = tt; return 1; }
        return 0;
    }
    if( DIST && ( kind == K_DIST_SPHERE || kind == K_DIST_TORUS ) ) return dist_events( sv, kind, n, ray, R( 0 ), s0, t0, t1 );
    // squaroid
    const M3<R> rax = node_rax( sv, n );
    const R qa = g0.w, qb = sv.geo[ n * GEO_STRIDE + 1 ].w, qc = sv.geo[ n * GEO_STRIDE + 2 ].w, qr = sv.geo[ n * GEO_STRIDE + 3 ].w;
    V3<R> p = mlv( rax, ray.p - pos );
    V3<R> d = mlv( rax, ray.d );
    V3<R> ad = v3<R>( qa * d.x, qb * d.y, qc * d.z );
    R f = dot( ad, d ), fs = dot( ad, p );
    R fq = qa * p.x * p.x + qb * p.y * p.y + qc * p.z * p.z + qr;
    *s0 = fq > R( 0 ) ? 0 : 1;
    if( f == R( 0 ) ) return 0;
    R ta, tb;
    if( !quadric_roots( qa, qb, qc, qr, p, d, ad, f, fs, fq, &ta, &tb ) ) return 0;
    int c = 0;
    if( ta >= R( 0 ) ) { *t0 = ta; c = 1; }
    if( tb >= R( 0 ) ) { if( c ) *t1 = tb; else *t0 = tb; c++; }
    return c;
}

// outward normal of a leaf at ray parameter t (the unshortened crossing), as its fp_ray_hit reports it
template <bool DIST, typename R, bool SH> __device__ __forceinline__ V3<R> leaf_normal( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, R t )
{
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );
    if( kind == K_PLANE ) return xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
    if( kind == K_SPHERE ) return unit( madd( ray.p - pos, ray.d, t - sv.eps ) );
    const M3<R> rax = node_rax( sv, n );
    if( DIST && ( kind == K_DIST_SPHERE || kind == K_DIST_TORUS ) )
    {   // gradient at the end point of the march = the unshortened crossing (objects.c:945-955)
        const V3<R> lp = mlv( rax, madd( ray.p, ray.d, t ) - pos ) * g0.w;
        return unit( tmlv( rax, dist_gradient( kind, sv.geo[ n * GEO_STRIDE + 1 ].w, lp, sv.eps ) ) );
    }
    V3<R> p = mlv( rax, ray.p - pos );
    V3<R> d = mlv( rax, ray.d );
    V3<R> x = madd( p, d, t );
    return unit( tmlv( rax, v3<R>( x.x * g0.w, x.y * sv.geo[ n * GEO_STRIDE + 1 ].w, x.z * sv.geo[ n * GEO_STRIDE + 2 ].w ) ) );
}

// state of the solid for a variable assignment: truth table, or the postfix program on a bit stack
// dead: the CLIP variables of the sub-envelopes the ray misses altogether (pass 1 of csg_eval).  Such a subtree is 0 whatever
// happens along the ray (objects.c:264), so the interpreter pushes a 0 and jumps over its words: a lamp of sixty butted
// pieces, of which a ray meets the envelopes of two, costs the words of two pieces per evaluation instead of all of them
// (hanging_lamps_in_row: the interpreter was 55 % of k_direct's instructions).
template <typename R, bool SH> __device__ __forceinline__ int csg_state( const SceneView<R, SH>& sv, const I4& pr, unsigned long long vars, unsigned long long dead = 0ull )
{
    if( pr.z >= 0 ) return ( sv.prog[ pr.z + ( int )( vars >> 5 ) ] >> ( ( unsigned int )vars & 31u ) ) & 1;
    unsigned int stk = 0;
    if( pr.z <= -2 )
    {   // evaluation program: truth tables over stretches of consecutive variables, joined by the operators of the chains they
        // were cut from (acn_tracer.cuh: build_eval_program)
        const int at = -2 - pr.z;
        const int end = at + 1 + sv.prog[ at ];
        const unsigned int vlo = ( unsigned int )vars, vhi = ( unsigned int )( vars >> 32 );
        // bits v .. v+31 of vars with 32-bit shifts (one funnel shift instead of a 64-bit shift by a register)
        #define ACN_VARS_FROM( v ) ( ( v ) < 32 ? __funnelshift_r( vlo, vhi, ( v ) ) : ( vhi >> ( ( v ) - 32 ) ) )
        #pragma unroll 1
        for( int pc = at + 1; pc < end; pc++ )
        {
            const int ins = sv.prog[ pc ];
            const int op = ins & 15;
            if( op == E_TAB )
            {   // one word: the table sits ( ins >> 14 ) words behind the program's length word
                const int nv = ( ins >> 4 ) & 15, v0 = ( ins >> 8 ) & 63;
                const unsigned int idx = ACN_VARS_FROM( v0 ) & ( ( 1u << nv ) - 1u );
                const int off = at + ( int )( ( unsigned int )ins >> 14 );
                stk = ( stk << 1 ) | ( ( ( unsigned int )sv.prog[ off + ( int )( idx >> 5 ) ] >> ( idx & 31u ) ) & 1u );
            }
            else if( op == E_VAR )  stk = ( stk << 1 ) | ( ACN_VARS_FROM( ins >> 4 ) & 1u );
            else if( op == E_CLIP ) stk &= ~1u | ( ACN_VARS_FROM( ins >> 4 ) & 1u );
            else if( op == E_NEG )  stk ^= 1u;
            else if( op == E_AND )  { const unsigned int t = stk & 1u; stk >>= 1; stk &= t | ~1u; }
            else                    { const unsigned int t = stk & 1u; stk >>= 1; stk |= t; }
        }
        #undef ACN_VARS_FROM
        return ( int )( stk & 1u );
    }
    int v = 0;
    #pragma unroll 1
    for( int pc = pr.x; pc < pr.x + pr.y; pc++ )
    {
        const int ins = sv.prog[ pc ];
        const int op = ins & 15;
        if( op == CSG_LEAF )      { stk = ( stk << 1 ) | ( unsigned int )( ( vars >> v ) & 1ull ); v++; }
        else if( op == CSG_RUN )  { stk = ( stk << 1 ) | ( unsigned int )( ( vars >> v ) & 1ull ); v++; pc += ins >> 4; }
        else if( op == CSG_CLIP ) { stk &= ~1u | ( unsigned int )( ( vars >> v ) & 1ull ); v++; }
        else if( op == CSG_NEG )  stk ^= 1u;
        else if( op == CSG_AND )  { const unsigned int t = stk & 1u; stk >>= 1; stk &= t | ~1u; }
        else if( op == CSG_OR )   { const unsigned int t = stk & 1u; stk >>= 1; stk |= t; }
        else if( op == CSG_ENV )
        {   // next word: words up to and including the subtree's CLIP | its variables << 16 (the CLIP variable is the last of them)
            const int w2 = sv.prog[ ++pc ];
            const int nvs = w2 >> 16;
            if( ( dead >> ( v + nvs - 1 ) ) & 1ull ) { stk <<= 1; v += nvs; pc += w2 & 0xFFFF; }
        }                                                // CSG_MORE: no effect
    }
    return ( int )( stk & 1u );
}

// inside-set of one leaf for t > 0 as an interval ( lo, hi ): lo < 0 when the origin is inside
template <typename R> __device__ __forceinline__ void member_interval( int s0, int c, R t0, R t1, R* lo, R* hi )
{
    const R inf = Num<R>::inf();
    if( s0 ) { *lo = R( -1 ); *hi = c >= 1 ? t0 : inf; }
    else     { *lo = c >= 1 ? t0 : inf; *hi = c == 2 ? t1 : inf; }
}

// obj_scale_s (objects.c:1418-1437): the ray in the frame of the scaled child.  The direction is renormalised, so a
// parameter s of the inner ray is the parameter s * fac of the outer one (fac = 1 / | M d |).
template <typename R, bool SH> __device__ __forceinline__ Ray<R> scale_ray( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, R* fac )
{
    const V3<R> pos = xyz( sv.geo[ n * GEO_STRIDE ] );
    const M3<R> rax = node_rax( sv, n );
    const V3<R> inv = v3<R>( sv.geo[ n * GEO_STRIDE ].w, sv.geo[ n * GEO_STRIDE + 1 ].w, sv.geo[ n * GEO_STRIDE + 2 ].w );
    Ray<R> rl;
    rl.p = mul( mlv( rax, ray.p - pos ), inv );
    rl.d = mul( mlv( rax, ray.d ), inv );
    const R len = r_sqrt( sqr( rl.d ) );
    const R f = len > R( 0 ) ? R( 1 ) / len : R( 0 );
    rl.d = rl.d * f;
    *fac = f;
    return rl;
}

// the crossing at tcur of the leaf with program-relative id `id` is the first boundary of the solid `root`: the hit
// distance (shortened by eps, as every fp_ray_hit reports it) and, if asked for, the normal of the leaf carried up the tree
template <bool DIST, typename R, bool SH> __device__ __forceinline__ R csg_report_hit( const SceneView<R, SH>& sv, int root, int prog_start, const Ray<R>& ray, R tcur, int id,
                                                                                      V3<R>* nor, HitCtx ctx )
{
    const R a = tcur - sv.eps;
    if( nor )
    {
        const int leaf = sv.prog[ prog_start + id ] >> 4;
        V3<R> nn;
        int sc = -1;                                // a scale node above the leaf (the builder sweeps at most one level of them)
        if( DIST && leaf != root ) for( int m = sv.parent[ leaf ]; m >= 0; m = sv.parent[ m ] ) { if( node_kind( sv.link[ m ] ) == K_SCALE ) sc = m; if( m == root ) break; }
        if( DIST && sc >= 0 )
        {
            R fac;
            const Ray<R> rl = scale_ray( sv, sc, ray, &fac );
            nn = leaf_normal<DIST>( sv, node_kind( sv.link[ leaf ] ), leaf, rl, fac > R( 0 ) ? tcur / fac : tcur );
        }
        else nn = leaf_normal<DIST>( sv, node_kind( sv.link[ leaf ] ), leaf, ray, tcur );
        // up the tree: roughness at every level that has it (objects.c:266), sign flip at negations, the normal carried
        // out of a scaled frame (objects.c:1431)
        for( int m = leaf; m != root && m >= 0; m = sv.parent[ m ] )
        {
            const I4 lk = sv.link[ m ];
            if( node_kind( lk ) == K_NEG ) nn = -nn;
            if( DIST && node_kind( lk ) == K_SCALE )
            {
                const V3<R> inv = v3<R>( sv.geo[ m * GEO_STRIDE ].w, sv.geo[ m * GEO_STRIDE + 1 ].w, sv.geo[ m * GEO_STRIDE + 2 ].w );
                nn = unit( tmlv( node_rax( sv, m ), mul( nn, inv ) ) );
            }
            if( node_flags( lk ) & F_ROUGH ) roughen( sv, m, ray, a, &nn, ctx );
        }
        if( node_kind( sv.link[ root ] ) == K_NEG ) nn = -nn;
        if( DIST && node_kind( sv.link[ root ] ) == K_SCALE )
        {
            const V3<R> inv = v3<R>( sv.geo[ root * GEO_STRIDE ].w, sv.geo[ root * GEO_STRIDE + 1 ].w, sv.geo[ root * GEO_STRIDE + 2 ].w );
            nn = unit( tmlv( node_rax( sv, root ), mul( nn, inv ) ) );
        }
        *nor = nn;
    }
    return a;
}

// returns the hit parameter or +inf for a miss.  A ray with more than CSG_E crossings is swept in rounds:
// each round keeps the CSG_E smallest crossings beyond t_floor; crossings at or before t_floor only
// toggle their variable (they were swept in an earlier round).
template <bool DIST, typename R, bool SH> __device__ ACN_CSG_INLINE R csg_eval( const SceneView<R, SH>& sv, int root, const Ray<R>& ray, V3<R>* nor, HitCtx ctx,
                                                             const CsgMem<R>& cm, const R t_far )
{
    // t_far: the caller's horizon (+ slack).  Crossings beyond it cannot become the reported hit and the state of
    // the solid beyond it is of no interest, so they are dropped at once; a convex run whose interval is empty or
    // starts beyond the horizon stops testing its remaining members (slab clipping's early exit).
    const R inf = Num<R>::inf();
    const I4 pr = sv.prog_ref[ root ];          // start, length, truth table offset (-1: interpret), variables
    R t_floor = R( 0 );
    int s = -1;                                 // state of the solid before the next crossing (-1: not yet evaluated)
    for( int round = 0; round < 16; round++ )
    {
        unsigned long long vars = 0, dead = 0;
        int nv = 0, ne = 0;
        bool dropped = false;
        R tmin0 = inf; int kmin0 = -1;              // smallest crossing appended so far (-1: unknown, scan)
        // ---- pass 1: classify the ray against every leaf.  The lanes of a warp that evaluate this object walk
        // the program in LOCKSTEP: a lane whose ray misses a sub-envelope does not jump ahead (it would then
        // execute other program words than its neighbours and serialise against them) but idles through the
        // subtree; the subtree is skipped only when no lane needs it.  The envelope's own crossings are the
        // CLIP variable's events, so the envelope is intersected once, at the ENV word.
        int skip_to = pr.x;                         // this lane idles while pc < skip_to
        R more_from = inf;                          // second crossing of the distance-field leaf just classified (CSG_MORE continues there)
        Ray<R> rc = ray; R fac = R( 1 );            // the ray in the frame of the words being classified, and what its parameter is worth outside (XFORM)
        #pragma unroll 1
        for( int pc = pr.x; pc < pr.x + pr.y; pc++ )
        {
            const int ins = sv.prog[ pc ];
            const int op = ins & 15, n = ins >> 4;
            const bool act = pc >= skip_to;
            R t0 = R( 0 ), t1 = R( 0 ); int s0 = 0, c = 0, id0 = CSG_VIRTUAL, id1 = CSG_VIRTUAL, var = nv;
            if( op == CSG_LEAF )
            {
                more_from = inf;
                if( act ) { c = leaf_events<DIST>( sv, node_kind( sv.link[ n ] ), n, DIST ? rc : ray, &s0, &t0, &t1 ); if( DIST ) { t0 *= fac; t1 *= fac; } id0 = id1 = pc - pr.x; if( DIST && c == 2 ) more_from = t1; }
                nv++;
            }
            else if( DIST && ( op == CSG_XFORM || op == CSG_XEND ) )
            {   // into / out of the frame of a scaled child (one level: the builder does not sweep nested scale nodes)
                if( op == CSG_XFORM ) rc = scale_ray( sv, n, ray, &fac ); else { rc = ray; fac = R( 1 ); }
                continue;
            }
            else if( DIST && op == CSG_MORE )
            {   // crossings 3 and 4 of the distance-field leaf in front (same variable); s0 collects toggles only
                var = nv - 1;
                if( act && more_from < t_far )
                {
                    int sd;
                    c = dist_events( sv, node_kind( sv.link[ n ] ), n, rc, ( more_from + sv.eps ) / fac, &sd, &t0, &t1 ); t0 *= fac; t1 *= fac;
                    id0 = id1 = pc - pr.x;
                }
                more_from = inf;
            }
            else if( op == CSG_RUN )
            {
                if( act )
                {
                    R lo = R( -1 ), hi = inf;
                    #pragma unroll 1
                    for( int m = 1; m <= n; m++ )
                    {
                        const int w = sv.prog[ pc + m ];
                        const int node = w >> 4;
                        R a0 = R( 0 ), a1 = R( 0 ); int ms0;
                        const int mc = leaf_events<false>( sv, node_kind( sv.link[ node ] ), node, DIST ? rc : ray, &ms0, &a0, &a1 );
                        if( DIST ) { a0 *= fac; a1 *= fac; }
                        if( ( w & 15 ) == CSG_MEMBER_NEG ) ms0 ^= 1;
                        R mlo, mhi;
                        member_interval( ms0, mc, a0, a1, &mlo, &mhi );
                        if( mlo > lo ) { lo = mlo; id0 = pc + m - pr.x; }
                        if( mhi < hi ) { hi = mhi; id1 = pc + m - pr.x; }
                        if( !( lo < hi ) || lo > t_far ) { lo = inf; break; }       // empty, or not before the horizon
                    }
                    if( lo < R( 0 ) )   { s0 = 1; if( hi < inf ) { c = 1; t0 = hi; id0 = id1; } }
                    else if( lo < hi )  { s0 = 0; c = 1; t0 = lo; if( hi < inf ) { c = 2; t1 = hi; } }
                }
                pc += n; nv++;
            }
            else if( op == CSG_ENV )
            {
                const int w2 = sv.prog[ ++pc ];
                const int sub_end = pc + 1 + ( w2 & 0xFFFF );            // first word behind the subtree's CLIP
                var = nv + ( w2 >> 16 ) - 1;                             // the CLIP variable closes the subtree
                if( act )
                {
                    const R4<R> e = sv.env[ n ];
                    c = sphere_events( xyz( e ), e.w, DIST ? rc : ray, &s0, &t0, &t1 );
                    if( DIST ) { t0 *= fac; t1 *= fac; }
                    if( c == 0 ) { s0 = 0; skip_to = sub_end; dead |= 1ull << var; }     // objects.c:264: the ray misses the envelope
                }
                if( !__any_sync( __activemask(), pc + 1 >= skip_to ) ) { nv += w2 >> 16; pc = sub_end - 1; }
            }
            else { if( op == CSG_CLIP ) nv++; continue; }               // CLIP: events were taken at ENV; NEG / AND / OR: pass 2
            // append the (<= 2) crossings: beyond the horizon -> dropped; at or before t_floor -> swept in an earlier
            // round, only toggles the variable; the smallest crossing of the round is tracked on the way
            #pragma unroll
            for( int k = 0; k < 2; k++ )
            {
                const R t = k ? t1 : t0;
                if( k >= c || t > t_far ) continue;
                if( t <= t_floor ) { s0 ^= 1; continue; }
                const unsigned short iv = ( unsigned short )( ( k ? id1 : id0 ) | ( var << 8 ) );
                if( ne < CSG_E )
                {
                    if( t < tmin0 ) { tmin0 = t; kmin0 = ne; }
                    cm.t[ ne * cm.stride ] = t; cm.iv[ ne * cm.stride ] = iv; ne++;
                }
                else
                {   // keep the CSG_E smallest
                    dropped = true;
                    int kmax = 0; R tmax = cm.t[ 0 ];
                    for( int q = 1; q < CSG_E; q++ ) { const R tq = cm.t[ q * cm.stride ]; if( tq > tmax ) { tmax = tq; kmax = q; } }
                    if( t < tmax ) { cm.t[ kmax * cm.stride ] = t; cm.iv[ kmax * cm.stride ] = iv; kmin0 = -1; }
                }
            }
            if( DIST && op == CSG_MORE ) vars ^= ( unsigned long long )s0 << var; else vars |= ( unsigned long long )s0 << var;
        }
        // ---- sweep: crossings in order of t until the solid's state flips at a real one.
        // The reference accepts the boundary point of one child when it lies on the wanted side of the OTHER child, and
        // takes that side at the shortened hit, eps in front of the boundary (objects.c:1063-1078,1220-1235).  A crossing is
        // therefore judged ALONE against the state eps before it: other crossings closer than eps ahead of it do not
        // count yet.  That matters where scripts butt pieces together on one cut plane: leaving piece A of A|B into
        // piece B through their common plane is two crossings at the same t, and the reference reports the seam (the exit
        // from A is tested against "not yet in B") — a sweep that applied the toggles one after the other saw the seam
        // or not depending on their order.  Crossings within eps of the first one of their group are judged against the
        // state before the group.
        // The grouping is compiled into the full-featured (MARCH) instantiation only; the host selects it for scenes in
        // which two leaves of one program share a surface (acn_tracer.cuh: CsgBuilder::coincident_leaves).
        int id = CSG_VIRTUAL;
        R tcur = t_floor;
        bool hit = false;
        if( DIST )
        {
            R t_group = -inf;
            unsigned long long vars_pre = vars;
            int s_pre = 0, s_last = -1;                 // F( vars_pre ); F( vars ) when known, else -1
            for( ;; )
            {
                int kmin = kmin0; R tmin = tmin0;
                if( kmin < 0 ) { tmin = inf; for( int q = 0; q < ne; q++ ) { const R tq = cm.t[ q * cm.stride ]; if( tq < tmin ) { tmin = tq; kmin = q; } } }
                kmin0 = -1;                             // only the first crossing is known in advance
                if( kmin < 0 ) break;
                const unsigned int iv = cm.iv[ kmin * cm.stride ];
                cm.t[ kmin * cm.stride ] = inf;
                const unsigned long long bit = 1ull << ( iv >> 8 );
                const bool first = !( tmin < t_group + sv.eps );
                if( first ) { vars_pre = vars; t_group = tmin; s_pre = s_last >= 0 ? s_last : csg_state( sv, pr, vars_pre, dead ); }
                vars ^= bit;
                tcur = tmin; id = ( int )( iv & 255u );
                s_last = -1;
                if( id != CSG_VIRTUAL )
                {
                    const int s1 = csg_state( sv, pr, vars_pre ^ bit, dead );
                    if( s1 != s_pre ) { hit = true; break; }
                    if( first ) s_last = s1;            // vars == vars_pre ^ bit
                }
            }
        }
        else
        {
            for( ;; )
            {
                const int s2 = csg_state( sv, pr, vars, dead );
                if( s >= 0 && s2 != s && id != CSG_VIRTUAL ) { hit = true; break; }
                s = s2;
                int kmin = kmin0; R tmin = tmin0;
                if( kmin < 0 ) { tmin = inf; for( int q = 0; q < ne; q++ ) { const R tq = cm.t[ q * cm.stride ]; if( tq < tmin ) { tmin = tq; kmin = q; } } }
                kmin0 = -1;                             // only the first crossing is known in advance
                if( kmin < 0 ) break;
                const unsigned int iv = cm.iv[ kmin * cm.stride ];
                cm.t[ kmin * cm.stride ] = inf;
                vars ^= 1ull << ( iv >> 8 );
                tcur = tmin; id = ( int )( iv & 255u );
            }
        }
        if( hit ) return csg_report_hit<DIST>( sv, root, pr.x, ray, tcur, id, nor, ctx );
        if( !dropped ) break;
        t_floor = tcur;                         // every kept crossing was swept: go on beyond the last one
    }
    return inf;
}

// objects the event sweep does not cover (scale nodes, CSG over distance fields): the reference's
// recursive march, out of line.  Only the MARCH instantiation of the kernels contains it — the mutually
// recursive obj_ray_hit / pair_hit / obj_side want ~250 registers, which would otherwise set the
// register count (and the occupancy) of every tracing kernel.
template <typename R, bool SH> __device__ __noinline__ R march_hit( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    const I4 lk = sv.link[ n ];
    return shape_hit( sv, lk, n, ray, nor, ctx );
}

// obj_ray_hit after the envelope test (objects.c:261-284): fp_ray_hit + roughness
// MARCH selects what the kernel instantiation carries: 0 the event sweep over planes, spheres and quadrics; 1 the
// full-featured sweep (distance-field leaves, coincident crossings, scale nodes; FP64: the reference's recursive march
// too); 2 no composite objects at all — a scene of planes, spheres and quadrics (many_spheres.acn) gets kernels whose
// register count and code size are set by the walk, not by the event sweep they never run.
template <typename R, int MARCH, bool SH> __device__ __forceinline__ R elem_hit( const SceneView<R, SH>& sv, const I4& lk, int c, const Ray<R>& ray, V3<R>* nor, HitCtx ctx,
                                                                         const CsgMem<R>& cm, const R t_far )
{
    const int kind = node_kind( lk );
    R a;
    if( kind == K_PLANE || kind == K_SPHERE || kind == K_SQUAROID ) a = prim_hit( sv, kind, c, ray, nor );
    else if( MARCH != 2 && kind >= K_PAIR_INSIDE && sv.prog_ref[ c ].y > 0 ) a = csg_eval<MARCH == 1>( sv, c, ray, nor, ctx, cm, t_far );
    else if( MARCH != 2 && ( kind == K_DIST_SPHERE || kind == K_DIST_TORUS ) ) a = dist_hit( sv, kind, c, ray, nor );
    else if( MARCH == 1 && sizeof( R ) == 8 ) a = march_hit( sv, c, ray, nor, ctx );     // the reference's recursive march: FP64 validation mode only
    else a = Num<R>::inf();                 // unreachable: the FP32 tracer refuses scenes the sweep does not cover (Tracer::init)
    if( nor && ( node_flags( lk ) & F_ROUGH ) && a < Num<R>::inf() ) roughen( sv, c, ray, a, nor, ctx );
    return a;
}

// ---------------------------------------------------------------------------------------------
// scene_query
//   Q_LIGHT / Q_MATTER  which root compounds are searched (lights first: strict '<' between the roots,
//                       so lights win ties — scene.c:362-382)
//   Q_TRANS             closest hit with the eps-merge of coincident surfaces of the root's ELEMENTS into
//                       one (exit_obj, enter_obj) transition (compound.c:246-299); a nested compound counts
//                       as one element whose hit is the plain closest hit of its contents (compound.c:215-244)
//   without Q_TRANS     "is anything at t <= t_any?": a hit with a <= t_any ends the search at once — exact
//                       for the shadow test, which only consumes (min > a) (scene.c:569), and for probes
// The recursion of the reference is an explicit stack of child ranges; a single running minimum with
// strict '<' selects the same element as the nested minima do (first in depth-first order wins ties).
// ---------------------------------------------------------------------------------------------
enum { Q_LIGHT = 1, Q_MATTER = 2, Q_TRANS = 4 };

// one traversal record.  From global memory (FP32: 32 bytes, 32-byte aligned) it is ONE 256-bit load (sm_100: LDG.E.256)
// instead of two 128-bit ones: the walk gathers a record per lane and step, and the L1 data pipe is what bounds it.
template <typename R, bool SH> __device__ __forceinline__ CRec<R> load_rec( const SceneView<R, SH>& sv, int i )
{
#if !defined(ACN_NO_LD256)
    if constexpr( !SH && sizeof( R ) == 4 )
    {
        CRec<R> r;
        const CRec<R>* p = sv.crec.p + i;
        asm( "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
             : "=f"( r.env.x ), "=f"( r.env.y ), "=f"( r.env.z ), "=f"( r.env.w ), "=r"( r.link.x ), "=r"( r.link.y ), "=r"( r.link.z ), "=r"( r.link.w )
             : "l"( p ) );
        return r;
    }
    else
#endif
    return sv.crec[ i ];
}

template <typename R, bool SH> __device__ __forceinline__ void trans_commit( const SceneView<R, SH>& sv, const Ray<R>& ray, R a, V3<R> nor, int obj, R* min_a, Trans<R>* tl )
{
    if( !( a < Num<R>::inf() ) ) return;
    if( a < *min_a - sv.eps )
    {
        *min_a = a;
        if( dot( nor, ray.d ) > R( 0 ) ) { tl->exit_nor = nor;  tl->exit_obj = obj; tl->enter_obj = -1; }
        else                             { tl->exit_nor = -nor; tl->exit_obj = -1;  tl->enter_obj = obj; }
    }
    else if( r_abs( a - *min_a ) < sv.eps )
    {
        *min_a = a < *min_a ? a : *min_a;
        if( dot( nor, ray.d ) > R( 0 ) ) tl->exit_obj = obj;
        else                             tl->enter_obj = obj;
    }
}

template <typename R, int MARCH, bool SH> __device__ __forceinline__ R scene_query( const SceneView<R, SH>& sv, const Ray<R>& ray, const int flags, const R t_far,
                                                                            Trans<R>* trans, HitCtx ctx, const CsgMem<R>& cm )
{
    // t_far: hits at a >= t_far are of no interest to the caller (the shadow test's light distance, a path
    // ray's max_path_length, +inf otherwise).  Without Q_TRANS a hit with a <= t_far ends the search at once.
    // Bounds are culled against the best candidate so far: nothing inside a ball that the ray enters
    // at t_env can be reported closer than t_env - eps, so skipping it when t_env > horizon + 2 eps changes nothing.
    //
    // The walk follows the threaded traversal records (CRec): one record per iteration, the position of a lane is one
    // index.  It is LOCKSTEP over the lanes that enter the query together: a warp reduction keeps everybody in the loop
    // until the last lane is through.  (Written as a plain while-loop with `continue`s the lanes drifted apart for good — a lane
    // whose ray misses a bound went round the loop on its own while its neighbours tested a sphere, and from then on
    // each drift group ran the same instructions at different times.)
    const R inf = Num<R>::inf();
    const bool want_trans = ( flags & Q_TRANS ) != 0;
    const R slack = R( 2 ) * sv.eps;
    const unsigned int mask = __activemask();
    R best = inf;
    bool found = false;                                     // any-hit search satisfied
    #pragma unroll 1
    for( int pass = 0; pass < 2; pass++ )
    {
        bool act = !found && ( flags & ( pass == 0 ? Q_LIGHT : Q_MATTER ) ) != 0;
        const int root = pass == 0 ? sv.light_root : sv.matter_root;
        const I4 rl = sv.link[ root ];
        const R far0 = r_min( t_far, best );                 // strict '<' between the roots: matter must beat the lights
        if( act && ( node_flags( rl ) & F_ENV ) && !envelope_hits_before( sv.env[ root ], ray, far0 + slack ) ) act = false;
        int cur = act ? ( pass == 0 ? sv.rec_light : sv.rec_matter ) : -1;
        R min_a = inf;
        Trans<R> tl; tl.exit_obj = tl.enter_obj = -1; tl.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
        // closest hit inside the current nested element of the root (a compound among the root's elements counts as ONE
        // element whose hit is the plain closest hit of its contents, compound.c:215-244)
        R el_a = inf; V3<R> el_n = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); int el_obj = -1;
        bool nested = false;
        for( ;; )
        {
            const bool more = !found && cur >= 0;
            // the records are laid out in pre-order (links point forward): the lanes at the LOWEST record take a step, a lane
            // that skipped a subtree waits at its next record until the others arrive or pass.  Lanes that meet the same
            // composite object then run its event sweep side by side (hanging_lamps_in_row: the sweep ran at 2.3 lanes
            // of 32 when every lane stepped in every iteration and the lanes spread over the records).
#ifndef ACN_WALK_EVERY_LANE
            const unsigned int low = __reduce_min_sync( mask, more ? ( unsigned int )cur : 0xFFFFFFFFu );
            if( low == 0xFFFFFFFFu ) break;
            if( more && ( unsigned int )cur == low )
#else
            if( !__any_sync( mask, more ) ) break;
            if( more )
#endif
            {
                const CRec<R> rec = load_rec( sv, cur );
                const I4 lk = rec.link;
                const int c = lk.w, fl = node_flags( lk );
                cur = lk.z;
                if( want_trans && nested && ( fl & F_TOP ) )
                {   // back among the root's elements: the nested element is through
                    trans_commit( sv, ray, el_a, el_n, el_obj, &min_a, &tl ); el_a = inf; el_obj = -1; nested = false;
                }
                // horizon: an element must come within eps of the root's minimum to matter (merge rule);
                // inside a nested element it must also beat that element's own minimum
                const R hor = ( want_trans ? r_min( r_min( min_a + sv.eps, el_a ), far0 ) : r_min( min_a, far0 ) ) + slack;
                // rec.env: the element's cull bound — a tight ball round the contents (CullBounds; F_SELF: the sphere itself;
                // F_ENV2: the reference's envelope must be met as well) or the reference's envelope
                bool ok = !( fl & ( F_ENV | F_SELF ) ) || envelope_hits_before( rec.env, ray, hor );
                if( ok && ( fl & F_ENV2 ) ) ok = envelope_hits( sv.env[ c ], ray );      // the ball sticks out of the envelope: the envelope is a gate only
                if( ok )
                {
                    if( node_kind( lk ) == K_COMPOUND || node_kind( lk ) == K_GROUP )
                    {
                        cur = lk.y;                          // first record of its list (its skip record when the list is empty); first member of a group
                        if( ( fl & F_TOP ) && node_kind( lk ) == K_COMPOUND ) nested = true;
                    }
                    else
                    {
                        V3<R> n;
                        R a;
                        if( fl & F_SELF )
                        {
                            a = sphere_hit<R>( xyz( rec.env ), rec.env.w, ray, sv.eps, want_trans ? &n : nullptr );
                            if( want_trans && ( fl & F_ROUGH ) && a < inf ) roughen( sv, c, ray, a, &n, ctx );
                        }
                        else a = elem_hit<R, MARCH>( sv, lk, c, ray, want_trans ? &n : nullptr, ctx, cm, hor );
                        if( !want_trans )
                        {
                            if( a < min_a ) { min_a = a; if( a <= t_far ) found = true; }
                        }
                        else if( nested )
                        {
                            if( a < el_a ) { el_a = a; el_n = n; el_obj = c; }
                        }
                        else trans_commit( sv, ray, a, n, c, &min_a, &tl );
                    }
                }
            }
        }
        if( want_trans && nested ) trans_commit( sv, ray, el_a, el_n, el_obj, &min_a, &tl );      // the root list ended inside a nested element
        if( min_a < best ) { best = min_a; if( want_trans ) *trans = tl; }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// The same walk one record at a time, for the kernels that keep 32 traversals going per warp and hand a lane a new ray
// the moment its old one is through (acn_kernels.cuh: refill loops).  The position of a traversal is `cur`; a lane can
// be interrupted between any two records.
//   walk_any_step     "is anything at a <= t_far?" (shadow rays, probes): returns the next record, WALK_END when the
//                     list is through without a find, WALK_FOUND at the first element hit at a <= t_far
//   WalkT, walk_step  closest hit with the eps-merge of the root's elements (Q_TRANS), or a probe: returns the next record,
//                     WALK_END or WALK_FOUND; walk_trans_finish closes a nested element that reached the end of the root list
// Same tests in the same order as scene_query, which the specialised builds and k_primary keep using.
// ---------------------------------------------------------------------------------------------
enum { WALK_END = -1, WALK_FOUND = -2 };

// In a refill loop 28 of 32 lanes test a record in every iteration, and whatever only the lanes do whose ray PASSED the
// bound runs at two or three lanes: thirty instructions of "compute the sphere hit and its normal" for one record in ten
// cost more issue slots than the bound tests of all of them (ncu, many_spheres).  The steps below therefore keep the
// common cases free of divergent branches: the ray / ball arithmetic is done once per record (RayBall), bound test,
// descent into a compound and a sphere held in its record (F_SELF) are selects on it, and the normal of a sphere inside
// a nested element is not computed when it becomes the element's closest hit but when the element is committed
// (WalkT::el_rec).  Everything else — planes, quadrics, composite objects, rough spheres, spheres among the root's own
// elements — takes the general path.
template <typename R> struct RayBall
{
    R s, q, disc, eps;                      // p.d, |p|^2 - r^2, r^2 - |p - (p.d) d|^2 (gmath.h:64-97, rejection-vector form)
    V3<R> p;
    __device__ __forceinline__ RayBall( const R4<R>& b, const Ray<R>& ray, R eps_ ) : eps( eps_ )
    {
        p = ray.p - xyz( b );
        s = dot( p, ray.d );
        q = sqr( p ) - b.w * b.w;
        const V3<R> l = p - ray.d * s;
        disc = b.w * b.w - sqr( l );
    }
    // envelope_hits_before
    __device__ __forceinline__ bool before( R t_far ) const
    {
        const R g = -s - t_far;
        return disc >= R( 0 ) && ( q < R( 0 ) || ( s < R( 0 ) && !( g > R( 0 ) && g * g > disc ) ) );
    }
    // sphere_hit without the normal
    __device__ __forceinline__ R hit() const
    {
        const R sq = r_sqrt( r_max( disc, R( 0 ) ) );
        const R a = ( s < R( 0 ) && q > R( 0 ) ) ? -s - sq - eps : -s + sq - eps;
        return ( disc >= R( 0 ) && ( s < R( 0 ) || q < R( 0 ) ) ) ? a : Num<R>::inf();
    }
    // is the sphere's hit distance <= x (STRICT: < x)?  No square root: entry -s - sq <= x + eps  <=>  g <= 0 or g^2 <= disc
    // with g = -s - x - eps; exit -s + sq <= x + eps  <=>  h >= 0 and disc <= h^2 with h = x + eps + s
    template <bool STRICT> __device__ __forceinline__ bool hit_within( R x ) const
    {
        const R h = x + eps + s;            // = -g
        const bool entry = s < R( 0 ) && q > R( 0 );
        const bool any = disc >= R( 0 ) && ( s < R( 0 ) || q < R( 0 ) );
        const bool e_in = STRICT ? ( h > R( 0 ) || h * h < disc ) : ( h >= R( 0 ) || h * h <= disc );
        const bool x_in = STRICT ? ( h > R( 0 ) && disc < h * h ) : ( h >= R( 0 ) && disc <= h * h );
        return any && ( entry ? e_in : x_in );
    }
};

// STRICT: a find needs a < t_far (probes, whose callers test a < t_lim) instead of a <= t_far (shadow rays: scene.c:569 min > a)
template <typename R, int MARCH, bool SH, bool STRICT> __device__ __forceinline__ int walk_any_step( const SceneView<R, SH>& sv, const Ray<R>& ray, const R t_far, const int cur,
                                                                                               HitCtx ctx, const CsgMem<R>& cm )
{
    const CRec<R> rec = load_rec( sv, cur );
    const I4 lk = rec.link;
    const int c = lk.w, fl = node_flags( lk );
    const R hor = t_far + R( 2 ) * sv.eps;
    const RayBall<R> rb( rec.env, ray, sv.eps );
    bool ok = !( fl & ( F_ENV | F_SELF ) ) || rb.before( hor );
    if( ok && ( fl & F_ENV2 ) ) ok = envelope_hits( sv.env[ c ], ray );
    const bool comp = node_kind( lk ) == K_COMPOUND || node_kind( lk ) == K_GROUP, self = ( fl & F_SELF ) != 0;
    int next = ( ok && comp ) ? lk.y : lk.z;
    if( ok && self && rb.template hit_within<STRICT>( t_far ) ) next = WALK_FOUND;
    if( ok && !comp && !self )
    {
        const R a = elem_hit<R, MARCH>( sv, lk, c, ray, ( V3<R>* )nullptr, ctx, cm, hor );
        if( STRICT ? a < t_far : ( a <= t_far && a < Num<R>::inf() ) ) next = WALK_FOUND;
    }
    return next;
}

template <typename R> struct WalkT
{
    R min_a, el_a; V3<R> el_n; int el_obj, el_rec; Trans<R> tl; bool nested;
    __device__ __forceinline__ void reset()
    {
        min_a = el_a = Num<R>::inf(); el_n = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); el_obj = el_rec = -1; nested = false;
        tl.exit_obj = tl.enter_obj = -1; tl.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    }
};

// the nested element of the root is through: its closest hit becomes one candidate of the root (compound.c:246-299).
// el_rec >= 0: that hit is a sphere held in its record, whose normal was left for now (sphere_ray_hit, gmath.h:64-97).
template <typename R, bool SH> __device__ __forceinline__ void walk_commit_nested( const SceneView<R, SH>& sv, const Ray<R>& ray, WalkT<R>& s )
{
    if( s.el_rec >= 0 && s.el_a < Num<R>::inf() )
    {
        const CRec<R> rec = sv.crec[ s.el_rec ];
        s.el_n = unit( madd( ray.p - xyz( rec.env ), ray.d, s.el_a ) );
        s.el_obj = rec.link.w;
    }
    trans_commit( sv, ray, s.el_a, s.el_n, s.el_obj, &s.min_a, &s.tl );
    s.el_a = Num<R>::inf(); s.el_obj = s.el_rec = -1; s.nested = false;
}

// one record of a walk that is a probe (STRICT any-hit) in some lanes and a closest-hit search with transitions in others.
// far0: min( the caller's t_far, the closest light hit ) — matter must beat the lights (scene.c:362-382)
template <typename R, int MARCH, bool SH> __device__ __forceinline__ int walk_step( const SceneView<R, SH>& sv, const Ray<R>& ray, const R far0, const bool want_trans, const int cur,
                                                                               WalkT<R>& s, HitCtx ctx, const CsgMem<R>& cm )
{
    const CRec<R> rec = load_rec( sv, cur );
    const I4 lk = rec.link;
    const int c = lk.w, fl = node_flags( lk );
    if( want_trans && s.nested && ( fl & F_TOP ) ) walk_commit_nested( sv, ray, s );         // back among the root's elements
    const R hor = ( want_trans ? r_min( r_min( s.min_a + sv.eps, s.el_a ), far0 ) : far0 ) + R( 2 ) * sv.eps;
    const RayBall<R> rb( rec.env, ray, sv.eps );
    bool ok = !( fl & ( F_ENV | F_SELF ) ) || rb.before( hor );
    if( ok && ( fl & F_ENV2 ) ) ok = envelope_hits( sv.env[ c ], ray );
    const bool comp = node_kind( lk ) == K_COMPOUND || node_kind( lk ) == K_GROUP;
    // a sphere held in its record whose hit needs no normal now: a probe, or inside a nested element
    const bool leaf = ( fl & ( F_SELF | F_ROUGH ) ) == F_SELF && ( !want_trans || s.nested );
    int next = ( ok && comp ) ? lk.y : lk.z;
    if( ok && ( fl & F_TOP ) && node_kind( lk ) == K_COMPOUND ) s.nested = true;
    const R a_leaf = rb.hit();
    if( ok && leaf )
    {
        if( !want_trans ) { if( a_leaf < far0 ) next = WALK_FOUND; }
        else if( a_leaf < s.el_a ) { s.el_a = a_leaf; s.el_rec = cur; }
    }
    if( ok && !comp && !leaf )
    {
        V3<R> n = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
        R a;
        if( fl & F_SELF )
        {
            a = sphere_hit<R>( xyz( rec.env ), rec.env.w, ray, sv.eps, &n );
            if( want_trans && ( fl & F_ROUGH ) && a < Num<R>::inf() ) roughen( sv, c, ray, a, &n, ctx );
        }
        else a = elem_hit<R, MARCH>( sv, lk, c, ray, want_trans ? &n : ( V3<R>* )nullptr, ctx, cm, hor );
        if( !want_trans ) { if( a < far0 ) next = WALK_FOUND; }
        else if( s.nested ) { if( a < s.el_a ) { s.el_a = a; s.el_n = n; s.el_obj = c; s.el_rec = -1; } }
        else trans_commit( sv, ray, a, n, c, &s.min_a, &s.tl );
    }
    return next;
}

template <typename R, bool SH> __device__ __forceinline__ void walk_trans_finish( const SceneView<R, SH>& sv, const Ray<R>& ray, WalkT<R>& s )
{
    if( s.nested ) walk_commit_nested( sv, ray, s );        // the root list ended inside a nested element
}

// may the walk of a root compound start at all?  (the root's own envelope, culled against the horizon)
// front: walk the copy of the matter records laid out front to back for the ray's octant (closest-hit searches; the shadow
// rays gain nothing from it — measured — and keep to the one canonical copy, which they then have the L1 for)
template <typename R, bool SH> __device__ __forceinline__ int walk_root( const SceneView<R, SH>& sv, const Ray<R>& ray, const bool light, const R far0, const bool front = true )
{
    const int root = light ? sv.light_root : sv.matter_root;
    const I4 rl = sv.link[ root ];
    if( ( node_flags( rl ) & F_ENV ) && !envelope_hits_before( sv.env[ root ], ray, far0 + R( 2 ) * sv.eps ) ) return WALK_END;
    if( light ) return sv.rec_light;
    if( !front ) return sv.rec_matter;
    // the copy of the matter records that is laid out front to back for this ray's octant
    return sv.rec_matter_oct[ ( ray.d.x < R( 0 ) ? 1 : 0 ) | ( ray.d.y < R( 0 ) ? 2 : 0 ) | ( ray.d.z < R( 0 ) ? 4 : 0 ) ];
}

// ---------------------------------------------------------------------------------------------
// Scene-specialised build (NVRTC, acn_spec.h): the generated header restates the structure of ONE scene — its top-level
// element lists and its CSG programs — as straight-line code over the same leaf functions, and replaces scene_query.
// ---------------------------------------------------------------------------------------------
#if defined(ACN_SPEC)
// the quadric leaf code once per kernel, out of line (everything by value: see dist_hit_ool)
template <typename R> struct Ev2 { R t0, t1; int c, s0; };
template <typename R, bool SH> __device__ __noinline__ Ev2<R> squaroid_events_ool_( SceneView<R, SH> sv, int n, Ray<R> ray )
{
    Ev2<R> e; e.t0 = e.t1 = R( 0 ); e.s0 = 0;
    e.c = leaf_events<false>( sv, K_SQUAROID, n, ray, &e.s0, &e.t0, &e.t1 );
    return e;
}
template <typename R, bool SH> __device__ __forceinline__ int squaroid_events_ool( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, int* s0, R* t0, R* t1 )
{
    const Ev2<R> e = squaroid_events_ool_( sv, n, ray );
    *s0 = e.s0; *t0 = e.t0; *t1 = e.t1;
    return e.c;
}
template <typename R, bool SH> __device__ __noinline__ HitN<R> squaroid_hit_ool( SceneView<R, SH> sv, int n, Ray<R> ray, bool want_nor )
{
    HitN<R> h; h.n = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    h.a = prim_hit( sv, K_SQUAROID, n, ray, want_nor ? &h.n : ( V3<R>* )nullptr );
    return h;
}
#include "acn_spec_gen.h"
#endif

template <typename R, int MARCH, bool SH> __device__ __forceinline__ R query( const SceneView<R, SH>& sv, const Ray<R>& ray, const int flags, const R t_far,
                                                                      Trans<R>* trans, HitCtx ctx, const CsgMem<R>& cm )
{
#if defined(ACN_SPEC_SCENE)
    return spec_scene_query<R, MARCH, SH>( sv, ray, flags, t_far, trans, ctx, cm );
#else
    return scene_query<R, MARCH, SH>( sv, ray, flags, t_far, trans, ctx, cm );
#endif
}

} // namespace acn
