// acn_isect.cuh — device-only scene queries of the wavefront tracer.
//
//   csg_eval      composite objects (obj_pair_inside_s / obj_pair_outside_s / obj_neg_s over plane, sphere
//                 and squaroid leaves) by interval lists held in SHARED memory
//   scene_query   ONE traversal routine for every kind of ray: scene_s_trans_hit (scene.c:362-382),
//                 compound_s_ray_trans_hit / compound_s_ray_hit (compound.c:215-299) and the shadow /
//                 probe "anything closer than t?" tests, selected by flags, so that each kernel holds a
//                 single copy of the intersection code (instruction-cache footprint) and none of it
//                 needs a call stack.
#pragma once

#include "acn_geom.h"

namespace acn {

// ---------------------------------------------------------------------------------------------
// CSG by interval lists.  The reference finds the first boundary of A&B / A|B by an alternating
// march over the children (objects.c:1052-1094,1209-1251), re-tracing whole subtrees for every
// rejected candidate: O(n^2) ray tests for an n-leaf solid and hopelessly divergent on a GPU.
// The same first boundary comes out of classifying the ray against every leaf ONCE:
//   * a leaf yields its state at the origin and its (<= 2) crossings for t > 0;
//   * '!' complements, '&' intersects, '|' unites the in/out state sequences (merge of two short
//     sorted lists, keeping only the crossings where the combined state flips);
//   * a node's envelope clips its inside-set (obj_side reports "outside" beyond the own envelope,
//     objects.c:365-370, also for negations) with VIRTUAL crossings that are never reported as hits
//     (obj_ray_hit only returns shape crossings), and gates the whole subtree (objects.c:264).
// A crossing of leaf X survives to the root exactly when every sibling on the way up is in the state
// the pair demands — the reference's acceptance test.  Programs are postfix; chains of the same
// operator are re-associated to the left (the solid is the same point set), so the evaluation stack
// stays at 2-3 lists however many facets an object has.
//
// The lists live in shared memory, element (slot, k) of thread tid at [(slot*CSG_K + k)*nthreads + tid]:
// conflict-free, and no local-memory frame (the previous register-array version cost 856 B of stack
// per thread and thrashed L1).  Crossing ids are the program-relative offset of the leaf instruction
// (one byte; CSG_VIRTUAL marks envelope crossings).
// ---------------------------------------------------------------------------------------------
template <typename R> struct CsgMem { R* t; unsigned char* id; int stride; };

template <typename R> __host__ __device__ inline size_t csg_mem_bytes( int nthreads )
{
    return ( size_t )( sizeof( R ) + 1 ) * CSG_SLOTS * CSG_K * nthreads;
}

template <typename R> __device__ __forceinline__ CsgMem<R> csg_mem( unsigned char* base, int nthreads, int tid )
{
    CsgMem<R> m;
    m.t = reinterpret_cast<R*>( base ) + tid;
    m.id = base + sizeof( R ) * CSG_SLOTS * CSG_K * nthreads + tid;
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

template <typename R> __device__ __forceinline__ int leaf_events( const SceneView<R>& sv, int kind, int n, const Ray<R>& ray, int* s0, R* t0, R* t1 )
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
    R fi = R( 1 ) / f;
    R tm = -fs * fi;
    V3<R> pm = madd( p, d, tm );
    R s = dot( ad, pm ) * fi;
    R q = ( qa * pm.x * pm.x + qb * pm.y * pm.y + qc * pm.z * pm.z + qr ) * fi;
    R r = s * s - q;
    if( r < R( 0 ) ) return 0;
    r = r_sqrt( r );
    R ta = tm - s - r, tb = tm - s + r;
    int c = 0;
    if( ta >= R( 0 ) ) { *t0 = ta; c = 1; }
    if( tb >= R( 0 ) ) { if( c ) *t1 = tb; else *t0 = tb; c++; }
    return c;
}

// outward normal of a leaf at ray parameter t (the unshortened crossing), as its fp_ray_hit reports it
template <typename R> __device__ __forceinline__ V3<R> leaf_normal( const SceneView<R>& sv, int kind, int n, const Ray<R>& ray, R t )
{
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );
    if( kind == K_PLANE ) return xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
    if( kind == K_SPHERE ) return unit( madd( ray.p - pos, ray.d, t - sv.eps ) );
    const M3<R> rax = node_rax( sv, n );
    V3<R> p = mlv( rax, ray.p - pos );
    V3<R> d = mlv( rax, ray.d );
    V3<R> x = madd( p, d, t );
    return unit( tmlv( rax, v3<R>( x.x * g0.w, x.y * sv.geo[ n * GEO_STRIDE + 1 ].w, x.z * sv.geo[ n * GEO_STRIDE + 2 ].w ) ) );
}

#define ACN_CSG_T( ph, k )  cm.t[ ( ( ph ) * CSG_K + ( k ) ) * cm.stride ]
#define ACN_CSG_ID( ph, k ) cm.id[ ( ( ph ) * CSG_K + ( k ) ) * cm.stride ]

// returns the hit parameter, +inf for a miss, or a value below -mag when a list overflowed before the
// first crossing was found (the caller then falls back to the reference march)
template <typename R> __device__ __forceinline__ R csg_eval( const SceneView<R>& sv, int root, const Ray<R>& ray, V3<R>* nor, HitCtx ctx,
                                                             const CsgMem<R>& cm )
{
    const R inf = Num<R>::inf();
    // registers: list headers (count 4 bits | origin state 1 bit, 6 bits per physical slot), the
    // logical -> physical slot permutation (4 bits each; logical CSG_S is the scratch slot)
    unsigned int hdr = 0, perm = 0x43210u;
    int sp = 0;
    R t_valid = inf;
    const int start = sv.prog_ref[ 2 * root ], len = sv.prog_ref[ 2 * root + 1 ];
    #pragma unroll 1
    for( int pc = start; pc < start + len; pc++ )
    {
        const int ins = sv.prog[ pc ];
        const int op = ins & 15, n = ins >> 4;
        bool merge = false, op_and = true;
        if( op == CSG_LEAF || op == CSG_CLIP )
        {
            R t0 = R( 0 ), t1 = R( 0 ); int s0, c;
            int id = CSG_VIRTUAL;
            if( op == CSG_LEAF ) { c = leaf_events( sv, node_kind( sv.link[ n ] ), n, ray, &s0, &t0, &t1 ); id = pc - start; }
            else { const R4<R> e = sv.env[ n ]; c = sphere_events( xyz( e ), e.w, ray, &s0, &t0, &t1 ); merge = true; }
            const unsigned int ph = ( perm >> ( 4 * sp ) ) & 15u;
            if( c > 0 ) { ACN_CSG_T( ph, 0 ) = t0; ACN_CSG_ID( ph, 0 ) = ( unsigned char )id; }
            if( c > 1 ) { ACN_CSG_T( ph, 1 ) = t1; ACN_CSG_ID( ph, 1 ) = ( unsigned char )id; }
            hdr = ( hdr & ~( 63u << ( 6 * ph ) ) ) | ( ( unsigned int )( c | ( s0 << 4 ) ) << ( 6 * ph ) );
            sp++;
        }
        else if( op == CSG_NEG )
        {
            const unsigned int ph = ( perm >> ( 4 * ( sp - 1 ) ) ) & 15u;
            hdr ^= 16u << ( 6 * ph );
        }
        else if( op == CSG_ENV )
        {
            const int skip = sv.prog[ ++pc ];
            if( !envelope_hits( sv.env[ n ], ray ) )
            {
                const unsigned int ph = ( perm >> ( 4 * sp ) ) & 15u;
                hdr &= ~( 63u << ( 6 * ph ) );          // empty list, outside at the origin
                sp++; pc += skip;
            }
        }
        else { merge = true; op_and = op == CSG_AND; }

        if( merge )     // combine the two topmost lists into the scratch slot, which becomes the new top
        {
            const unsigned int pa = ( perm >> ( 4 * ( sp - 2 ) ) ) & 15u, pb = ( perm >> ( 4 * ( sp - 1 ) ) ) & 15u, po = ( perm >> ( 4 * CSG_S ) ) & 15u;
            const int na = ( hdr >> ( 6 * pa ) ) & 15, nb = ( hdr >> ( 6 * pb ) ) & 15;
            int sa = ( hdr >> ( 6 * pa + 4 ) ) & 1, sb = ( hdr >> ( 6 * pb + 4 ) ) & 1;
            int s = op_and ? ( sa & sb ) : ( sa | sb );
            const int s_init = s;
            int i = 0, j = 0, o = 0;
            R ta = na > 0 ? ACN_CSG_T( pa, 0 ) : inf, tb = nb > 0 ? ACN_CSG_T( pb, 0 ) : inf;
            R t_last = R( 0 );
            while( i < na || j < nb )
            {
                const bool take_a = j >= nb || ( i < na && ta <= tb );
                R t; unsigned char id;
                if( take_a ) { t = ta; id = ACN_CSG_ID( pa, i ); i++; sa ^= 1; ta = i < na ? ACN_CSG_T( pa, i ) : inf; }
                else         { t = tb; id = ACN_CSG_ID( pb, j ); j++; sb ^= 1; tb = j < nb ? ACN_CSG_T( pb, j ) : inf; }
                const int s2 = op_and ? ( sa & sb ) : ( sa | sb );
                if( s2 != s )
                {
                    if( o < CSG_K ) { ACN_CSG_T( po, o ) = t; ACN_CSG_ID( po, o ) = id; t_last = t; }
                    o++; s = s2;
                }
            }
            if( o > CSG_K ) { o = CSG_K; t_valid = r_min( t_valid, t_last ); }   // exact up to the last kept crossing
            hdr = ( hdr & ~( 63u << ( 6 * po ) ) ) | ( ( unsigned int )( o | ( s_init << 4 ) ) << ( 6 * po ) );
            // logical sp-2 <- scratch; scratch <- old physical slot of A
            perm = ( perm & ~( ( 15u << ( 4 * ( sp - 2 ) ) ) | ( 15u << ( 4 * CSG_S ) ) ) ) | ( po << ( 4 * ( sp - 2 ) ) ) | ( pa << ( 4 * CSG_S ) );
            sp--;
        }
    }
    // first real crossing of the root list
    const unsigned int p0 = perm & 15u;
    const int n0 = ( hdr >> ( 6 * p0 ) ) & 15;
    for( int k = 0; k < n0; k++ )
    {
        const int id = ACN_CSG_ID( p0, k );
        if( id == CSG_VIRTUAL ) continue;
        const R t = ACN_CSG_T( p0, k );
        if( t > t_valid ) break;
        const R a = t - sv.eps;
        if( nor )
        {
            const int leaf = sv.prog[ start + id ] >> 4;
            V3<R> nn = leaf_normal( sv, node_kind( sv.link[ leaf ] ), leaf, ray, t );
            // up the tree: roughness at every level that has it (objects.c:266), sign flip at negations
            for( int m = leaf; m != root && m >= 0; m = sv.parent[ m ] )
            {
                const I4 lk = sv.link[ m ];
                if( node_kind( lk ) == K_NEG ) nn = -nn;
                if( node_flags( lk ) & F_ROUGH ) roughen( sv, m, ray, a, &nn, ctx );
            }
            if( node_kind( sv.link[ root ] ) == K_NEG ) nn = -nn;
            *nor = nn;
        }
        return a;
    }
    if( t_valid < inf ) return R( -2 ) * Num<R>::mag();
    return inf;
}

// objects the interval evaluator does not cover (distance fields, scale nodes, CSG over those, list
// overflow): the reference's recursive march, out of line
template <typename R> __device__ __noinline__ R march_hit( const SceneView<R>& sv, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    const I4 lk = sv.link[ n ];
    return shape_hit( sv, lk, n, ray, nor, ctx );
}

// obj_ray_hit after the envelope test (objects.c:261-284): fp_ray_hit + roughness
template <typename R> __device__ __forceinline__ R elem_hit( const SceneView<R>& sv, const I4& lk, int c, const Ray<R>& ray, V3<R>* nor, HitCtx ctx,
                                                             const CsgMem<R>& cm )
{
    const int kind = node_kind( lk );
    R a;
    if( kind == K_PLANE || kind == K_SPHERE || kind == K_SQUAROID ) a = prim_hit( sv, kind, c, ray, nor );
    else
    {
        a = R( -2 ) * Num<R>::mag();
        if( kind >= K_PAIR_INSIDE && sv.prog_ref[ 2 * c + 1 ] > 0 ) a = csg_eval( sv, c, ray, nor, ctx, cm );
        if( !( a > -Num<R>::mag() ) ) a = march_hit( sv, c, ray, nor, ctx );
    }
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

template <typename R> __device__ __forceinline__ void trans_commit( const SceneView<R>& sv, const Ray<R>& ray, R a, V3<R> nor, int obj, R* min_a, Trans<R>* tl )
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

template <typename R> __device__ __forceinline__ R scene_query( const SceneView<R>& sv, const Ray<R>& ray, const int flags, const R t_any,
                                                                Trans<R>* trans, HitCtx ctx, const CsgMem<R>& cm )
{
    const R inf = Num<R>::inf();
    const bool want_trans = ( flags & Q_TRANS ) != 0;
    R best = inf;
    #pragma unroll 1
    for( int pass = 0; pass < 2; pass++ )
    {
        if( !( flags & ( pass == 0 ? Q_LIGHT : Q_MATTER ) ) ) continue;
        const int root = pass == 0 ? sv.light_root : sv.matter_root;
        const I4 rl = sv.link[ root ];
        if( ( node_flags( rl ) & F_ENV ) && !envelope_hits( sv.env[ root ], ray ) ) continue;
        int sb[ COMPOUND_STACK ], se[ COMPOUND_STACK ];
        int sp = 0;
        int beg = rl.y, end = rl.y + rl.z;
        R min_a = inf;
        Trans<R> tl; tl.exit_obj = tl.enter_obj = -1; tl.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
        R el_a = inf; V3<R> el_n = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); int el_obj = -1;     // closest hit inside the current nested element
        for( ;; )
        {
            while( beg < end )
            {
                const int c = sv.children[ beg++ ];
                const I4 lk = sv.link[ c ];
                if( ( node_flags( lk ) & F_ENV ) && !envelope_hits( sv.env[ c ], ray ) ) continue;
                if( node_kind( lk ) == K_COMPOUND )
                {
                    if( sp < COMPOUND_STACK ) { sb[ sp ] = beg; se[ sp ] = end; sp++; beg = lk.y; end = lk.y + lk.z; }
                    continue;
                }
                V3<R> n;
                const R a = elem_hit( sv, lk, c, ray, want_trans ? &n : nullptr, ctx, cm );
                if( !want_trans )
                {
                    if( a < min_a ) { min_a = a; if( a <= t_any ) return a; }
                }
                else if( sp > 0 )
                {
                    if( a < el_a ) { el_a = a; el_n = n; el_obj = c; }
                }
                else trans_commit( sv, ray, a, n, c, &min_a, &tl );
            }
            if( sp == 0 ) break;
            sp--; beg = sb[ sp ]; end = se[ sp ];
            if( sp == 0 && want_trans ) { trans_commit( sv, ray, el_a, el_n, el_obj, &min_a, &tl ); el_a = inf; el_obj = -1; }
        }
        if( min_a < best ) { best = min_a; if( want_trans ) *trans = tl; }
    }
    return best;
}

} // namespace acn
