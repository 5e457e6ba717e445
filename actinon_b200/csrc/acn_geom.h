// acn_geom.h — packed scene view + ray/shape intersection, CSG evaluation and compound traversal.
//
// Host+device templates on the real type R.  On the device this is the intersection stage of the
// wavefront tracer (R = float, or double in validation mode); on the host (R = double) it serves
// scene construction only (Monte-Carlo envelope estimation, objects.c:312-363) — never rendering.
//
// Node table layout (one entry per compound_s / obj_*_s, indices instead of pointers):
//   env [n]      R4  envelope centre xyz, radius              (objects.c:35-40)
//   link[n]      I4  kind | flags<<8, a, b, material          a,b: CSG children, or compound child range
//   geo [5n+0]   R4  pos.xyz, tail0
//   geo [5n+1]   R4  rax.x (row), tail1
//   geo [5n+2]   R4  rax.y (row), tail2
//   geo [5n+3]   R4  rax.z (row), tail3
//   geo [5n+4]   R4  surface_roughness, -, -, -
// A sphere test touches env+link+geo0 = 48 B (f32); a plane adds geo3.
#pragma once

#include "acn_math.h"

namespace acn {

template <typename R> struct alignas( sizeof( R ) * 4 ) R4 { R x, y, z, w; };
struct alignas( 16 ) I4 { int x, y, z, w; };
// one element of a child list, packed for the traversal: its cull bound (envelope or tighter ball) and
// link = ( kind | flags << 8, first record of the element's own list (compounds), skip record, node index );
// skip = the record a ray goes to when it is through with this element: the next sibling or, behind the last one of
// a list, the skip record of the list's compound (-1: end of the root list).  The records of a scene form a tree
// laid out depth-first (acn_tracer.cuh), so a traversal is one index: no stack.
template <typename R> struct CRec { R4<R> env; I4 link; };

enum
{
    K_COMPOUND = 0, K_PLANE = 1, K_SPHERE = 2, K_SQUAROID = 3, K_DIST_SPHERE = 4, K_DIST_TORUS = 5,
    K_PAIR_INSIDE = 6, K_PAIR_OUTSIDE = 7, K_NEG = 8, K_SCALE = 9,
    K_GROUP = 10        // traversal records only: a pure bound over a run of consecutive elements of a list (acn_tracer.cuh)
};
// flags of a node; the last three occur in traversal records only (acn_tracer.cuh: CullBounds, threaded records) —
// F_SELF: the record's ball IS the sphere; F_ENV2: test env[node] as well; F_TOP: element of a root compound
enum { F_ENV = 1, F_ROUGH = 2, F_SELF = 4, F_ENV2 = 8, F_TOP = 16 };
enum { GEO_STRIDE = 5 };
enum { SEED_POSITION_HASH = 0, SEED_INDEX_KEYED = 1 };
enum { CSG_MAX_STEPS = 512, COMPOUND_STACK = 16 };
enum { CSG_LEAF = 0, CSG_NEG = 1, CSG_AND = 2, CSG_OR = 3, CSG_CLIP = 4, CSG_ENV = 5, CSG_RUN = 6, CSG_MEMBER = 7, CSG_MEMBER_NEG = 8, CSG_MORE = 9, CSG_XFORM = 10, CSG_XEND = 11 };   // word = op | arg << 4, see acn_isect.cuh
enum { E_TAB = 0, E_VAR = 1, E_CLIP = 2, E_NEG = 3, E_AND = 4, E_OR = 5 };      // words of an evaluation program (acn_tracer.cuh: build_eval_program)
enum { CSG_E = 16, CSG_VIRTUAL = 255, CSG_MAX_VARS = 64, CSG_TABLE_VARS = 12 };   // crossings kept per ray, id of envelope crossings, variable limits

// A scene table: element i by value.  SH = false: a pointer (device global memory, or host memory in scene
// construction).  SH = true: the table was staged into shared memory and is addressed by its 32-bit
// shared-window byte address — the compiler keeps such bases in uniform registers and emits LDS, so the
// seven tables cost the tracing kernels no vector registers (as generic 64-bit pointers they were spilled
// to local memory at the 96-register cap and re-read before every table access).
template <typename T, bool SH> struct Tab
{
    const T* p;
    ACN_HD T operator[]( int i ) const { return p[ i ]; }
    ACN_HD Tab& operator=( const T* q ) { p = q; return *this; }
};
#if defined(__CUDACC__)
template <typename T> struct Tab<T, true>
{
    unsigned int a;
    __device__ __forceinline__ T operator[]( int i ) const
    {
        return *reinterpret_cast<const T*>( __cvta_shared_to_generic( a + ( unsigned int )i * ( unsigned int )sizeof( T ) ) );
    }
};
#endif

template <typename R, bool SH = false> struct SceneView
{
    Tab<R4<R>, SH>   env;
    Tab<I4, SH>      link;
    Tab<R4<R>, SH>   geo;
    Tab<int, SH>     children;  // child node indices of all compounds (host / march)
    Tab<CRec<R>, SH> crec;      // the same lists as packed records (device traversal)
    Tab<int, SH>     prog;      // postfix CSG programs (interval evaluator), see acn_isect.cuh: csg_eval
    Tab<I4, SH>      prog_ref;  // per node: program start, length (0: none -> reference march), truth table offset (-1: none), variables
    Tab<int, SH>     parent;    // per node: CSG parent (-1 at the top of an object)
    R   eps;            // shell thickness (f3_eps, vectors.h:33)
    int light_root;
    int matter_root;
    int rec_light;      // first traversal record of each root list (-1: empty)
    int rec_matter;
    int rec_matter_oct[ 8 ];    // the matter list front to back for each octant of ray directions (acn_tracer.cuh); = rec_matter when not built
    int seed_mode;
};

template <typename R> struct Ray { V3<R> p, d; };

// per-ray context: key of the ray in the ray tree (index-keyed roughness seeding)
struct HitCtx { u64 key; };

// trans_data_s (compound.h:31-36); -1 = no object
template <typename R> struct Trans { V3<R> exit_nor; int exit_obj; int enter_obj; };

template <typename R> ACN_HD V3<R> xyz( const R4<R>& v ) { return v3<R>( v.x, v.y, v.z ); }
ACN_HD int node_kind( const I4& l )  { return l.x & 0xFF; }
ACN_HD int node_flags( const I4& l ) { return l.x >> 8; }

// ---------------------------------------------------------------------------------------------
// sphere: sphere_ray_hit / sphere_observer_side (gmath.h:64-97) with the discriminant taken from the
// rejection vector (r^2 - |p - (p.d)d|^2), which keeps FP32 accurate when |p| >> r.
// ---------------------------------------------------------------------------------------------
template <typename R> ACN_HD bool envelope_hits( const R4<R>& e, const Ray<R>& ray )                  // objects.c:90-93
{
    V3<R> p = ray.p - xyz( e );
    R s = dot( p, ray.d );
    R q = sqr( p ) - e.w * e.w;
    V3<R> l = p - ray.d * s;
    R disc = e.w * e.w - sqr( l );
    return disc >= R( 0 ) && ( s < R( 0 ) || q < R( 0 ) );
}

// envelope test with a horizon: false as well when the ray enters the envelope beyond t_far (nothing inside can be
// hit before t_far then).  -s - sqrt(disc) > t_far  <=>  -s - t_far > 0 and ( -s - t_far )^2 > disc: no square root.
template <typename R> ACN_HD bool envelope_hits_before( const R4<R>& e, const Ray<R>& ray, R t_far )
{
    V3<R> p = ray.p - xyz( e );
    R s = dot( p, ray.d );
    R q = sqr( p ) - e.w * e.w;
    V3<R> l = p - ray.d * s;
    R disc = e.w * e.w - sqr( l );
    if( !( disc >= R( 0 ) ) ) return false;
    if( q < R( 0 ) ) return true;               // origin inside
    if( !( s < R( 0 ) ) ) return false;
    R g = -s - t_far;
    return !( g > R( 0 ) && g * g > disc );
}

template <typename R> ACN_HD R sphere_hit( V3<R> c, R r, const Ray<R>& ray, R eps, V3<R>* nor )
{
    V3<R> p = ray.p - c;
    R s = dot( p, ray.d );
    R q = sqr( p ) - r * r;
    V3<R> l = p - ray.d * s;
    R disc = r * r - sqr( l );
    if( disc < R( 0 ) ) return Num<R>::inf();
    R offs;
    if( s < R( 0 ) && q > R( 0 ) )      offs = -s - r_sqrt( disc ) - eps;   // entry
    else if( s < R( 0 ) || q < R( 0 ) ) offs = -s + r_sqrt( disc ) - eps;   // exit
    else return Num<R>::inf();
    if( nor ) *nor = unit( madd( p, ray.d, offs ) );
    return offs;
}

template <typename R> ACN_HD int sphere_side( V3<R> c, R r, V3<R> pos ) { return sqr( pos - c ) > r * r ? 1 : -1; }

// ---------------------------------------------------------------------------------------------
// distance functions (distance.c:39-42,83-92)
// ---------------------------------------------------------------------------------------------
template <typename R> ACN_HD R dist_fn( int kind, R ex_radius, V3<R> p )
{
    if( kind == K_DIST_SPHERE ) return r_sqrt( sqr( p ) ) - R( 1 );
    R f = r_sqrt( p.x * p.x + p.y * p.y );
    R fi = f > R( 0 ) ? R( 1 ) / f : R( 1 );
    R x = p.x * fi - p.x, y = p.y * fi - p.y;
    return r_sqrt( x * x + y * y + p.z * p.z ) - ex_radius;
}

// gradient of a distance function at p (object frame): the normal of a distance-field hit before it is rotated back
template <typename R> ACN_HD V3<R> dist_gradient( int kind, R ex_radius, V3<R> p, R eps )
{
    V3<R> g;
    if( sizeof( R ) == 8 )
    {
        // the reference's forward differences with step eps (objects.c:947-953)
        R d0 = dist_fn( kind, ex_radius, p );
        g.x = ( dist_fn( kind, ex_radius, v3<R>( p.x + eps, p.y, p.z ) ) - d0 ) / eps;
        g.y = ( dist_fn( kind, ex_radius, v3<R>( p.x, p.y + eps, p.z ) ) - d0 ) / eps;
        g.z = ( dist_fn( kind, ex_radius, v3<R>( p.x, p.y, p.z + eps ) ) - d0 ) / eps;
    }
    else
    {
        // FP32: a difference quotient over eps would carry ~1e-3 rounding noise; both
        // distance functions have a closed-form gradient, which the quotient approximates
        // to O(eps) — use it.  torus: unit vector from the nearest point of the core circle.
        g = p;
        if( kind == K_DIST_TORUS )
        {
            R f = r_sqrt( p.x * p.x + p.y * p.y );
            R fi = f > R( 0 ) ? R( 1 ) / f : R( 1 );
            g = v3<R>( p.x - p.x * fi, p.y - p.y * fi, p.z );
        }
    }
    return g;
}

template <typename R, bool SH> ACN_HD M3<R> node_rax( const SceneView<R, SH>& sv, int n )
{
    M3<R> m;
    m.x = xyz( sv.geo[ n * GEO_STRIDE + 1 ] );
    m.y = xyz( sv.geo[ n * GEO_STRIDE + 2 ] );
    m.z = xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
    return m;
}

// distance-field objects (objects.c:903-959): sphere tracing in the scaled object frame.  Kept out of line:
// it is the heaviest and most divergent primitive and must not bloat every intersection site.
template <typename R, bool SH> ACN_HD R dist_hit_body( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, V3<R>* nor )
{
    const R inf = Num<R>::inf();
    const R eps = sv.eps;
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );
    const M3<R> rax = node_rax( sv, n );

    const R inv_scale = g0.w;
    const R ex_radius = sv.geo[ n * GEO_STRIDE + 1 ].w;
    const int cycles  = ( int )sv.geo[ n * GEO_STRIDE + 2 ].w;
    const I4 lk = sv.link[ n ];
    Ray<R> rl = ray;
    R offs0 = R( 0 );
    R lim = inf;                // march parameter (object frame) at which the ray has left the object's own envelope
    if( node_flags( lk ) & F_ENV )
    {
        const R4<R> e = sv.env[ n ];
        if( sphere_side( xyz( e ), e.w, ray.p ) == 1 )
        {
            offs0 = sphere_hit<R>( xyz( e ), e.w, ray, eps, nullptr );
            if( !( offs0 < inf ) ) return inf;
            rl.p = madd( ray.p, ray.d, offs0 );
        }
        // The reference marches on for all `cycles` steps after the ray has passed the shape (objects.c:925-933) and
        // then finds |dist| > eps: a miss.  The shape lies inside its envelope, so once the march is beyond the far
        // side of the envelope (+1 % of its radius) the outcome is settled: stop there.  Same result, and a ray that
        // crosses the envelope of a chain link without touching the torus costs ~10 steps instead of 200.
        const V3<R> q = rl.p - xyz( e );
        const R sq = dot( q, rl.d ), dq = e.w * e.w - sqr( q - rl.d * sq );
        if( dq >= R( 0 ) ) lim = ( -sq + r_sqrt( dq ) + R( 0.01 ) * e.w ) * inv_scale;
    }
    V3<R> lp = mlv( rax, rl.p - pos ) * inv_scale;
    V3<R> ld = mlv( rax, rl.d );
    R offs1 = R( 0 );
    R dist = dist_fn( kind, ex_radius, lp );
    if( dist > R( 0 ) )
    {
        for( int i = 0; i < cycles; i++ )
        {
            offs1 += dist + eps;
            dist = dist_fn( kind, ex_radius, madd( lp, ld, offs1 ) );
            if( dist < R( 0 ) || dist > Num<R>::mag() ) break;
            if( offs1 > lim && dist > eps ) break;
        }
    }
    else
    {
        for( int i = 0; i < cycles; i++ )
        {
            offs1 -= dist - eps;
            dist = dist_fn( kind, ex_radius, madd( lp, ld, offs1 ) );
            if( dist > R( 0 ) || dist < -Num<R>::mag() ) break;
        }
    }
    if( r_abs( dist ) <= eps )
    {
        if( nor )
        {
            *nor = unit( tmlv( rax, dist_gradient( kind, ex_radius, madd( lp, ld, offs1 ), eps ) ) );
        }
        return offs0 + offs1 / inv_scale - eps;
    }
    return inf;
}

// Out-of-line entry.  Everything crosses the call BY VALUE: a reference parameter would take the address of the
// caller's ray / scene view / normal, which pins them to local memory for the whole kernel — the hot loops then
// re-read the ray from the stack before every envelope test (seen as LDL in the SASS of every tracing kernel).
template <typename R> struct HitN { R a; V3<R> n; };

template <typename R, bool SH> ACN_NOINLINE HitN<R> dist_hit_ool( SceneView<R, SH> sv, int kind, int n, Ray<R> ray, bool want_nor )
{
    HitN<R> h;
    h.n = v3<R>( R( 0 ), R( 0 ), R( 0 ) );
    h.a = dist_hit_body( sv, kind, n, ray, want_nor ? &h.n : nullptr );
    return h;
}

template <typename R, bool SH> ACN_HD R dist_hit( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, V3<R>* nor )
{
    const HitN<R> h = dist_hit_ool( sv, kind, n, ray, nor != nullptr );
    if( nor ) *nor = h.n;
    return h.a;
}

// Roots ta <= tb of  f t^2 + 2 fs t + fq = 0  for the quadric qa x^2 + qb y^2 + qc z^2 + qr (p, d in the object frame,
// ad = (qa dx, qb dy, qc dz), f = ad.d != 0).  s*s - q cancels catastrophically in FP32 when the origin is far from the
// quadric (both terms ~ |p|^2), so the quadratic is re-expanded around the point of closest approach pm = p + t0*d,
// where its coefficients are of the size of the object.  That needs t0 = -fs/f to be a sensible distance: for cones and
// hyperboloids (a coefficient < 0) f itself cancels when the ray runs nearly along a generating line, t0 goes to
// infinity and pm loses the near root.  In FP32 those rays take the textbook cancellation-free form
// q = -(fs + sign(fs) sqrt(fs^2 - f fq)), roots q/f and fq/q, which stays accurate as f -> 0.
template <typename R> ACN_HD bool quadric_roots( R qa, R qb, R qc, R qr, V3<R> p, V3<R> d, V3<R> ad, R f, R fs, R fq, R* ta, R* tb )
{
    if( sizeof( R ) == 4 && r_abs( f ) < R( 0.05 ) * ( r_abs( ad.x * d.x ) + r_abs( ad.y * d.y ) + r_abs( ad.z * d.z ) ) )
    {
        R disc = fs * fs - f * fq;
        if( disc < R( 0 ) ) return false;
        R sq = r_sqrt( disc );
        R qq = -( fs + ( fs < R( 0 ) ? -sq : sq ) );
        R t1 = qq / f;
        R t2 = qq != R( 0 ) ? fq / qq : t1;
        *ta = r_min( t1, t2 ); *tb = r_max( t1, t2 );
        return true;
    }
    R fi = R( 1 ) / f;
    R t0 = -fs * fi;
    V3<R> pm = madd( p, d, t0 );
    R s = dot( ad, pm ) * fi;                                                  // ~ 0
    R q = ( qa * pm.x * pm.x + qb * pm.y * pm.y + qc * pm.z * pm.z + qr ) * fi;
    R r = s * s - q;
    if( r < R( 0 ) ) return false;
    r = r_sqrt( r );
    *ta = t0 - s - r; *tb = t0 - s + r;
    return true;
}

// ---------------------------------------------------------------------------------------------
// primitives: fp_ray_hit of plane / sphere / squaroid / distance objects
// (gmath.h:38-45, objects.c:529-537,649-657,778-821,903-959).  No envelope test, no roughness.
// ---------------------------------------------------------------------------------------------
template <typename R, bool SH> ACN_HD R prim_hit( const SceneView<R, SH>& sv, int kind, int n, const Ray<R>& ray, V3<R>* nor )
{
    const R inf = Num<R>::inf();
    const R eps = sv.eps;
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );

    if( kind == K_SPHERE ) return sphere_hit( pos, g0.w, ray, eps, nor );

    if( kind == K_PLANE )
    {
        V3<R> nz = xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
        R div = dot( nz, ray.d );
        if( div == R( 0 ) ) return inf;
        R offs = dot( pos - ray.p, nz ) / div;
        if( nor ) *nor = nz;
        return offs > R( 0 ) ? offs - eps : inf;
    }

    const M3<R> rax = node_rax( sv, n );

    if( kind == K_SQUAROID )
    {
        const R qa = g0.w;
        const R qb = sv.geo[ n * GEO_STRIDE + 1 ].w;
        const R qc = sv.geo[ n * GEO_STRIDE + 2 ].w;
        const R qr = sv.geo[ n * GEO_STRIDE + 3 ].w;
        V3<R> p = mlv( rax, ray.p - pos );
        V3<R> d = mlv( rax, ray.d );
        V3<R> ad = v3<R>( qa * d.x, qb * d.y, qc * d.z );
        R f  = dot( ad, d );
        R fs = dot( ad, p );
        R fq = qa * p.x * p.x + qb * p.y * p.y + qc * p.z * p.z + qr;
        R a;
        if( f != R( 0 ) )
        {
            R ta, tb;
            if( !quadric_roots( qa, qb, qc, qr, p, d, ad, f, fs, fq, &ta, &tb ) ) return inf;
            a = ta;
            if( a < R( 0 ) ) a = tb;
            if( a < R( 0 ) ) return inf;
        }
        else
        {
            if( fq == R( 0 ) ) return inf;
            a = -fs / ( R( 2 ) * fq );          // sic: objects.c:802
        }
        if( !( a < inf ) ) return inf;
        if( nor )
        {
            V3<R> x = madd( p, d, a );
            *nor = unit( tmlv( rax, v3<R>( x.x * qa, x.y * qb, x.z * qc ) ) );
        }
        return a - eps;
    }

    return dist_hit( sv, kind, n, ray, nor );
}

// fp_side of the primitives (gmath.h:52-55,93-97, objects.c:823-827,961-966)
template <typename R, bool SH> ACN_HD int prim_side( const SceneView<R, SH>& sv, int kind, int n, V3<R> x )
{
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    const V3<R> pos = xyz( g0 );
    if( kind == K_SPHERE ) return sphere_side( pos, g0.w, x );
    if( kind == K_PLANE )  return dot( x - pos, xyz( sv.geo[ n * GEO_STRIDE + 3 ] ) ) > R( 0 ) ? 1 : -1;
    const M3<R> rax = node_rax( sv, n );
    V3<R> p = mlv( rax, x - pos );
    if( kind == K_SQUAROID )
    {
        R v = g0.w * p.x * p.x + sv.geo[ n * GEO_STRIDE + 1 ].w * p.y * p.y +
              sv.geo[ n * GEO_STRIDE + 2 ].w * p.z * p.z + sv.geo[ n * GEO_STRIDE + 3 ].w;
        return v > R( 0 ) ? 1 : -1;
    }
    return dist_fn( kind, sv.geo[ n * GEO_STRIDE + 1 ].w, p * g0.w ) > R( 0 ) ? 1 : -1;
}

// ---------------------------------------------------------------------------------------------
// obj_ray_hit / obj_side (objects.c:261-284,365-370) over the whole object algebra.
// CSG nodes recurse (device stack); the recursion depth is the CSG nesting depth.
// ---------------------------------------------------------------------------------------------
template <typename R, bool SH> ACN_HDN int obj_side( const SceneView<R, SH>& sv, int n, V3<R> x );
template <typename R, bool SH> ACN_HDN R   obj_ray_hit( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx );

// roughness perturbation of the normal (objects.c:266-282); out of line, arguments and result by value (see dist_hit_ool)
template <typename R> ACN_NOINLINE V3<R> roughen_ool( R rough, int seed_mode, V3<R> hit_pos, V3<R> v, u64 key, int n )
{
    u64 rv = seed_mode == SEED_POSITION_HASH ? random_seed( hit_pos, ( u64 )1246 )
                                             : mix64( mix64( key, KEY_ROUGH ), ( u64 )n );
    R f;
    f = rnd0<R>( &rv ) * R( 0.99 ); v.x += rough * r_log( ( R( 1 ) - f ) / ( R( 1 ) + f ) );
    f = rnd0<R>( &rv ) * R( 0.99 ); v.y += rough * r_log( ( R( 1 ) - f ) / ( R( 1 ) + f ) );
    f = rnd0<R>( &rv ) * R( 0.99 ); v.z += rough * r_log( ( R( 1 ) - f ) / ( R( 1 ) + f ) );
    return unit( v );
}

template <typename R, bool SH> ACN_HD void roughen( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, R a, V3<R>* nor, HitCtx ctx )
{
    *nor = roughen_ool<R>( sv.geo[ n * GEO_STRIDE + 4 ].x, sv.seed_mode, madd( ray.p, ray.d, a ), *nor, ctx.key, n );
}

// A&B (want = -1) and A|B (want = +1): first boundary point of either child lying on the wanted
// side of the other; alternating march with 2*eps steps (objects.c:1052-1094,1209-1251)
template <typename R, bool SH> ACN_HDN R pair_hit( const SceneView<R, SH>& sv, int o1, int o2, int want, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    const R inf = Num<R>::inf();
    V3<R> n1, n2;
    R a1 = obj_ray_hit( sv, o1, ray, nor ? &n1 : nullptr, ctx );
    R a2 = obj_ray_hit( sv, o2, ray, nor ? &n2 : nullptr, ctx );
    if( a1 < a2 && obj_side( sv, o2, madd( ray.p, ray.d, a1 ) ) == want )
    {
        if( nor ) *nor = n1;
        return a1;
    }
    if( !( a2 < inf ) ) return inf;
    if( obj_side( sv, o1, madd( ray.p, ray.d, a2 ) ) == want )
    {
        if( nor ) *nor = n2;
        return a2;
    }
    R offs = a2;
    int cur = o1, other = o2;
    Ray<R> r2; r2.d = ray.d;
    for( int it = 0; it < CSG_MAX_STEPS && offs < inf; it++ )
    {
        r2.p = madd( ray.p, ray.d, offs );
        R a = obj_ray_hit( sv, cur, r2, nor ? &n1 : nullptr, ctx );
        if( !( a < inf ) ) return inf;
        if( obj_side( sv, other, madd( r2.p, r2.d, a ) ) == want )
        {
            if( nor ) *nor = n1;
            return offs + a;
        }
        offs += a + R( 2 ) * sv.eps;
        int t = cur; cur = other; other = t;
    }
    return inf;
}


// fp_ray_hit dispatch without the own-envelope test and without roughness
template <typename R, bool SH> ACN_HD R shape_hit( const SceneView<R, SH>& sv, const I4& lk, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    const int kind = node_kind( lk );
    if( kind <= K_DIST_TORUS ) return prim_hit( sv, kind, n, ray, nor );
    if( kind == K_PAIR_INSIDE )  return pair_hit( sv, lk.y, lk.z, -1, ray, nor, ctx );
    if( kind == K_PAIR_OUTSIDE ) return pair_hit( sv, lk.y, lk.z, +1, ray, nor, ctx );
    if( kind == K_NEG )                                                                  // objects.c:1329-1339
    {
        R a = obj_ray_hit( sv, lk.y, ray, nor, ctx );
        if( a < Num<R>::inf() && nor ) *nor = -( *nor );
        return a;
    }
    // K_SCALE (objects.c:1418-1437)
    {
        const V3<R> pos = xyz( sv.geo[ n * GEO_STRIDE ] );
        const M3<R> rax = node_rax( sv, n );
        const V3<R> inv = v3<R>( sv.geo[ n * GEO_STRIDE ].w, sv.geo[ n * GEO_STRIDE + 1 ].w, sv.geo[ n * GEO_STRIDE + 2 ].w );
        Ray<R> rl;
        rl.p = mul( mlv( rax, ray.p - pos ), inv );
        rl.d = mul( mlv( rax, ray.d ), inv );
        R len = r_sqrt( sqr( rl.d ) );
        R fac = len > R( 0 ) ? R( 1 ) / len : R( 0 );
        rl.d = rl.d * fac;
        V3<R> n1;
        R a1 = obj_ray_hit( sv, lk.y, rl, nor ? &n1 : nullptr, ctx ) + sv.eps;
        if( a1 < Num<R>::inf() )
        {
            if( nor ) *nor = unit( tmlv( rax, mul( n1, inv ) ) );
            return a1 * fac - sv.eps;
        }
        return Num<R>::inf();
    }
}

// obj_ray_hit body after the envelope test: shape + roughness
template <typename R, bool SH> ACN_HD R obj_hit_noenv( const SceneView<R, SH>& sv, const I4& lk, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    R a = shape_hit( sv, lk, n, ray, nor, ctx );
    if( nor && ( node_flags( lk ) & F_ROUGH ) && a < Num<R>::inf() ) roughen( sv, n, ray, a, nor, ctx );
    return a;
}

template <typename R, bool SH> ACN_HDN R obj_ray_hit( const SceneView<R, SH>& sv, int n, const Ray<R>& ray, V3<R>* nor, HitCtx ctx )
{
    const I4 lk = sv.link[ n ];
    if( ( node_flags( lk ) & F_ENV ) && !envelope_hits( sv.env[ n ], ray ) ) return Num<R>::inf();
    return obj_hit_noenv( sv, lk, n, ray, nor, ctx );
}

// obj_side (objects.c:365-370): "outside" whenever outside the own envelope — also for negations
template <typename R, bool SH> ACN_HDN int obj_side( const SceneView<R, SH>& sv, int n, V3<R> x )
{
    const I4 lk = sv.link[ n ];
    if( node_flags( lk ) & F_ENV )
    {
        const R4<R> e = sv.env[ n ];
        if( sphere_side( xyz( e ), e.w, x ) == 1 ) return 1;
    }
    const int kind = node_kind( lk );
    if( kind <= K_DIST_TORUS ) return prim_side( sv, kind, n, x );
    if( kind == K_PAIR_INSIDE )  return ( obj_side( sv, lk.y, x ) + obj_side( sv, lk.z, x ) == -2 ) ? -1 : 1;   // objects.c:1096-1099
    if( kind == K_PAIR_OUTSIDE ) return ( obj_side( sv, lk.y, x ) + obj_side( sv, lk.z, x ) ==  2 ) ? 1 : -1;   // objects.c:1253-1256
    if( kind == K_NEG ) return -obj_side( sv, lk.y, x );                                                        // objects.c:1341-1344
    {                                                                                                           // objects.c:1439-1443
        const V3<R> pos = xyz( sv.geo[ n * GEO_STRIDE ] );
        const M3<R> rax = node_rax( sv, n );
        const V3<R> inv = v3<R>( sv.geo[ n * GEO_STRIDE ].w, sv.geo[ n * GEO_STRIDE + 1 ].w, sv.geo[ n * GEO_STRIDE + 2 ].w );
        return obj_side( sv, lk.y, mul( mlv( rax, x - pos ), inv ) );
    }
}

// ---------------------------------------------------------------------------------------------
// obj_fov (objects.c:254-259): cone from pos that contains the whole light.
// sphere objects.c:619-637; plane :520-527; pairs :1035-1044,1192-1201 (envelope_s_fov :70-88)
// returns false for shapes without a fov function (rejected at upload for lights)
// ---------------------------------------------------------------------------------------------
template <typename R, bool SH> ACN_HD bool obj_fov( const SceneView<R, SH>& sv, int n, V3<R> pos, V3<R>* axis, R* cos_rs )
{
    const I4 lk = sv.link[ n ];
    const int kind = node_kind( lk );
    const R4<R> g0 = sv.geo[ n * GEO_STRIDE ];
    if( kind == K_SPHERE || ( ( kind == K_PAIR_INSIDE || kind == K_PAIR_OUTSIDE ) && ( node_flags( lk ) & F_ENV ) ) )
    {
        V3<R> c; R r;
        if( kind == K_SPHERE ) { c = xyz( g0 ); r = g0.w; }
        else { const R4<R> e = sv.env[ n ]; c = xyz( e ); r = e.w; }
        V3<R> diff = c - pos;
        R d2 = sqr( diff ), r2 = r * r;
        *axis = unit( diff );
        *cos_rs = d2 > r2 ? r_sqrt( R( 1 ) - r2 / d2 ) : R( -1 );
        return true;
    }
    if( kind == K_PLANE )
    {
        V3<R> d = -xyz( sv.geo[ n * GEO_STRIDE + 3 ] );
        *axis = d;
        *cos_rs = dot( xyz( g0 ) - pos, d ) > R( 0 ) ? R( 0 ) : R( 1 );
        return true;
    }
    if( kind == K_PAIR_INSIDE || kind == K_PAIR_OUTSIDE )
    {
        *axis = unit( xyz( g0 ) - pos );
        *cos_rs = R( 0 );
        return true;
    }
    return false;
}

// obj_projection for the chess texture (objects.c:514-518,602-617,892-895)
template <typename R, bool SH> ACN_HD void obj_projection( const SceneView<R, SH>& sv, int n, V3<R> pos, R* u, R* v )
{
    const int kind = node_kind( sv.link[ n ] );
    const V3<R> c = xyz( sv.geo[ n * GEO_STRIDE ] );
    const M3<R> rax = node_rax( sv, n );
    if( kind == K_PLANE )
    {
        V3<R> p = pos - c;
        *u = dot( p, rax.x ); *v = dot( p, rax.y );
    }
    else if( kind == K_SPHERE )
    {
        V3<R> r = unit( pos - c );
        R x = dot( r, rax.x );
        R y = dot( r, cross( rax.z, rax.x ) );
        R z = r_min( r_max( dot( r, rax.z ), R( -1 ) ), R( 1 ) );
        *u = r_atan2( x, y ); *v = r_asin( z );
    }
    else { *u = R( 0 ); *v = R( 0 ); }
}

} // namespace acn
