// acn_oracle.cpp — CPU oracle for the Actinon sample tracer.  TEST INFRASTRUCTURE ONLY.
//
// A plain FP64, recursive restatement of the reference's per-sample loop.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
// product (actinon_b200/) never does.
//
// Follows, function by function:
//   vectors.h:45-48,148-206,238-241,315-322,372-384   rnd, of_length, von, con, seeds, cap sampling, reflection, con_z, cl_s_sat
//   gmath.h:38-97, gmath.c:68-113                      plane/sphere hit + side, Fresnel, refraction
//   objects.c:90-103,261-284,365-370,411-422           envelope, obj_ray_hit (+roughness), obj_side, obj_color
//   objects.c:514-537,602-657,778-827,903-966          plane, sphere, squaroid, distance objects
//   objects.c:1035-1099,1192-1256,1329-1344,1418-1443  pair_inside, pair_outside, neg, scale
//   distance.c:39-42,83-92, textures.c:99-102,142-148  distance functions, textures
//   compound.c:215-299                                 compound_s_ray_hit, compound_s_ray_trans_hit
//   scene.c:362-382,394-416,420-667,956-1013           scene_s_trans_hit, oren_nayar_weight, scene_s_lum, lum_machine_s_func
//   scene.c:804-813                                    lum_image_s_push (oracle_accumulate)
//
// PARITY PINNING.  The reference has no tests, golden vectors or stored hashes for this path
// (SURVEY.md §4, §8c), and its three LCGs live in the absent, unpinned dependency
// github.com/johsteffens/beth (bcore_lcg00/01/02_u3).  The constants below are placeholders
// (full-period 64-bit LCGs): "parity unpinned" at the bit level for anything that consumes random
// numbers.  What IS pinned: every deterministic leaf function is checked against the reference's
// own vectors.h / gmath.h / gmath.c compiled in place (oracle/_ref, see oracle/Makefile and
// tests/test_oracle_vs_reference_leaf.py), analytic known answers, and the shipped images
// statistically.
//
// Two switches that the reference does not have:
//   seed_mode 0  reference behaviour (position-hash seeding, scene.c:537, objects.c:269)
//   seed_mode 1  index-keyed seeding shared with the CUDA tracer for the 1e-3 per-pixel check
//   eps          shell thickness; 1e-6 = reference (vectors.h:33)
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <thread>
#include <vector>
#include <chrono>

#include "../include/actinon_b200.h"

typedef uint64_t u3_t;
typedef int64_t  s3_t;

// ---------------------------------------------------------------------------------------------
// counters: one per event class of SURVEY.md §8(d); FLOPs = sum count_k * F_k
// ---------------------------------------------------------------------------------------------
enum
{
    C_SPHERE_MISS, C_SPHERE_HIT, C_SPHERE_NOR, C_PLANE, C_SQUAROID, C_SQUAROID_NOR, C_DIST_STEP, C_DIST_NOR,
    C_SIDE_SPHERE, C_SIDE_PLANE, C_SIDE_SQUAROID, C_SIDE_DIST,
    C_FRESNEL, C_REFRACT, C_MIRROR, C_CAP_SAMPLE, C_OREN_NAYAR, C_DIRECT_BOOK, C_ROUGHNESS, C_CAMERA, C_GAMMA, C_ABSORB,
    R_PRIMARY, R_REFLECT, R_CHROMATIC, R_REFRACT, R_PATH, R_SHADOW, R_LIGHTHIT, R_DIFFUSE,
    C_COUNT
};

static const double flops_per_event[ C_COUNT ] =
{
    16, 20, 20, 15, 66, 35, 24, 80,
    9, 8, 27, 24,
    48, 24, 23, 32, 32, 26, 20, 40, 0, 3,
    0, 0, 0, 0, 0, 0, 0, 0
};
// transcendental (special-function) calls per event, reported separately
static const double sf_per_event[ C_COUNT ] =
{
    0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0,
    0, 0, 0, 2, 3, 0, 3, 0, 3, 3,
    0, 0, 0, 0, 0, 0, 0, 0
};

// counters are kept per phase of the ray tree so that FLOPs can be attributed to the tracer's kernels:
// 0 primary rays, 1 reflection/chromatic/refraction rays, 2 path rays, 3 direct-light loop
enum { PH_PRIMARY = 0, PH_SPEC = 1, PH_PATH = 2, PH_DIRECT = 3, PH_COUNT = 4 };
struct counters_t { uint64_t c[ PH_COUNT ][ C_COUNT ]; };

// ---------------------------------------------------------------------------------------------
// v3d_s / m3d_s
// ---------------------------------------------------------------------------------------------
struct v3d { double x, y, z; };
struct m3d { v3d x, y, z; };
struct ray_t { v3d p, d; };

static inline v3d V( double x, double y, double z ) { v3d v = { x, y, z }; return v; }
static inline v3d add( v3d a, v3d b ) { return V( a.x + b.x, a.y + b.y, a.z + b.z ); }
static inline v3d sub( v3d a, v3d b ) { return V( a.x - b.x, a.y - b.y, a.z - b.z ); }
static inline v3d mlf( v3d a, double f ) { return V( a.x * f, a.y * f, a.z * f ); }
static inline v3d neg( v3d a ) { return V( -a.x, -a.y, -a.z ); }
static inline double mlv( v3d a, v3d b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double vsqr( v3d a ) { return mlv( a, a ); }
static inline v3d mlx( v3d o, v3d f ) { return V( o.y * f.z - o.z * f.y, o.z * f.x - o.x * f.z, o.x * f.y - o.y * f.x ); }
static inline v3d mld( v3d a, v3d b ) { return V( a.x * b.x, a.y * b.y, a.z * b.z ); }
static inline v3d ray_pos( const ray_t* r, double offs ) { return add( r->p, mlf( r->d, offs ) ); }

// vectors.h:148-154
static inline v3d of_length( v3d o, double a )
{
    double r_sqr = vsqr( o );
    if( fabs( r_sqr - 1.0 ) < 1E-8 ) return o;
    double f = r_sqr > 0 ? ( a / sqrt( r_sqr ) ) : 0;
    return V( o.x * f, o.y * f, o.z * f );
}

// vectors.h:157-175
static inline v3d von( v3d o, v3d v )
{
    v3d o_n = of_length( o, 1.0 );
    v = sub( v, mlf( o_n, mlv( o_n, v ) ) );
    return of_length( v, 1.0 );
}
static inline v3d con( v3d o )
{
    double xx = o.x * o.x, yy = o.y * o.y, zz = o.z * o.z;
    v3d v;
    v.x = ( ( xx <= yy ) && ( xx <= zz ) ) ? 1 : 0;
    v.y = ( ( yy <= xx ) && ( yy <= zz ) ) ? 1 : 0;
    v.z = ( ( zz <= xx ) && ( zz <= yy ) ) ? 1 : 0;
    return von( o, v );
}
static inline v3d orthogonal_projection( v3d o, v3d nor ) { double f = mlv( o, nor ); return V( o.x - nor.x * f, o.y - nor.y * f, o.z - nor.z * f ); }
static inline v3d reflection( v3d dir, v3d nor ) { return of_length( sub( dir, mlf( nor, 2.0 * mlv( dir, nor ) ) ), 1.0 ); }

static inline v3d m_mlv( const m3d* o, v3d v ) { return V( mlv( o->x, v ), mlv( o->y, v ), mlv( o->z, v ) ); }
static inline v3d m_tmlv( const m3d* o, v3d v )
{
    return V( o->x.x * v.x + o->y.x * v.y + o->z.x * v.z, o->x.y * v.x + o->y.y * v.y + o->z.y * v.z, o->x.z * v.x + o->y.z * v.y + o->z.z * v.z );
}
static inline m3d m_transposed( m3d o ) { m3d m = { V( o.x.x, o.y.x, o.z.x ), V( o.x.y, o.y.y, o.z.y ), V( o.x.z, o.y.z, o.z.z ) }; return m; }
static inline m3d m_con_z( v3d v ) { m3d m; m.z = of_length( v, 1.0 ); m.x = con( v ); m.y = mlx( m.z, m.x ); return m; }   // vectors.h:315-322

// ---------------------------------------------------------------------------------------------
// RNG — beth's bcore_lcg00/01/02_u3 are not available: PLACEHOLDER constants (see header)
// ---------------------------------------------------------------------------------------------
static inline u3_t lcg00( u3_t v ) { return v * 6364136223846793005ull + 1442695040888963407ull; }
static inline u3_t lcg01( u3_t v ) { return v * 3935559000370003845ull + 2691343689449507681ull; }
static inline u3_t lcg02( u3_t v ) { return v * 2862933555777941757ull + 3037000493ull; }

static inline double rnd0( u3_t* rv ) { return ( *rv = lcg00( *rv ) ) * ( 2.0 / 0xFFFFFFFFFFFFFFFFull ) - 1.0; }   // vectors.h:45
static inline double rnd1( u3_t* rv ) { return ( *rv = lcg00( *rv ) ) * ( 1.0 / 0xFFFFFFFFFFFFFFFFull ); }         // vectors.h:48

static inline u3_t seed_from_f3( double v )     // vectors.h:177-182
{
    int exp = 0;
    s3_t seed_s3 = ( s3_t )( frexp( v, &exp ) * ( double )0x7FFFFFFFFFFFFFFFll );
    return ( u3_t )seed_s3 * 27362149ull;
}
static inline u3_t random_seed( v3d o, u3_t rv )   // vectors.h:185-190
{
    return seed_from_f3( o.x ) * lcg00( rv ) + seed_from_f3( o.y ) * lcg01( rv ) + seed_from_f3( o.z ) * lcg02( rv );
}

static inline v3d random_sphere_cap( u3_t* rv, double h )     // vectors.h:197-206
{
    v3d v;
    double phi = 2.0 * M_PI * rnd1( rv );
    v.z = 1.0 - rnd1( rv ) * h;
    double scale = sqrt( 1.0 - v.z * v.z );
    v.x = sin( phi ) * scale;
    v.y = cos( phi ) * scale;
    return v;
}

// index-keyed seeding (not in the reference): splitmix64 finaliser over (key, salt)
static inline u3_t mix64( u3_t key, u3_t salt )
{
    u3_t z = key + salt * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    return z ^ ( z >> 31 );
}
enum { KEY_REFLECT = 1, KEY_CHROMATIC = 2, KEY_REFRACT = 3, KEY_DIFFUSE = 4, KEY_ROUGH = 5, KEY_PATH0 = 16 };

// ---------------------------------------------------------------------------------------------
// object graph rebuilt from the flat scene
// ---------------------------------------------------------------------------------------------
struct obj_t
{
    int kind, index;
    const acn_flat_node* nd;
    const acn_flat_material* mat;
    v3d pos; m3d rax;
    bool has_env; v3d env_pos; double env_r;
    double rough;
    obj_t* o1; obj_t* o2;
};

struct cmp_t;
struct elem_t { obj_t* obj; cmp_t* cmp; };
struct cmp_t { bool has_env; v3d env_pos; double env_r; std::vector<elem_t> el; };

struct trans_t { v3d exit_nor; const obj_t* exit_obj; const obj_t* enter_obj; };

struct oscene_t
{
    acn_flat_params prm;
    cmp_t* light; cmp_t* matter;
    std::vector<obj_t*> objs; std::vector<cmp_t*> cmps;
    double eps; int seed_mode;
    ~oscene_t() { for( obj_t* o : objs ) delete o; for( cmp_t* c : cmps ) delete c; }
};

struct ctx_t { const oscene_t* sc; counters_t* cnt; double eps; double inf; int phase; };

#define CNT( ctx, k ) ( ( ctx )->cnt->c[ ( ctx )->phase ][ k ]++ )

static obj_t* build_obj( oscene_t* sc, const acn_flat_scene* fs, int i )
{
    const acn_flat_node* nd = &fs->nodes[ i ];
    obj_t* o = new obj_t();
    sc->objs.push_back( o );
    o->kind = nd->kind; o->index = i; o->nd = nd; o->mat = &fs->materials[ nd->material ];
    o->pos = V( nd->pos[ 0 ], nd->pos[ 1 ], nd->pos[ 2 ] );
    o->rax.x = V( nd->rax[ 0 ], nd->rax[ 1 ], nd->rax[ 2 ] );
    o->rax.y = V( nd->rax[ 3 ], nd->rax[ 4 ], nd->rax[ 5 ] );
    o->rax.z = V( nd->rax[ 6 ], nd->rax[ 7 ], nd->rax[ 8 ] );
    o->has_env = nd->has_envelope != 0;
    o->env_pos = V( nd->env_pos[ 0 ], nd->env_pos[ 1 ], nd->env_pos[ 2 ] ); o->env_r = nd->env_radius;
    o->rough = nd->surface_roughness;
    o->o1 = o->o2 = NULL;
    if( nd->kind == ACN_KIND_PAIR_INSIDE || nd->kind == ACN_KIND_PAIR_OUTSIDE ) { o->o1 = build_obj( sc, fs, nd->child0 ); o->o2 = build_obj( sc, fs, nd->child1 ); }
    if( nd->kind == ACN_KIND_NEG || nd->kind == ACN_KIND_SCALE ) o->o1 = build_obj( sc, fs, nd->child0 );
    return o;
}

static cmp_t* build_cmp( oscene_t* sc, const acn_flat_scene* fs, int i )
{
    const acn_flat_node* nd = &fs->nodes[ i ];
    cmp_t* c = new cmp_t();
    sc->cmps.push_back( c );
    c->has_env = nd->has_envelope != 0;
    c->env_pos = V( nd->env_pos[ 0 ], nd->env_pos[ 1 ], nd->env_pos[ 2 ] ); c->env_r = nd->env_radius;
    for( int k = 0; k < nd->child1; k++ )
    {
        int ci = fs->children[ nd->child0 + k ];
        elem_t e = { NULL, NULL };
        if( fs->nodes[ ci ].kind == ACN_KIND_COMPOUND ) e.cmp = build_cmp( sc, fs, ci ); else e.obj = build_obj( sc, fs, ci );
        c->el.push_back( e );
    }
    return c;
}

// ---------------------------------------------------------------------------------------------
// gmath.h
// ---------------------------------------------------------------------------------------------
static double plane_ray_hit( ctx_t* cx, v3d pos, v3d nor, const ray_t* ray, v3d* p_nor )     // gmath.h:38-45
{
    CNT( cx, C_PLANE );
    double div = mlv( nor, ray->d );
    if( div == 0 ) return cx->inf;
    double offs = mlv( sub( pos, ray->p ), nor ) / div;
    if( p_nor ) *p_nor = nor;
    return ( offs > 0 ) ? offs - cx->eps : cx->inf;
}

static double sphere_ray_hit( ctx_t* cx, v3d pos, double r, const ray_t* ray, v3d* p_nor )   // gmath.h:64-85
{
    v3d p = sub( ray->p, pos );
    double s = mlv( p, ray->d );
    double q = vsqr( p ) - ( r * r );
    double s2 = s * s;
    if( s2 < q ) { CNT( cx, C_SPHERE_MISS ); return cx->inf; }
    double offs = cx->inf;
    if( s < 0 && q > 0 )      offs = -s - sqrt( s2 - q ) - cx->eps;
    else if( s < 0 || q < 0 ) offs = -s + sqrt( s2 - q ) - cx->eps;
    if( offs < cx->inf ) CNT( cx, C_SPHERE_HIT ); else CNT( cx, C_SPHERE_MISS );
    if( offs < cx->inf && p_nor ) { CNT( cx, C_SPHERE_NOR ); *p_nor = of_length( sub( ray_pos( ray, offs ), pos ), 1.0 ); }
    return offs;
}

static inline int sphere_observer_side( v3d pos, double r, v3d observer ) { return ( vsqr( sub( observer, pos ) ) > r * r ) ? 1 : -1; }   // gmath.h:93-97
static inline int plane_observer_side( v3d pos, v3d nor, v3d observer ) { return mlv( sub( observer, pos ), nor ) > 0 ? 1 : -1; }        // gmath.h:52-55

// gmath.c:68-91
static double fresnel_reflection( v3d dir_i, v3d exit_nor, double trix, v3d* dir )
{
    double c = mlv( dir_i, exit_nor );
    double f = c < 0 ? trix : 1.0 / trix;
    double cos_ai = fabs( c );
    cos_ai = cos_ai > 1.0 ? 1.0 : cos_ai;
    double sin_ai = sqrt( 1.0 - cos_ai * cos_ai );
    double sin_at = sin_ai * f;
    double reflectance = 1.0;
    if( sin_at < 1 )
    {
        double cos_at = sqrt( 1.0 - sin_at * sin_at );
        double rs = ( f * cos_ai - cos_at ) / ( f * cos_ai + cos_at ); rs *= rs;
        double rp = ( f * cos_at - cos_ai ) / ( f * cos_at + cos_ai ); rp *= rp;
        reflectance = ( rs + rp ) * 0.5;
    }
    if( dir ) *dir = reflection( dir_i, exit_nor );
    return reflectance;
}

// gmath.c:94-113
static void fresnel_refraction( v3d dir_i, v3d exit_nor, double trix, v3d* dir )
{
    double c = mlv( dir_i, exit_nor );
    double f = c < 0 ? trix : 1.0 / trix;
    double q = f * f * ( 1.0 - c * c );
    if( q < 1.0 )
    {
        double b = -f * c + ( c > 0 ? sqrt( 1.0 - q ) : -sqrt( 1.0 - q ) );
        *dir = add( mlf( dir_i, f ), mlf( exit_nor, b ) );
    }
    else *dir = dir_i;
}

// ---------------------------------------------------------------------------------------------
// objects.c
// ---------------------------------------------------------------------------------------------
static bool envelope_ray_hits( ctx_t* cx, v3d pos, double r, const ray_t* ray ) { return sphere_ray_hit( cx, pos, r, ray, NULL ) < cx->inf; }   // :90-93

static double distance_fn( const obj_t* o, v3d pos )     // distance.c:39-42,83-92
{
    if( o->kind == ACN_KIND_DIST_SPHERE ) return sqrt( pos.x * pos.x + pos.y * pos.y + pos.z * pos.z ) - 1.0;
    double x = pos.x, y = pos.y;
    double f = sqrt( x * x + y * y );
    double f_inv = ( f > 0 ) ? ( 1.0 / f ) : 1.0;
    x *= f_inv; y *= f_inv;
    double dx = x - pos.x, dy = y - pos.y;
    return sqrt( dx * dx + dy * dy + pos.z * pos.z ) - o->nd->tail[ 1 ];
}

static double obj_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* ray, v3d* p_nor, u3_t key );
static int obj_side( ctx_t* cx, const obj_t* o, v3d pos );

static double squaroid_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* r, v3d* p_nor )     // :778-821
{
    CNT( cx, C_SQUAROID );
    const double qa = o->nd->tail[ 0 ], qb = o->nd->tail[ 1 ], qc = o->nd->tail[ 2 ], qr = o->nd->tail[ 3 ];
    v3d p = m_mlv( &o->rax, sub( r->p, o->pos ) );
    v3d d = m_mlv( &o->rax, r->d );
    double f  = qa * d.x * d.x + qb * d.y * d.y + qc * d.z * d.z;
    double fs = qa * d.x * p.x + qb * d.y * p.y + qc * d.z * p.z;
    double fq = qa * p.x * p.x + qb * p.y * p.y + qc * p.z * p.z + qr;
    double a = cx->inf;
    if( f != 0 )
    {
        double f_inv = 1.0 / f;
        double s = fs * f_inv, q = fq * f_inv;
        double rr = s * s - q;
        if( rr < 0 ) return cx->inf;
        rr = sqrt( rr );
        a = -s - rr;
        if( a < 0 ) a = -s + rr;
        if( a < 0 ) a = cx->inf;
    }
    else
    {
        a = ( fq != 0 ) ? -fs / ( 2 * fq ) : cx->inf;
    }
    if( a == cx->inf ) return cx->inf;
    if( p_nor )
    {
        CNT( cx, C_SQUAROID_NOR );
        v3d n1 = V( ( p.x + a * d.x ) * qa, ( p.y + a * d.y ) * qb, ( p.z + a * d.z ) * qc );
        *p_nor = of_length( m_tmlv( &o->rax, n1 ), 1.0 );
    }
    return a - cx->eps;
}

static double distance_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* r, v3d* p_nor )     // :903-959
{
    const double inv_scale = o->nd->tail[ 0 ];
    const uint64_t cycles = ( uint64_t )o->nd->tail[ 2 ];
    ray_t ray = *r;
    double offs0 = 0;
    if( o->has_env )
    {
        if( sphere_observer_side( o->env_pos, o->env_r, r->p ) == 1 )
        {
            offs0 = sphere_ray_hit( cx, o->env_pos, o->env_r, &ray, NULL );
            if( offs0 >= cx->inf ) return cx->inf;
            ray.p = ray_pos( &ray, offs0 );
        }
    }
    ray.p = mlf( m_mlv( &o->rax, sub( ray.p, o->pos ) ), inv_scale );
    ray.d = m_mlv( &o->rax, ray.d );
    double offs1 = 0;
    double dist = distance_fn( o, ray.p );
    CNT( cx, C_DIST_STEP );
    if( dist > 0 )
    {
        for( uint64_t i = 0; i < cycles; i++ )
        {
            offs1 += dist + cx->eps;
            dist = distance_fn( o, ray_pos( &ray, offs1 ) );
            CNT( cx, C_DIST_STEP );
            if( dist < 0 || dist > 1E+30 ) break;
        }
    }
    else
    {
        for( uint64_t i = 0; i < cycles; i++ )
        {
            offs1 -= dist - cx->eps;
            dist = distance_fn( o, ray_pos( &ray, offs1 ) );
            CNT( cx, C_DIST_STEP );
            if( dist > 0 || dist < -1E+30 ) break;
        }
    }
    if( fabs( dist ) <= cx->eps )
    {
        if( p_nor )
        {
            CNT( cx, C_DIST_NOR );
            v3d p = ray_pos( &ray, offs1 );
            double d0 = distance_fn( o, p );
            v3d n;
            n.x = ( distance_fn( o, V( p.x + cx->eps, p.y, p.z ) ) - d0 ) / cx->eps;
            n.y = ( distance_fn( o, V( p.x, p.y + cx->eps, p.z ) ) - d0 ) / cx->eps;
            n.z = ( distance_fn( o, V( p.x, p.y, p.z + cx->eps ) ) - d0 ) / cx->eps;
            *p_nor = of_length( m_tmlv( &o->rax, n ), 1.0 );
        }
        return offs0 + ( offs1 / inv_scale ) - cx->eps;
    }
    return cx->inf;
}

// pair_inside (want -1, :1052-1094) and pair_outside (want +1, :1209-1251)
static double pair_ray_hit( ctx_t* cx, const obj_t* o, int want, const ray_t* r, v3d* p_nor, u3_t key )
{
    v3d n1, n2;
    double a1 = obj_ray_hit( cx, o->o1, r, &n1, key );
    double a2 = obj_ray_hit( cx, o->o2, r, &n2, key );
    if( a1 < a2 && obj_side( cx, o->o2, ray_pos( r, a1 ) ) == want ) { if( p_nor ) *p_nor = n1; return a1; }
    if( a2 >= cx->inf ) return cx->inf;
    if( obj_side( cx, o->o1, ray_pos( r, a2 ) ) == want ) { if( p_nor ) *p_nor = n2; return a2; }
    double offs = a2;
    ray_t ray;
    ray.d = r->d;
    ray.p = ray_pos( r, offs );
    const obj_t* obj1 = o->o1;
    const obj_t* obj2 = o->o2;
    while( offs < cx->inf )
    {
        double a = obj_ray_hit( cx, obj1, &ray, &n1, key );
        if( a >= cx->inf ) return cx->inf;
        if( obj_side( cx, obj2, ray_pos( &ray, a ) ) == want ) { if( p_nor ) *p_nor = n1; return offs + a; }
        offs += a + 2 * cx->eps;
        ray.p = ray_pos( r, offs );
        const obj_t* tmp = obj2; obj2 = obj1; obj1 = tmp;
    }
    return cx->inf;
}

static double scale_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* r, v3d* p_nor, u3_t key )     // :1418-1437
{
    v3d inv = V( o->nd->tail[ 0 ], o->nd->tail[ 1 ], o->nd->tail[ 2 ] );
    ray_t ray;
    ray.p = mld( m_mlv( &o->rax, sub( r->p, o->pos ) ), inv );
    ray.d = mld( m_mlv( &o->rax, r->d ), inv );
    double d_length = sqrt( vsqr( ray.d ) );
    double d_factor = ( d_length > 0 ) ? ( 1.0 / d_length ) : 0;
    ray.d = mlf( ray.d, d_factor );
    v3d n1;
    double a1 = obj_ray_hit( cx, o->o1, &ray, &n1, key ) + cx->eps;
    if( a1 < cx->inf )
    {
        n1 = mld( n1, inv );
        if( p_nor ) *p_nor = of_length( m_tmlv( &o->rax, n1 ), 1.0 );
        return a1 * d_factor - cx->eps;
    }
    return cx->inf;
}

static double fp_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* ray, v3d* p_nor, u3_t key )
{
    switch( o->kind )
    {
        case ACN_KIND_PLANE:        return plane_ray_hit( cx, o->pos, o->rax.z, ray, p_nor );                  // :529-532
        case ACN_KIND_SPHERE:       return sphere_ray_hit( cx, o->pos, o->nd->tail[ 0 ], ray, p_nor );         // :649-652
        case ACN_KIND_SQUAROID:     return squaroid_ray_hit( cx, o, ray, p_nor );
        case ACN_KIND_DIST_SPHERE:
        case ACN_KIND_DIST_TORUS:   return distance_ray_hit( cx, o, ray, p_nor );
        case ACN_KIND_PAIR_INSIDE:  return pair_ray_hit( cx, o, -1, ray, p_nor, key );
        case ACN_KIND_PAIR_OUTSIDE: return pair_ray_hit( cx, o, +1, ray, p_nor, key );
        case ACN_KIND_NEG:                                                                                    // :1329-1339
        {
            v3d n1;
            double a1 = obj_ray_hit( cx, o->o1, ray, &n1, key );
            if( a1 < cx->inf ) { if( p_nor ) *p_nor = neg( n1 ); return a1; }
            return cx->inf;
        }
        case ACN_KIND_SCALE:        return scale_ray_hit( cx, o, ray, p_nor, key );
        default: return cx->inf;
    }
}

// objects.c:261-284
static double obj_ray_hit( ctx_t* cx, const obj_t* o, const ray_t* ray, v3d* p_nor, u3_t key )
{
    if( o->has_env && !envelope_ray_hits( cx, o->env_pos, o->env_r, ray ) ) return cx->inf;
    double a = fp_ray_hit( cx, o, ray, p_nor, key );
    if( a < cx->inf && o->rough > 0 && p_nor )
    {
        CNT( cx, C_ROUGHNESS );
        v3d n = *p_nor;
        u3_t rv = cx->sc->seed_mode == ACN_SEED_POSITION_HASH ? random_seed( ray_pos( ray, a ), 1246 )
                                                             : mix64( mix64( key, KEY_ROUGH ), ( u3_t )o->index );
        double f;
        f = rnd0( &rv ) * 0.99; n.x += o->rough * log( ( 1.0 - f ) / ( 1.0 + f ) );
        f = rnd0( &rv ) * 0.99; n.y += o->rough * log( ( 1.0 - f ) / ( 1.0 + f ) );
        f = rnd0( &rv ) * 0.99; n.z += o->rough * log( ( 1.0 - f ) / ( 1.0 + f ) );
        *p_nor = of_length( n, 1.0 );
    }
    return a;
}

// objects.c:365-370 + the per-type fp_side
static int obj_side( ctx_t* cx, const obj_t* o, v3d pos )
{
    if( o->has_env && sphere_observer_side( o->env_pos, o->env_r, pos ) == 1 ) return 1;
    switch( o->kind )
    {
        case ACN_KIND_PLANE:  CNT( cx, C_SIDE_PLANE );  return plane_observer_side( o->pos, o->rax.z, pos );
        case ACN_KIND_SPHERE: CNT( cx, C_SIDE_SPHERE ); return sphere_observer_side( o->pos, o->nd->tail[ 0 ], pos );
        case ACN_KIND_SQUAROID:                                                                               // :823-827
        {
            CNT( cx, C_SIDE_SQUAROID );
            v3d p = m_mlv( &o->rax, sub( pos, o->pos ) );
            return ( o->nd->tail[ 0 ] * p.x * p.x + o->nd->tail[ 1 ] * p.y * p.y + o->nd->tail[ 2 ] * p.z * p.z + o->nd->tail[ 3 ] ) > 0 ? 1 : -1;
        }
        case ACN_KIND_DIST_SPHERE: case ACN_KIND_DIST_TORUS:                                                  // :961-966
        {
            CNT( cx, C_SIDE_DIST );
            v3d p = mlf( m_mlv( &o->rax, sub( pos, o->pos ) ), o->nd->tail[ 0 ] );
            return distance_fn( o, p ) > 0 ? 1 : -1;
        }
        case ACN_KIND_PAIR_INSIDE:  return ( obj_side( cx, o->o1, pos ) + obj_side( cx, o->o2, pos ) == -2 ) ? -1 : 1;
        case ACN_KIND_PAIR_OUTSIDE: return ( obj_side( cx, o->o1, pos ) + obj_side( cx, o->o2, pos ) ==  2 ) ? 1 : -1;
        case ACN_KIND_NEG:          return -1 * obj_side( cx, o->o1, pos );
        case ACN_KIND_SCALE:                                                                                  // :1439-1443
        {
            v3d p = m_mlv( &o->rax, sub( pos, o->pos ) );
            return obj_side( cx, o->o1, mld( p, V( o->nd->tail[ 0 ], o->nd->tail[ 1 ], o->nd->tail[ 2 ] ) ) );
        }
        default: return 1;
    }
}

// obj_fov: sphere :619-637, plane :520-527, pairs :1035-1044 / :1192-1201 (envelope_s_fov :70-88)
static bool obj_fov( const obj_t* o, v3d pos, v3d* axis, double* cos_rs )
{
    if( o->kind == ACN_KIND_SPHERE || ( ( o->kind == ACN_KIND_PAIR_INSIDE || o->kind == ACN_KIND_PAIR_OUTSIDE ) && o->has_env ) )
    {
        v3d c = o->kind == ACN_KIND_SPHERE ? o->pos : o->env_pos;
        double r = o->kind == ACN_KIND_SPHERE ? o->nd->tail[ 0 ] : o->env_r;
        v3d diff = sub( c, pos );
        *axis = of_length( diff, 1.0 );
        double diff_sqr = vsqr( diff ), radius_sqr = r * r;
        *cos_rs = diff_sqr > radius_sqr ? sqrt( 1.0 - ( radius_sqr / diff_sqr ) ) : -1;
        return true;
    }
    if( o->kind == ACN_KIND_PLANE )
    {
        *axis = neg( o->rax.z );
        *cos_rs = mlv( sub( o->pos, pos ), *axis ) > 0 ? 0 : 1;
        return true;
    }
    if( o->kind == ACN_KIND_PAIR_INSIDE || o->kind == ACN_KIND_PAIR_OUTSIDE )
    {
        *axis = of_length( sub( o->pos, pos ), 1.0 );
        *cos_rs = 0;
        return true;
    }
    return false;
}

// obj_color :411-422, textures.c:99-102,142-148, projections objects.c:514-518,602-617,892-895
static v3d obj_color( const obj_t* o, v3d pos )
{
    const acn_flat_material* m = o->mat;
    if( m->texture_kind == ACN_TEX_NONE )  return V( m->color[ 0 ], m->color[ 1 ], m->color[ 2 ] );
    if( m->texture_kind == ACN_TEX_PLAIN ) return V( m->tex_color1[ 0 ], m->tex_color1[ 1 ], m->tex_color1[ 2 ] );
    double u = 0, v = 0;
    if( o->kind == ACN_KIND_PLANE )
    {
        v3d p = sub( pos, o->pos );
        u = mlv( p, o->rax.x ); v = mlv( p, o->rax.y );
    }
    else if( o->kind == ACN_KIND_SPHERE )
    {
        v3d r = of_length( sub( pos, o->pos ), 1.0 );
        double x = mlv( r, o->rax.x );
        double y = mlv( r, mlx( o->rax.z, o->rax.x ) );
        double z = mlv( r, o->rax.z );
        z = z > 1.0 ? 1.0 : z; z = z < -1.0 ? -1.0 : z;
        u = atan2( x, y ); v = asin( z );
    }
    long long x = llrint( u * m->tex_scale ), y = llrint( v * m->tex_scale );
    return ( ( x ^ y ) & 1 ) ? V( m->tex_color1[ 0 ], m->tex_color1[ 1 ], m->tex_color1[ 2 ] ) : V( m->tex_color2[ 0 ], m->tex_color2[ 1 ], m->tex_color2[ 2 ] );
}

// ---------------------------------------------------------------------------------------------
// compound.c
// ---------------------------------------------------------------------------------------------
static double compound_ray_hit( ctx_t* cx, const cmp_t* o, const ray_t* ray, v3d* p_nor, const obj_t** hit_obj, u3_t key )     // :215-244
{
    if( o->has_env && !envelope_ray_hits( cx, o->env_pos, o->env_r, ray ) ) return cx->inf;
    v3d nor;
    double min_a = cx->inf;
    for( size_t i = 0; i < o->el.size(); i++ )
    {
        const obj_t* hit_obj_l = NULL;
        double a;
        if( o->el[ i ].cmp ) a = compound_ray_hit( cx, o->el[ i ].cmp, ray, &nor, &hit_obj_l, key );
        else { hit_obj_l = o->el[ i ].obj; a = obj_ray_hit( cx, hit_obj_l, ray, &nor, key ); }
        if( a < min_a )
        {
            min_a = a;
            if( p_nor ) *p_nor = nor;
            if( hit_obj ) *hit_obj = hit_obj_l;
        }
    }
    return min_a;
}

static double compound_ray_trans_hit( ctx_t* cx, const cmp_t* o, const ray_t* ray, trans_t* trans, u3_t key )     // :246-299
{
    if( o->has_env && !envelope_ray_hits( cx, o->env_pos, o->env_r, ray ) ) return cx->inf;
    v3d nor;
    double min_a = cx->inf;
    for( size_t i = 0; i < o->el.size(); i++ )
    {
        const obj_t* hit_obj = NULL;
        double a;
        if( o->el[ i ].cmp ) a = compound_ray_hit( cx, o->el[ i ].cmp, ray, &nor, &hit_obj, key );
        else { hit_obj = o->el[ i ].obj; a = obj_ray_hit( cx, hit_obj, ray, &nor, key ); }
        if( a < cx->inf )
        {
            if( a < min_a - cx->eps )
            {
                min_a = a;
                if( mlv( nor, ray->d ) > 0 ) { trans->exit_nor = nor; trans->exit_obj = hit_obj; trans->enter_obj = NULL; }
                else { trans->exit_nor = neg( nor ); trans->exit_obj = NULL; trans->enter_obj = hit_obj; }
            }
            else if( fabs( a - min_a ) < cx->eps )
            {
                min_a = a < min_a ? a : min_a;
                if( mlv( nor, ray->d ) > 0 ) trans->exit_obj = hit_obj; else trans->enter_obj = hit_obj;
            }
        }
    }
    return min_a;
}

// ---------------------------------------------------------------------------------------------
// scene.c
// ---------------------------------------------------------------------------------------------
static double scene_trans_hit( ctx_t* cx, const ray_t* r, trans_t* trans, u3_t key )     // :362-382
{
    double min_a = cx->inf, a;
    trans_t trans_l;
    memset( &trans_l, 0, sizeof( trans_l ) );
    if( ( a = compound_ray_trans_hit( cx, cx->sc->light, r, &trans_l, key ) ) < min_a ) { min_a = a; *trans = trans_l; }
    if( ( a = compound_ray_trans_hit( cx, cx->sc->matter, r, &trans_l, key ) ) < min_a ) { min_a = a; *trans = trans_l; }
    return min_a;
}

static double oren_nayar_weight( double weight, double theta_i, double on_a, double on_b, v3d out_d, v3d nor, v3d ray_prj )     // :394-416
{
    double theta_r = acos( weight );
    double cos_phi = -mlv( of_length( orthogonal_projection( out_d, nor ), 1.0 ), ray_prj );
    double mx = theta_i > theta_r ? theta_i : theta_r, mn = theta_i < theta_r ? theta_i : theta_r;
    return weight * ( on_a + ( on_b * ( cos_phi > 0 ? cos_phi : 0 ) * sin( mx ) * tan( mn ) ) );
}

static v3d scene_lum( ctx_t* cx, const ray_t* ray, double offs, trans_t* trans, uint64_t depth, double intensity, u3_t key )     // :420-667
{
    const oscene_t* scene = cx->sc;
    const acn_flat_params* prm = &scene->prm;
    const v3d bg = V( prm->background_color[ 0 ], prm->background_color[ 1 ], prm->background_color[ 2 ] );
    v3d lum = { 0, 0, 0 };
    if( depth == 0 || intensity < prm->trace_min_intensity ) return lum;

    v3d pos = ray_pos( ray, offs );

    if( trans->enter_obj && trans->enter_obj->mat->radiance > 0 )
    {
        CNT( cx, R_LIGHTHIT );
        v3d dp = sub( pos, trans->enter_obj->pos );
        double diff_sqr = vsqr( dp );
        double light_intensity = ( diff_sqr > 0 ) ? ( trans->enter_obj->mat->radiance / diff_sqr ) : 1E+30;
        return mlf( obj_color( trans->enter_obj, pos ), light_intensity * intensity );
    }

    double trans_refractive_index = 1.0, fresnel_reflectivity = 0, chromatic_reflectivity = 0, diffuse_reflectivity = 0;
    double on_a = 1.0, on_b = 0.0;
    bool transparent = false;

    if( trans->enter_obj )
    {
        const acn_flat_material* m = trans->enter_obj->mat;
        trans_refractive_index = m->refractive_index;
        fresnel_reflectivity   = ( m->fresnel_reflectivity != 0 && m->refractive_index != 1.0 ) ? 1.0 : 0.0;
        chromatic_reflectivity = m->chromatic_reflectivity;
        diffuse_reflectivity   = m->diffuse_reflectivity;
        transparent = ( m->transparency[ 0 ] * m->transparency[ 0 ] + m->transparency[ 1 ] * m->transparency[ 1 ] + m->transparency[ 2 ] * m->transparency[ 2 ] ) > 0;
        double sigma = m->sigma;
        if( sigma > 0 )
        {
            double sigma_sqr = sigma * sigma;
            on_a = 1.0 - 0.5 * sigma_sqr / ( sigma_sqr + 0.33 );
            on_b = 0.45 * sigma_sqr / ( sigma_sqr + 0.09 );
        }
    }

    if( trans->exit_obj )
    {
        trans_refractive_index /= trans->exit_obj->mat->refractive_index;
        fresnel_reflectivity = 1.0;
        diffuse_reflectivity = chromatic_reflectivity = 0;
        transparent = true;
    }

    /// fresnel reflection
    if( fresnel_reflectivity > 0 && intensity >= prm->trace_min_intensity )
    {
        CNT( cx, C_FRESNEL ); CNT( cx, R_REFLECT );
        ray_t out;
        out.p = pos;
        double reflectance = fresnel_reflection( ray->d, trans->exit_nor, trans_refractive_index, &out.d ) * fresnel_reflectivity;
        u3_t ckey = mix64( key, KEY_REFLECT );
        const int ph_save = cx->phase; cx->phase = PH_SPEC;
        trans_t trans_l; memset( &trans_l, 0, sizeof( trans_l ) );
        double a;
        v3d lum_l;
        if( ( a = scene_trans_hit( cx, &out, &trans_l, ckey ) ) < cx->inf ) lum_l = scene_lum( cx, &out, a, &trans_l, depth - 1, reflectance * intensity, ckey );
        else lum_l = mlf( bg, reflectance * intensity );
        cx->phase = ph_save;
        lum = add( lum, lum_l );
        intensity *= ( 1.0 - reflectance );
    }

    /// chromatic reflection
    if( chromatic_reflectivity > 0 && intensity >= prm->trace_min_intensity )
    {
        CNT( cx, C_MIRROR ); CNT( cx, R_CHROMATIC );
        ray_t out;
        out.p = pos;
        out.d = reflection( ray->d, trans->exit_nor );
        u3_t ckey = mix64( key, KEY_CHROMATIC );
        const int ph_save = cx->phase; cx->phase = PH_SPEC;
        trans_t trans_l; memset( &trans_l, 0, sizeof( trans_l ) );
        double a;
        v3d lum_l;
        if( ( a = scene_trans_hit( cx, &out, &trans_l, ckey ) ) < cx->inf ) lum_l = scene_lum( cx, &out, a, &trans_l, depth - 1, chromatic_reflectivity * intensity, ckey );
        else lum_l = mlf( bg, chromatic_reflectivity * intensity );
        cx->phase = ph_save;
        v3d cl = obj_color( trans->enter_obj, pos );
        lum = add( lum, mld( lum_l, cl ) );
        intensity *= ( 1.0 - chromatic_reflectivity );
    }

    /// diffuse reflection  (guard: the reference would dereference NULL for exit-only hits with Imin = 0)
    if( trans->enter_obj && intensity * diffuse_reflectivity >= prm->trace_min_intensity )
    {
        CNT( cx, R_DIFFUSE );
        double diffuse_intensity = intensity * diffuse_reflectivity;
        ray_t surface; surface.p = pos; surface.d = neg( trans->exit_nor );

        double theta_i = acos( -mlv( ray->d, surface.d ) );
        v3d ray_projection = of_length( orthogonal_projection( ray->d, surface.d ), 1.0 );

        u3_t rv = scene->seed_mode == ACN_SEED_POSITION_HASH
                      ? random_seed( surface.p, 3294479285ull ) + random_seed( surface.d, 3247146734ull )
                      : mix64( key, KEY_DIFFUSE );

        v3d lum_l = { 0, 0, 0 };
        const int ph_hit = cx->phase;
        cx->phase = PH_DIRECT;

        for( size_t i = 0; i < scene->light->el.size(); i++ )
        {
            v3d cl_sum = { 0, 0, 0 };
            ray_t out = surface;
            const obj_t* light_src = scene->light->el[ i ].obj;
            v3d axis; double cos_rs;
            obj_fov( light_src, pos, &axis, &cos_rs );
            m3d src_con = m_transposed( m_con_z( axis ) );
            double cyl_hgt = 1 - cos_rs;                                                  // areal_coverage, vectors.h:362
            v3d color = obj_color( light_src, light_src->pos );
            uint64_t direct_samples = ( uint64_t )( prm->direct_samples * diffuse_intensity );
            direct_samples = ( direct_samples == 0 ) ? 1 : direct_samples;

            for( uint64_t j = 0; j < direct_samples; j++ )
            {
                CNT( cx, C_CAP_SAMPLE );
                out.d = m_mlv( &src_con, random_sphere_cap( &rv, cyl_hgt ) );
                double weight = mlv( out.d, surface.d );
                if( weight <= 0 ) continue;

                CNT( cx, R_SHADOW );
                double a = obj_ray_hit( cx, light_src, &out, NULL, 0 );
                if( a >= cx->inf ) continue;

                if( on_b > 0 ) { CNT( cx, C_OREN_NAYAR ); weight = oren_nayar_weight( weight, theta_i, on_a, on_b, out.d, surface.d, ray_projection ); }

                CNT( cx, R_SHADOW );
                if( compound_ray_hit( cx, scene->matter, &out, NULL, NULL, 0 ) > a )
                {
                    CNT( cx, C_DIRECT_BOOK );
                    v3d hit_pos = ray_pos( &out, a );
                    double diff_sqr = vsqr( sub( hit_pos, light_src->pos ) );
                    double local_intensity = ( diff_sqr > 0 ) ? ( light_src->mat->radiance / diff_sqr ) : 1E+30;
                    cl_sum = add( cl_sum, mlf( color, local_intensity * weight * diffuse_intensity ) );
                }
            }
            lum_l = add( lum_l, mlf( cl_sum, 2.0 * cyl_hgt / direct_samples ) );
        }

        cx->phase = ph_hit;

        // path tracing
        if( prm->path_samples && depth > 10 )
        {
            cx->phase = PH_PATH;
            v3d cl_sum = { 0, 0, 0 };
            ray_t out = surface;
            m3d out_con = m_transposed( m_con_z( surface.d ) );
            uint64_t path_samples = ( uint64_t )( prm->path_samples * diffuse_intensity );
            path_samples = ( path_samples == 0 ) ? 1 : path_samples;

            for( uint64_t i = 0; i < path_samples; i++ )
            {
                CNT( cx, C_CAP_SAMPLE );
                out.d = m_mlv( &out_con, random_sphere_cap( &rv, 1.0 ) );
                double weight = mlv( out.d, surface.d );
                if( weight <= 0 ) continue;
                if( on_b > 0 ) { CNT( cx, C_OREN_NAYAR ); weight = oren_nayar_weight( weight, theta_i, on_a, on_b, out.d, surface.d, ray_projection ); }

                CNT( cx, R_PATH );
                u3_t ckey = mix64( key, KEY_PATH0 + i );
                trans_t trans_l; memset( &trans_l, 0, sizeof( trans_l ) );
                double a = compound_ray_trans_hit( cx, scene->matter, &out, &trans_l, ckey );
                if( a < prm->max_path_length ) cl_sum = add( cl_sum, scene_lum( cx, &out, a, &trans_l, depth - 10, weight * diffuse_intensity, ckey ) );
                else cl_sum = add( cl_sum, mlf( bg, weight * diffuse_intensity ) );
            }
            lum_l = add( lum_l, mlf( cl_sum, 2.0 / path_samples ) );
            cx->phase = ph_hit;
        }

        v3d cl = obj_color( trans->enter_obj, pos );
        lum = add( lum, mld( lum_l, cl ) );
        intensity *= ( 1.0 - diffuse_reflectivity );
    }

    /// refraction
    if( transparent && intensity >= prm->trace_min_intensity )
    {
        CNT( cx, C_REFRACT ); CNT( cx, R_REFRACT );
        ray_t out;
        out.p = ray_pos( ray, offs + 2.0 * cx->eps );
        fresnel_refraction( ray->d, trans->exit_nor, trans_refractive_index, &out.d );
        u3_t ckey = mix64( key, KEY_REFRACT );
        const int ph_save = cx->phase; cx->phase = PH_SPEC;
        trans_t trans_l; memset( &trans_l, 0, sizeof( trans_l ) );
        double a;
        v3d lum_l;
        if( ( a = scene_trans_hit( cx, &out, &trans_l, ckey ) ) < cx->inf ) lum_l = scene_lum( cx, &out, a, &trans_l, depth - 1, intensity, ckey );
        else lum_l = mlf( bg, intensity );
        cx->phase = ph_save;
        lum = add( lum, lum_l );
    }

    /// exiting object
    if( trans->exit_obj )
    {
        CNT( cx, C_ABSORB );
        const double* t = trans->exit_obj->mat->transparency;
        double rf = offs > 0 ? pow( t[ 0 ], offs ) : 1.0;
        double gf = offs > 0 ? pow( t[ 1 ], offs ) : 1.0;
        double bf = offs > 0 ? pow( t[ 2 ], offs ) : 1.0;
        lum.x *= rf; lum.y *= gf; lum.z *= bf;
    }
    return lum;
}

static inline v3d cl_sat( v3d o, double gamma )     // vectors.h:372-384
{
    double x = pow( o.x, gamma ), y = pow( o.y, gamma ), z = pow( o.z, gamma );
    x = x > 0.0 ? x < 1.0 ? x : 1.0 : 0.0;
    y = y > 0.0 ? y < 1.0 ? y : 1.0 : 0.0;
    z = z > 0.0 ? z < 1.0 ? z : 1.0 : 0.0;
    return V( x, y, z );
}

// lum_machine_s_func (scene.c:956-1013) for one worker
struct machine_t
{
    const oscene_t* sc;
    const double* xy; double* rgb; double* lin;
    uint64_t n, index_base;
    std::atomic<uint64_t> index;
    machine_t() : index( 0 ) {}
};

static void machine_func( machine_t* o, counters_t* cnt )
{
    const oscene_t* sc = o->sc;
    const acn_flat_params* prm = &sc->prm;
    ctx_t cx; cx.sc = sc; cx.cnt = cnt; cx.eps = sc->eps; cx.inf = INFINITY; cx.phase = PH_PRIMARY;
    uint64_t width = prm->image_width, height = prm->image_height;
    uint64_t unit_sz = ( height >> 1 );
    double unit_f = 1.0 / unit_sz;

    m3d camera_rotation;
    {
        v3d ry = of_length( V( prm->camera_view_direction[ 0 ], prm->camera_view_direction[ 1 ], prm->camera_view_direction[ 2 ] ), 1 );
        v3d rz = of_length( V( prm->camera_top_direction[ 0 ], prm->camera_top_direction[ 1 ], prm->camera_top_direction[ 2 ] ), 1 );
        rz = von( ry, rz );
        v3d rx = mlx( ry, rz );
        camera_rotation.x = rx; camera_rotation.y = ry; camera_rotation.z = rz;
        camera_rotation = m_transposed( camera_rotation );
    }

    uint64_t index;
    while( ( index = o->index.fetch_add( 1 ) ) < o->n )
    {
        CNT( &cx, C_CAMERA ); CNT( &cx, R_PRIMARY ); CNT( &cx, C_GAMMA );
        double monitor_x = o->xy[ 2 * index ], monitor_y = o->xy[ 2 * index + 1 ];
        double z = unit_f * ( ( double )( height >> 1 ) - monitor_y );
        double x = unit_f * ( monitor_x - ( double )( width >> 1 ) );
        v3d d = of_length( V( x, prm->camera_focal_length, z ), 1.0 );
        ray_t ray;
        ray.p = V( prm->camera_position[ 0 ], prm->camera_position[ 1 ], prm->camera_position[ 2 ] );
        ray.d = m_mlv( &camera_rotation, d );
        v3d out_clr = V( prm->background_color[ 0 ], prm->background_color[ 1 ], prm->background_color[ 2 ] );
        u3_t key = mix64( o->index_base + index, 0x5EEDull );
        trans_t trans_l; memset( &trans_l, 0, sizeof( trans_l ) );
        double offs = scene_trans_hit( &cx, &ray, &trans_l, key );
        if( offs < cx.inf ) out_clr = scene_lum( &cx, &ray, offs, &trans_l, prm->trace_depth, 1.0, key );
        if( o->lin ) { o->lin[ 3 * index ] = out_clr.x; o->lin[ 3 * index + 1 ] = out_clr.y; o->lin[ 3 * index + 2 ] = out_clr.z; }
        v3d c = cl_sat( out_clr, prm->gamma );
        o->rgb[ 3 * index ] = c.x; o->rgb[ 3 * index + 1 ] = c.y; o->rgb[ 3 * index + 2 ] = c.z;
    }
}

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench only)
// ---------------------------------------------------------------------------------------------
extern "C" {

int oracle_counter_count( void ) { return C_COUNT; }
double oracle_flops_per_event( int k ) { return ( k >= 0 && k < C_COUNT ) ? flops_per_event[ k ] : 0; }
double oracle_sf_per_event( int k ) { return ( k >= 0 && k < C_COUNT ) ? sf_per_event[ k ] : 0; }

// lum_machine_s_run (scene.c:1017-1028).  rgb: post-gamma, clamped (double[3n]); lin: optional
// pre-gamma linear radiance.  counters: uint64[C_COUNT] summed over threads.  returns wall seconds
// in *seconds.
int oracle_render( const acn_flat_scene* fs, const double* xy, uint64_t n, uint64_t index_base, int seed_mode, double eps,
                   int threads, double* rgb, double* lin, uint64_t* counters, double* seconds, double* phase_flops )
{
    if( !fs || ( n && ( !xy || !rgb ) ) ) return -1;
    oscene_t sc;
    sc.prm = fs->params;
    sc.eps = eps > 0 ? eps : 1E-6;
    sc.seed_mode = seed_mode;
    sc.light = build_cmp( &sc, fs, fs->light_root );
    sc.matter = build_cmp( &sc, fs, fs->matter_root );
    for( size_t i = 0; i < sc.light->el.size(); i++ )
    {
        v3d ax; double c;
        if( !sc.light->el[ i ].obj || !obj_fov( sc.light->el[ i ].obj, V( 0, 0, 0 ), &ax, &c ) ) return -5;
    }
    machine_t m;
    m.sc = &sc; m.xy = xy; m.rgb = rgb; m.lin = lin; m.n = n; m.index_base = index_base;
    if( threads < 1 ) threads = 1;
    std::vector<counters_t> cnt( threads );
    for( auto& c : cnt ) memset( &c, 0, sizeof( c ) );
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for( int i = 1; i < threads; i++ ) th.emplace_back( machine_func, &m, &cnt[ i ] );
    machine_func( &m, &cnt[ 0 ] );
    for( auto& t : th ) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if( seconds ) *seconds = std::chrono::duration<double>( t1 - t0 ).count();
    if( counters )
    {
        for( int k = 0; k < C_COUNT; k++ ) { counters[ k ] = 0; for( int i = 0; i < threads; i++ ) for( int ph = 0; ph < PH_COUNT; ph++ ) counters[ k ] += cnt[ i ].c[ ph ][ k ]; }
    }
    if( phase_flops )
    {
        for( int ph = 0; ph < PH_COUNT; ph++ )
        {
            phase_flops[ ph ] = 0;
            for( int k = 0; k < C_COUNT; k++ ) for( int i = 0; i < threads; i++ ) phase_flops[ ph ] += ( double )cnt[ i ].c[ ph ][ k ] * flops_per_event[ k ];
        }
    }
    return 0;
}

// lum_image_s_push_arr (scene.c:804-820): sums[w*h*4] = r,g,b,weight
int oracle_accumulate( const double* xy, const double* rgb, uint64_t n, int width, int height, double* sums )
{
    for( uint64_t i = 0; i < n; i++ )
    {
        int x = ( int )xy[ 2 * i ], y = ( int )xy[ 2 * i + 1 ];
        if( x >= 0 && x < width && y >= 0 && y < height )
        {
            double* p = sums + 4 * ( ( size_t )y * width + x );
            p[ 0 ] += rgb[ 3 * i ]; p[ 1 ] += rgb[ 3 * i + 1 ]; p[ 2 ] += rgb[ 3 * i + 2 ]; p[ 3 ] += 1.0;
        }
    }
    return 0;
}

// ---- leaf functions exposed for pinning against oracle/_ref (the reference's own code) ----
double oracle_sphere_ray_hit( const double pos[ 3 ], double r, const double rp[ 3 ], const double rd[ 3 ], double eps, double nor[ 3 ] )
{
    counters_t c; memset( &c, 0, sizeof( c ) );
    ctx_t cx; cx.sc = NULL; cx.cnt = &c; cx.eps = eps; cx.inf = INFINITY; cx.phase = 0;
    ray_t ray = { V( rp[ 0 ], rp[ 1 ], rp[ 2 ] ), V( rd[ 0 ], rd[ 1 ], rd[ 2 ] ) };
    v3d n = V( 0, 0, 0 );
    double a = sphere_ray_hit( &cx, V( pos[ 0 ], pos[ 1 ], pos[ 2 ] ), r, &ray, &n );
    if( nor ) { nor[ 0 ] = n.x; nor[ 1 ] = n.y; nor[ 2 ] = n.z; }
    return a;
}
double oracle_plane_ray_hit( const double pos[ 3 ], const double pn[ 3 ], const double rp[ 3 ], const double rd[ 3 ], double eps )
{
    counters_t c; memset( &c, 0, sizeof( c ) );
    ctx_t cx; cx.sc = NULL; cx.cnt = &c; cx.eps = eps; cx.inf = INFINITY; cx.phase = 0;
    ray_t ray = { V( rp[ 0 ], rp[ 1 ], rp[ 2 ] ), V( rd[ 0 ], rd[ 1 ], rd[ 2 ] ) };
    return plane_ray_hit( &cx, V( pos[ 0 ], pos[ 1 ], pos[ 2 ] ), V( pn[ 0 ], pn[ 1 ], pn[ 2 ] ), &ray, NULL );
}
double oracle_fresnel_reflection( const double d[ 3 ], const double n[ 3 ], double trix, double out_dir[ 3 ] )
{
    v3d o;
    double r = fresnel_reflection( V( d[ 0 ], d[ 1 ], d[ 2 ] ), V( n[ 0 ], n[ 1 ], n[ 2 ] ), trix, &o );
    out_dir[ 0 ] = o.x; out_dir[ 1 ] = o.y; out_dir[ 2 ] = o.z;
    return r;
}
void oracle_fresnel_refraction( const double d[ 3 ], const double n[ 3 ], double trix, double out_dir[ 3 ] )
{
    v3d o;
    fresnel_refraction( V( d[ 0 ], d[ 1 ], d[ 2 ] ), V( n[ 0 ], n[ 1 ], n[ 2 ] ), trix, &o );
    out_dir[ 0 ] = o.x; out_dir[ 1 ] = o.y; out_dir[ 2 ] = o.z;
}
void oracle_con_z( const double v[ 3 ], double m[ 9 ] )
{
    m3d r = m_con_z( V( v[ 0 ], v[ 1 ], v[ 2 ] ) );
    m[ 0 ] = r.x.x; m[ 1 ] = r.x.y; m[ 2 ] = r.x.z; m[ 3 ] = r.y.x; m[ 4 ] = r.y.y; m[ 5 ] = r.y.z; m[ 6 ] = r.z.x; m[ 7 ] = r.z.y; m[ 8 ] = r.z.z;
}
uint64_t oracle_random_seed( const double v[ 3 ], uint64_t rv ) { return random_seed( V( v[ 0 ], v[ 1 ], v[ 2 ] ), rv ); }
void oracle_random_sphere_cap( uint64_t* rv, double h, double out[ 3 ] ) { v3d v = random_sphere_cap( rv, h ); out[ 0 ] = v.x; out[ 1 ] = v.y; out[ 2 ] = v.z; }
void oracle_cl_sat( const double c[ 3 ], double gamma, double out[ 3 ] ) { v3d v = cl_sat( V( c[ 0 ], c[ 1 ], c[ 2 ] ), gamma ); out[ 0 ] = v.x; out[ 1 ] = v.y; out[ 2 ] = v.z; }
double oracle_oren_nayar_weight( double weight, double theta_i, double on_a, double on_b, const double out_d[ 3 ], const double nor[ 3 ], const double prj[ 3 ] )
{
    return oren_nayar_weight( weight, theta_i, on_a, on_b, V( out_d[ 0 ], out_d[ 1 ], out_d[ 2 ] ), V( nor[ 0 ], nor[ 1 ], nor[ 2 ] ), V( prj[ 0 ], prj[ 1 ], prj[ 2 ] ) );
}

} // extern "C"
