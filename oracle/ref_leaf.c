/* ref_leaf.c — thin C wrappers around the REFERENCE'S OWN leaf math, compiled in place from
 * /root/reference/src (vectors.h, gmath.h, gmath.c) against oracle/shim/bcore_std.h.
 * Output: oracle/_ref/libacn_refleaf.so (git-ignored).  Used by tests to pin the oracle's
 * restatement of these functions.  No reference source is copied into this repository.
 * TEST INFRASTRUCTURE ONLY. */
#include "vectors.h"
#include "gmath.h"

static v3d_s v3( const double* p ) { v3d_s v = { p[ 0 ], p[ 1 ], p[ 2 ] }; return v; }
static void put( double* o, v3d_s v ) { o[ 0 ] = v.x; o[ 1 ] = v.y; o[ 2 ] = v.z; }

double ref_sphere_ray_hit( const double pos[ 3 ], double r, const double rp[ 3 ], const double rd[ 3 ], double nor[ 3 ] )
{
    ray_s ray = { v3( rp ), v3( rd ) };
    v3d_s n = { 0, 0, 0 };
    double a = sphere_ray_hit( v3( pos ), r, &ray, &n );
    if( nor ) put( nor, n );
    return a;
}
double ref_plane_ray_hit( const double pos[ 3 ], const double pn[ 3 ], const double rp[ 3 ], const double rd[ 3 ] )
{
    ray_s ray = { v3( rp ), v3( rd ) };
    return plane_ray_hit( v3( pos ), v3( pn ), &ray, NULL );
}
int ref_sphere_observer_side( const double pos[ 3 ], double r, const double o[ 3 ] ) { return sphere_observer_side( v3( pos ), r, v3( o ) ); }
int ref_plane_observer_side( const double pos[ 3 ], const double n[ 3 ], const double o[ 3 ] ) { return plane_observer_side( v3( pos ), v3( n ), v3( o ) ); }
double ref_fresnel_reflection( const double d[ 3 ], const double n[ 3 ], double trix, double out_dir[ 3 ] )
{
    v3d_s o;
    double r = fresnel_reflection( v3( d ), v3( n ), trix, &o );
    put( out_dir, o );
    return r;
}
void ref_fresnel_refraction( const double d[ 3 ], const double n[ 3 ], double trix, double out_dir[ 3 ] )
{
    v3d_s o;
    fresnel_refraction( v3( d ), v3( n ), trix, &o );
    put( out_dir, o );
}
void ref_con_z( const double v[ 3 ], double m[ 9 ] )
{
    m3d_s r = m3d_s_con_z( v3( v ) );
    put( m, r.x ); put( m + 3, r.y ); put( m + 6, r.z );
}
void ref_con( const double v[ 3 ], double o[ 3 ] ) { put( o, v3d_s_con( v3( v ) ) ); }
void ref_of_length( const double v[ 3 ], double a, double o[ 3 ] ) { put( o, v3d_s_of_length( v3( v ), a ) ); }
void ref_reflection( const double d[ 3 ], const double n[ 3 ], double o[ 3 ] ) { put( o, v3d_s_reflection( v3( d ), v3( n ) ) ); }
uint64_t ref_random_seed( const double v[ 3 ], uint64_t rv ) { return v3d_s_random_seed( v3( v ), rv ); }
uint64_t ref_seed_from_f3( double v ) { return v3d_s_seed_from_f3( v ); }
void ref_random_sphere_cap( uint64_t* rv, double h, double out[ 3 ] ) { put( out, v3d_s_random_sphere_cap( rv, h ) ); }
double ref_rnd0( uint64_t* rv ) { return f3_rnd0( rv ); }
double ref_rnd1( uint64_t* rv ) { return f3_rnd1( rv ); }
void ref_cl_sat( const double c[ 3 ], double gamma, double out[ 3 ] ) { put( out, cl_s_sat( v3( c ), gamma ) ); }
double ref_eps( void ) { return f3_eps; }
