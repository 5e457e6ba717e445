// acn_math.h — vector algebra, RNG and surface-response leaf math of the sample tracer.
//
// Host+device templates on the real type R (float = product path, double = validation mode and
// host-side envelope estimation).  Each function names the reference code whose *behaviour* it
// reproduces; the formulations are chosen for FP32 (rejection-vector discriminants, algebraic
// Oren-Nayar, rsqrt normalisation), not transcribed.
#pragma once

#if defined(__CUDACC_RTC__)
// run-time compilation (NVRTC, acn_spec.h): no host headers; the math functions are built in
typedef unsigned long long uint64_t;
typedef long long          int64_t;
typedef unsigned int       uint32_t;
typedef int                int32_t;
#ifndef INFINITY
#define INFINITY ( __int_as_float( 0x7f800000 ) )
#endif
#else
#include <stdint.h>
#include <math.h>
#endif

#if defined(__CUDACC__)
#define ACN_HD  __host__ __device__ __forceinline__
#define ACN_HDN __host__ __device__ __noinline__
#define ACN_NOINLINE __host__ __device__ __noinline__
#else
#define ACN_HD  inline
#define ACN_HDN
#define ACN_NOINLINE
#endif

namespace acn {

typedef uint64_t u64;
typedef int64_t  s64;

// ---------------------------------------------------------------------------------------------
// real-type traits
// ---------------------------------------------------------------------------------------------
template <typename R> struct Num;
template <> struct Num<float>
{
    static ACN_HD float inf() { return INFINITY; }
    static ACN_HD float mag() { return 1E+30f; }             // f3_mag, vectors.h:32
    static ACN_HD float unit_tol() { return 0.0f; }          // of_length shortcut disabled in f32
};
template <> struct Num<double>
{
    static ACN_HD double inf() { return INFINITY; }
    static ACN_HD double mag() { return 1E+30; }
    static ACN_HD double unit_tol() { return 1E-8; }         // vectors.h:151
};

ACN_HD float  r_sqrt( float v )  { return sqrtf( v ); }
ACN_HD double r_sqrt( double v ) { return sqrt( v ); }
ACN_HD float  r_abs( float v )   { return fabsf( v ); }
ACN_HD double r_abs( double v )  { return fabs( v ); }
ACN_HD float  r_min( float a, float b )   { return fminf( a, b ); }
ACN_HD double r_min( double a, double b ) { return fmin( a, b ); }
ACN_HD float  r_max( float a, float b )   { return fmaxf( a, b ); }
ACN_HD double r_max( double a, double b ) { return fmax( a, b ); }
ACN_HD float  r_pow( float a, float b )   { return powf( a, b ); }
ACN_HD double r_pow( double a, double b ) { return pow( a, b ); }
ACN_HD float  r_log( float a )  { return logf( a ); }
ACN_HD double r_log( double a ) { return log( a ); }
ACN_HD float  r_acos( float a )  { return acosf( a ); }
ACN_HD double r_acos( double a ) { return acos( a ); }
ACN_HD float  r_atan2( float a, float b )   { return atan2f( a, b ); }
ACN_HD double r_atan2( double a, double b ) { return atan2( a, b ); }
ACN_HD float  r_asin( float a )  { return asinf( a ); }
ACN_HD double r_asin( double a ) { return asin( a ); }

// 1/sqrt: hardware MUFU.RSQ + one Newton step in f32 device code
ACN_HD float r_rsqrt( float v )
{
#if defined(__CUDA_ARCH__)
    float y = rsqrtf( v );
    return y * ( 1.5f - 0.5f * v * y * y );
#else
    return 1.0f / sqrtf( v );
#endif
}
ACN_HD double r_rsqrt( double v ) { return 1.0 / sqrt( v ); }

// sin/cos of 2*pi*u
ACN_HD void r_sincos_2pi( float u, float* s, float* c )
{
#if defined(__CUDA_ARCH__)
    sincospif( 2.0f * u, s, c );
#else
    float phi = 6.283185307179586f * u; *s = sinf( phi ); *c = cosf( phi );
#endif
}
ACN_HD void r_sincos_2pi( double u, double* s, double* c )
{
    double phi = 2.0 * 3.14159265358979323846 * u; *s = sin( phi ); *c = cos( phi );
}

// ---------------------------------------------------------------------------------------------
// v3d_s (vectors.h:106-175)
// ---------------------------------------------------------------------------------------------
template <typename R> struct V3 { R x, y, z; };

template <typename R> ACN_HD V3<R> v3( R x, R y, R z ) { V3<R> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename R> ACN_HD V3<R> operator+( V3<R> a, V3<R> b ) { return v3<R>( a.x + b.x, a.y + b.y, a.z + b.z ); }
template <typename R> ACN_HD V3<R> operator-( V3<R> a, V3<R> b ) { return v3<R>( a.x - b.x, a.y - b.y, a.z - b.z ); }
template <typename R> ACN_HD V3<R> operator-( V3<R> a ) { return v3<R>( -a.x, -a.y, -a.z ); }
template <typename R> ACN_HD V3<R> operator*( V3<R> a, R f ) { return v3<R>( a.x * f, a.y * f, a.z * f ); }
template <typename R> ACN_HD V3<R> mul( V3<R> a, V3<R> b ) { return v3<R>( a.x * b.x, a.y * b.y, a.z * b.z ); }   // v3d_s_mld
template <typename R> ACN_HD R dot( V3<R> a, V3<R> b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> ACN_HD R sqr( V3<R> a ) { return dot( a, a ); }
template <typename R> ACN_HD V3<R> cross( V3<R> o, V3<R> f )                                                     // v3d_s_mlx
{
    return v3<R>( o.y * f.z - o.z * f.y, o.z * f.x - o.x * f.z, o.x * f.y - o.y * f.x );
}
template <typename R> ACN_HD V3<R> madd( V3<R> p, V3<R> d, R t ) { return v3<R>( p.x + d.x * t, p.y + d.y * t, p.z + d.z * t ); } // ray_s_pos

// v3d_s_of_length( o, 1 ) (vectors.h:148-154): unchanged if already unit within 1e-8 (f64), zero stays zero
template <typename R> ACN_HD V3<R> unit( V3<R> o )
{
    R r2 = sqr( o );
    if( Num<R>::unit_tol() > R( 0 ) && r_abs( r2 - R( 1 ) ) < Num<R>::unit_tol() ) return o;
    R f = r2 > R( 0 ) ? r_rsqrt( r2 ) : R( 0 );
    return o * f;
}

// v3d_s_von (vectors.h:157-162): v made orthonormal to o inside the plane (o,v)
template <typename R> ACN_HD V3<R> von( V3<R> o, V3<R> v )
{
    V3<R> on = unit( o );
    return unit( v - on * dot( on, v ) );
}

// v3d_s_con (vectors.h:165-175): canonical orthonormal — unit axes of the smallest |component|
template <typename R> ACN_HD V3<R> con( V3<R> o )
{
    R xx = o.x * o.x, yy = o.y * o.y, zz = o.z * o.z;
    V3<R> v;
    v.x = ( xx <= yy && xx <= zz ) ? R( 1 ) : R( 0 );
    v.y = ( yy <= xx && yy <= zz ) ? R( 1 ) : R( 0 );
    v.z = ( zz <= xx && zz <= yy ) ? R( 1 ) : R( 0 );
    return von( o, v );
}

// v3d_s_reflection (vectors.h:238-241)
template <typename R> ACN_HD V3<R> reflect( V3<R> dir, V3<R> nor )
{
    return unit( dir - nor * ( R( 2 ) * dot( dir, nor ) ) );
}

// ---------------------------------------------------------------------------------------------
// m3d_s (vectors.h:246-332): rows x,y,z
// ---------------------------------------------------------------------------------------------
template <typename R> struct M3 { V3<R> x, y, z; };

template <typename R> ACN_HD V3<R> mlv( const M3<R>& m, V3<R> v )  { return v3<R>( dot( m.x, v ), dot( m.y, v ), dot( m.z, v ) ); }
template <typename R> ACN_HD V3<R> tmlv( const M3<R>& m, V3<R> v ) { return m.x * v.x + m.y * v.y + m.z * v.z; }

// columns of transposed( m3d_s_con_z( v ) ) as used at scene.c:550,588:  world = X*a + Y*b + Z*c
template <typename R> struct Basis { V3<R> X, Y, Z; };
template <typename R> ACN_HD Basis<R> basis_con_z( V3<R> v )
{
    Basis<R> b;
    b.Z = unit( v );
    b.X = con( v );
    b.Y = cross( b.Z, b.X );
    return b;
}
template <typename R> ACN_HD V3<R> from_basis( const Basis<R>& b, V3<R> v ) { return b.X * v.x + b.Y * v.y + b.Z * v.z; }

// ---------------------------------------------------------------------------------------------
// RNG (vectors.h:45-48,177-190).  The three LCGs live in beth (bcore_lcg00/01/02_u3), which is not
// part of the reference tree; these constants are PLACEHOLDERS (Knuth MMIX / L'Ecuyer full-period
// multipliers): statistically equivalent, not bit-compatible with upstream.  Shared with oracle/.
// ---------------------------------------------------------------------------------------------
#define ACN_LCG00_A 6364136223846793005ull
#define ACN_LCG00_C 1442695040888963407ull
#define ACN_LCG01_A 3935559000370003845ull
#define ACN_LCG01_C 2691343689449507681ull
#define ACN_LCG02_A 2862933555777941757ull
#define ACN_LCG02_C 3037000493ull

ACN_HD u64 lcg00( u64 v ) { return v * ACN_LCG00_A + ACN_LCG00_C; }
ACN_HD u64 lcg01( u64 v ) { return v * ACN_LCG01_A + ACN_LCG01_C; }
ACN_HD u64 lcg02( u64 v ) { return v * ACN_LCG02_A + ACN_LCG02_C; }

// u64 -> [0,1]: rv * (1.0 / 0xFFFFFFFFFFFFFFFF); the divisor rounds to 2^64 in double (vectors.h:48)
ACN_HD float u64_to_unit( u64 v, float )
{
#if defined(__CUDA_ARCH__)
    return __ull2float_rn( v ) * 5.42101086242752217e-20f;
#else
    return ( float )v * 5.42101086242752217e-20f;
#endif
}
ACN_HD double u64_to_unit( u64 v, double ) { return ( double )v * 5.42101086242752217e-20; }

template <typename R> ACN_HD R rnd1( u64* rv ) { *rv = lcg00( *rv ); return u64_to_unit( *rv, R( 0 ) ); }            // f3_rnd1
template <typename R> ACN_HD R rnd0( u64* rv ) { *rv = lcg00( *rv ); return u64_to_unit( *rv, R( 0 ) ) * R( 2 ) - R( 1 ); } // f3_rnd0

// LCG skip-ahead: state after n steps of lcg00 (Brown, "Random number generation with arbitrary strides")
ACN_HD u64 lcg00_skip( u64 v, u64 n )
{
    u64 a = ACN_LCG00_A, c = ACN_LCG00_C, A = 1, C = 0;
    while( n )
    {
        if( n & 1 ) { A *= a; C = C * a + c; }
        c *= ( a + 1 );
        a *= a;
        n >>= 1;
    }
    return v * A + C;
}

// v3d_s_seed_from_f3 (vectors.h:177-182): (s64)( frexp(v).mantissa * 0x7FFF...F ) * 27362149.
// 0x7FFFFFFFFFFFFFFF converts to 2^63 as a double, so the product is the 53-bit significand
// shifted left by 10, with sign — taken straight from the bits of the double.
ACN_HD u64 seed_from_f3( double v )
{
    union { double d; u64 u; } cv; cv.d = v;
    u64 bits = cv.u;
    int  e   = ( int )( ( bits >> 52 ) & 0x7FF );
    u64  m   = bits & 0xFFFFFFFFFFFFFull;
    s64  s;
    if( e == 0 )
    {
        if( m == 0 ) return 0;                       // frexp(0) = 0
        // subnormal: normalise the significand
        int sh = 0; while( !( m & 0x10000000000000ull ) ) { m <<= 1; sh++; }
        s = ( s64 )( m << 10 );
    }
    else if( e == 0x7FF )
    {
        return 0;                                    // inf/nan: undefined in the reference; pin to 0
    }
    else
    {
        s = ( s64 )( ( m | 0x10000000000000ull ) << 10 );
    }
    if( bits >> 63 ) s = -s;
    return ( u64 )s * 27362149ull;              // two's-complement wrap like the reference's s3_t *=
}

// v3d_s_random_seed (vectors.h:185-190)
template <typename R> ACN_HD u64 random_seed( V3<R> o, u64 rv )
{
    return seed_from_f3( ( double )o.x ) * lcg00( rv ) +
           seed_from_f3( ( double )o.y ) * lcg01( rv ) +
           seed_from_f3( ( double )o.z ) * lcg02( rv );
}

// index-keyed seeding (ACN_SEED_INDEX_KEYED): splitmix64 finaliser over (key, salt)
ACN_HD u64 mix64( u64 key, u64 salt )
{
    u64 z = key + salt * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    return z ^ ( z >> 31 );
}

// branch codes of the ray tree for index-keyed seeding
enum { KEY_REFLECT = 1, KEY_CHROMATIC = 2, KEY_REFRACT = 3, KEY_DIFFUSE = 4, KEY_ROUGH = 5, KEY_PATH0 = 16 };

// v3d_s_random_sphere_cap (vectors.h:197-206): uniform on the cap of height h around +z, 2 draws
template <typename R> ACN_HD V3<R> sphere_cap( u64* rv, R h )
{
    R u = rnd1<R>( rv );
    R z = R( 1 ) - rnd1<R>( rv ) * h;
    R sc = r_sqrt( r_max( R( 1 ) - z * z, R( 0 ) ) );
    R s, c;
    r_sincos_2pi( u, &s, &c );
    return v3<R>( s * sc, c * sc, z );
}

// ---------------------------------------------------------------------------------------------
// Fresnel / refraction (gmath.c:68-113)
// ---------------------------------------------------------------------------------------------
// unpolarised reflectance; trix = refractive-index ratio in ray direction; TIR -> 1
template <typename R> ACN_HD R fresnel_reflectance( V3<R> dir, V3<R> exit_nor, R trix )
{
    R c = dot( dir, exit_nor );
    R f = c < R( 0 ) ? trix : R( 1 ) / trix;
    R cos_ai = r_min( r_abs( c ), R( 1 ) );
    R sin_at = r_sqrt( R( 1 ) - cos_ai * cos_ai ) * f;
    if( !( sin_at < R( 1 ) ) ) return R( 1 );
    R cos_at = r_sqrt( R( 1 ) - sin_at * sin_at );
    R rs = ( f * cos_ai - cos_at ) / ( f * cos_ai + cos_at );
    R rp = ( f * cos_at - cos_ai ) / ( f * cos_at + cos_ai );
    return ( rs * rs + rp * rp ) * R( 0.5 );
}

// Snell direction; falls back to the incident direction beyond the critical angle (gmath.c:94-113)
template <typename R> ACN_HD V3<R> refract( V3<R> dir, V3<R> exit_nor, R trix )
{
    R c = dot( dir, exit_nor );
    R f = c < R( 0 ) ? trix : R( 1 ) / trix;
    R q = f * f * ( R( 1 ) - c * c );
    if( q < R( 1 ) )
    {
        R sq = r_sqrt( R( 1 ) - q );
        R b = -f * c + ( c > R( 0 ) ? sq : -sq );
        return dir * f + exit_nor * b;
    }
    return dir;
}

// ---------------------------------------------------------------------------------------------
// Oren-Nayar weight (scene.c:394-416).  theta = acos(cos) for both angles, so
// sin(max theta) * tan(min theta) = sqrt(1-cmin^2) * sqrt(1-cmax^2) / cmax  — no trig calls.
//   w       cos(theta_r) = out_d . nor  (> 0)
//   cos_i   cos(theta_i) = -ray.d . nor
//   ray_prj unit projection of the incoming direction onto the surface
// ---------------------------------------------------------------------------------------------
template <typename R> ACN_HD R oren_nayar( R w, R cos_i, R on_a, R on_b, V3<R> out_d, V3<R> nor, V3<R> ray_prj )
{
    V3<R> op = unit( out_d - nor * dot( out_d, nor ) );
    R cos_phi = -dot( op, ray_prj );
    R ci = r_min( r_max( cos_i, R( -1 ) ), R( 1 ) );
    R cmax = r_max( ci, w ), cmin = r_min( ci, w );
    R st = r_sqrt( r_max( R( 1 ) - cmin * cmin, R( 0 ) ) ) * r_sqrt( r_max( R( 1 ) - cmax * cmax, R( 0 ) ) ) / cmax;
    return w * ( on_a + on_b * r_max( cos_phi, R( 0 ) ) * st );
}

// Oren-Nayar A,B from sigma (scene.c:455-461)
template <typename R> ACN_HD void oren_nayar_ab( R sigma, R* a, R* b )
{
    *a = R( 1 ); *b = R( 0 );
    if( sigma > R( 0 ) )
    {
        R s2 = sigma * sigma;
        *a = R( 1 ) - R( 0.5 ) * s2 / ( s2 + R( 0.33 ) );
        *b = R( 0.45 ) * s2 / ( s2 + R( 0.09 ) );
    }
}

} // namespace acn
