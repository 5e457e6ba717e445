/* Stand-in for beth's bcore_std.h — just enough for the reference's vectors.h / gmath.h / gmath.c to
 * compile IN PLACE from /root/reference/src (see oracle/Makefile).  beth (github.com/johsteffens/beth,
 * version unpinned by the reference's makefile) is not available here; nothing in this file comes
 * from it.  The three LCGs are the oracle's placeholders (same constants as oracle/acn_oracle.cpp).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ACN_SHIM_BCORE_STD_H
#define ACN_SHIM_BCORE_STD_H

#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <math.h>

typedef double   f3_t;
typedef float    f2_t;
typedef uint64_t u3_t;
typedef uint32_t u2_t;
typedef uint8_t  u0_t;
typedef int64_t  s3_t;
typedef int32_t  s2_t;
typedef size_t   uz_t;
typedef bool     bl_t;
typedef uint64_t tp_t;
typedef uint64_t aware_t;
typedef void*       vd_t;
typedef const void* vc_t;
typedef const char* sc_t;

typedef struct bcore_signal_s { int unused; } bcore_signal_s;
typedef struct bcore_array_dyn_solid_static_s { void* data; uz_t size, space; } bcore_array_dyn_solid_static_s;

#define BCORE_DECLARE_FUNCTIONS_OBJ( name )
#define TYPEOF_init1 1
/* beth's typeof("name") is a type hash; `typeof` is a GNU keyword, so shadow it with a macro */
#define typeof( name ) ( ( tp_t )0 )
static inline tp_t bcore_signal_s_handle_type( const bcore_signal_s* o, tp_t t ) { (void)o; (void)t; return 0; }

static inline f3_t f3_sqr( f3_t v ) { return v * v; }
static inline f3_t f3_abs( f3_t v ) { return v < 0 ? -v : v; }
static inline f3_t f3_max( f3_t a, f3_t b ) { return a > b ? a : b; }
static inline f3_t f3_min( f3_t a, f3_t b ) { return a < b ? a : b; }

static inline u3_t bcore_lcg00_u3( u3_t v ) { return v * 6364136223846793005ull + 1442695040888963407ull; }
static inline u3_t bcore_lcg01_u3( u3_t v ) { return v * 3935559000370003845ull + 2691343689449507681ull; }
static inline u3_t bcore_lcg02_u3( u3_t v ) { return v * 2862933555777941757ull + 3037000493ull; }

#endif
