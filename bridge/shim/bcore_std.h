/* Stand-in for beth's bcore_std.h, just enough for `gcc -fsyntax-only` of bridge/acn_bridge.c against the reference's own
 * objects.h / compound.h / scene.h where they lie (bridge/Makefile).  beth (github.com/johsteffens/beth) is not available
 * here; nothing in this file comes from it: the types below only have to make the reference's DECLARATIONS parse.
 * With the real beth headers on the include path this directory is not used. */
#ifndef ACN_BRIDGE_SHIM_BCORE_STD_H
#define ACN_BRIDGE_SHIM_BCORE_STD_H
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <string.h>
#include <signal.h>
#include <math.h>
#include <time.h>
#include <stdarg.h>
typedef double   f3_t;
typedef float    f2_t;
typedef uint64_t u3_t;
typedef uint32_t u2_t;
typedef uint8_t  u0_t;
typedef int64_t  s3_t;
typedef int32_t  s2_t;
typedef size_t   uz_t;
typedef size_t   sz_t;
typedef bool     bl_t;
typedef uint64_t tp_t;
typedef uint64_t aware_t;
typedef void*       vd_t;
typedef const void* vc_t;
typedef const char* sc_t;
typedef char*       sd_t;
typedef struct sr_s { vd_t o; vc_t p; tp_t f; } sr_s;
typedef struct st_s { aware_t _; sd_t data; uz_t size, space; } st_s;
typedef struct bcore_signal_s { int unused; } bcore_signal_s;
typedef struct bcore_source_s bcore_source_s;
typedef struct bcore_sink_s bcore_sink_s;
typedef struct bcore_hmap_tpto_s bcore_hmap_tpto_s;
typedef struct bcore_hmap_tp_sr_s bcore_hmap_tp_sr_s;
typedef struct bcore_arr_tp_s bcore_arr_tp_s;
typedef struct bcore_arr_sr_s bcore_arr_sr_s;
typedef struct bcore_arr_st_s bcore_arr_st_s;
typedef struct bcore_mutex_s { int unused; } bcore_mutex_s;
typedef struct bclos_frame_s bclos_frame_s;
typedef struct bclos_signature_s bclos_signature_s;
typedef struct bcore_array_dyn_solid_static_s { void* data; uz_t size, space; } bcore_array_dyn_solid_static_s;
typedef struct bcore_array_dyn_link_static_s { void* data; uz_t size, space; } bcore_array_dyn_link_static_s;
#define BCORE_DECLARE_FUNCTIONS_OBJ( name )
#define BCORE_DEFINE_INLINE_SPECT_GET_TYPED_CACHED( name )
#define BCORE_DEFINE_INLINE_SPECT_GET_AWARE( name )
#define BCORE_DECLARE_OBJECT( name ) typedef struct name name; struct name
#define TYPEOF_init1 1
#define typeof( name ) ( ( tp_t )0 )
static inline f3_t f3_sqr( f3_t v ) { return v * v; }
static inline f3_t f3_abs( f3_t v ) { return v < 0 ? -v : v; }
static inline f3_t f3_max( f3_t a, f3_t b ) { return a > b ? a : b; }
static inline f3_t f3_min( f3_t a, f3_t b ) { return a < b ? a : b; }
static inline u3_t bcore_lcg00_u3( u3_t v ) { return v * 6364136223846793005ull + 1442695040888963407ull; }
static inline u3_t bcore_lcg01_u3( u3_t v ) { return v * 3935559000370003845ull + 2691343689449507681ull; }
static inline u3_t bcore_lcg02_u3( u3_t v ) { return v * 2862933555777941757ull + 3037000493ull; }
void  bcore_err_fa( sc_t format, ... );
vd_t  bcore_u_alloc( uz_t unit, vd_t p, uz_t n, uz_t* granted );
void  bcore_free( vd_t p );
#endif
