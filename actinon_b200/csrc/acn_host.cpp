// acn_host.cpp — host half of the C ABI: scene-description API, flattening, pass controller,
// .pnm output and resume.  Pure host code (works without a GPU); it never renders.
#include "acn_model.h"

#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>

namespace acn { void set_error( const char* fmt, ... ); }

using namespace acnh;

struct acn_scene { Scene sc; std::string last_name; };

namespace {

Value* get( acn_scene* s, acn_obj h )
{
    if( !s || h < 0 || ( size_t )h >= s->sc.handles.size() || !s->sc.handles[ h ] ) { acn::set_error( "bad object handle %d", ( int )h ); return nullptr; }
    return s->sc.handles[ h ].get();
}

Obj* get_obj( acn_scene* s, acn_obj h )
{
    Value* v = get( s, h );
    if( !v ) return nullptr;
    if( v->type != Value::OBJ ) { acn::set_error( "handle %d is not a shape object", ( int )h ); return nullptr; }
    return v->obj.get();
}

acn_obj put( acn_scene* s, std::unique_ptr<Obj> o ) { return s->sc.add_handle( make_obj( std::move( o ) ) ); }

V3d v3p( const double* p ) { return vec3( p[ 0 ], p[ 1 ], p[ 2 ] ); }

} // namespace

extern "C" {

int acn_scene_create( acn_scene** out )
{
    if( !out ) return ACN_ERR_INVALID_ARG;
    *out = new acn_scene();
    return ACN_OK;
}

void acn_scene_destroy( acn_scene* s ) { delete s; }

acn_flat_params* acn_scene_params( acn_scene* s ) { return s ? &s->sc.params : nullptr; }

#define ACN_NEED_SCENE( s ) do { if( !( s ) ) { acn::set_error( "null scene" ); return ACN_ERR_INVALID_ARG; } } while( 0 )

acn_obj acn_create_plane( acn_scene* s ) { ACN_NEED_SCENE( s ); return put( s, make_plane() ); }
acn_obj acn_create_sphere( acn_scene* s, double radius ) { ACN_NEED_SCENE( s ); return put( s, make_sphere( radius ) ); }
acn_obj acn_create_squaroid( acn_scene* s, double a, double b, double c, double r ) { ACN_NEED_SCENE( s ); return put( s, make_squaroid( a, b, c, r ) ); }
acn_obj acn_create_ellipsoid( acn_scene* s, double rx, double ry, double rz ) { ACN_NEED_SCENE( s ); return put( s, make_ellipsoid( rx, ry, rz ) ); }
acn_obj acn_create_cylinder( acn_scene* s, double rx, double ry ) { ACN_NEED_SCENE( s ); return put( s, make_cylinder( rx, ry ) ); }
acn_obj acn_create_cone( acn_scene* s, double rx, double ry, double rz ) { ACN_NEED_SCENE( s ); return put( s, make_cone( rx, ry, rz ) ); }
acn_obj acn_create_hyperboloid1( acn_scene* s, double rx, double ry, double rz ) { ACN_NEED_SCENE( s ); return put( s, make_hyperboloid1( rx, ry, rz ) ); }
acn_obj acn_create_hyperboloid2( acn_scene* s, double rx, double ry, double rz ) { ACN_NEED_SCENE( s ); return put( s, make_hyperboloid2( rx, ry, rz ) ); }
acn_obj acn_create_torus( acn_scene* s, double r1, double r2 )
{
    ACN_NEED_SCENE( s );
    if( r1 == 0 ) { acn::set_error( "create_torus: radius1 must not be 0" ); return ACN_ERR_INVALID_ARG; }
    return put( s, make_torus( r1, r2 ) );
}
acn_obj acn_create_distance_sphere( acn_scene* s ) { ACN_NEED_SCENE( s ); return put( s, make_distance_sphere() ); }

acn_obj acn_clone( acn_scene* s, acn_obj o )
{
    Value* v = get( s, o ); if( !v ) return ACN_ERR_INVALID_ARG;
    return s->sc.add_handle( v->clone() );
}

acn_obj acn_pair_inside( acn_scene* s, acn_obj o1, acn_obj o2 )
{
    Obj* a = get_obj( s, o1 ); Obj* b = get_obj( s, o2 ); if( !a || !b ) return ACN_ERR_INVALID_ARG;
    return put( s, make_pair_inside( *a, *b ) );
}
acn_obj acn_pair_outside( acn_scene* s, acn_obj o1, acn_obj o2 )
{
    Obj* a = get_obj( s, o1 ); Obj* b = get_obj( s, o2 ); if( !a || !b ) return ACN_ERR_INVALID_ARG;
    return put( s, make_pair_outside( *a, *b ) );
}
acn_obj acn_neg( acn_scene* s, acn_obj o1 )
{
    Obj* a = get_obj( s, o1 ); if( !a ) return ACN_ERR_INVALID_ARG;
    return put( s, make_neg( *a ) );
}
acn_obj acn_scale_object( acn_scene* s, acn_obj o1, const double scale[ 3 ] )
{
    Obj* a = get_obj( s, o1 ); if( !a || !scale ) return ACN_ERR_INVALID_ARG;
    return put( s, make_scale( *a, v3p( scale ) ) );
}

acn_obj acn_list_create( acn_scene* s ) { ACN_NEED_SCENE( s ); return s->sc.add_handle( make_list() ); }

int acn_list_push( acn_scene* s, acn_obj list, acn_obj item )
{
    Value* l = get( s, list ); Value* it = get( s, item );
    if( !l || !it ) return ACN_ERR_INVALID_ARG;
    if( l->type != Value::LIST ) { acn::set_error( "handle %d is not a list", ( int )list ); return ACN_ERR_INVALID_ARG; }
    l->list.push_back( it->clone() );
    return ACN_OK;
}

acn_obj acn_list_inside_composite( acn_scene* s, acn_obj list )
{
    Value* l = get( s, list ); if( !l || l->type != Value::LIST ) return ACN_ERR_INVALID_ARG;
    std::string err;
    auto o = list_inside_composite( l->list, 0, l->list.size(), &err );
    if( !o ) { acn::set_error( "%s", err.c_str() ); return ACN_ERR_INVALID_ARG; }
    return put( s, std::move( o ) );
}
acn_obj acn_list_outside_composite( acn_scene* s, acn_obj list )
{
    Value* l = get( s, list ); if( !l || l->type != Value::LIST ) return ACN_ERR_INVALID_ARG;
    std::string err;
    auto o = list_outside_composite( l->list, 0, l->list.size(), &err );
    if( !o ) { acn::set_error( "%s", err.c_str() ); return ACN_ERR_INVALID_ARG; }
    return put( s, std::move( o ) );
}
acn_obj acn_list_create_compound( acn_scene* s, acn_obj list )
{
    Value* l = get( s, list ); if( !l || l->type != Value::LIST ) return ACN_ERR_INVALID_ARG;
    std::unique_ptr<Compound> c( new Compound() );
    std::string err;
    for( const VP& e : l->list ) if( e && !compound_push_value( *c, *e, &err ) ) { acn::set_error( "%s", err.c_str() ); return ACN_ERR_INVALID_ARG; }
    return s->sc.add_handle( make_cmp( std::move( c ) ) );
}

int acn_move( acn_scene* s, acn_obj o, const double v[ 3 ] )
{
    Value* x = get( s, o ); if( !x || !v ) return ACN_ERR_INVALID_ARG;
    return value_move( *x, v3p( v ) ) ? ACN_OK : ACN_ERR_INVALID_ARG;
}
int acn_rotate( acn_scene* s, acn_obj o, const double m[ 9 ] )
{
    Value* x = get( s, o ); if( !x || !m ) return ACN_ERR_INVALID_ARG;
    M3d r; r.x = v3p( m ); r.y = v3p( m + 3 ); r.z = v3p( m + 6 );
    return value_rotate( *x, r ) ? ACN_OK : ACN_ERR_INVALID_ARG;
}
int acn_scale( acn_scene* s, acn_obj o, double f )
{
    Value* x = get( s, o ); if( !x ) return ACN_ERR_INVALID_ARG;
    return value_scale( *x, f ) ? ACN_OK : ACN_ERR_INVALID_ARG;
}

#define ACN_OBJ_SETTER( NAME, EXPR ) \
    int NAME( acn_scene* s, acn_obj o, double v ) { Obj* x = get_obj( s, o ); if( !x ) return ACN_ERR_INVALID_ARG; EXPR; return ACN_OK; }

int acn_set_color( acn_scene* s, acn_obj o, const double rgb[ 3 ] ) { Obj* x = get_obj( s, o ); if( !x || !rgb ) return ACN_ERR_INVALID_ARG; x->prp.color = v3p( rgb ); return ACN_OK; }
int acn_set_transparency( acn_scene* s, acn_obj o, const double rgb[ 3 ] ) { Obj* x = get_obj( s, o ); if( !x || !rgb ) return ACN_ERR_INVALID_ARG; x->prp.transparency = v3p( rgb ); return ACN_OK; }
ACN_OBJ_SETTER( acn_set_refractive_index, x->set_refractive_index( v ) )
ACN_OBJ_SETTER( acn_set_radiance, x->prp.radiance = v )
ACN_OBJ_SETTER( acn_set_fresnel_reflectivity, x->prp.fresnel_reflectivity = v )
ACN_OBJ_SETTER( acn_set_chromatic_reflectivity, x->prp.chromatic_reflectivity = v )
ACN_OBJ_SETTER( acn_set_diffuse_reflectivity, x->prp.diffuse_reflectivity = v )
ACN_OBJ_SETTER( acn_set_sigma, x->prp.sigma = v )
ACN_OBJ_SETTER( acn_set_surface_roughness, x->prp.surface_roughness = v )

int acn_set_material( acn_scene* s, acn_obj o, const char* name )
{
    Obj* x = get_obj( s, o ); if( !x || !name ) return ACN_ERR_INVALID_ARG;
    if( !x->set_material( name ) ) { acn::set_error( "set_surface: Unknown material specification '%s'.", name ); return ACN_ERR_INVALID_ARG; }
    return ACN_OK;
}

int acn_set_envelope( acn_scene* s, acn_obj o, const double pos[ 3 ], double radius )
{
    Value* x = get( s, o ); if( !x || !pos ) return ACN_ERR_INVALID_ARG;
    Envelope e{ v3p( pos ), radius };
    if( x->type == Value::OBJ ) { x->obj->prp.has_envelope = true; x->obj->prp.envelope = e; return ACN_OK; }
    if( x->type == Value::CMP ) { x->cmp->has_envelope = true; x->cmp->envelope = e; return ACN_OK; }
    acn::set_error( "set_envelope: not an object or compound" );
    return ACN_ERR_INVALID_ARG;
}

int acn_set_auto_envelope( acn_scene* s, acn_obj o )
{
    Value* x = get( s, o ); if( !x ) return ACN_ERR_INVALID_ARG;
    if( x->type == Value::OBJ ) { x->obj->set_auto_envelope(); return ACN_OK; }
    if( x->type == Value::CMP ) { x->cmp->set_auto_envelope(); return ACN_OK; }
    acn::set_error( "set_auto_envelope: not an object or compound" );
    return ACN_ERR_INVALID_ARG;
}

int acn_set_bounding_envelope( acn_scene* s, acn_obj o )
{
    Obj* x = get_obj( s, o ); if( !x ) return ACN_ERR_INVALID_ARG;
    if( !x->set_bounding_envelope() ) { acn::set_error( "set_bounding_envelope: the shape is unbounded (or a scale node): no analytic bound" ); return ACN_ERR_UNSUPPORTED; }
    return ACN_OK;
}

int acn_set_texture_plain( acn_scene* s, acn_obj o, const double rgb[ 3 ] )
{
    Obj* x = get_obj( s, o ); if( !x || !rgb ) return ACN_ERR_INVALID_ARG;
    x->prp.tex.kind = ACN_TEX_PLAIN; x->prp.tex.c1 = v3p( rgb );
    return ACN_OK;
}
int acn_set_texture_chess( acn_scene* s, acn_obj o, const double rgb1[ 3 ], const double rgb2[ 3 ], double scale )
{
    Obj* x = get_obj( s, o ); if( !x || !rgb1 || !rgb2 ) return ACN_ERR_INVALID_ARG;
    if( x->kind != ACN_KIND_PLANE && x->kind != ACN_KIND_SPHERE && x->kind != ACN_KIND_DIST_SPHERE && x->kind != ACN_KIND_DIST_TORUS )
    {
        acn::set_error( "chess texture needs a projection function (plane, sphere, distance object; objects.c:247-252)" );
        return ACN_ERR_UNSUPPORTED;
    }
    x->prp.tex.kind = ACN_TEX_CHESS; x->prp.tex.c1 = v3p( rgb1 ); x->prp.tex.c2 = v3p( rgb2 ); x->prp.tex.scale = scale;
    return ACN_OK;
}

int acn_scene_clear( acn_scene* s ) { ACN_NEED_SCENE( s ); s->sc.clear(); return ACN_OK; }

int acn_scene_push( acn_scene* s, acn_obj o )
{
    Value* x = get( s, o ); if( !x ) return ACN_ERR_INVALID_ARG;
    std::string err;
    if( !s->sc.push( *x, &err ) ) { acn::set_error( "%s", err.c_str() ); return ACN_ERR_INVALID_ARG; }
    return ACN_OK;
}

int acn_scene_load_acn( acn_scene* s, const char* path, int argc, const char* const* argv, int* n_images )
{
    ACN_NEED_SCENE( s );
    if( !path ) return ACN_ERR_INVALID_ARG;
    std::vector<std::string> args;
    for( int i = 0; i < argc; i++ ) args.push_back( argv[ i ] ? argv[ i ] : "" );
    std::string err;
    int rc = interpret_file( s->sc, path, args, &err );
    if( n_images ) *n_images = ( int )s->sc.images.size();
    if( rc ) { acn::set_error( "%s", err.c_str() ); return rc; }
    return ACN_OK;
}

const char* acn_scene_image_name( acn_scene* s, int image_index )
{
    if( !s || image_index < 0 || ( size_t )image_index >= s->sc.images.size() ) return nullptr;
    return s->sc.images[ image_index ].name.c_str();
}

int acn_scene_select_image( acn_scene* s, int image_index )
{
    ACN_NEED_SCENE( s );
    if( image_index < 0 || ( size_t )image_index >= s->sc.images.size() ) { acn::set_error( "image index %d out of range", image_index ); return ACN_ERR_INVALID_ARG; }
    const RecordedImage& im = s->sc.images[ image_index ];
    s->sc.params = im.params;
    s->sc.light = std::move( *im.light->clone() );
    s->sc.matter = std::move( *im.matter->clone() );
    return ACN_OK;
}

int acn_scene_flatten( acn_scene* s, const acn_flat_scene** out )
{
    ACN_NEED_SCENE( s );
    if( !out ) return ACN_ERR_INVALID_ARG;
    s->sc.flatten();
    *out = &s->sc.flat;
    return ACN_OK;
}

// ---------------------------------------------------------------------------------------------
// lum_image_s + pass controller (scene.c:682-885,1032-1165)
// ---------------------------------------------------------------------------------------------
} // extern "C"

struct acn_image
{
    int32_t width = 0, height = 0;
    int32_t cycle = 0;              // next gradient cycle to render
    uint64_t rval = 21943294ull;    // jitter stream state at the start of that cycle (scene.c:799)
    uint64_t rval_next = 21943294ull;
    std::vector<double> arr;        // per pixel: pos.x, pos.y, clr.r, clr.g, clr.b, weight
    std::vector<double> pass_xy;    // sample list of the pass being rendered
};

namespace {

inline void img_avg( const acn_image* im, int x, int y, double c[ 3 ] )     // lum_image_s_get_avg
{
    c[ 0 ] = c[ 1 ] = c[ 2 ] = 0;
    if( x < 0 || x >= im->width || y < 0 || y >= im->height ) return;
    const double* p = &im->arr[ 6 * ( ( size_t )y * im->width + x ) ];
    double f = p[ 5 ] > 0 ? 1.0 / p[ 5 ] : 1.0;
    c[ 0 ] = p[ 2 ] * f; c[ 1 ] = p[ 3 ] * f; c[ 2 ] = p[ 4 ] * f;
}

inline double img_dev( const acn_image* im, const double ref[ 3 ], int x, int y )     // lum_image_s_clr_dev
{
    if( x < 0 || x >= im->width || y < 0 || y >= im->height ) return 0;
    double c[ 3 ]; img_avg( im, x, y, c );
    double d0 = ref[ 0 ] - c[ 0 ], d1 = ref[ 1 ] - c[ 1 ], d2 = ref[ 2 ] - c[ 2 ];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

inline double img_sqr_grad( const acn_image* im, int x, int y )     // lum_image_s_sqr_grad
{
    double v[ 3 ]; img_avg( im, x, y, v );
    double g0 = 0, g1;
    for( int dx = -1; dx <= 1; dx++ )
        for( int dy = -1; dy <= 1; dy++ )
        {
            if( dx == 0 && dy == 0 ) continue;
            g1 = img_dev( im, v, x + dx, y + dy ); g0 = g1 > g0 ? g1 : g0;
        }
    return g0;
}

inline uint8_t pack8( double c ) { return c > 0.0 ? ( c < 1.0 ? ( uint8_t )( c * 256 ) : 255 ) : 0; }     // cps_from_cl, scene.c:76-82

} // namespace

extern "C" {

int acn_image_create( int32_t width, int32_t height, acn_image** out )
{
    if( !out || width <= 0 || height <= 0 ) { acn::set_error( "acn_image_create: bad arguments" ); return ACN_ERR_INVALID_ARG; }
    acn_image* im = new acn_image();
    im->width = width; im->height = height;
    im->arr.assign( ( size_t )width * height * 6, 0.0 );
    *out = im;
    return ACN_OK;
}

void acn_image_destroy( acn_image* im ) { delete im; }
int acn_image_size( const acn_image* im, int32_t* width, int32_t* height )
{
    if( !im ) return ACN_ERR_INVALID_ARG;
    if( width ) *width = im->width;
    if( height ) *height = im->height;
    return ACN_OK;
}
int32_t  acn_image_cycle( const acn_image* im ) { return im ? im->cycle : -1; }
uint64_t acn_image_rval( const acn_image* im ) { return im ? im->rval : 0; }

int acn_image_next_pass( acn_image* im, const acn_flat_params* prm, const double** xy, uint64_t* n )
{
    if( !im || !prm || !xy || !n ) return ACN_ERR_INVALID_ARG;
    im->pass_xy.clear();
    *xy = nullptr; *n = 0;
    if( im->cycle > prm->gradient_cycles ) return ACN_OK;      // all gradient_cycles + 1 passes done (scene.c:1103)
    uint64_t rval = im->rval;
    if( im->cycle == 0 )
    {
        im->pass_xy.reserve( ( size_t )im->width * im->height * 2 );
        for( int j = 0; j < im->height; j++ )
            for( int i = 0; i < im->width; i++ ) { im->pass_xy.push_back( i + 0.5 ); im->pass_xy.push_back( j + 0.5 ); }
    }
    else
    {
        const double thr2 = prm->gradient_threshold * prm->gradient_threshold;
        for( int j = 0; j < im->height; j++ )
            for( int i = 0; i < im->width; i++ )
                if( img_sqr_grad( im, i, j ) > thr2 )
                    for( int k = 0; k < prm->gradient_samples; k++ )
                    {
                        double dx = acn::rnd1<double>( &rval );
                        double dy = acn::rnd1<double>( &rval );
                        im->pass_xy.push_back( i + dx ); im->pass_xy.push_back( j + dy );
                    }
    }
    im->rval_next = rval;
    *xy = im->pass_xy.data();
    *n = im->pass_xy.size() / 2;
    return ACN_OK;
}

int acn_image_push( acn_image* im, const double* xy, const float* rgb, uint64_t n )
{
    if( !im || ( n && ( !xy || !rgb ) ) ) return ACN_ERR_INVALID_ARG;
    for( uint64_t i = 0; i < n; i++ )
    {
        int x = ( int )xy[ 2 * i ], y = ( int )xy[ 2 * i + 1 ];           // weight = 1 (scene.c:806-807)
        if( x >= 0 && x < im->width && y >= 0 && y < im->height )
        {
            double* p = &im->arr[ 6 * ( ( size_t )y * im->width + x ) ];
            p[ 0 ] += xy[ 2 * i ]; p[ 1 ] += xy[ 2 * i + 1 ];
            p[ 2 ] += rgb[ 3 * i ]; p[ 3 ] += rgb[ 3 * i + 1 ]; p[ 4 ] += rgb[ 3 * i + 2 ];
            p[ 5 ] += 1.0;
        }
    }
    im->rval = im->rval_next;
    im->cycle++;
    return ACN_OK;
}

int acn_image_average( const acn_image* im, float* rgb )
{
    if( !im || !rgb ) return ACN_ERR_INVALID_ARG;
    for( int y = 0; y < im->height; y++ )
        for( int x = 0; x < im->width; x++ )
        {
            double c[ 3 ]; img_avg( im, x, y, c );
            float* o = rgb + 3 * ( ( size_t )y * im->width + x );
            o[ 0 ] = ( float )c[ 0 ]; o[ 1 ] = ( float )c[ 1 ]; o[ 2 ] = ( float )c[ 2 ];
        }
    return ACN_OK;
}

int acn_image_sums( const acn_image* im, double* sums )
{
    if( !im || !sums ) return ACN_ERR_INVALID_ARG;
    memcpy( sums, im->arr.data(), im->arr.size() * sizeof( double ) );
    return ACN_OK;
}

int acn_image_add_sums( acn_image* im, const double* sums )
{
    if( !im || !sums ) return ACN_ERR_INVALID_ARG;
    for( size_t i = 0; i < im->arr.size(); i++ ) im->arr[ i ] += sums[ i ];
    return ACN_OK;
}

int acn_image_set_state( acn_image* im, int32_t cycle, uint64_t rval, const double* sums )
{
    if( !im || !sums || cycle < 0 ) return ACN_ERR_INVALID_ARG;
    memcpy( im->arr.data(), sums, im->arr.size() * sizeof( double ) );
    im->cycle = cycle; im->rval = im->rval_next = rval;
    return ACN_OK;
}

int acn_image_write_pnm( const acn_image* im, const char* path, uint64_t* hash )
{
    if( !im ) return ACN_ERR_INVALID_ARG;
    std::vector<uint8_t> px( ( size_t )im->width * im->height * 3 );
    uint64_t h = 0xcbf29ce484222325ull;                                  // FNV-1a 64 over the packed 0x00BBGGRR words
    for( int y = 0; y < im->height; y++ )
        for( int x = 0; x < im->width; x++ )
        {
            double c[ 3 ]; img_avg( im, x, y, c );
            uint8_t* o = &px[ 3 * ( ( size_t )y * im->width + x ) ];
            o[ 0 ] = pack8( c[ 0 ] ); o[ 1 ] = pack8( c[ 1 ] ); o[ 2 ] = pack8( c[ 2 ] );
            const uint8_t w[ 4 ] = { o[ 0 ], o[ 1 ], o[ 2 ], 0 };
            for( int k = 0; k < 4; k++ ) { h ^= w[ k ]; h *= 0x100000001b3ull; }
        }
    if( hash ) *hash = h;
    if( path )
    {
        FILE* f = fopen( path, "wb" );
        if( !f ) { acn::set_error( "cannot open '%s' for writing", path ); return ACN_ERR_IO; }
        fprintf( f, "P6\n%d %d\n255\n", im->width, im->height );
        size_t wr = fwrite( px.data(), 1, px.size(), f );
        fclose( f );
        if( wr != px.size() ) { acn::set_error( "short write on '%s'", path ); return ACN_ERR_IO; }
    }
    return ACN_OK;
}

int acn_image_save( const acn_image* im, const char* path )
{
    if( !im || !path ) return ACN_ERR_INVALID_ARG;
    FILE* f = fopen( path, "wb" );
    if( !f ) { acn::set_error( "cannot open '%s' for writing", path ); return ACN_ERR_IO; }
    const char magic[ 8 ] = { 'A', 'C', 'N', 'L', 'U', 'M', '0', '1' };
    bool ok = fwrite( magic, 1, 8, f ) == 8;
    ok = ok && fwrite( &im->width, sizeof( int32_t ), 1, f ) == 1 && fwrite( &im->height, sizeof( int32_t ), 1, f ) == 1;
    ok = ok && fwrite( &im->cycle, sizeof( int32_t ), 1, f ) == 1 && fwrite( &im->rval, sizeof( uint64_t ), 1, f ) == 1;
    ok = ok && fwrite( im->arr.data(), sizeof( double ), im->arr.size(), f ) == im->arr.size();
    fclose( f );
    if( !ok ) { acn::set_error( "short write on '%s'", path ); return ACN_ERR_IO; }
    return ACN_OK;
}

int acn_image_load( const char* path, acn_image** out )
{
    if( !path || !out ) return ACN_ERR_INVALID_ARG;
    FILE* f = fopen( path, "rb" );
    if( !f ) { acn::set_error( "cannot open '%s'", path ); return ACN_ERR_IO; }
    char magic[ 8 ]; int32_t w = 0, h = 0, cyc = 0; uint64_t rval = 0;
    bool ok = fread( magic, 1, 8, f ) == 8 && memcmp( magic, "ACNLUM01", 8 ) == 0;
    ok = ok && fread( &w, sizeof( int32_t ), 1, f ) == 1 && fread( &h, sizeof( int32_t ), 1, f ) == 1;
    ok = ok && fread( &cyc, sizeof( int32_t ), 1, f ) == 1 && fread( &rval, sizeof( uint64_t ), 1, f ) == 1;
    ok = ok && w > 0 && h > 0 && ( int64_t )w * h < ( 1ll << 31 );
    acn_image* im = nullptr;
    if( ok )
    {
        im = new acn_image();
        im->width = w; im->height = h; im->cycle = cyc; im->rval = im->rval_next = rval;
        im->arr.resize( ( size_t )w * h * 6 );
        ok = fread( im->arr.data(), sizeof( double ), im->arr.size(), f ) == im->arr.size();
    }
    fclose( f );
    if( !ok ) { delete im; acn::set_error( "'%s' is not a valid lum-image file", path ); return ACN_ERR_IO; }
    *out = im;
    return ACN_OK;
}

} // extern "C"
