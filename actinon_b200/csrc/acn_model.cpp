// acn_model.cpp — host-side scene model, scene-edit semantics and the flattener.
#include "acn_model.h"

#include <math.h>
#include <string.h>

namespace acnh {

using acn::dot; using acn::unit; using acn::mlv; using acn::mul; using acn::sqr;

M3d mat_rot_x( double a ) { double s = sin( a ), c = cos( a ); M3d m; m.x = vec3( 1, 0, 0 ); m.y = vec3( 0, c, -s ); m.z = vec3( 0, s, c ); return m; }
M3d mat_rot_y( double a ) { double s = sin( a ), c = cos( a ); M3d m; m.x = vec3( c, 0, s ); m.y = vec3( 0, 1, 0 ); m.z = vec3( -s, 0, c ); return m; }
M3d mat_rot_z( double a ) { double s = sin( a ), c = cos( a ); M3d m; m.x = vec3( c, -s, 0 ); m.y = vec3( s, c, 0 ); m.z = vec3( 0, 0, 1 ); return m; }

// smallest sphere around two spheres (objects.c:113-136)
Envelope envelope_of_pair( const Envelope& e1, const Envelope& e2 )
{
    const double r1 = e1.radius, r2 = e2.radius;
    const V3d diff = e1.pos - e2.pos;
    const double d = sqrt( sqr( diff ) );
    const double rmax = r1 > r2 ? r1 : r2, rmin = r1 < r2 ? r1 : r2;
    if( rmin + d <= rmax ) return r1 > r2 ? e1 : e2;
    auto of_length = []( V3d o, double a ) { double r2_ = sqr( o ); if( fabs( r2_ - 1.0 ) < 1E-8 ) return o; double f = r2_ > 0 ? a / sqrt( r2_ ) : 0; return o * f; };
    const V3d p1 = e1.pos + of_length( diff, r1 );
    const V3d p2 = e2.pos - of_length( diff, r2 );
    Envelope e;
    e.pos = ( p1 + p2 ) * 0.5;
    e.radius = ( r1 + r2 + d ) * 0.5;
    return e;
}

// ---------------------------------------------------------------------------------------------
void Props::move( const V3d& v ) { pos = pos + v; if( has_envelope ) envelope.pos = envelope.pos + v; }
void Props::rotate( const M3d& m )
{
    rax = mat_mlm( m, rax );
    pos = mlv( m, pos );
    if( has_envelope ) envelope.pos = mlv( m, envelope.pos );
}
void Props::scale( double f )
{
    pos = pos * f;
    if( has_envelope ) { envelope.pos = envelope.pos * f; envelope.radius *= f; }
}

static bool is_pair( int k ) { return k == ACN_KIND_PAIR_INSIDE || k == ACN_KIND_PAIR_OUTSIDE; }
static bool is_dist( int k ) { return k == ACN_KIND_DIST_SPHERE || k == ACN_KIND_DIST_TORUS; }

std::unique_ptr<Obj> Obj::clone() const
{
    std::unique_ptr<Obj> o( new Obj() );
    o->kind = kind; o->prp = prp;
    memcpy( o->tail, tail, sizeof( tail ) );
    if( o1 ) o->o1 = o1->clone();
    if( o2 ) o->o2 = o2->clone();
    return o;
}

void Obj::move( const V3d& v )
{
    prp.move( v );
    if( is_pair( kind ) || kind == ACN_KIND_NEG ) { if( o1 ) o1->move( v ); if( o2 ) o2->move( v ); }   // objects.c:1101-1106,1346
}

void Obj::rotate( const M3d& m )
{
    prp.rotate( m );
    if( is_pair( kind ) || kind == ACN_KIND_NEG ) { if( o1 ) o1->rotate( m ); if( o2 ) o2->rotate( m ); }
}

void Obj::scale( double f )
{
    prp.scale( f );
    switch( kind )
    {
        case ACN_KIND_SPHERE:   tail[ 0 ] *= f; break;                                   // objects.c:661
        case ACN_KIND_SQUAROID: tail[ 3 ] *= f * f; break;                               // objects.c:831
        case ACN_KIND_DIST_SPHERE: case ACN_KIND_DIST_TORUS: tail[ 0 ] *= 1.0 / f; break; // objects.c:970
        case ACN_KIND_PAIR_INSIDE: case ACN_KIND_PAIR_OUTSIDE: case ACN_KIND_NEG:
            if( o1 ) o1->scale( f ); if( o2 ) o2->scale( f ); break;
        case ACN_KIND_SCALE:                                                             // objects.c:1455-1459
        {
            double g = f != 0 ? 1.0 / f : 1.0;
            tail[ 0 ] *= g; tail[ 1 ] *= g; tail[ 2 ] *= g;
        }
        break;
        default: break;
    }
}

void Obj::set_refractive_index( double n )
{
    prp.refractive_index = n;
    prp.fresnel_reflectivity = ( n == 1.0 ) ? 0.0 : 1.0;
}

bool Obj::set_material( const std::string& name )
{
    struct M { const char* name; double n; double t[ 3 ]; double fr, ch, df; double sigma; bool has_color; double col[ 3 ]; };
    static const M tbl[] =
    {
        { "transparent",      1.0,  { 1, 1, 1 },          1, 0, 0, -1,   false, { 0, 0, 0 } },
        { "glass",            1.46, { 0.8, 0.9, 0.9 },    1, 0, 0, -1,   false, { 0, 0, 0 } },
        { "water",            1.32, { 0.5, 0.9, 0.99 },   1, 0, 0, -1,   false, { 0, 0, 0 } },
        { "sapphire",         1.76, { 0.7, 0.7, 0.7 },    1, 0, 0, -1,   false, { 0, 0, 0 } },
        { "diamond",          2.42, { 0.8, 0.8, 0.8 },    1, 0, 0, -1,   false, { 0, 0, 0 } },
        { "diffuse",          1.0,  { 0, 0, 0 },          0, 0, 1, 0.29, false, { 0, 0, 0 } },
        { "diffuse_polished", 1.5,  { 0, 0, 0 },          1, 0, 1, 0.29, false, { 0, 0, 0 } },
        { "perfect_mirror",   1.0,  { 0, 0, 0 },          0, 1, 0, -1,   true,  { 1, 1, 1 } },
        { "mirror",           1.0,  { 0, 0, 0 },          0, 1, 0, -1,   true,  { 0.92, 0.94, 0.87 } },
        { "gold",             1.0,  { 0, 0, 0 },          0, 1, 0, -1,   true,  { 0.83, 0.69, 0.22 } },
        { "silver",           1.0,  { 0, 0, 0 },          0, 1, 0, -1,   true,  { 0.8, 0.8, 0.8 } },
    };
    for( const M& m : tbl )
    {
        if( name == m.name )
        {
            prp.refractive_index = m.n;
            prp.transparency = vec3( m.t[ 0 ], m.t[ 1 ], m.t[ 2 ] );
            prp.fresnel_reflectivity = m.fr; prp.chromatic_reflectivity = m.ch; prp.diffuse_reflectivity = m.df;
            if( m.sigma >= 0 ) prp.sigma = m.sigma;
            if( m.has_color ) prp.color = vec3( m.col[ 0 ], m.col[ 1 ], m.col[ 2 ] );
            return true;
        }
    }
    return false;
}

bool analytic_envelope( const Obj& o, Envelope* out )
{
    if( o.prp.has_envelope ) { *out = o.prp.envelope; return true; }
    switch( o.kind )
    {
        case ACN_KIND_SPHERE: out->pos = o.prp.pos; out->radius = o.tail[ 0 ]; return true;
        case ACN_KIND_SQUAROID:
        {   // a x^2 + b y^2 + c z^2 + r <= 0 with a, b, c > 0 > r: an ellipsoid with half axes sqrt( -r / a ), ...
            const double a = o.tail[ 0 ], b = o.tail[ 1 ], c = o.tail[ 2 ], r = o.tail[ 3 ];
            if( !( a > 0 && b > 0 && c > 0 && r < 0 ) ) return false;
            const double m = a < b ? ( a < c ? a : c ) : ( b < c ? b : c );
            out->pos = o.prp.pos; out->radius = sqrt( -r / m );
            return true;
        }
        case ACN_KIND_DIST_SPHERE: out->pos = o.prp.pos; out->radius = 1.0 / o.tail[ 0 ]; return true;                     // unit sphere scaled by 1 / inv_scale
        case ACN_KIND_DIST_TORUS:  out->pos = o.prp.pos; out->radius = ( 1.0 + o.tail[ 1 ] ) / o.tail[ 0 ]; return true;   // major radius 1, tube radius ex_radius
        case ACN_KIND_PAIR_INSIDE:
        {
            Envelope e1, e2;
            const bool b1 = o.o1 && analytic_envelope( *o.o1, &e1 ), b2 = o.o2 && analytic_envelope( *o.o2, &e2 );
            if( b1 && b2 ) { *out = e1.radius <= e2.radius ? e1 : e2; return true; }
            if( b1 ) { *out = e1; return true; }
            if( b2 ) { *out = e2; return true; }
            return false;
        }
        case ACN_KIND_PAIR_OUTSIDE:
        {
            Envelope e1, e2;
            if( !( o.o1 && o.o2 && analytic_envelope( *o.o1, &e1 ) && analytic_envelope( *o.o2, &e2 ) ) ) return false;
            *out = envelope_of_pair( e1, e2 );
            return true;
        }
        default: return false;
    }
}

bool Obj::set_bounding_envelope()
{
    Envelope e;
    if( !analytic_envelope( *this, &e ) ) return false;
    e.radius *= 1.0 + 1E-9;          // the shell of thickness eps around the surface stays inside
    e.radius += 4E-6;
    prp.has_envelope = true; prp.envelope = e;
    return true;
}

void Obj::set_auto_envelope()
{
    Envelope e = estimate_envelope( *this, 1000, 123, 1.1 );
    prp.has_envelope = true; prp.envelope = e;
}

// ---------------------------------------------------------------------------------------------
static std::unique_ptr<Obj> new_obj( int kind ) { std::unique_ptr<Obj> o( new Obj() ); o->kind = kind; return o; }

std::unique_ptr<Obj> make_plane() { return new_obj( ACN_KIND_PLANE ); }
std::unique_ptr<Obj> make_sphere( double radius ) { auto o = new_obj( ACN_KIND_SPHERE ); o->tail[ 0 ] = radius; return o; }
std::unique_ptr<Obj> make_squaroid( double a, double b, double c, double r )
{
    auto o = new_obj( ACN_KIND_SQUAROID ); o->tail[ 0 ] = a; o->tail[ 1 ] = b; o->tail[ 2 ] = c; o->tail[ 3 ] = r; return o;
}
static double inv_sqr( double r ) { return r != 0 ? 1.0 / ( r * r ) : 1.0; }
std::unique_ptr<Obj> make_ellipsoid( double rx, double ry, double rz )    { return make_squaroid( inv_sqr( rx ), inv_sqr( ry ),  inv_sqr( rz ), -1 ); }   // objects.c:723-736
std::unique_ptr<Obj> make_hyperboloid1( double rx, double ry, double rz ) { return make_squaroid( inv_sqr( rx ), inv_sqr( ry ), -inv_sqr( rz ), -1 ); }   // :738-746
std::unique_ptr<Obj> make_hyperboloid2( double rx, double ry, double rz ) { return make_squaroid( inv_sqr( rx ), inv_sqr( ry ), -inv_sqr( rz ),  1 ); }   // :748-756
std::unique_ptr<Obj> make_cone( double rx, double ry, double rz )         { return make_squaroid( inv_sqr( rx ), inv_sqr( ry ), -inv_sqr( rz ),  0 ); }   // :758-766
std::unique_ptr<Obj> make_cylinder( double rx, double ry )                { return make_squaroid( inv_sqr( rx ), inv_sqr( ry ), 0, -1 ); }                // :768-776

std::unique_ptr<Obj> make_distance_sphere()
{
    auto o = new_obj( ACN_KIND_DIST_SPHERE ); o->tail[ 0 ] = 1.0; o->tail[ 1 ] = 0; o->tail[ 2 ] = 200; return o;
}

std::unique_ptr<Obj> make_torus( double r1, double r2 )
{
    auto o = new_obj( ACN_KIND_DIST_TORUS );
    o->tail[ 0 ] = 1.0; o->tail[ 1 ] = r2 / r1; o->tail[ 2 ] = 200;
    o->scale( r1 );
    o->prp.has_envelope = true;
    o->prp.envelope.pos = vec3( 0, 0, 0 );
    o->prp.envelope.radius = ( r1 + r2 ) * 1.01;
    return o;
}

std::unique_ptr<Obj> make_pair_inside( const Obj& a, const Obj& b )
{
    auto o = new_obj( ACN_KIND_PAIR_INSIDE );
    o->prp = a.prp; o->o1 = a.clone(); o->o2 = b.clone();
    return o;
}

std::unique_ptr<Obj> make_pair_outside( const Obj& a, const Obj& b )
{
    auto o = new_obj( ACN_KIND_PAIR_OUTSIDE );
    o->prp = a.prp; o->o1 = a.clone(); o->o2 = b.clone();
    o->prp.has_envelope = false;       // o2 is outside o1: the true envelope would be bigger (objects.c:1169-1173)
    return o;
}

std::unique_ptr<Obj> make_neg( const Obj& a )
{
    auto o = new_obj( ACN_KIND_NEG );
    o->prp = a.prp; o->o1 = a.clone();
    return o;
}

std::unique_ptr<Obj> make_scale( const Obj& a, const V3d& s )
{
    auto o = new_obj( ACN_KIND_SCALE );
    o->prp = a.prp;
    o->prp.pos = vec3( 0, 0, 0 ); o->prp.rax = mat_ident();
    if( o->prp.has_envelope )
    {
        o->prp.envelope.pos = mul( o->prp.envelope.pos, s );
        double m = s.x > s.y ? s.x : s.y; m = m > s.z ? m : s.z;
        o->prp.envelope.radius *= m;
    }
    o->o1 = a.clone();
    o->tail[ 0 ] = s.x != 0 ? 1.0 / s.x : 1.0;
    o->tail[ 1 ] = s.y != 0 ? 1.0 / s.y : 1.0;
    o->tail[ 2 ] = s.z != 0 ? 1.0 / s.z : 1.0;
    return o;
}

// ---------------------------------------------------------------------------------------------
// flattening helpers
// ---------------------------------------------------------------------------------------------
struct Flattener
{
    std::vector<acn_flat_node>& nodes;
    std::vector<int32_t>& children;
    std::vector<acn_flat_material>& materials;
    std::map<std::string, int> mat_index;

    Flattener( std::vector<acn_flat_node>& n, std::vector<int32_t>& c, std::vector<acn_flat_material>& m ) : nodes( n ), children( c ), materials( m ) {}

    int material( const Props& p )
    {
        acn_flat_material m;
        memset( &m, 0, sizeof( m ) );
        m.color[ 0 ] = p.color.x; m.color[ 1 ] = p.color.y; m.color[ 2 ] = p.color.z;
        m.radiance = p.radiance; m.refractive_index = p.refractive_index;
        m.fresnel_reflectivity = p.fresnel_reflectivity; m.chromatic_reflectivity = p.chromatic_reflectivity;
        m.diffuse_reflectivity = p.diffuse_reflectivity; m.sigma = p.sigma;
        m.transparency[ 0 ] = p.transparency.x; m.transparency[ 1 ] = p.transparency.y; m.transparency[ 2 ] = p.transparency.z;
        m.texture_kind = p.tex.kind;
        m.tex_color1[ 0 ] = p.tex.c1.x; m.tex_color1[ 1 ] = p.tex.c1.y; m.tex_color1[ 2 ] = p.tex.c1.z;
        m.tex_color2[ 0 ] = p.tex.c2.x; m.tex_color2[ 1 ] = p.tex.c2.y; m.tex_color2[ 2 ] = p.tex.c2.z;
        m.tex_scale = p.tex.scale;
        std::string key( reinterpret_cast<const char*>( &m ), sizeof( m ) );
        auto it = mat_index.find( key );
        if( it != mat_index.end() ) return it->second;
        int idx = ( int )materials.size();
        materials.push_back( m );
        mat_index[ key ] = idx;
        return idx;
    }

    int add_obj( const Obj& o )
    {
        int idx = ( int )nodes.size();
        nodes.emplace_back();
        acn_flat_node nd;
        memset( &nd, 0, sizeof( nd ) );
        nd.kind = o.kind; nd.child0 = -1; nd.child1 = -1;
        nd.material = material( o.prp );
        nd.has_envelope = o.prp.has_envelope ? 1 : 0;
        nd.env_pos[ 0 ] = o.prp.envelope.pos.x; nd.env_pos[ 1 ] = o.prp.envelope.pos.y; nd.env_pos[ 2 ] = o.prp.envelope.pos.z;
        nd.env_radius = o.prp.envelope.radius;
        nd.pos[ 0 ] = o.prp.pos.x; nd.pos[ 1 ] = o.prp.pos.y; nd.pos[ 2 ] = o.prp.pos.z;
        const V3d* rows[ 3 ] = { &o.prp.rax.x, &o.prp.rax.y, &o.prp.rax.z };
        for( int r = 0; r < 3; r++ ) { nd.rax[ 3 * r ] = rows[ r ]->x; nd.rax[ 3 * r + 1 ] = rows[ r ]->y; nd.rax[ 3 * r + 2 ] = rows[ r ]->z; }
        nd.surface_roughness = o.prp.surface_roughness;
        memcpy( nd.tail, o.tail, sizeof( nd.tail ) );
        if( o.o1 ) nd.child0 = add_obj( *o.o1 );
        if( o.o2 ) nd.child1 = add_obj( *o.o2 );
        nodes[ idx ] = nd;
        return idx;
    }

    int add_compound( const Compound& c )
    {
        int idx = ( int )nodes.size();
        nodes.emplace_back();
        acn_flat_node nd;
        memset( &nd, 0, sizeof( nd ) );
        nd.kind = ACN_KIND_COMPOUND; nd.material = -1;
        nd.has_envelope = c.has_envelope ? 1 : 0;
        nd.env_pos[ 0 ] = c.envelope.pos.x; nd.env_pos[ 1 ] = c.envelope.pos.y; nd.env_pos[ 2 ] = c.envelope.pos.z;
        nd.env_radius = c.envelope.radius;
        nd.rax[ 0 ] = nd.rax[ 4 ] = nd.rax[ 8 ] = 1.0;
        nd.child0 = ( int )children.size();
        nd.child1 = ( int )c.items.size();
        children.resize( children.size() + c.items.size(), -1 );
        for( size_t i = 0; i < c.items.size(); i++ )
        {
            int ci = c.items[ i ].obj ? add_obj( *c.items[ i ].obj ) : add_compound( *c.items[ i ].cmp );
            children[ nd.child0 + i ] = ci;
        }
        nodes[ idx ] = nd;
        return idx;
    }
};

// packed double view of a node list for the host instantiation of acn_geom.h
struct HostView
{
    std::vector<acn::R4<double>> env, geo;
    std::vector<acn::I4> link;
    acn::SceneView<double> sv;

    void build( const std::vector<acn_flat_node>& nodes, const std::vector<int32_t>& children, double eps )
    {
        const size_t n = nodes.size();
        env.resize( n ); geo.resize( n * acn::GEO_STRIDE ); link.resize( n );
        for( size_t i = 0; i < n; i++ )
        {
            const acn_flat_node& nd = nodes[ i ];
            int flags = 0;
            if( nd.has_envelope ) flags |= acn::F_ENV;
            if( nd.surface_roughness > 0 && nd.kind != ACN_KIND_COMPOUND ) flags |= acn::F_ROUGH;
            env[ i ].x = nd.env_pos[ 0 ]; env[ i ].y = nd.env_pos[ 1 ]; env[ i ].z = nd.env_pos[ 2 ]; env[ i ].w = nd.has_envelope ? nd.env_radius : -1;
            link[ i ].x = nd.kind | ( flags << 8 ); link[ i ].y = nd.child0; link[ i ].z = nd.child1; link[ i ].w = nd.material;
            acn::R4<double>* g = &geo[ i * acn::GEO_STRIDE ];
            g[ 0 ].x = nd.pos[ 0 ]; g[ 0 ].y = nd.pos[ 1 ]; g[ 0 ].z = nd.pos[ 2 ]; g[ 0 ].w = nd.tail[ 0 ];
            for( int r = 0; r < 3; r++ ) { g[ 1 + r ].x = nd.rax[ 3 * r ]; g[ 1 + r ].y = nd.rax[ 3 * r + 1 ]; g[ 1 + r ].z = nd.rax[ 3 * r + 2 ]; g[ 1 + r ].w = nd.tail[ 1 + r ]; }
            g[ 4 ].x = nd.surface_roughness; g[ 4 ].y = g[ 4 ].z = g[ 4 ].w = 0;
        }
        sv.env = env.data(); sv.geo = geo.data(); sv.link = link.data(); sv.children = children.empty() ? nullptr : children.data();
        sv.prog = nullptr; sv.prog_ref = nullptr; sv.parent = nullptr;
        sv.eps = eps; sv.light_root = 0; sv.matter_root = 0; sv.seed_mode = acn::SEED_POSITION_HASH;
    }
};

// Monte-Carlo bounding sphere used only while BUILDING a scene (objects.c:286-363): random rays
// from a running centre estimate, exit point of each, centre = mean, radius = max distance * factor.
Envelope estimate_envelope( const Obj& o, int samples, uint32_t rseed, double radius_factor )
{
    using namespace acn;
    std::vector<acn_flat_node> nodes; std::vector<int32_t> children; std::vector<acn_flat_material> mats;
    Flattener fl( nodes, children, mats );
    const int root = fl.add_obj( o );
    HostView hv;
    const double eps = 1E-6;
    hv.build( nodes, children, eps );
    const double inf = Num<double>::inf();
    HitCtx ctx; ctx.key = 0;

    auto ray_exit = [ & ]( const Ray<double>& ray ) -> double          // obj_ray_exit, objects.c:286-310
    {
        V3d nor = vec3( 0, 0, 0 );
        double a = obj_ray_hit<double>( hv.sv, root, ray, &nor, ctx );
        if( !( a < inf ) ) return inf;
        Ray<double> rl = ray;
        double sum = 0;
        int guard = 0;
        while( a < inf && guard++ < 100000 )
        {
            a += eps * 2;
            sum += a;
            rl.p = madd( rl.p, rl.d, a );
            a = obj_ray_hit<double>( hv.sv, root, rl, &nor, ctx );
        }
        return dot( nor, ray.d ) > 0 ? sum : inf;
    };

    u64 rv = rseed;
    Ray<double> ray;
    ray.p = o.prp.pos;
    std::vector<V3d> pts;
    V3d sum = vec3( 0, 0, 0 );
    for( int i = 0; i < samples; i++ )
    {
        // v3d_s_random_sphere_belt( &rv, 1.0 ) (vectors.h:209-218)
        double phi = 2.0 * 3.14159265358979323846 * rnd1<double>( &rv );
        double z = rnd0<double>( &rv ) * 1.0;
        double sc = sqrt( 1.0 - z * z );
        ray.d = vec3( sin( phi ) * sc, cos( phi ) * sc, z );
        double a = ray_exit( ray );
        if( a < inf )
        {
            V3d pos = madd( ray.p, ray.d, a );
            pts.push_back( pos );
            sum = sum + pos;
            ray.p = sum * ( 1.0 / ( double )pts.size() );
            ray.p.x += eps * rnd0<double>( &rv );
            ray.p.y += eps * rnd0<double>( &rv );
            ray.p.z += eps * rnd0<double>( &rv );
        }
    }
    Envelope e;
    e.pos = ray.p;
    e.radius = Num<double>::mag();
    if( !pts.empty() )
    {
        double max_r2 = 0;
        for( const V3d& p : pts ) { double r = sqr( ray.p - p ); if( r > max_r2 ) max_r2 = r; }
        e.radius = sqrt( max_r2 ) * radius_factor;
    }
    return e;
}

// ---------------------------------------------------------------------------------------------
// Compound
// ---------------------------------------------------------------------------------------------
std::unique_ptr<Compound> Compound::clone() const
{
    std::unique_ptr<Compound> c( new Compound() );
    c->has_envelope = has_envelope; c->envelope = envelope;
    c->items.reserve( items.size() );
    for( const Elem& e : items )
    {
        Elem n;
        if( e.obj ) n.obj = e.obj->clone();
        if( e.cmp ) n.cmp = e.cmp->clone();
        c->items.push_back( std::move( n ) );
    }
    return c;
}

void Compound::push_obj( const Obj& o )
{
    Elem e; e.obj = o.clone();
    const Props& p = e.obj->prp;
    if( has_envelope )
    {
        if( p.has_envelope ) envelope = envelope_of_pair( envelope, p.envelope );
        else has_envelope = false;
    }
    else if( items.empty() )      // becomes the first element
    {
        has_envelope = p.has_envelope;
        if( p.has_envelope ) envelope = p.envelope;
    }
    items.push_back( std::move( e ) );
}

void Compound::push_compound( const Compound& c )
{
    if( c.has_envelope )
    {
        Elem e; e.cmp = c.clone();
        items.push_back( std::move( e ) );
    }
    else
    {
        for( const Elem& e : c.items )
        {
            if( e.obj ) push_obj( *e.obj );
            else if( e.cmp ) push_compound( *e.cmp );
        }
    }
}

void Compound::move( const V3d& v )
{
    if( has_envelope ) envelope.pos = envelope.pos + v;
    for( Elem& e : items ) { if( e.obj ) e.obj->move( v ); else if( e.cmp ) e.cmp->move( v ); }
}
void Compound::rotate( const M3d& m )
{
    if( has_envelope ) envelope.pos = mlv( m, envelope.pos );
    for( Elem& e : items ) { if( e.obj ) e.obj->rotate( m ); else if( e.cmp ) e.cmp->rotate( m ); }
}
void Compound::scale( double f )
{
    if( has_envelope ) { envelope.pos = envelope.pos * f; envelope.radius *= f; }
    for( Elem& e : items ) { if( e.obj ) e.obj->scale( f ); else if( e.cmp ) e.cmp->scale( f ); }
}

void Compound::set_auto_envelope()
{
    has_envelope = false;
    for( Elem& e : items )
    {
        Envelope env{ vec3( 0, 0, 0 ), 0 };
        if( e.cmp )
        {
            if( !e.cmp->has_envelope ) e.cmp->set_auto_envelope();
            if( e.cmp->has_envelope ) env = e.cmp->envelope;
        }
        else if( e.obj )
        {
            if( !e.obj->prp.has_envelope ) e.obj->set_auto_envelope();
            env = e.obj->prp.envelope;
        }
        if( has_envelope ) envelope = envelope_of_pair( envelope, env );
        else { has_envelope = true; envelope = env; }
    }
}

// ---------------------------------------------------------------------------------------------
// Values
// ---------------------------------------------------------------------------------------------
VP Value::clone() const
{
    VP c( new Value() );
    c->type = type; c->b = b; c->i = i; c->f = f; c->s = s; c->v = v; c->m = m; c->func = func; c->builtin = builtin;
    if( obj ) c->obj = obj->clone();
    if( cmp ) c->cmp = cmp->clone();
    c->list.reserve( list.size() );
    for( const VP& e : list ) c->list.push_back( e ? e->clone() : VP() );
    for( const auto& kv : map ) c->map.push_back( std::make_pair( kv.first, kv.second ? kv.second->clone() : VP() ) );
    return c;
}

VP Value::map_get( const std::string& k ) const
{
    for( const auto& kv : map ) if( kv.first == k ) return kv.second;
    return VP();
}
void Value::map_set( const std::string& k, VP val )
{
    for( auto& kv : map ) if( kv.first == k ) { kv.second = val; return; }
    map.push_back( std::make_pair( k, val ) );
}

VP make_nil() { return VP( new Value() ); }
VP make_bool( bool b ) { VP v( new Value() ); v->type = Value::BOOL; v->b = b; return v; }
VP make_int( long long i ) { VP v( new Value() ); v->type = Value::INT; v->i = i; return v; }
VP make_num( double f ) { VP v( new Value() ); v->type = Value::NUM; v->f = f; return v; }
VP make_str( const std::string& s ) { VP v( new Value() ); v->type = Value::STR; v->s = s; return v; }
VP make_vec( const V3d& x ) { VP v( new Value() ); v->type = Value::VEC; v->v = x; return v; }
VP make_mat( const M3d& m ) { VP v( new Value() ); v->type = Value::MAT; v->m = m; return v; }
VP make_obj( std::unique_ptr<Obj> o ) { VP v( new Value() ); v->type = Value::OBJ; v->obj = std::move( o ); return v; }
VP make_cmp( std::unique_ptr<Compound> c ) { VP v( new Value() ); v->type = Value::CMP; v->cmp = std::move( c ); return v; }
VP make_list() { VP v( new Value() ); v->type = Value::LIST; return v; }
VP make_map() { VP v( new Value() ); v->type = Value::MAP; return v; }

bool value_move( Value& v, const V3d& d )
{
    switch( v.type )
    {
        case Value::OBJ: v.obj->move( d ); return true;
        case Value::CMP: v.cmp->move( d ); return true;
        case Value::LIST: for( VP& e : v.list ) if( e ) value_move( *e, d ); return true;
        case Value::MAP:  for( auto& kv : v.map ) if( kv.second ) value_move( *kv.second, d ); return true;
        case Value::VEC: v.v = v.v + d; return true;
        default: return false;
    }
}
bool value_rotate( Value& v, const M3d& m )
{
    switch( v.type )
    {
        case Value::OBJ: v.obj->rotate( m ); return true;
        case Value::CMP: v.cmp->rotate( m ); return true;
        case Value::LIST: for( VP& e : v.list ) if( e ) value_rotate( *e, m ); return true;
        case Value::MAP:  for( auto& kv : v.map ) if( kv.second ) value_rotate( *kv.second, m ); return true;
        case Value::VEC: v.v = mlv( m, v.v ); return true;
        default: return false;
    }
}
bool value_scale( Value& v, double f )
{
    switch( v.type )
    {
        case Value::OBJ: v.obj->scale( f ); return true;
        case Value::CMP: v.cmp->scale( f ); return true;
        case Value::LIST: for( VP& e : v.list ) if( e ) value_scale( *e, f ); return true;
        case Value::MAP:  for( auto& kv : v.map ) if( kv.second ) value_scale( *kv.second, f ); return true;
        case Value::VEC: v.v = v.v * f; return true;
        default: return false;
    }
}

// balanced composites (container.c:376-410)
static std::unique_ptr<Obj> composite( const std::vector<VP>& l, size_t start, size_t size, bool inside, std::string* err )
{
    if( size > l.size() ) size = l.size();
    if( size == 0 || start >= l.size() ) { if( err ) *err = "composite of an empty list"; return nullptr; }
    if( size == 1 )
    {
        if( !l[ start ] || l[ start ]->type != Value::OBJ ) { if( err ) *err = "composite: list element is not an object"; return nullptr; }
        return l[ start ]->obj->clone();
    }
    auto a = composite( l, start, size >> 1, inside, err );
    auto b = composite( l, start + ( size >> 1 ), size - ( size >> 1 ), inside, err );
    if( !a || !b ) return nullptr;
    return inside ? make_pair_inside( *a, *b ) : make_pair_outside( *a, *b );
}
std::unique_ptr<Obj> list_inside_composite( const std::vector<VP>& l, size_t start, size_t size, std::string* err )  { return composite( l, start, size, true, err ); }
std::unique_ptr<Obj> list_outside_composite( const std::vector<VP>& l, size_t start, size_t size, std::string* err ) { return composite( l, start, size, false, err ); }

bool compound_push_value( Compound& c, const Value& v, std::string* err )
{
    switch( v.type )
    {
        case Value::OBJ: c.push_obj( *v.obj ); return true;
        case Value::CMP: c.push_compound( *v.cmp ); return true;
        case Value::LIST: for( const VP& e : v.list ) if( e && !compound_push_value( c, *e, err ) ) return false; return true;
        case Value::MAP:  for( const auto& kv : v.map ) if( kv.second && !compound_push_value( c, *kv.second, err ) ) return false; return true;
        default: if( err ) *err = "cannot push this value to a compound"; return false;
    }
}

// ---------------------------------------------------------------------------------------------
// Scene
// ---------------------------------------------------------------------------------------------
void default_params( acn_flat_params* p )
{
    memset( p, 0, sizeof( *p ) );
    p->threads = 10; p->image_width = 800; p->image_height = 600; p->gamma = 1.0;          // scene.c:185-213
    p->gradient_threshold = 0.1; p->gradient_samples = 10; p->gradient_cycles = 1;
    p->camera_focal_length = 1.0;
    p->trace_depth = 11; p->trace_min_intensity = 0; p->direct_samples = 100; p->path_samples = 0;
    p->max_path_length = 1E+30;
}

Scene::Scene() { default_params( &params ); memset( &flat, 0, sizeof( flat ) ); }

bool Scene::push( const Value& v, std::string* err )
{
    switch( v.type )
    {
        case Value::OBJ:
            if( v.obj->prp.radiance > 0 ) light.push_obj( *v.obj ); else matter.push_obj( *v.obj );
            return true;
        case Value::CMP: matter.push_compound( *v.cmp ); return true;
        case Value::LIST: for( const VP& e : v.list ) if( e && !push( *e, err ) ) return false; return true;
        case Value::MAP:  for( const auto& kv : v.map ) if( kv.second && !push( *kv.second, err ) ) return false; return true;
        default: return true;    // scene_s_push ignores other types (scene.c:278)
    }
}

void Scene::flatten()
{
    f_nodes.clear(); f_children.clear(); f_materials.clear();
    Flattener fl( f_nodes, f_children, f_materials );
    // the scene-level compounds never carry an envelope of their own that matters for culling: the
    // reference keeps whatever compound_s_push_q left there (compound.c:149-164) — reproduce it
    int lr = fl.add_compound( light );
    int mr = fl.add_compound( matter );
    memset( &flat, 0, sizeof( flat ) );
    flat.params = params;
    flat.n_nodes = ( int )f_nodes.size(); flat.n_children = ( int )f_children.size(); flat.n_materials = ( int )f_materials.size();
    flat.light_root = lr; flat.matter_root = mr;
    flat.nodes = f_nodes.data(); flat.children = f_children.data(); flat.materials = f_materials.data();
}

} // namespace acnh
