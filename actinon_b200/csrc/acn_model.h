// acn_model.h — host-side scene model: the part of the reference's object system that the
// flattener has to understand (objects.h:51-97, objects.c:141-197,1463-1716, compound.c:36-207,
// container.c:376-421, scene.c:153-279).  Pure host C++; no rendering here.
#pragma once

#include <memory>
#include <string>
#include <vector>
#include <map>
#include <functional>

#include "acn_geom.h"
#include "../../include/actinon_b200.h"

namespace acnh {

typedef acn::V3<double> V3d;
typedef acn::M3<double> M3d;

inline V3d vec3( double x, double y, double z ) { return acn::v3<double>( x, y, z ); }
inline M3d mat_ident() { M3d m; m.x = vec3( 1, 0, 0 ); m.y = vec3( 0, 1, 0 ); m.z = vec3( 0, 0, 1 ); return m; }
// m3d_s_mlm (vectors.h:278-281): rows of a transformed by o
inline M3d mat_mlm( const M3d& o, const M3d& a ) { M3d m; m.x = acn::mlv( o, a.x ); m.y = acn::mlv( o, a.y ); m.z = acn::mlv( o, a.z ); return m; }
M3d mat_rot_x( double a );   // radians (vectors.h:289-307)
M3d mat_rot_y( double a );
M3d mat_rot_z( double a );

struct Envelope { V3d pos; double radius; };
Envelope envelope_of_pair( const Envelope& e1, const Envelope& e2 );          // objects.c:113-136

struct Texture { int kind = ACN_TEX_NONE; V3d c1 = vec3( 0.7, 0.7, 0.7 ), c2 = vec3( 0, 0, 0 ); double scale = 1.0; };

// properties_s (objects.h:51-78), defaults objects.c:141-177
struct Props
{
    V3d pos = vec3( 0, 0, 0 );
    M3d rax = mat_ident();
    Texture tex;
    V3d color = vec3( 0.7, 0.7, 0.7 );
    double radiance = 0, refractive_index = 1, fresnel_reflectivity = 1, chromatic_reflectivity = 0,
           diffuse_reflectivity = 1, sigma = 0, surface_roughness = 0;
    V3d transparency = vec3( 0, 0, 0 );
    bool has_envelope = false;
    Envelope envelope{ vec3( 0, 0, 0 ), 0 };

    void move( const V3d& v );       // objects.c:179-183
    void rotate( const M3d& m );     // objects.c:185-190
    void scale( double f );          // objects.c:192-196
};

struct Obj
{
    int    kind = ACN_KIND_SPHERE;
    Props  prp;
    double tail[ 4 ] = { 0, 0, 0, 0 };
    std::unique_ptr<Obj> o1, o2;

    std::unique_ptr<Obj> clone() const;
    void move( const V3d& v );
    void rotate( const M3d& m );
    void scale( double f );
    void set_refractive_index( double n );       // objects.c:436-448
    bool set_material( const std::string& name ); // objects.c:1589-1682
    void set_auto_envelope();                    // objects.c:470-476
    bool set_bounding_envelope();                // analytic bound instead of the Monte-Carlo estimate; false: shape is unbounded
};

// Conservative bounding sphere of a shape where one follows from its parameters: sphere, ellipsoid,
// distance-field sphere / torus, A&B (the smaller of the children's bounds: either contains the intersection), A|B
// (envelope_of_pair of the children's bounds, the reference's own rule, objects.c:113-136).  A shape's own envelope, if
// set, counts as its bound (obj_side reports "outside" beyond it, objects.c:365-370).  false for unbounded shapes
// (plane, cylinder, cone, hyperboloids, negations) and for scale nodes.
bool analytic_envelope( const Obj& o, Envelope* out );

std::unique_ptr<Obj> make_plane();
std::unique_ptr<Obj> make_sphere( double radius );
std::unique_ptr<Obj> make_squaroid( double a, double b, double c, double r );
std::unique_ptr<Obj> make_ellipsoid( double rx, double ry, double rz );
std::unique_ptr<Obj> make_cylinder( double rx, double ry );
std::unique_ptr<Obj> make_cone( double rx, double ry, double rz );
std::unique_ptr<Obj> make_hyperboloid1( double rx, double ry, double rz );
std::unique_ptr<Obj> make_hyperboloid2( double rx, double ry, double rz );
std::unique_ptr<Obj> make_torus( double r1, double r2 );                  // closures.c:568-591
std::unique_ptr<Obj> make_distance_sphere();
std::unique_ptr<Obj> make_pair_inside( const Obj& a, const Obj& b );      // objects.c:1011-1018
std::unique_ptr<Obj> make_pair_outside( const Obj& a, const Obj& b );     // objects.c:1161-1176
std::unique_ptr<Obj> make_neg( const Obj& a );                            // objects.c:1315-1321
std::unique_ptr<Obj> make_scale( const Obj& a, const V3d& scale );        // objects.c:1388-1407

Envelope estimate_envelope( const Obj& o, int samples, uint32_t rseed, double radius_factor );   // objects.c:312-363

struct Compound;
struct Elem
{
    std::unique_ptr<Obj> obj;
    std::unique_ptr<Compound> cmp;
};

// compound_s (compound.c:36-50)
struct Compound
{
    bool has_envelope = false;
    Envelope envelope{ vec3( 0, 0, 0 ), 0 };
    std::vector<Elem> items;

    std::unique_ptr<Compound> clone() const;
    void clear() { items.clear(); }          // compound_s_clear keeps the envelope (compound.c:114-117)
    void push_obj( const Obj& o );           // compound.c:144-165
    void push_compound( const Compound& c ); // compound.c:166-182
    void move( const V3d& v );
    void rotate( const M3d& m );
    void scale( double f );
    void set_auto_envelope();                // compound.c:73-107
    size_t count() const { return items.size(); }
};

// ---------------------------------------------------------------------------------------------
// script values (interpreter.c / container.c): value semantics with explicit clone()
// ---------------------------------------------------------------------------------------------
struct Value;
typedef std::shared_ptr<Value> VP;
struct Closure;
struct Frame;

struct Value
{
    enum Type { NIL, BOOL, INT, NUM, STR, VEC, MAT, OBJ, CMP, LIST, MAP, FUNC, SCENE, BUILTIN };
    Type type = NIL;
    bool b = false;
    long long i = 0;
    double f = 0;
    std::string s;
    V3d v = vec3( 0, 0, 0 );
    M3d m = mat_ident();
    std::unique_ptr<Obj> obj;
    std::unique_ptr<Compound> cmp;
    std::vector<VP> list;
    std::vector<std::pair<std::string, VP>> map;     // insertion-ordered
    std::shared_ptr<Closure> func;
    int builtin = -1;

    VP clone() const;
    VP map_get( const std::string& k ) const;
    void map_set( const std::string& k, VP v );
};

VP make_nil();
VP make_bool( bool b );
VP make_int( long long i );
VP make_num( double f );
VP make_str( const std::string& s );
VP make_vec( const V3d& v );
VP make_mat( const M3d& m );
VP make_obj( std::unique_ptr<Obj> o );
VP make_cmp( std::unique_ptr<Compound> c );
VP make_list();
VP make_map();

// recursive edits on any value (container.c:289-374 arr_s_move/rotate/scale, map_s_*)
bool value_move( Value& v, const V3d& d );
bool value_rotate( Value& v, const M3d& m );
bool value_scale( Value& v, double f );

std::unique_ptr<Obj> list_inside_composite( const std::vector<VP>& l, size_t start, size_t size, std::string* err );
std::unique_ptr<Obj> list_outside_composite( const std::vector<VP>& l, size_t start, size_t size, std::string* err );
bool compound_push_value( Compound& c, const Value& v, std::string* err );     // compound_s_push_q

// ---------------------------------------------------------------------------------------------
// scene_s (scene.c:153-213)
// ---------------------------------------------------------------------------------------------
struct RecordedImage
{
    std::string name;
    acn_flat_params params;
    std::unique_ptr<Compound> light, matter;
};

struct Scene
{
    acn_flat_params params;
    Compound light, matter;
    std::vector<VP> handles;                 // C-API value table
    std::vector<RecordedImage> images;       // create_image calls recorded by the .acn front-end

    // flatten output (owned)
    std::vector<acn_flat_node> f_nodes;
    std::vector<int32_t> f_children;
    std::vector<acn_flat_material> f_materials;
    acn_flat_scene flat;

    Scene();
    void clear() { light.clear(); matter.clear(); }
    bool push( const Value& v, std::string* err );     // scene.c:238-279
    void flatten();
    int  add_handle( VP v ) { handles.push_back( v ); return ( int )handles.size() - 1; }
};

void default_params( acn_flat_params* p );

// .acn front-end (acn_interp.cpp)
int interpret_file( Scene& scene, const std::string& path, const std::vector<std::string>& args, std::string* err );

} // namespace acnh
