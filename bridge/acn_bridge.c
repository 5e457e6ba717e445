/* bridge/acn_bridge.c — the reference-side binding of libactinon_b200.so: what a maintainer of johsteffens/actinon adds to
 * the reference tree so that `actinon <script.acn>` renders on B200s.  Built with the reference (-DACN_B200, -Iinclude of
 * this repository, -lactinon_b200); here it is only SYNTAX-CHECKED against the reference's own headers where they lie
 * (bridge/Makefile: gcc -fsyntax-only with bridge/shim standing in for beth, which is not in the reference tree).
 *
 * Two parts:
 *   1. acn_flatten_scene: walks scene_s -> compound_s -> obj_*_s by type tag (quicktypes.h:38-47) and fills acn_flat_scene.
 *   2. lum_machine_s_run (src/scene.c:1017-1028) and the pass loop around it on the device.
 *
 * The shape structs, compound_s, scene_s and lum_s are private to objects.c / compound.c / scene.c.  The mirrors below
 * repeat their layouts (file:line given at each); inside the reference tree they are replaced by including this file at
 * the end of the owning .c file, or by moving the typedefs into the headers.  Field access is all this file needs. */
#include <stdlib.h>
#include <string.h>
#include <signal.h>

#include "actinon_b200.h"
#include "objects.h"
#include "compound.h"
#include "distance.h"
#include "scene.h"

/* ---- layout mirrors (reference file:line) -------------------------------------------------------------------------- */
typedef struct { obj_hdr_s hdr; }                                   acnb_plane_s;        /* objects.c:481-493  */
typedef struct { obj_hdr_s hdr; f3_t radius; }                      acnb_sphere_s;       /* objects.c:556-569  */
typedef struct { obj_hdr_s hdr; f3_t a, b, c, r; }                  acnb_squaroid_s;     /* objects.c:669-683  */
typedef struct { obj_hdr_s hdr; f3_t inv_scale; uz_t cycles; vd_t distance; } acnb_distance_s;   /* objects.c:836-851 */
typedef struct { obj_hdr_s hdr; vd_t o1; vd_t o2; }                 acnb_pair_s;         /* objects.c:975-989, 1125-1139 */
typedef struct { obj_hdr_s hdr; vd_t o1; }                          acnb_neg_s;          /* objects.c:1282-1295 */
typedef struct { obj_hdr_s hdr; v3d_s inv_scale; vd_t o1; }         acnb_scale_s;        /* objects.c:1353-1367 */
typedef struct { aware_t _; envelope_s* envelope; vd_t* data; uz_t size; uz_t space; } acnb_compound_s;   /* compound.c:36-50 */
typedef struct { aware_t _; distance_fp fp_distance; f3_t ex_radius; } acnb_distance_torus_s;              /* distance.c:61-66 */
typedef struct { aware_t _; vc_t p; cl_s color; }                   acnb_txm_plain_s;    /* textures.c:75-80   */
typedef struct { aware_t _; vc_t p; cl_s color1; cl_s color2; f3_t scale; } acnb_txm_chess_s;              /* textures.c:121-128 */
typedef struct                                                                            /* scene.c:153-183    */
{
    aware_t _;
    uz_t threads, image_width, image_height;
    f3_t gamma, gradient_threshold;
    uz_t gradient_samples, gradient_cycles;
    cl_s background_color;
    v3d_s camera_position, camera_view_direction, camera_top_direction;
    f3_t camera_focal_length;
    uz_t trace_depth;
    f3_t trace_min_intensity;
    uz_t direct_samples, path_samples;
    f3_t max_path_length;
    acnb_compound_s* light;
    acnb_compound_s* matter;
    s3_t experimental_level;
} acnb_scene_s;
typedef struct { v2d_s pos; cl_s clr; f3_t weight; } acnb_lum_s;                          /* scene.c:682-687    */
typedef struct { aware_t _; acnb_lum_s* data; uz_t size; uz_t space; } acnb_lum_arr_s;    /* scene.c:700-720    */

#ifndef TYPEOF_distance_torus_s
#define TYPEOF_distance_torus_s typeof( "distance_torus_s" )                               /* distance.c:59      */
#endif

/* ---- growing arrays of the flat scene ------------------------------------------------------------------------------ */
typedef struct
{
    acn_flat_node* nodes; int32_t n, cap;
    int32_t* kids; int32_t nk, ck;
    acn_flat_material* mats; int32_t nm, cm;
    int failed;
} acn_bridge_s;

static void* bridge_grow( void* p, int32_t* cap, int32_t need, size_t unit )
{
    if( need <= *cap ) return p;
    int32_t c = *cap ? *cap : 64;
    while( c < need ) c *= 2;
    *cap = c;
    return realloc( p, ( size_t )c * unit );
}

static int32_t bridge_push_node( acn_bridge_s* b, int32_t kind )
{
    b->nodes = bridge_grow( b->nodes, &b->cap, b->n + 1, sizeof( acn_flat_node ) );
    acn_flat_node* n = &b->nodes[ b->n ];
    memset( n, 0, sizeof( *n ) );
    n->kind = kind; n->child0 = n->child1 = n->material = -1;
    n->rax[ 0 ] = n->rax[ 4 ] = n->rax[ 8 ] = 1.0;
    return b->n++;
}

static int32_t bridge_push_children( acn_bridge_s* b, const int32_t* idx, int32_t count )
{
    b->kids = bridge_grow( b->kids, &b->ck, b->nk + count, sizeof( int32_t ) );
    memcpy( b->kids + b->nk, idx, ( size_t )count * sizeof( int32_t ) );
    b->nk += count;
    return b->nk - count;
}

static void bridge_envelope( acn_flat_node* n, const envelope_s* e )                      /* objects.h:30-34 */
{
    n->has_envelope = e != NULL;
    if( e ) { n->env_pos[ 0 ] = e->pos.x; n->env_pos[ 1 ] = e->pos.y; n->env_pos[ 2 ] = e->pos.z; n->env_radius = e->radius; }
}

static int32_t bridge_material( acn_bridge_s* b, const properties_s* p )                  /* objects.h:51-78 */
{
    b->mats = bridge_grow( b->mats, &b->cm, b->nm + 1, sizeof( acn_flat_material ) );
    acn_flat_material* m = &b->mats[ b->nm ];
    memset( m, 0, sizeof( *m ) );
    m->color[ 0 ] = p->color.x; m->color[ 1 ] = p->color.y; m->color[ 2 ] = p->color.z;
    m->radiance = p->radiance;
    m->refractive_index = p->refractive_index;
    m->fresnel_reflectivity = p->fresnel_reflectivity;
    m->chromatic_reflectivity = p->chromatic_reflectivity;
    m->diffuse_reflectivity = p->diffuse_reflectivity;
    m->sigma = p->sigma;
    m->transparency[ 0 ] = p->transparency.x; m->transparency[ 1 ] = p->transparency.y; m->transparency[ 2 ] = p->transparency.z;
    m->texture_kind = ACN_TEX_NONE;
    if( p->texture_field )                                                                /* textures.c:75-148 */
    {
        const tp_t t = *( const aware_t* )p->texture_field;
        if( t == TYPEOF_txm_plain_s )
        {
            const acnb_txm_plain_s* x = p->texture_field;
            m->texture_kind = ACN_TEX_PLAIN;
            m->tex_color1[ 0 ] = x->color.x; m->tex_color1[ 1 ] = x->color.y; m->tex_color1[ 2 ] = x->color.z;
        }
        else if( t == TYPEOF_txm_chess_s )
        {
            const acnb_txm_chess_s* x = p->texture_field;
            m->texture_kind = ACN_TEX_CHESS;
            m->tex_color1[ 0 ] = x->color1.x; m->tex_color1[ 1 ] = x->color1.y; m->tex_color1[ 2 ] = x->color1.z;
            m->tex_color2[ 0 ] = x->color2.x; m->tex_color2[ 1 ] = x->color2.y; m->tex_color2[ 2 ] = x->color2.z;
            m->tex_scale = x->scale;
        }
        else b->failed = 1;
    }
    return b->nm++;
}

/* one compound_s or obj_*_s -> its node index (the nodes of its children follow it) */
static int32_t bridge_object( acn_bridge_s* b, vc_t obj )
{
    const tp_t type = *( const aware_t* )obj;
    if( type == TYPEOF_compound_s )
    {
        const acnb_compound_s* c = obj;
        const int32_t self = bridge_push_node( b, ACN_KIND_COMPOUND );
        int32_t* tmp = malloc( ( c->size ? c->size : 1 ) * sizeof( int32_t ) );
        for( uz_t i = 0; i < c->size; i++ ) tmp[ i ] = bridge_object( b, c->data[ i ] );
        const int32_t first = bridge_push_children( b, tmp, ( int32_t )c->size );
        free( tmp );
        b->nodes[ self ].child0 = first;                  /* first index into acn_flat_scene.children */
        b->nodes[ self ].child1 = ( int32_t )c->size;
        bridge_envelope( &b->nodes[ self ], c->envelope );
        return self;
    }
    const obj_hdr_s* h = obj;
    int32_t kind = -1;
    if(      type == TYPEOF_obj_plane_s        ) kind = ACN_KIND_PLANE;
    else if( type == TYPEOF_obj_sphere_s       ) kind = ACN_KIND_SPHERE;
    else if( type == TYPEOF_obj_squaroid_s     ) kind = ACN_KIND_SQUAROID;
    else if( type == TYPEOF_obj_distance_s     )
        kind = *( const aware_t* )( ( const acnb_distance_s* )obj )->distance == TYPEOF_distance_torus_s ? ACN_KIND_DIST_TORUS : ACN_KIND_DIST_SPHERE;
    else if( type == TYPEOF_obj_pair_inside_s  ) kind = ACN_KIND_PAIR_INSIDE;
    else if( type == TYPEOF_obj_pair_outside_s ) kind = ACN_KIND_PAIR_OUTSIDE;
    else if( type == TYPEOF_obj_neg_s          ) kind = ACN_KIND_NEG;
    else if( type == TYPEOF_obj_scale_s        ) kind = ACN_KIND_SCALE;
    if( kind < 0 ) { b->failed = 1; return -1; }
    const int32_t self = bridge_push_node( b, kind );
    const int32_t mat = bridge_material( b, &h->prp );
    int32_t c0 = -1, c1 = -1;
    double tail[ 4 ] = { 0, 0, 0, 0 };
    switch( kind )
    {
        case ACN_KIND_SPHERE:   tail[ 0 ] = ( ( const acnb_sphere_s* )obj )->radius; break;
        case ACN_KIND_SQUAROID: { const acnb_squaroid_s* q = obj; tail[ 0 ] = q->a; tail[ 1 ] = q->b; tail[ 2 ] = q->c; tail[ 3 ] = q->r; } break;
        case ACN_KIND_DIST_SPHERE:
        case ACN_KIND_DIST_TORUS:
        {
            const acnb_distance_s* d = obj;
            tail[ 0 ] = d->inv_scale; tail[ 2 ] = ( double )d->cycles;
            if( kind == ACN_KIND_DIST_TORUS ) tail[ 1 ] = ( ( const acnb_distance_torus_s* )d->distance )->ex_radius;
        }
        break;
        case ACN_KIND_PAIR_INSIDE:
        case ACN_KIND_PAIR_OUTSIDE: { const acnb_pair_s* p = obj; c0 = bridge_object( b, p->o1 ); c1 = bridge_object( b, p->o2 ); } break;
        case ACN_KIND_NEG:      c0 = bridge_object( b, ( ( const acnb_neg_s* )obj )->o1 ); break;
        case ACN_KIND_SCALE:    { const acnb_scale_s* s = obj; tail[ 0 ] = s->inv_scale.x; tail[ 1 ] = s->inv_scale.y; tail[ 2 ] = s->inv_scale.z; c0 = bridge_object( b, s->o1 ); } break;
        default: break;
    }
    acn_flat_node* n = &b->nodes[ self ];               /* taken after the recursion: the array may have moved */
    n->pos[ 0 ] = h->prp.pos.x; n->pos[ 1 ] = h->prp.pos.y; n->pos[ 2 ] = h->prp.pos.z;
    memcpy( n->rax, &h->prp.rax, 9 * sizeof( double ) );                                  /* m3d_s = rows x, y, z (vectors.h:246) */
    n->surface_roughness = h->prp.surface_roughness;
    n->material = mat; n->child0 = c0; n->child1 = c1;
    memcpy( n->tail, tail, sizeof( tail ) );
    bridge_envelope( n, h->prp.envelope );
    return self;
}

void acn_bridge_down( acn_bridge_s* b ) { free( b->nodes ); free( b->kids ); free( b->mats ); memset( b, 0, sizeof( *b ) ); }

/* scene_s (scene.c:153-183) -> acn_flat_scene; the arrays stay owned by *b */
int acn_flatten_scene( const void* scene, acn_bridge_s* b, acn_flat_scene* out )
{
    const acnb_scene_s* s = scene;
    memset( out, 0, sizeof( *out ) );
    acn_flat_params* p = &out->params;
    p->image_width = ( int32_t )s->image_width; p->image_height = ( int32_t )s->image_height;
    p->gamma = s->gamma;
    p->background_color[ 0 ] = s->background_color.x; p->background_color[ 1 ] = s->background_color.y; p->background_color[ 2 ] = s->background_color.z;
    p->camera_position[ 0 ] = s->camera_position.x; p->camera_position[ 1 ] = s->camera_position.y; p->camera_position[ 2 ] = s->camera_position.z;
    p->camera_view_direction[ 0 ] = s->camera_view_direction.x; p->camera_view_direction[ 1 ] = s->camera_view_direction.y; p->camera_view_direction[ 2 ] = s->camera_view_direction.z;
    p->camera_top_direction[ 0 ] = s->camera_top_direction.x; p->camera_top_direction[ 1 ] = s->camera_top_direction.y; p->camera_top_direction[ 2 ] = s->camera_top_direction.z;
    p->camera_focal_length = s->camera_focal_length;
    p->trace_depth = ( int32_t )s->trace_depth; p->trace_min_intensity = s->trace_min_intensity;
    p->direct_samples = ( int32_t )s->direct_samples; p->path_samples = ( int32_t )s->path_samples; p->max_path_length = s->max_path_length;
    p->gradient_samples = ( int32_t )s->gradient_samples; p->gradient_cycles = ( int32_t )s->gradient_cycles;
    p->gradient_threshold = s->gradient_threshold; p->threads = ( int32_t )s->threads;
    out->light_root  = bridge_object( b, s->light );
    out->matter_root = bridge_object( b, s->matter );
    out->nodes = b->nodes; out->n_nodes = b->n;
    out->children = b->kids; out->n_children = b->nk;
    out->materials = b->mats; out->n_materials = b->nm;
    return ( b->failed || out->light_root < 0 || out->matter_root < 0 ) ? ACN_ERR_BAD_SCENE : ACN_OK;
}

/* ---- the call site: src/scene.c:1017-1028 (the pthread implementation stays as the #else branch) ------------------- */
#ifdef ACN_B200
extern volatile sig_atomic_t signal_received_g;                       /* scene.c:893-902 */

typedef struct lum_arr_s lum_arr_s;                                   /* private to scene.c (:700-720); see acnb_lum_arr_s */

static acn_tracer* acn_tracer_g = NULL;           /* one per process; the scene is immutable during create_image */

void lum_machine_s_run( const scene_s* scene, lum_arr_s* lum_arr_ )
{
    acnb_lum_arr_s* lum_arr = ( acnb_lum_arr_s* )lum_arr_;
    if( !acn_tracer_g )
    {
        acn_bridge_s bridge; acn_flat_scene flat; acn_options opt;
        memset( &bridge, 0, sizeof( bridge ) );
        acn_options_default( &opt );
        opt.specialize = ACN_SPECIALIZE_ON;                           /* kernels compiled for this scene's structure */
        if( acn_flatten_scene( scene, &bridge, &flat ) ) bcore_err_fa( "actinon_b200: unknown object type in the scene\n" );
        int rc = acn_tracer_create( &flat, &opt, &acn_tracer_g );
        if( rc == ACN_ERR_UNSUPPORTED ) { opt.specialize = ACN_SPECIALIZE_OFF; rc = acn_tracer_create( &flat, &opt, &acn_tracer_g ); }
        if( rc ) bcore_err_fa( "actinon_b200: #<sc_t>\n", acn_last_error() );
        acn_bridge_down( &bridge );
    }
    const uz_t n = lum_arr->size;
    double* xy  = malloc( 2 * n * sizeof( double ) );
    float*  rgb = malloc( 3 * n * sizeof( float ) );
    for( uz_t i = 0; i < n; i++ ) { xy[ 2 * i ] = lum_arr->data[ i ].pos.x; xy[ 2 * i + 1 ] = lum_arr->data[ i ].pos.y; }
    static volatile int cancel = 0;
    cancel = ( signal_received_g == SIGINT );                         /* polled between wavefront iterations, like scene.c:978 */
    const int rc = acn_render_samples( acn_tracer_g, xy, n, 0, rgb, &cancel, NULL );
    if( rc == ACN_OK )
        for( uz_t i = 0; i < n; i++ ) { cl_s c = { rgb[ 3 * i ], rgb[ 3 * i + 1 ], rgb[ 3 * i + 2 ] }; lum_arr->data[ i ].clr = c; }
    else if( rc != ACN_ERR_CANCELLED ) bcore_err_fa( "actinon_b200: #<sc_t>\n", acn_last_error() );
    free( xy ); free( rgb );
}

/* Several GPUs, and the pass loop of scene_s_create_image_file (scene.c:1103-1159) on the devices: instead of calling
 * lum_machine_s_run once per pass, the image lives on the GPUs (acn_group_*: pixel tiles dealt to the devices, per-pass
 * sums exchanged through peer memory, bit-identical for any number of devices) and comes back after every pass for the
 * .pnm writer (scene.c:866-885), which stays as it is. */
int acn_bridge_render( const scene_s* scene, int n_devices, acn_image* image, void ( *after_pass )( const acn_image*, int pass, void* ), void* arg )
{
    acn_bridge_s bridge; acn_flat_scene flat; acn_options opt; acn_group* g = NULL;
    memset( &bridge, 0, sizeof( bridge ) );
    acn_options_default( &opt );
    if( acn_flatten_scene( scene, &bridge, &flat ) ) return ACN_ERR_BAD_SCENE;
    opt.specialize = ACN_SPECIALIZE_ON;
    int rc = acn_group_create( &flat, &opt, NULL, n_devices, &g );
    if( rc == ACN_ERR_UNSUPPORTED ) { opt.specialize = ACN_SPECIALIZE_OFF; rc = acn_group_create( &flat, &opt, NULL, n_devices, &g ); }
    acn_bridge_down( &bridge );
    if( rc ) return rc;
    static volatile int cancel = 0;
    uint64_t base = 0, n = 0;
    for( int pass = 0; ; pass++ )
    {
        cancel = ( signal_received_g == SIGINT );
        rc = acn_group_render_pass( g, base, &n, &cancel, NULL );
        if( rc || n == 0 ) break;
        base += n;
        if( ( rc = acn_group_download( g, image ) ) ) break;
        if( after_pass ) after_pass( image, pass, arg );
    }
    acn_group_destroy( g );
    return rc;
}
#endif
