// acn_tracer.cu — C ABI of the device tracer (see include/actinon_b200.h) and the FP32 peak probe.
// There is deliberately no CPU rendering path in this library: without a CUDA device every
// tracer entry point fails with ACN_ERR_NO_DEVICE.
#include "acn_tracer.cuh"
#include "acn_dimage.cuh"
#include "acn_group.cuh"

#include <stdarg.h>
#include <mutex>

namespace acn {

static thread_local char g_error[ 1024 ] = "";

void set_error( const char* fmt, ... )
{
    va_list ap;
    va_start( ap, fmt );
    vsnprintf( g_error, sizeof( g_error ), fmt, ap );
    va_end( ap );
}

int validate_flat_scene( const acn_flat_scene* fs )
{
    if( !fs || !fs->nodes || fs->n_nodes <= 0 ) { set_error( "flat scene: no nodes" ); return ACN_ERR_BAD_SCENE; }
    if( fs->n_children < 0 || ( fs->n_children > 0 && !fs->children ) ) { set_error( "flat scene: bad child list" ); return ACN_ERR_BAD_SCENE; }
    if( fs->n_materials < 0 || ( fs->n_materials > 0 && !fs->materials ) ) { set_error( "flat scene: bad material list" ); return ACN_ERR_BAD_SCENE; }
    const int n = fs->n_nodes;
    auto node_ok = [ & ]( int i ) { return i >= 0 && i < n; };
    if( !node_ok( fs->light_root ) || !node_ok( fs->matter_root ) ||
        fs->nodes[ fs->light_root ].kind != ACN_KIND_COMPOUND || fs->nodes[ fs->matter_root ].kind != ACN_KIND_COMPOUND )
    {
        set_error( "flat scene: light_root/matter_root must be compound nodes" ); return ACN_ERR_BAD_SCENE;
    }
    for( int i = 0; i < n; i++ )
    {
        const acn_flat_node& nd = fs->nodes[ i ];
        if( nd.kind < 0 || nd.kind >= ACN_KIND_COUNT ) { set_error( "flat scene: node %d has unknown kind %d", i, nd.kind ); return ACN_ERR_BAD_SCENE; }
        if( nd.kind == ACN_KIND_COMPOUND )
        {
            if( nd.child1 < 0 || nd.child0 < 0 || ( long long )nd.child0 + nd.child1 > fs->n_children )
            { set_error( "flat scene: compound %d child range out of bounds", i ); return ACN_ERR_BAD_SCENE; }
            for( int k = 0; k < nd.child1; k++ )
            {
                int c = fs->children[ nd.child0 + k ];
                if( !node_ok( c ) || c == i ) { set_error( "flat scene: compound %d has bad child %d", i, c ); return ACN_ERR_BAD_SCENE; }
            }
        }
        else
        {
            if( nd.material < 0 || nd.material >= fs->n_materials ) { set_error( "flat scene: node %d has bad material %d", i, nd.material ); return ACN_ERR_BAD_SCENE; }
            const bool pair = nd.kind == ACN_KIND_PAIR_INSIDE || nd.kind == ACN_KIND_PAIR_OUTSIDE;
            const bool unary = nd.kind == ACN_KIND_NEG || nd.kind == ACN_KIND_SCALE;
            if( ( pair || unary ) && ( !node_ok( nd.child0 ) || nd.child0 == i || fs->nodes[ nd.child0 ].kind == ACN_KIND_COMPOUND ) )
            { set_error( "flat scene: node %d has bad child0 %d", i, nd.child0 ); return ACN_ERR_BAD_SCENE; }
            if( pair && ( !node_ok( nd.child1 ) || nd.child1 == i || fs->nodes[ nd.child1 ].kind == ACN_KIND_COMPOUND ) )
            { set_error( "flat scene: node %d has bad child1 %d", i, nd.child1 ); return ACN_ERR_BAD_SCENE; }
            if( ( nd.kind == ACN_KIND_DIST_SPHERE || nd.kind == ACN_KIND_DIST_TORUS ) && !( nd.tail[ 0 ] != 0 ) )
            { set_error( "flat scene: distance object %d has inv_scale 0", i ); return ACN_ERR_BAD_SCENE; }
        }
    }
    if( compound_depth( fs, fs->light_root, 0 ) > COMPOUND_STACK || compound_depth( fs, fs->matter_root, 0 ) > COMPOUND_STACK )
    { set_error( "flat scene: compounds nested deeper than %d", ( int )COMPOUND_STACK ); return ACN_ERR_BAD_SCENE; }
    if( csg_depth( fs, fs->light_root, 0 ) >= 64 || csg_depth( fs, fs->matter_root, 0 ) >= 64 )
    { set_error( "flat scene: CSG nesting deeper than 63 (or cyclic)" ); return ACN_ERR_BAD_SCENE; }
    const acn_flat_params& p = fs->params;
    if( p.image_width <= 0 || p.image_height <= 1 ) { set_error( "flat scene: bad image size %dx%d", p.image_width, p.image_height ); return ACN_ERR_BAD_SCENE; }
    if( p.direct_samples < 0 || p.path_samples < 0 || p.trace_depth < 0 ) { set_error( "flat scene: negative sample counts / depth" ); return ACN_ERR_BAD_SCENE; }
    return ACN_OK;
}

// ---------------------------------------------------------------------------------------------
// FP32 peak probe: 8 independent FFMA chains per thread, every SM saturated
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 ) k_fma_peak( float* out, int iters, float a, float b )
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for( int i = 0; i < iters; i++ )
    {
        #pragma unroll
        for( int u = 0; u < 16; u++ )
        {
            x0 = fmaf( x0, a, b ); x1 = fmaf( x1, a, b ); x2 = fmaf( x2, a, b ); x3 = fmaf( x3, a, b );
            x4 = fmaf( x4, a, b ); x5 = fmaf( x5, a, b ); x6 = fmaf( x6, a, b ); x7 = fmaf( x7, a, b );
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if( s == 123.456f ) out[ 0 ] = s;
}

} // namespace acn

using namespace acn;

static_assert( TEX_NONE == ACN_TEX_NONE && TEX_PLAIN == ACN_TEX_PLAIN && TEX_CHESS == ACN_TEX_CHESS, "texture kinds of the device code = C ABI" );
static_assert( K_COMPOUND == ACN_KIND_COMPOUND && K_PLANE == ACN_KIND_PLANE && K_SCALE == ACN_KIND_SCALE, "node kinds of the device code = C ABI" );

extern "C" {

const char* acn_last_error( void ) { return g_error; }
const char* acn_version( void ) { return "actinon_b200 0.1 (sm_100a wavefront tracer)"; }

void acn_options_default( acn_options* opt )
{
    if( !opt ) return;
    memset( opt, 0, sizeof( *opt ) );
    opt->seed_mode = ACN_SEED_POSITION_HASH;
    opt->precision = ACN_PRECISION_F32;
    opt->eps = 0;
    opt->wave_budget = 0;
    opt->device = -1;
    opt->csg_mode = ACN_CSG_AUTO;
    opt->specialize = ACN_SPECIALIZE_AUTO;
}

int acn_device_count( void )
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount( &n );
    if( e != cudaSuccess || n <= 0 )
    {
        set_error( "no CUDA device: %s (this library has no CPU rendering path)", cudaGetErrorString( e ) );
        cudaGetLastError();
        return ACN_ERR_NO_DEVICE;
    }
    return n;
}

int acn_tracer_create( const acn_flat_scene* scene, const acn_options* opt, acn_tracer** out )
{
    if( !scene || !out ) { set_error( "acn_tracer_create: null argument" ); return ACN_ERR_INVALID_ARG; }
    *out = nullptr;
    acn_options o;
    if( opt ) o = *opt; else acn_options_default( &o );
    int rc = validate_flat_scene( scene );
    if( rc ) return rc;
    if( getenv( "ACN_DUMP_PROGRAMS" ) ) dump_programs( scene );
    int nd = acn_device_count();
    if( nd < 0 ) return nd;
    int dev = o.device;
    if( dev < 0 ) { if( cudaGetDevice( &dev ) != cudaSuccess ) dev = 0; }
    if( dev >= nd ) { set_error( "device %d out of range (%d devices)", dev, nd ); return ACN_ERR_INVALID_ARG; }
    TracerBase* t = nullptr;
    if( o.precision == ACN_PRECISION_F64 )
    {
        Tracer<double>* tr = new Tracer<double>(); tr->device = dev; rc = tr->init( scene, &o ); t = tr;
    }
    else
    {
        Tracer<float>* tr = new Tracer<float>(); tr->device = dev; rc = tr->init( scene, &o ); t = tr;
    }
    if( rc ) { delete t; return rc; }
    *out = reinterpret_cast<acn_tracer*>( t );
    return ACN_OK;
}

void acn_tracer_destroy( acn_tracer* t )
{
    if( t ) delete reinterpret_cast<TracerBase*>( t );
}

int acn_render_samples_device( acn_tracer* t, const double* d_xy, uint64_t n, uint64_t index_base,
                               float* d_rgb, void* stream, const volatile int* cancel, acn_stats* stats )
{
    if( !t || ( n && ( !d_xy || !d_rgb ) ) ) { set_error( "acn_render_samples_device: null argument" ); return ACN_ERR_INVALID_ARG; }
    TracerBase* tb = reinterpret_cast<TracerBase*>( t );
    // stream 0 is what it is everywhere in CUDA: the legacy default stream (which is also PyTorch's default stream), so
    // the kernels are ordered against the caller's other work on it.  The tracer's private stream: acn_tracer_stream().
    return tb->render( d_xy, n, index_base, d_rgb, ( cudaStream_t )stream, cancel, stats );
}

int acn_spec_probe( const acn_flat_scene* scene, const acn_options* opt, int compile, char* src, uint64_t cap, uint64_t* len, double* seconds )
{
    if( !scene ) { set_error( "acn_spec_probe: null scene" ); return ACN_ERR_INVALID_ARG; }
    acn_options o;
    if( opt ) o = *opt; else acn_options_default( &o );
    int rc = validate_flat_scene( scene );
    if( rc ) return rc;
    const bool f64 = o.precision == ACN_PRECISION_F64;
    CsgBuilder cb;
    cb.build( scene, o.csg_mode == ACN_CSG_INTERVALS || ( o.csg_mode == ACN_CSG_AUTO && !f64 ) );
    if( getenv( "ACN_VERBOSE" ) )
    {   // which composite objects the event sweep covers
        int swept = 0, marched = 0, maxv = 0, maxw = 0;
        std::function<void( int )> cnt = [ & ]( int c )
        {
            const acn_flat_node& cn = scene->nodes[ c ];
            for( int i = 0; i < cn.child1; i++ )
            {
                const int e = scene->children[ cn.child0 + i ];
                const acn_flat_node& nd = scene->nodes[ e ];
                if( nd.kind == ACN_KIND_COMPOUND ) { cnt( e ); continue; }
                if( nd.kind < ACN_KIND_PAIR_INSIDE ) continue;
                if( cb.prog_ref[ e ].y > 0 ) { swept++; if( cb.prog_ref[ e ].w > maxv ) maxv = cb.prog_ref[ e ].w; if( cb.prog_ref[ e ].y > maxw ) maxw = cb.prog_ref[ e ].y; }
                else { marched++; fprintf( stderr, "acn: object %d (kind %d) runs the reference march\n", e, nd.kind ); }
            }
        };
        cnt( scene->light_root ); cnt( scene->matter_root );
        fprintf( stderr, "acn: %d composite objects swept (largest: %d variables, %d words), %d marched; dist leaves %d, coincident %d\n",
                 swept, maxv, maxw, marched, ( int )cb.has_dist_leaf, ( int )cb.has_coincident );
    }
    const SpecPlan pl = f64 ? plan_spec<double>( scene, cb ) : plan_spec<float>( scene, cb );
    if( len ) *len = pl.src.size();
    if( seconds ) *seconds = 0;
    if( src && cap ) { const size_t k = pl.src.size() < cap - 1 ? pl.src.size() : ( size_t )cap - 1; memcpy( src, pl.src.data(), k ); src[ k ] = 0; }
    if( compile && !pl.src.empty() )
    {
        const char* xo = getenv( "ACN_SPEC_OPTS" );
        std::shared_ptr<SpecBinary> bin = spec_compile( pl.src, f64, pl.march, pl.sh, "sm_100a", xo ? xo : "" );
        if( !bin ) return ACN_ERR_UNSUPPORTED;
        if( seconds ) *seconds = bin->compile_seconds;
    }
    return ACN_OK;
}

void* acn_tracer_stream( acn_tracer* t )
{
    return t ? ( void* )reinterpret_cast<TracerBase*>( t )->own_stream : nullptr;
}

int acn_render_samples( acn_tracer* t, const double* xy, uint64_t n, uint64_t index_base,
                        float* rgb, const volatile int* cancel, acn_stats* stats )
{
    if( !t || ( n && ( !xy || !rgb ) ) ) { set_error( "acn_render_samples: null argument" ); return ACN_ERR_INVALID_ARG; }
    TracerBase* tb = reinterpret_cast<TracerBase*>( t );
    if( n == 0 ) { if( stats ) memset( stats, 0, sizeof( *stats ) ); return ACN_OK; }
    if( n > 0x7FFFFFFFull ) { set_error( "at most 2^31-1 samples per call" ); return ACN_ERR_INVALID_ARG; }      // before any staging allocation
    ACN_CUDA( cudaSetDevice( tb->device ) );
    if( tb->stage_cap < n )
    {
        cudaFree( tb->d_xy_stage ); cudaFree( tb->d_rgb_stage ); tb->d_xy_stage = nullptr; tb->d_rgb_stage = nullptr; tb->stage_cap = 0;
        int rc;
        if( ( rc = dev_alloc( &tb->d_xy_stage, ( size_t )n * 2 ) ) ) return rc;
        if( ( rc = dev_alloc( &tb->d_rgb_stage, ( size_t )n * 3 ) ) ) return rc;
        tb->stage_cap = n;
    }
    cudaStream_t st = tb->own_stream;
    ACN_CUDA( cudaMemcpyAsync( tb->d_xy_stage, xy, ( size_t )n * 2 * sizeof( double ), cudaMemcpyHostToDevice, st ) );
    int rc = tb->render( tb->d_xy_stage, n, index_base, tb->d_rgb_stage, st, cancel, stats );
    if( rc ) return rc;
    ACN_CUDA( cudaMemcpyAsync( rgb, tb->d_rgb_stage, ( size_t )n * 3 * sizeof( float ), cudaMemcpyDeviceToHost, st ) );
    ACN_CUDA( cudaStreamSynchronize( st ) );
    return ACN_OK;
}

int acn_accumulate_device( acn_tracer* t, const double* d_xy, const float* d_rgb, uint64_t n, float* d_accum, void* stream )
{
    if( !t || ( n && ( !d_xy || !d_rgb || !d_accum ) ) ) { set_error( "acn_accumulate_device: null argument" ); return ACN_ERR_INVALID_ARG; }
    TracerBase* tb = reinterpret_cast<TracerBase*>( t );
    if( n == 0 ) return ACN_OK;
    ACN_CUDA( cudaSetDevice( tb->device ) );
    k_accumulate<<< grid_for( n, 256 ), 256, 0, ( cudaStream_t )stream >>>( d_xy, d_rgb, n, tb->width, tb->height, d_accum );
    ACN_CUDA( cudaGetLastError() );
    return ACN_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident image + pass controller (acn_dimage.cuh)
// ---------------------------------------------------------------------------------------------
int acn_dimage_create( int device, int32_t width, int32_t height, acn_dimage** out )
{
    if( !out || width <= 0 || height <= 0 || ( int64_t )width * height >= ( 1ll << 31 ) ) { set_error( "acn_dimage_create: bad arguments" ); return ACN_ERR_INVALID_ARG; }
    *out = nullptr;
    int nd = acn_device_count();
    if( nd < 0 ) return nd;
    if( device < 0 ) { if( cudaGetDevice( &device ) != cudaSuccess ) device = 0; }
    if( device >= nd ) { set_error( "device %d out of range (%d devices)", device, nd ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( device ) );
    DImage* di = new DImage();
    di->device = device; di->width = width; di->height = height;
    int rc = ACN_OK;
    if( !rc ) rc = dev_alloc( &di->d_tot, di->words() );
    if( !rc ) rc = dev_alloc( &di->d_delta, di->words() );
    if( !rc ) rc = dev_alloc( &di->d_cnt, ( size_t )width * height );
    if( !rc ) rc = dev_alloc( &di->d_blk, ( size_t )di->blocks() );
    if( !rc ) rc = dev_alloc( &di->d_counts, 2 );
    if( !rc && cudaMallocHost( ( void** )&di->h_counts, 2 * sizeof( unsigned long long ) ) != cudaSuccess ) rc = ACN_ERR_OUT_OF_MEMORY;
    if( !rc && cudaStreamCreateWithFlags( &di->stream, cudaStreamNonBlocking ) != cudaSuccess ) rc = ACN_ERR_CUDA;
    if( !rc && ( cudaMemset( di->d_tot, 0, di->words() * 8 ) != cudaSuccess || cudaMemset( di->d_delta, 0, di->words() * 8 ) != cudaSuccess ) ) rc = ACN_ERR_CUDA;
    if( rc ) { delete di; return rc; }
    *out = reinterpret_cast<acn_dimage*>( di );
    return ACN_OK;
}

void acn_dimage_destroy( acn_dimage* d ) { if( d ) delete reinterpret_cast<DImage*>( d ); }

int acn_dimage_set_shard( acn_dimage* d, int32_t n_ranks, int32_t rank, int32_t tile )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || n_ranks < 1 || rank < 0 || rank >= n_ranks || tile < 1 || di->in_pass ) { set_error( "acn_dimage_set_shard: bad arguments" ); return ACN_ERR_INVALID_ARG; }
    di->n_ranks = n_ranks; di->rank = rank; di->tile = tile;
    return ACN_OK;
}

int32_t acn_pixel_owner( int32_t x, int32_t y, int32_t tile, int32_t n_ranks )
{
    if( x < 0 || y < 0 || tile < 1 || n_ranks < 1 ) return -1;
    return pixel_owner( x, y, tile, n_ranks );
}

int acn_dimage_reset( acn_dimage* d )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di ) { set_error( "acn_dimage_reset: null image" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    ACN_CUDA( cudaMemsetAsync( di->d_tot, 0, di->words() * 8, di->stream ) );
    ACN_CUDA( cudaMemsetAsync( di->d_delta, 0, di->words() * 8, di->stream ) );
    ACN_CUDA( cudaStreamSynchronize( di->stream ) );
    di->cycle = 0; di->rval = 21943294ull; di->in_pass = false;
    return ACN_OK;
}

int32_t  acn_dimage_cycle( const acn_dimage* d ) { return d ? reinterpret_cast<const DImage*>( d )->cycle : -1; }
uint64_t acn_dimage_rval( const acn_dimage* d ) { return d ? reinterpret_cast<const DImage*>( d )->rval : 0; }
void*    acn_dimage_stream( acn_dimage* d ) { return d ? ( void* )reinterpret_cast<DImage*>( d )->stream : nullptr; }

int acn_dimage_begin_pass( acn_dimage* d, const acn_flat_params* prm, const double** d_xy, uint64_t* n_local, uint64_t* n_total )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !prm || !n_local ) { set_error( "acn_dimage_begin_pass: null argument" ); return ACN_ERR_INVALID_ARG; }
    if( di->in_pass ) { set_error( "acn_dimage_begin_pass: the previous pass was not ended" ); return ACN_ERR_INVALID_ARG; }
    *n_local = 0; if( n_total ) *n_total = 0; if( d_xy ) *d_xy = nullptr;
    if( di->cycle > prm->gradient_cycles ) return ACN_OK;              // all gradient_cycles + 1 passes done (scene.c:1103)
    if( prm->gradient_samples < 0 || prm->gradient_samples > 0xFFFF ) { set_error( "gradient_samples out of range" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    const int W = di->width, H = di->height, nb = di->blocks();
    const double thr2 = prm->gradient_threshold * prm->gradient_threshold;
    k_img_select<<< nb, 256, 0, di->stream >>>( di->d_tot, W, H, di->cycle, thr2, prm->gradient_samples, di->tile, di->n_ranks, di->rank, di->d_cnt, di->d_blk );
    k_img_scan_blocks<<< 1, 1024, 0, di->stream >>>( di->d_blk, nb, di->d_counts );
    ACN_CUDA( cudaMemcpyAsync( di->h_counts, di->d_counts, 2 * sizeof( unsigned long long ), cudaMemcpyDeviceToHost, di->stream ) );
    ACN_CUDA( cudaStreamSynchronize( di->stream ) );
    di->pass_total = di->h_counts[ 0 ]; di->pass_local = di->h_counts[ 1 ];
    if( di->pass_local > di->xy_cap )
    {
        cudaFree( di->d_xy ); di->d_xy = nullptr; di->xy_cap = 0;
        const uint64_t cap = di->pass_local + di->pass_local / 4 + 1024;
        int rc = dev_alloc( &di->d_xy, ( size_t )cap * 2 );
        if( rc ) return rc;
        di->xy_cap = cap;
    }
    if( di->pass_local ) k_img_emit<<< nb, 256, 0, di->stream >>>( di->d_cnt, di->d_blk, W, H, di->cycle, di->rval, di->d_xy );
    ACN_CUDA( cudaStreamSynchronize( di->stream ) );
    di->in_pass = true;
    *n_local = di->pass_local; if( n_total ) *n_total = di->pass_total; if( d_xy ) *d_xy = di->d_xy;
    return ACN_OK;
}

int acn_dimage_accumulate( acn_dimage* d, const double* d_xy, const float* d_rgb, uint64_t n, void* stream )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || ( n && ( !d_xy || !d_rgb ) ) ) { set_error( "acn_dimage_accumulate: null argument" ); return ACN_ERR_INVALID_ARG; }
    if( n == 0 ) return ACN_OK;
    ACN_CUDA( cudaSetDevice( di->device ) );
    k_img_accumulate<<< grid_for( n, 256 ), 256, 0, ( cudaStream_t )stream >>>( d_xy, d_rgb, n, di->width, di->height, di->d_delta );
    ACN_CUDA( cudaGetLastError() );
    return ACN_OK;
}

uint64_t* acn_dimage_delta( acn_dimage* d, uint64_t* n_words )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di ) return nullptr;
    if( n_words ) *n_words = di->words();
    return ( uint64_t* )di->d_delta;
}

int acn_dimage_read_pass_xy( acn_dimage* d, double* xy )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || ( di->pass_local && !xy ) ) { set_error( "acn_dimage_read_pass_xy: null argument" ); return ACN_ERR_INVALID_ARG; }
    if( di->pass_local == 0 ) return ACN_OK;
    ACN_CUDA( cudaSetDevice( di->device ) );
    ACN_CUDA( cudaMemcpy( xy, di->d_xy, ( size_t )di->pass_local * 2 * sizeof( double ), cudaMemcpyDeviceToHost ) );
    return ACN_OK;
}

// delta <-> a caller-owned device buffer of the same size (a torch tensor to all-reduce over NCCL, one process per GPU)
int acn_dimage_copy_delta( acn_dimage* d, uint64_t* d_dst, void* stream )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !d_dst ) { set_error( "acn_dimage_copy_delta: null argument" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    ACN_CUDA( cudaMemcpyAsync( d_dst, di->d_delta, di->words() * 8, cudaMemcpyDeviceToDevice, ( cudaStream_t )stream ) );
    return ACN_OK;
}
int acn_dimage_set_delta( acn_dimage* d, const uint64_t* d_src, void* stream )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !d_src ) { set_error( "acn_dimage_set_delta: null argument" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    ACN_CUDA( cudaMemcpyAsync( di->d_delta, d_src, di->words() * 8, cudaMemcpyDeviceToDevice, ( cudaStream_t )stream ) );
    return ACN_OK;
}

int acn_dimage_end_pass( acn_dimage* d, void* stream )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !di->in_pass ) { set_error( "acn_dimage_end_pass: no pass in progress" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    k_img_commit<<< grid_for( di->words(), 256 ), 256, 0, ( cudaStream_t )stream >>>( di->d_tot, di->d_delta, di->words() );
    ACN_CUDA( cudaStreamSynchronize( ( cudaStream_t )stream ) );
    if( di->cycle > 0 ) di->rval = lcg00_skip( di->rval, 2ull * di->pass_total );      // the jitter stream of the pass (scene.c:1130-1131)
    di->cycle++;
    di->in_pass = false;
    return ACN_OK;
}

int acn_dimage_render_pass( acn_dimage* d, acn_tracer* t, const acn_flat_params* prm, uint64_t index_base, uint64_t* n_local, uint64_t* n_total,
                            const volatile int* cancel, acn_stats* stats )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    TracerBase* tb = reinterpret_cast<TracerBase*>( t );
    if( !di || !tb || !prm ) { set_error( "acn_dimage_render_pass: null argument" ); return ACN_ERR_INVALID_ARG; }
    if( tb->device != di->device || tb->width != di->width || tb->height != di->height ) { set_error( "acn_dimage_render_pass: tracer and image differ in device or size" ); return ACN_ERR_INVALID_ARG; }
    const double* xy = nullptr; uint64_t nl = 0, nt = 0;
    int rc = acn_dimage_begin_pass( d, prm, &xy, &nl, &nt );
    if( n_local ) *n_local = nl; if( n_total ) *n_total = nt;
    if( stats ) memset( stats, 0, sizeof( *stats ) );
    if( rc || !di->in_pass ) return rc;
    if( nl > di->rgb_cap )
    {
        cudaFree( di->d_rgb ); di->d_rgb = nullptr; di->rgb_cap = 0;
        const uint64_t cap = nl + nl / 4 + 1024;
        if( ( rc = dev_alloc( &di->d_rgb, ( size_t )cap * 3 ) ) ) return rc;
        di->rgb_cap = cap;
    }
    if( nl )
    {
        rc = tb->render( xy, nl, index_base, di->d_rgb, di->stream, cancel, stats );
        if( rc ) { di->in_pass = false; cudaMemsetAsync( di->d_delta, 0, di->words() * 8, di->stream ); cudaStreamSynchronize( di->stream ); return rc; }   // the pass is discarded (scene.c:1143-1153)
        if( ( rc = acn_dimage_accumulate( d, xy, di->d_rgb, nl, di->stream ) ) ) return rc;
    }
    if( di->n_ranks == 1 ) return acn_dimage_end_pass( d, di->stream );
    ACN_CUDA( cudaStreamSynchronize( di->stream ) );       // several ranks: the caller sums the deltas (acn_dimage_delta), then acn_dimage_end_pass
    return ACN_OK;
}

int acn_dimage_download( acn_dimage* d, acn_image* im )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !im ) { set_error( "acn_dimage_download: null argument" ); return ACN_ERR_INVALID_ARG; }
    int32_t w = 0, h = 0; acn_image_size( im, &w, &h );
    if( w != di->width || h != di->height ) { set_error( "acn_dimage_download: image size differs" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    std::vector<unsigned long long> raw( di->words() );
    ACN_CUDA( cudaMemcpy( raw.data(), di->d_tot, raw.size() * 8, cudaMemcpyDeviceToHost ) );
    std::vector<double> sums( raw.size() );
    for( size_t i = 0; i < raw.size(); i++ ) sums[ i ] = ( i % PIX_STRIDE ) == 5 ? ( double )raw[ i ] : ( double )raw[ i ] * ACN_PIX_INV;
    return acn_image_set_state( im, di->cycle, di->rval, sums.data() );
}

int acn_dimage_upload( acn_dimage* d, const acn_image* im )
{
    DImage* di = reinterpret_cast<DImage*>( d );
    if( !di || !im || di->in_pass ) { set_error( "acn_dimage_upload: bad argument" ); return ACN_ERR_INVALID_ARG; }
    int32_t w = 0, h = 0; acn_image_size( im, &w, &h );
    if( w != di->width || h != di->height ) { set_error( "acn_dimage_upload: image size differs" ); return ACN_ERR_INVALID_ARG; }
    ACN_CUDA( cudaSetDevice( di->device ) );
    std::vector<double> sums( di->words() );
    int rc = acn_image_sums( im, sums.data() );
    if( rc ) return rc;
    std::vector<unsigned long long> raw( sums.size() );
    for( size_t i = 0; i < raw.size(); i++ ) raw[ i ] = ( i % PIX_STRIDE ) == 5 ? ( unsigned long long )llrint( sums[ i ] ) : ( unsigned long long )llrint( sums[ i ] * ACN_PIX_SCALE );
    ACN_CUDA( cudaMemcpy( di->d_tot, raw.data(), raw.size() * 8, cudaMemcpyHostToDevice ) );
    di->cycle = acn_image_cycle( im ); di->rval = acn_image_rval( im );
    return ACN_OK;
}

// ---------------------------------------------------------------------------------------------
// one image on several GPUs in one process (acn_group.cuh)
// ---------------------------------------------------------------------------------------------
int acn_group_create( const acn_flat_scene* scene, const acn_options* opt, const int32_t* devices, int32_t n, acn_group** out )
{
    if( !scene || !out || n < 1 || n > GROUP_MAX ) { set_error( "acn_group_create: bad arguments (1..%d devices)", ( int )GROUP_MAX ); return ACN_ERR_INVALID_ARG; }
    *out = nullptr;
    int nd = acn_device_count();
    if( nd < 0 ) return nd;
    acn_options o;
    if( opt ) o = *opt; else acn_options_default( &o );
    Group* g = new Group();
    g->n = n; g->prm = scene->params;
    g->rc.assign( n, 0 ); g->stats.resize( n ); g->n_local.assign( n, 0 ); g->n_total.assign( n, 0 ); g->err.resize( n );
    g->bar.n = n;
    int rc = ACN_OK;
    for( int r = 0; r < n && !rc; r++ )
    {
        const int dev = devices ? devices[ r ] : r;
        if( dev < 0 || dev >= nd ) { set_error( "acn_group_create: device %d out of range (%d devices)", dev, nd ); rc = ACN_ERR_INVALID_ARG; break; }
        g->devices.push_back( dev );
        o.device = dev;
        acn_tracer* t = nullptr; acn_dimage* d = nullptr;
        if( ( rc = acn_tracer_create( scene, &o, &t ) ) ) break;
        g->tracers.push_back( t );
        if( ( rc = acn_dimage_create( dev, scene->params.image_width, scene->params.image_height, &d ) ) ) break;
        g->images.push_back( d );
        if( ( rc = acn_dimage_set_shard( d, n, r, 4 ) ) ) break;
    }
    // peer access between every pair of distinct devices
    for( int r = 0; r < n && !rc; r++ )
        for( int k = 0; k < n; k++ )
        {
            if( g->devices[ k ] == g->devices[ r ] ) continue;
            int can = 0;
            cudaSetDevice( g->devices[ r ] );
            if( cudaDeviceCanAccessPeer( &can, g->devices[ r ], g->devices[ k ] ) != cudaSuccess || !can ) { g->peer = false; continue; }
            const cudaError_t e = cudaDeviceEnablePeerAccess( g->devices[ k ], 0 );
            if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled ) g->peer = false;
            cudaGetLastError();
        }
    if( !rc && !g->peer )
    {
        g->stage.assign( n, nullptr );
        for( int r = 0; r < n && !rc; r++ ) { cudaSetDevice( g->devices[ r ] ); rc = dev_alloc( &g->stage[ r ], g->img( r )->words() ); }
    }
    if( rc ) { delete g; return rc; }
    for( int r = 0; r < n; r++ ) g->threads.emplace_back( [ g, r ] { g->worker( r ); } );
    *out = reinterpret_cast<acn_group*>( g );
    return ACN_OK;
}

void acn_group_destroy( acn_group* gp ) { if( gp ) delete reinterpret_cast<Group*>( gp ); }

int acn_group_size( const acn_group* gp ) { return gp ? reinterpret_cast<const Group*>( gp )->n : 0; }

int acn_group_uses_peer_access( const acn_group* gp ) { return gp && reinterpret_cast<const Group*>( gp )->peer ? 1 : 0; }

int acn_group_render_pass( acn_group* gp, uint64_t index_base, uint64_t* n_samples, const volatile int* cancel, acn_stats* stats )
{
    Group* g = reinterpret_cast<Group*>( gp );
    if( !g ) { set_error( "acn_group_render_pass: null group" ); return ACN_ERR_INVALID_ARG; }
    if( n_samples ) *n_samples = 0;
    if( stats ) memset( stats, 0, sizeof( *stats ) );
    {
        std::unique_lock<std::mutex> lk( g->mu );
        g->index_base = index_base; g->cancel = cancel; g->done = 0; g->cmd = 1; g->seq++;
        g->cv_cmd.notify_all();
        g->cv_done.wait( lk, [ & ] { return g->done == g->n; } );
    }
    for( int r = 0; r < g->n; r++ )
        if( g->rc[ r ] ) { set_error( "rank %d (device %d): %s", r, g->devices[ r ], g->err[ r ].c_str() ); return g->rc[ r ]; }
    if( n_samples ) *n_samples = g->n_total[ 0 ];
    if( stats )
        for( int r = 0; r < g->n; r++ )
        {
            const acn_stats& s = g->stats[ r ];
            stats->samples += s.samples; stats->rays_primary += s.rays_primary; stats->rays_reflection += s.rays_reflection;
            stats->rays_chromatic += s.rays_chromatic; stats->rays_refraction += s.rays_refraction; stats->rays_path += s.rays_path;
            stats->rays_shadow += s.rays_shadow; stats->rays_light += s.rays_light; stats->diffuse_hits += s.diffuse_hits;
            stats->kernel_launches += s.kernel_launches; stats->waves = s.waves > stats->waves ? s.waves : stats->waves;
            stats->device_ms = s.device_ms > stats->device_ms ? s.device_ms : stats->device_ms;
        }
    return ACN_OK;
}

int acn_group_download( acn_group* gp, acn_image* im )
{
    Group* g = reinterpret_cast<Group*>( gp );
    if( !g || !im ) { set_error( "acn_group_download: null argument" ); return ACN_ERR_INVALID_ARG; }
    return acn_dimage_download( g->images[ 0 ], im );       // every rank holds the whole image
}

int acn_group_upload( acn_group* gp, const acn_image* im )
{
    Group* g = reinterpret_cast<Group*>( gp );
    if( !g || !im ) { set_error( "acn_group_upload: null argument" ); return ACN_ERR_INVALID_ARG; }
    for( int r = 0; r < g->n; r++ ) { const int rc = acn_dimage_upload( g->images[ r ], im ); if( rc ) return rc; }
    return ACN_OK;
}

acn_dimage* acn_group_image( acn_group* gp, int32_t rank )
{
    Group* g = reinterpret_cast<Group*>( gp );
    return g && rank >= 0 && rank < g->n ? g->images[ rank ] : nullptr;
}

double acn_measure_fp32_peak_tflops( int device )
{
    int nd = acn_device_count();
    if( nd < 0 ) return ( double )nd;
    if( device < 0 ) { if( cudaGetDevice( &device ) != cudaSuccess ) device = 0; }
    if( cudaSetDevice( device ) != cudaSuccess ) return ( double )ACN_ERR_CUDA;
    cudaDeviceProp pr;
    if( cudaGetDeviceProperties( &pr, device ) != cudaSuccess ) return ( double )ACN_ERR_CUDA;
    float* d = nullptr;
    if( cudaMalloc( &d, 256 ) != cudaSuccess ) return ( double )ACN_ERR_OUT_OF_MEMORY;
    const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
    k_fma_peak<<< blocks, threads >>>( d, 64, 0.999f, 0.001f );       // warm-up
    double best = 0;
    for( int rep = 0; rep < 5; rep++ )
    {
        cudaEventRecord( e0 );
        k_fma_peak<<< blocks, threads >>>( d, iters, 0.999f, 0.001f );
        cudaEventRecord( e1 );
        cudaEventSynchronize( e1 );
        float ms = 0; cudaEventElapsedTime( &ms, e0, e1 );
        double flops = 2.0 * 8 * 16 * ( double )iters * blocks * threads;
        double tf = flops / ( ms * 1e-3 ) / 1e12;
        if( tf > best ) best = tf;
    }
    cudaEventDestroy( e0 ); cudaEventDestroy( e1 ); cudaFree( d );
    if( cudaGetLastError() != cudaSuccess ) return ( double )ACN_ERR_CUDA;
    return best;
}

} // extern "C"
