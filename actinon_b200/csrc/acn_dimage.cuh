// acn_dimage.cuh — the pass controller of the reference on the device.
//
// scene_s_create_image_file (reference src/scene.c:1103-1159) alternates "build the sample list of the pass" and
// "lum_machine_s_run": pass 0 samples every pixel centre, passes 1..gradient_cycles re-sample, with gradient_samples
// jittered positions each, the pixels whose 3x3 neighbourhood of running averages differs from them by more than the
// threshold (lum_image_s_sqr_grad, scene.c:848-862), drawing the jitter from ONE sequential LCG stream in raster order
// (scene.c:1124-1138), and accumulates the results per pixel (lum_image_s_push, scene.c:804-813).
//
// Here the running sums live in device memory, the selection is a stencil kernel, the sample list comes out of a prefix
// sum over the selected pixels with the LCG advanced by skip-ahead (sample s of the pass uses draws 2s and 2s+1 of the
// stream, so the list is bit-identical to the sequential one), and the results are accumulated where they are.  Per pass
// the host reads back two counters.
//
// Sums are 64-bit fixed point (Q20.44: every float32 sample value >= 2^-20 is represented exactly, as in the reference's
// double sums) and integer addition is associative, so an image is bit-identical however its samples are spread over
// passes' waves or over GPUs.  For several GPUs the pixels are dealt to the ranks in small tiles along a Morton curve
// (every aligned group of n_ranks tiles holds each rank once); a rank traces and accumulates only its own pixels into
// the pass DELTA, the deltas are summed over the ranks (disjoint support, integer sum) and added to every rank's totals,
// which keeps the selection of the next pass identical everywhere.
#pragma once

#include "acn_kernels.cuh"

namespace acn {

#define ACN_PIX_SCALE 17592186044416.0           // 2^44
#define ACN_PIX_INV   5.6843418860808015e-14     // 2^-44
enum { PIX_STRIDE = 6 };                         // pos.x, pos.y, r, g, b, weight

__host__ __device__ __forceinline__ unsigned int morton2( unsigned int x, unsigned int y )
{
    unsigned int r = 0;
    for( int b = 0; b < 16; b++ ) r |= ( ( x >> b ) & 1u ) << ( 2 * b ) | ( ( y >> b ) & 1u ) << ( 2 * b + 1 );
    return r;
}
__host__ __device__ __forceinline__ int pixel_owner( int x, int y, int tile, int n_ranks )
{
    return n_ranks <= 1 ? 0 : ( int )( morton2( ( unsigned int )( x / tile ), ( unsigned int )( y / tile ) ) % ( unsigned int )n_ranks );
}

// lum_image_s_get_avg (scene.c:824-835)
__device__ __forceinline__ void pix_avg( const unsigned long long* __restrict__ tot, int W, int H, int x, int y, double c[ 3 ] )
{
    const unsigned long long* p = tot + ( size_t )PIX_STRIDE * ( ( size_t )y * W + x );
    const unsigned long long w = p[ 5 ];
    const double f = w > 0 ? 1.0 / ( double )w : 1.0;
    c[ 0 ] = ( double )p[ 2 ] * ACN_PIX_INV * f; c[ 1 ] = ( double )p[ 3 ] * ACN_PIX_INV * f; c[ 2 ] = ( double )p[ 4 ] * ACN_PIX_INV * f;
}

// Per pixel: how many samples the pass draws for it (pass 0: one; later passes: gradient_samples where the squared colour
// distance to one of the 8 neighbours exceeds threshold^2, strictly — scene.c:848-862,1127) and whether this rank owns
// it.  cnt[ p ] = count | owned << 31; blk[ b ] = ( count summed over the block ) << 32 | ( the owned part of it ).
__global__ void __launch_bounds__( 256 )
k_img_select( const unsigned long long* __restrict__ tot, int W, int H, int cycle, double thr2, int gs, int tile, int n_ranks, int rank,
              unsigned int* __restrict__ cnt, unsigned long long* __restrict__ blk )
{
    __shared__ unsigned int sg[ 8 ], sl[ 8 ];
    const size_t p = ( size_t )blockIdx.x * 256 + threadIdx.x;
    unsigned int c = 0, own = 0;
    if( p < ( size_t )W * H )
    {
        const int x = ( int )( p % W ), y = ( int )( p / W );
        if( cycle == 0 ) c = 1;
        else
        {
            double v[ 3 ]; pix_avg( tot, W, H, x, y, v );
            double g0 = 0;
            for( int dx = -1; dx <= 1; dx++ )
                for( int dy = -1; dy <= 1; dy++ )
                {
                    if( dx == 0 && dy == 0 ) continue;
                    const int xx = x + dx, yy = y + dy;
                    if( xx < 0 || xx >= W || yy < 0 || yy >= H ) continue;          // lum_image_s_clr_dev returns 0 outside
                    double q[ 3 ]; pix_avg( tot, W, H, xx, yy, q );
                    const double d0 = v[ 0 ] - q[ 0 ], d1 = v[ 1 ] - q[ 1 ], d2 = v[ 2 ] - q[ 2 ];
                    const double g1 = d0 * d0 + d1 * d1 + d2 * d2;
                    g0 = g1 > g0 ? g1 : g0;
                }
            if( g0 > thr2 ) c = ( unsigned int )gs;
        }
        own = pixel_owner( x, y, tile, n_ranks ) == rank ? 1u : 0u;
        cnt[ p ] = c | ( own << 31 );
    }
    unsigned int g = c, l = own ? c : 0;
    #pragma unroll
    for( int o = 16; o > 0; o >>= 1 ) { g += __shfl_down_sync( ACN_FULL, g, o ); l += __shfl_down_sync( ACN_FULL, l, o ); }
    if( ( threadIdx.x & 31 ) == 0 ) { sg[ threadIdx.x >> 5 ] = g; sl[ threadIdx.x >> 5 ] = l; }
    __syncthreads();
    if( threadIdx.x == 0 )
    {
        unsigned long long G = 0, L = 0;
        for( int k = 0; k < 8; k++ ) { G += sg[ k ]; L += sl[ k ]; }
        blk[ blockIdx.x ] = ( G << 32 ) | L;
    }
}

// exclusive scan of the block sums (both halves at once: neither half reaches 2^32), in place; totals -> out[ 0..1 ]
__global__ void __launch_bounds__( 1024 )
k_img_scan_blocks( unsigned long long* __restrict__ blk, int n, unsigned long long* __restrict__ out )
{
    __shared__ unsigned long long part[ 1024 ];
    const int per = ( n + 1023 ) / 1024, lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    unsigned long long s = 0;
    for( int i = lo; i < hi; i++ ) s += blk[ i ];
    part[ threadIdx.x ] = s;
    __syncthreads();
    if( threadIdx.x == 0 )
    {
        unsigned long long run = 0;
        for( int k = 0; k < 1024; k++ ) { const unsigned long long v = part[ k ]; part[ k ] = run; run += v; }
        out[ 0 ] = run >> 32; out[ 1 ] = run & 0xFFFFFFFFull;
    }
    __syncthreads();
    unsigned long long run = part[ threadIdx.x ];
    for( int i = lo; i < hi; i++ ) { const unsigned long long v = blk[ i ]; blk[ i ] = run; run += v; }
}

// the sample positions of the pass, this rank's part: sample s (global index, raster order, gradient_samples per selected
// pixel) of a gradient pass is ( x + rnd1, y + rnd1 ) with draws 2s and 2s+1 of the stream that starts at rval
__global__ void __launch_bounds__( 256 )
k_img_emit( const unsigned int* __restrict__ cnt, const unsigned long long* __restrict__ blk, int W, int H, int cycle, u64 rval, double* __restrict__ xy )
{
    __shared__ unsigned long long ws[ 8 ];
    const size_t p = ( size_t )blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int c = 0, own = 0;
    if( p < ( size_t )W * H ) { const unsigned int v = cnt[ p ]; c = v & 0x7FFFFFFFu; own = v >> 31; }
    unsigned long long mine = ( ( unsigned long long )c << 32 ) | ( own ? c : 0u ), inc = mine;
    #pragma unroll
    for( int o = 1; o < 32; o <<= 1 ) { const unsigned long long v = __shfl_up_sync( ACN_FULL, inc, o ); if( lane >= o ) inc += v; }
    if( lane == 31 ) ws[ wid ] = inc;
    __syncthreads();
    unsigned long long base = blk[ blockIdx.x ];
    for( int k = 0; k < wid; k++ ) base += ws[ k ];
    const unsigned long long excl = base + inc - mine;
    if( !own || c == 0 ) return;
    const unsigned long long g0 = excl >> 32, l0 = excl & 0xFFFFFFFFull;
    const int x = ( int )( p % W ), y = ( int )( p / W );
    if( cycle == 0 ) { xy[ 2 * l0 ] = x + 0.5; xy[ 2 * l0 + 1 ] = y + 0.5; return; }
    u64 rv = lcg00_skip( rval, 2ull * g0 );
    for( unsigned int k = 0; k < c; k++ )
    {
        const double dx = rnd1<double>( &rv ), dy = rnd1<double>( &rv );
        xy[ 2 * ( l0 + k ) ] = x + dx; xy[ 2 * ( l0 + k ) + 1 ] = y + dy;
    }
}

// lum_image_s_push (scene.c:804-813): pixel = truncation of the sample position (weight 1)
__global__ void k_img_accumulate( const double* __restrict__ xy, const float* __restrict__ rgb, unsigned long long n, int W, int H, unsigned long long* __restrict__ acc )
{
    const unsigned long long i = ( unsigned long long )blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    const double px = xy[ 2 * i ], py = xy[ 2 * i + 1 ];
    const int x = ( int )px, y = ( int )py;
    if( x < 0 || x >= W || y < 0 || y >= H ) return;
    unsigned long long* a = acc + ( size_t )PIX_STRIDE * ( ( size_t )y * W + x );
    atomicAdd( a + 0, ( unsigned long long )__double2ll_rn( px * ACN_PIX_SCALE ) );
    atomicAdd( a + 1, ( unsigned long long )__double2ll_rn( py * ACN_PIX_SCALE ) );
    atomicAdd( a + 2, ( unsigned long long )__double2ll_rn( ( double )rgb[ 3 * i + 0 ] * ACN_PIX_SCALE ) );
    atomicAdd( a + 3, ( unsigned long long )__double2ll_rn( ( double )rgb[ 3 * i + 1 ] * ACN_PIX_SCALE ) );
    atomicAdd( a + 4, ( unsigned long long )__double2ll_rn( ( double )rgb[ 3 * i + 2 ] * ACN_PIX_SCALE ) );
    atomicAdd( a + 5, 1ull );
}

// totals += delta; delta = 0
__global__ void k_img_commit( unsigned long long* __restrict__ tot, unsigned long long* __restrict__ delta, size_t n )
{
    const size_t i = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= n ) return;
    const unsigned long long d = delta[ i ];
    if( d ) { tot[ i ] += d; delta[ i ] = 0; }
}

struct DImage
{
    int device = 0, width = 0, height = 0;
    int cycle = 0;
    u64 rval = 21943294ull;                     // lum_image_s_reset, scene.c:799
    int n_ranks = 1, rank = 0, tile = 4;
    unsigned long long* d_tot = nullptr;        // [ W*H*6 ]
    unsigned long long* d_delta = nullptr;      // [ W*H*6 ], the pass in progress
    unsigned int* d_cnt = nullptr;              // [ W*H ]
    unsigned long long* d_blk = nullptr;        // [ blocks ]
    unsigned long long* d_counts = nullptr;     // [ 2 ]: samples of the pass, this rank's part
    unsigned long long* h_counts = nullptr;     // pinned
    double* d_xy = nullptr; uint64_t xy_cap = 0;
    float* d_rgb = nullptr; uint64_t rgb_cap = 0;
    uint64_t pass_total = 0, pass_local = 0;
    bool in_pass = false;
    cudaStream_t stream = nullptr;

    size_t words() const { return ( size_t )width * height * PIX_STRIDE; }
    int blocks() const { return ( int )( ( ( size_t )width * height + 255 ) / 256 ); }

    ~DImage()
    {
        cudaSetDevice( device );
        cudaFree( d_tot ); cudaFree( d_delta ); cudaFree( d_cnt ); cudaFree( d_blk ); cudaFree( d_counts ); cudaFree( d_xy ); cudaFree( d_rgb );
        if( h_counts ) cudaFreeHost( h_counts );
        if( stream ) cudaStreamDestroy( stream );
    }
};

} // namespace acn
