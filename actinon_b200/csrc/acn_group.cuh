// acn_group.cuh — one image on several GPUs of one box, inside ONE process (the reference is a single C program: this is
// what its maintainer links against to get more than one device behind lum_machine_s_run).
//
// One worker thread, one tracer (the flat scene replicated) and one device image per GPU.  Pixels are dealt to the ranks in
// 4x4 tiles along a Morton curve (acn_dimage.cuh).  Per pass every rank builds the same selection on its replica of the
// totals, traces its own pixels and accumulates them into its pass delta; then — the one exchange step of the path — every
// rank adds, for each pixel, the delta of the pixel's OWNER to its totals, reading it straight out of the owner's memory
// over NVLink / NVSwitch (peer loads in the commit kernel: no staging buffer, no collective library, no atomics; the deltas
// have disjoint support and are integers, so the result does not depend on any order).  Two host barriers per pass
// bracket the peer reads.  Without peer access between some pair of devices the deltas travel by cudaMemcpyPeer instead.
#pragma once

#include <thread>
#include <mutex>
#include <condition_variable>
#include <vector>

#include "acn_dimage.cuh"

namespace acn {

enum { GROUP_MAX = 16 };
struct DeltaPtrs { const unsigned long long* p[ GROUP_MAX ]; };

// totals[ pixel ] += delta of the pixel's owner (peer memory)
__global__ void k_img_commit_owner( unsigned long long* __restrict__ tot, DeltaPtrs deltas, int W, int H, int tile, int n_ranks )
{
    const size_t p = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( p >= ( size_t )W * H ) return;
    const int x = ( int )( p % W ), y = ( int )( p / W );
    const unsigned long long* d = deltas.p[ pixel_owner( x, y, tile, n_ranks ) ] + p * PIX_STRIDE;
    // 48 bytes per pixel as three 16-byte peer loads
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>( d ), b = *reinterpret_cast<const ulonglong2*>( d + 2 ), c = *reinterpret_cast<const ulonglong2*>( d + 4 );
    if( ( a.x | a.y | b.x | b.y | c.x | c.y ) == 0 ) return;
    unsigned long long* t = tot + p * PIX_STRIDE;
    t[ 0 ] += a.x; t[ 1 ] += a.y; t[ 2 ] += b.x; t[ 3 ] += b.y; t[ 4 ] += c.x; t[ 5 ] += c.y;
}

__global__ void k_img_add( unsigned long long* __restrict__ tot, const unsigned long long* __restrict__ src, size_t n )
{
    const size_t i = ( size_t )blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) { const unsigned long long v = src[ i ]; if( v ) tot[ i ] += v; }
}

struct HostBarrier
{
    std::mutex mu; std::condition_variable cv; int n = 1, waiting = 0; unsigned long long gen = 0;
    void wait()
    {
        std::unique_lock<std::mutex> lk( mu );
        const unsigned long long g = gen;
        if( ++waiting == n ) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait( lk, [ & ] { return gen != g; } );
    }
};

struct Group
{
    int n = 0;
    std::vector<int> devices;
    std::vector<acn_tracer*> tracers;
    std::vector<acn_dimage*> images;
    acn_flat_params prm;
    bool peer = true;
    std::vector<unsigned long long*> stage;       // per rank, only without peer access
    // command channel
    std::vector<std::thread> threads;
    std::mutex mu; std::condition_variable cv_cmd, cv_done;
    unsigned long long seq = 0; int cmd = 0, done = 0;      // cmd: 1 = pass, 2 = quit
    HostBarrier bar;
    uint64_t index_base = 0;
    const volatile int* cancel = nullptr;
    std::vector<int> rc; std::vector<acn_stats> stats; std::vector<uint64_t> n_local, n_total;
    std::vector<std::string> err;
    bool failed = false;

    DImage* img( int r ) { return reinterpret_cast<DImage*>( images[ r ] ); }

    void worker( int r )
    {
        cudaSetDevice( devices[ r ] );
        unsigned long long seen = 0;
        for( ;; )
        {
            int c;
            { std::unique_lock<std::mutex> lk( mu ); cv_cmd.wait( lk, [ & ] { return seq != seen; } ); seen = seq; c = cmd; }
            if( c == 2 ) return;
            pass( r );
            { std::lock_guard<std::mutex> lk( mu ); if( ++done == n ) cv_done.notify_all(); }
        }
    }

    void pass( int r )
    {
        DImage* di = img( r );
        rc[ r ] = acn_dimage_render_pass( images[ r ], tracers[ r ], &prm, index_base, &n_local[ r ], &n_total[ r ], cancel, &stats[ r ] );
        if( rc[ r ] ) err[ r ] = acn_last_error();
        bar.wait();                                                      // every delta is complete (or its rank failed)
        bool any_fail = false;
        for( int k = 0; k < n; k++ ) any_fail = any_fail || rc[ k ] != 0;
        const bool active = !any_fail && di->in_pass;
        if( active )
        {
            const int W = di->width, H = di->height;
            if( peer )
            {
                DeltaPtrs dp;
                for( int k = 0; k < GROUP_MAX; k++ ) dp.p[ k ] = k < n ? img( k )->d_delta : nullptr;
                k_img_commit_owner<<< grid_for( ( size_t )W * H, 256 ), 256, 0, di->stream >>>( di->d_tot, dp, W, H, di->tile, n );
            }
            else
            {
                for( int k = 0; k < n; k++ )
                {
                    const unsigned long long* src = img( k )->d_delta;
                    if( k != r ) { cudaMemcpyPeerAsync( stage[ r ], devices[ r ], src, devices[ k ], di->words() * 8, di->stream ); src = stage[ r ]; }
                    k_img_add<<< grid_for( di->words(), 256 ), 256, 0, di->stream >>>( di->d_tot, src, di->words() );
                }
            }
            if( cudaStreamSynchronize( di->stream ) != cudaSuccess ) { rc[ r ] = ACN_ERR_CUDA; err[ r ] = "group commit failed"; }
        }
        bar.wait();                                                      // every rank has read every delta
        if( di->in_pass )
        {
            cudaMemsetAsync( di->d_delta, 0, di->words() * 8, di->stream );
            cudaStreamSynchronize( di->stream );
            if( active )
            {
                if( di->cycle > 0 ) di->rval = lcg00_skip( di->rval, 2ull * di->pass_total );
                di->cycle++;
            }
            di->in_pass = false;                                         // a failed or cancelled pass is discarded on every rank (scene.c:1143-1153)
        }
    }

    ~Group()
    {
        if( !threads.empty() )
        {
            { std::lock_guard<std::mutex> lk( mu ); cmd = 2; seq++; }
            cv_cmd.notify_all();
            for( auto& t : threads ) t.join();
        }
        for( size_t r = 0; r < stage.size(); r++ ) if( stage[ r ] ) { cudaSetDevice( devices[ r ] ); cudaFree( stage[ r ] ); }
        for( auto d : images ) acn_dimage_destroy( d );
        for( auto t : tracers ) acn_tracer_destroy( t );
    }
};

} // namespace acn
