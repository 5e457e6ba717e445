// acn_rtc.h — run-time compilation and loading of the scene-specialised kernels (host side).
//
// NVRTC (libnvrtc.so.12, dlopen'ed: the library itself loads on machines without it) compiles acn_kernels.cuh — the very
// sources of the ahead-of-time kernels, embedded in the library at build time (build/acn_embed.inc) — together with the
// header SpecGen wrote for the scene, for the device's own architecture (sm_100a on B200).  The cubin is cached in this
// process and on disk (ACN_CACHE_DIR, default ~/.cache/actinon_b200) under a hash of everything that went into it.  The
// module is loaded and launched through the driver API, whose entry points come from cudaGetDriverEntryPoint (no libcuda
// at link time).  Any failure here is reported and the tracer keeps its generic kernels: specialisation changes speed,
// never results.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <stdlib.h>
#include <unistd.h>
#include <time.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <memory>

namespace acn {

void set_error( const char* fmt, ... );

// the device sources, embedded at build time (tools/embed_src.py -> build/acn_embed.cpp)
extern const int ACN_EMBED_COUNT;
extern const char* const ACN_EMBED_NAMES[];
extern const char* const ACN_EMBED_SRC[];

struct NvrtcApi
{
    decltype( &nvrtcCreateProgram )      CreateProgram = nullptr;
    decltype( &nvrtcDestroyProgram )     DestroyProgram = nullptr;
    decltype( &nvrtcCompileProgram )     CompileProgram = nullptr;
    decltype( &nvrtcGetProgramLogSize )  GetProgramLogSize = nullptr;
    decltype( &nvrtcGetProgramLog )      GetProgramLog = nullptr;
    decltype( &nvrtcGetCUBINSize )       GetCUBINSize = nullptr;
    decltype( &nvrtcGetCUBIN )           GetCUBIN = nullptr;
    decltype( &nvrtcAddNameExpression )  AddNameExpression = nullptr;
    decltype( &nvrtcGetLoweredName )     GetLoweredName = nullptr;
    decltype( &nvrtcGetErrorString )     GetErrorString = nullptr;
    bool ok = false;

    static NvrtcApi& get()
    {
        static NvrtcApi api;
        static std::once_flag once;
        std::call_once( once, []()
        {
            void* h = nullptr;
            const char* names[] = { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so" };
            for( const char* n : names ) { h = dlopen( n, RTLD_NOW | RTLD_LOCAL ); if( h ) break; }
            if( !h ) return;
            bool all = true;
            auto sym = [ & ]( const char* n ) { void* s = dlsym( h, n ); if( !s ) all = false; return s; };
            api.CreateProgram     = ( decltype( api.CreateProgram ) )sym( "nvrtcCreateProgram" );
            api.DestroyProgram    = ( decltype( api.DestroyProgram ) )sym( "nvrtcDestroyProgram" );
            api.CompileProgram    = ( decltype( api.CompileProgram ) )sym( "nvrtcCompileProgram" );
            api.GetProgramLogSize = ( decltype( api.GetProgramLogSize ) )sym( "nvrtcGetProgramLogSize" );
            api.GetProgramLog     = ( decltype( api.GetProgramLog ) )sym( "nvrtcGetProgramLog" );
            api.GetCUBINSize      = ( decltype( api.GetCUBINSize ) )sym( "nvrtcGetCUBINSize" );
            api.GetCUBIN          = ( decltype( api.GetCUBIN ) )sym( "nvrtcGetCUBIN" );
            api.AddNameExpression = ( decltype( api.AddNameExpression ) )sym( "nvrtcAddNameExpression" );
            api.GetLoweredName    = ( decltype( api.GetLoweredName ) )sym( "nvrtcGetLoweredName" );
            api.GetErrorString    = ( decltype( api.GetErrorString ) )sym( "nvrtcGetErrorString" );
            api.ok = all;
        } );
        return api;
    }
};

struct DriverApi
{
    CUresult ( *ModuleLoadData )( CUmodule*, const void* ) = nullptr;
    CUresult ( *ModuleUnload )( CUmodule ) = nullptr;
    CUresult ( *ModuleGetFunction )( CUfunction*, CUmodule, const char* ) = nullptr;
    CUresult ( *FuncSetAttribute )( CUfunction, CUfunction_attribute, int ) = nullptr;
    CUresult ( *FuncGetAttribute )( int*, CUfunction_attribute, CUfunction ) = nullptr;
    CUresult ( *OccupancyMaxActiveBlocksPerMultiprocessor )( int*, CUfunction, int, size_t ) = nullptr;
    CUresult ( *LaunchKernel )( CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void** ) = nullptr;
    bool ok = false;

    static DriverApi& get()
    {
        static DriverApi api;
        static std::once_flag once;
        std::call_once( once, []()
        {
            bool all = true;
            auto ent = [ & ]( const char* n ) -> void*
            {
                void* f = nullptr;
                cudaDriverEntryPointQueryResult q;
                if( cudaGetDriverEntryPoint( n, &f, cudaEnableDefault, &q ) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f ) { all = false; f = nullptr; }
                return f;
            };
            api.ModuleLoadData    = ( decltype( api.ModuleLoadData ) )ent( "cuModuleLoadData" );
            api.ModuleUnload      = ( decltype( api.ModuleUnload ) )ent( "cuModuleUnload" );
            api.ModuleGetFunction = ( decltype( api.ModuleGetFunction ) )ent( "cuModuleGetFunction" );
            api.FuncSetAttribute  = ( decltype( api.FuncSetAttribute ) )ent( "cuFuncSetAttribute" );
            api.FuncGetAttribute  = ( decltype( api.FuncGetAttribute ) )ent( "cuFuncGetAttribute" );
            api.OccupancyMaxActiveBlocksPerMultiprocessor = ( decltype( api.OccupancyMaxActiveBlocksPerMultiprocessor ) )ent( "cuOccupancyMaxActiveBlocksPerMultiprocessor" );
            api.LaunchKernel      = ( decltype( api.LaunchKernel ) )ent( "cuLaunchKernel" );
            cudaGetLastError();
            api.ok = all;
        } );
        return api;
    }
};

enum { SPEC_K_PRIMARY = 0, SPEC_K_RAYS, SPEC_K_PATH, SPEC_K_DIRECT, SPEC_K_SHADE, SPEC_K_COUNT };

// a compiled module: the cubin and the lowered names of its five kernels
struct SpecBinary
{
    std::vector<char> cubin;
    std::string name[ SPEC_K_COUNT ];
    double compile_seconds = 0;
    bool from_disk = false;
};

// loaded on one device
struct SpecModule
{
    CUmodule mod = nullptr;
    CUfunction fn[ SPEC_K_COUNT ] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    int regs[ SPEC_K_COUNT ] = { 0, 0, 0, 0, 0 };
    ~SpecModule() { if( mod && DriverApi::get().ok ) DriverApi::get().ModuleUnload( mod ); }
};

static inline unsigned long long fnv1a( const void* data, size_t n, unsigned long long h = 0xcbf29ce484222325ull )
{
    const unsigned char* p = ( const unsigned char* )data;
    for( size_t i = 0; i < n; i++ ) { h ^= p[ i ]; h *= 0x100000001b3ull; }
    return h;
}

static inline std::string spec_cache_dir()
{
    const char* e = getenv( "ACN_CACHE_DIR" );
    std::string d;
    if( e && e[ 0 ] ) d = e;
    else { const char* h = getenv( "HOME" ); d = std::string( h && h[ 0 ] ? h : "/tmp" ) + "/.cache/actinon_b200"; }
    std::string cur;
    for( size_t i = 0; i <= d.size(); i++ )     // mkdir -p
        if( i == d.size() || ( d[ i ] == '/' && i > 0 ) ) { cur = d.substr( 0, i ); mkdir( cur.c_str(), 0755 ); }
    return d;
}

// kernel name expressions of the instantiation ( R, MARCH, SH )
static inline void spec_kernel_exprs( bool f64, bool march, bool sh, std::string out[ SPEC_K_COUNT ] )
{
    const std::string r = f64 ? "double" : "float", m = march ? "true" : "false", s = sh ? "true" : "false";
    out[ SPEC_K_PRIMARY ] = "acn::k_primary<" + r + ", " + m + ", " + s + ">";
    out[ SPEC_K_RAYS ]    = "acn::k_rays<" + r + ", " + m + ", " + s + ">";
    out[ SPEC_K_PATH ]    = "acn::k_path<" + r + ", " + m + ", " + s + ">";
    out[ SPEC_K_DIRECT ]  = "acn::k_direct<" + r + ", " + m + ", " + s + ">";
    out[ SPEC_K_SHADE ]   = "acn::k_shade<" + r + ", " + s + ">";
}

// compiles (or fetches from the caches) the kernels specialised by `gen`; arch e.g. "sm_100a".  nullptr + set_error on failure.
static inline std::shared_ptr<SpecBinary> spec_compile( const std::string& gen, bool f64, bool march, bool sh, const std::string& arch, const std::string& extra_opts )
{
    static std::mutex mu;
    static std::map<unsigned long long, std::shared_ptr<SpecBinary>> mem;
    std::string exprs[ SPEC_K_COUNT ];
    spec_kernel_exprs( f64, march, sh, exprs );
    unsigned long long key = fnv1a( gen.data(), gen.size() );
    for( int i = 0; i < ACN_EMBED_COUNT; i++ ) key = fnv1a( ACN_EMBED_SRC[ i ], strlen( ACN_EMBED_SRC[ i ] ), key );
    for( int i = 0; i < SPEC_K_COUNT; i++ ) key = fnv1a( exprs[ i ].data(), exprs[ i ].size(), key );
    key = fnv1a( arch.data(), arch.size(), key ); key = fnv1a( extra_opts.data(), extra_opts.size(), key );
    std::lock_guard<std::mutex> lock( mu );
    auto it = mem.find( key );
    if( it != mem.end() ) return it->second;

    auto bin = std::make_shared<SpecBinary>();
    char kname[ 64 ]; snprintf( kname, sizeof( kname ), "%016llx", key );
    const bool use_disk = !getenv( "ACN_NO_DISK_CACHE" );
    const std::string base = use_disk ? spec_cache_dir() + "/" + kname : std::string();
    if( use_disk )
    {   // <key>.cubin + <key>.names (one lowered name per line)
        FILE* fc = fopen( ( base + ".cubin" ).c_str(), "rb" ); FILE* fn = fopen( ( base + ".names" ).c_str(), "r" );
        if( fc && fn )
        {
            fseek( fc, 0, SEEK_END ); const long sz = ftell( fc ); fseek( fc, 0, SEEK_SET );
            bin->cubin.resize( sz > 0 ? ( size_t )sz : 0 );
            bool good = sz > 0 && fread( bin->cubin.data(), 1, ( size_t )sz, fc ) == ( size_t )sz;
            char line[ 1024 ];
            for( int i = 0; i < SPEC_K_COUNT && good; i++ )
            {
                if( !fgets( line, sizeof( line ), fn ) ) { good = false; break; }
                line[ strcspn( line, "\r\n" ) ] = 0; bin->name[ i ] = line;
            }
            if( good ) { fclose( fc ); fclose( fn ); bin->from_disk = true; mem[ key ] = bin; return bin; }
        }
        if( fc ) fclose( fc ); if( fn ) fclose( fn );
    }

    NvrtcApi& rtc = NvrtcApi::get();
    if( !rtc.ok ) { set_error( "scene specialisation: libnvrtc.so.12 not found" ); return nullptr; }
    std::vector<const char*> hsrc, hname;
    for( int i = 0; i < ACN_EMBED_COUNT; i++ ) { hsrc.push_back( ACN_EMBED_SRC[ i ] ); hname.push_back( ACN_EMBED_NAMES[ i ] ); }
    hsrc.push_back( gen.c_str() ); hname.push_back( "acn_spec_gen.h" );
    const std::string main_src = "#define ACN_SPEC 1\n#include \"acn_kernels.cuh\"\n";
    nvrtcProgram prog = nullptr;
    nvrtcResult r = rtc.CreateProgram( &prog, main_src.c_str(), "acn_spec.cu", ( int )hsrc.size(), hsrc.data(), hname.data() );
    if( r != NVRTC_SUCCESS ) { set_error( "nvrtcCreateProgram: %s", rtc.GetErrorString( r ) ); return nullptr; }
    for( int i = 0; i < SPEC_K_COUNT; i++ ) rtc.AddNameExpression( prog, exprs[ i ].c_str() );
    const std::string a = "--gpu-architecture=" + arch;
    std::vector<std::string> ov = { a, "-std=c++17", "-use_fast_math", "-lineinfo" };
    {   // extra options, blank separated
        size_t i = 0;
        while( i < extra_opts.size() )
        {
            size_t j = extra_opts.find( ' ', i ); if( j == std::string::npos ) j = extra_opts.size();
            if( j > i ) ov.push_back( extra_opts.substr( i, j - i ) );
            i = j + 1;
        }
    }
    std::vector<const char*> opts; for( auto& o : ov ) opts.push_back( o.c_str() );
    timespec t0, t1; clock_gettime( CLOCK_MONOTONIC, &t0 );
    r = rtc.CompileProgram( prog, ( int )opts.size(), opts.data() );
    clock_gettime( CLOCK_MONOTONIC, &t1 );
    bin->compile_seconds = ( t1.tv_sec - t0.tv_sec ) + 1e-9 * ( t1.tv_nsec - t0.tv_nsec );
    if( r != NVRTC_SUCCESS )
    {
        size_t ls = 0; rtc.GetProgramLogSize( prog, &ls );
        std::string log( ls, 0 ); if( ls ) rtc.GetProgramLog( prog, &log[ 0 ] );
        set_error( "scene specialisation failed to compile (%s): %.800s", rtc.GetErrorString( r ), log.c_str() );
        if( getenv( "ACN_VERBOSE" ) ) fprintf( stderr, "acn: NVRTC log:\n%s\n", log.c_str() );
        rtc.DestroyProgram( &prog );
        return nullptr;
    }
    bool good = true;
    for( int i = 0; i < SPEC_K_COUNT; i++ )
    {
        const char* ln = nullptr;
        if( rtc.GetLoweredName( prog, exprs[ i ].c_str(), &ln ) != NVRTC_SUCCESS || !ln ) { good = false; break; }
        bin->name[ i ] = ln;
    }
    size_t cs = 0;
    if( good && ( rtc.GetCUBINSize( prog, &cs ) != NVRTC_SUCCESS || cs == 0 ) ) good = false;
    if( good ) { bin->cubin.resize( cs ); good = rtc.GetCUBIN( prog, bin->cubin.data() ) == NVRTC_SUCCESS; }
    rtc.DestroyProgram( &prog );
    if( !good ) { set_error( "scene specialisation: no cubin / lowered names from NVRTC" ); return nullptr; }
    if( use_disk )
    {   // written under a temporary name and renamed: concurrent processes (one per GPU) may compile the same key
        char tmp[ 64 ]; snprintf( tmp, sizeof( tmp ), ".tmp%d", ( int )getpid() );
        FILE* fc = fopen( ( base + ".cubin" + tmp ).c_str(), "wb" );
        if( fc ) { fwrite( bin->cubin.data(), 1, bin->cubin.size(), fc ); fclose( fc ); rename( ( base + ".cubin" + tmp ).c_str(), ( base + ".cubin" ).c_str() ); }
        FILE* fn = fopen( ( base + ".names" + tmp ).c_str(), "w" );
        if( fn ) { for( int i = 0; i < SPEC_K_COUNT; i++ ) fprintf( fn, "%s\n", bin->name[ i ].c_str() ); fclose( fn ); rename( ( base + ".names" + tmp ).c_str(), ( base + ".names" ).c_str() ); }
    }
    mem[ key ] = bin;
    return bin;
}

// loads a binary into the current device's primary context
static inline std::shared_ptr<SpecModule> spec_load( const SpecBinary& bin )
{
    DriverApi& drv = DriverApi::get();
    if( !drv.ok ) { set_error( "scene specialisation: driver entry points unavailable" ); return nullptr; }
    auto m = std::make_shared<SpecModule>();
    CUresult r = drv.ModuleLoadData( &m->mod, bin.cubin.data() );
    if( r != CUDA_SUCCESS ) { m->mod = nullptr; set_error( "cuModuleLoadData failed (%d)", ( int )r ); return nullptr; }
    for( int i = 0; i < SPEC_K_COUNT; i++ )
    {
        r = drv.ModuleGetFunction( &m->fn[ i ], m->mod, bin.name[ i ].c_str() );
        if( r != CUDA_SUCCESS ) { set_error( "cuModuleGetFunction(%s) failed (%d)", bin.name[ i ].c_str(), ( int )r ); return nullptr; }
        drv.FuncGetAttribute( &m->regs[ i ], CU_FUNC_ATTRIBUTE_NUM_REGS, m->fn[ i ] );
    }
    return m;
}

} // namespace acn
