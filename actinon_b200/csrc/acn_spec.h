// acn_spec.h — run-time specialisation of the tracing kernels to ONE scene's structure (host side).
//
// The generic kernels (acn_kernels.cuh compiled by nvcc) INTERPRET the scene: scene_query walks child-record lists with a
// range stack, csg_eval decodes postfix program words, looks every leaf's kind up, appends crossings to an event list in
// shared memory and sorts it by repeated minimum search.  On the headline scenes three integer/logic instructions are
// issued per floating-point one and the lanes of a warp sit in different program words (profiles/r01_r1p_*).
//
// A scene's STRUCTURE — which elements the light and matter compounds hold, their kinds and flags, the CSG programs — is
// fixed for a whole render (and across the frames of an animation, which only move geometry).  SpecGen restates it as
// straight-line C++ over the same leaf functions (leaf_events, sphere_events, prim_hit, trans_commit: the arithmetic is
// shared with the generic path, not duplicated):
//   * spec_scene_query: the element lists unrolled, kinds and node indices literal, so every table access is an LDS with
//     an immediate offset and every kind dispatch is folded;
//   * csg_spec_<node>: pass 1 of the event sweep unrolled (sub-envelope gates become nested ifs), the crossings of each
//     variable in REGISTERS (no event list, no shared memory), the sweep = a minimum search over the live crossings per
//     round with a done-mask, the boolean function a truth-table lookup or a literal bit expression.
// Geometry stays in the staged tables, so all scenes of one structure share one compiled module.  The source is compiled
// with NVRTC for sm_100a (acn_spec.cpp), cached in memory and on disk by a hash of the generated text.
#pragma once

#include <string>
#include <vector>
#include <stdio.h>

namespace acn {

struct SpecGen
{
    const acn_flat_scene* fs = nullptr;
    const std::vector<int>* prog = nullptr;          // CsgBuilder::prog
    const std::vector<I4>*  prog_ref = nullptr;      // CsgBuilder::prog_ref
    std::string out;
    int n_elements = 0, n_programs = 0;

    static const int MAX_ELEMENTS = 64;              // unrolled element tests per scene query
    // Code size: the instruction caches hold 32 KB (L1.5) = 2 000 instructions; straight-line code that every ray walks
    // end to end streams from L2, and warps in different places thrash it (diamond_video: 8 000-instruction kernels ran
    // 2.3x SLOWER than the interpreter).  Long convex runs are therefore LOOPS over their member words, one loop per
    // segment of members of one kind (the kind stays literal); the quadric leaf code can be kept out of line.
    // Measured (B200, ms per pass 0, profiles/r02_spec_sweep.txt): diamond 12.6 unrolled / 10.6 with segments longer than 3
    // as loops; diamond_video frame 49: 15.9 / 12.8 (39.9 before any size control); quadrics out of line: no difference.
    int run_unroll = 3;                              // segments of more members than this become loops
    bool ool_squaroid = false;                       // squaroid leaves / elements through one out-of-line function

    void read_env()
    {
        const char* e = getenv( "ACN_SPEC_RUN_UNROLL" ); if( e && e[ 0 ] ) run_unroll = atoi( e );
        e = getenv( "ACN_SPEC_OOL_SQUAROID" ); if( e && e[ 0 ] ) ool_squaroid = e[ 0 ] != '0';
    }

    void p( const char* fmt, ... )
    {
        char buf[ 2048 ];
        va_list ap; va_start( ap, fmt ); vsnprintf( buf, sizeof( buf ), fmt, ap ); va_end( ap );
        out += buf;
    }

    int  kind( int n ) const { return fs->nodes[ n ].kind; }
    bool has_env( int n ) const { return fs->nodes[ n ].has_envelope != 0; }
    // nodes whose contents are known to stick out of their envelope (CullBounds, acn_tracer.cuh): a hit in front of the envelope
    // is possible, so the envelope is only a gate (objects.c:264) and must not be culled against the horizon
    const std::vector<char>* env_gate_only = nullptr;
    const char* env_test( int n ) const { return env_gate_only && ( *env_gate_only )[ n ] ? "envelope_hits( sv.env[ %d ], ray )" : "envelope_hits_before( sv.env[ %d ], ray, hor )"; }
    bool rough( int n ) const { return fs->nodes[ n ].surface_roughness > 0 && kind( n ) != ACN_KIND_COMPOUND; }
    static bool simple_leaf( int k ) { return k == ACN_KIND_PLANE || k == ACN_KIND_SPHERE || k == ACN_KIND_SQUAROID; }
    static int  max_crossings( int k ) { return k == ACN_KIND_PLANE ? 1 : 2; }

    // ---- can this program be restated?  plane / sphere / squaroid leaves only (distance-field leaves and groups of
    // coincident crossings stay with the generic evaluator of the full-featured kernels)
    bool program_ok( int root ) const
    {
        const I4 pr = ( *prog_ref )[ root ];
        if( pr.y <= 0 || pr.w > 32 ) return false;      // the done-mask of the register sweep holds 2 x 32 crossings
        for( int pc = pr.x; pc < pr.x + pr.y; pc++ )
        {
            const int ins = ( *prog )[ pc ], op = ins & 15, n = ins >> 4;
            if( op == CSG_MORE || op == CSG_XFORM || op == CSG_XEND ) return false;
            if( op == CSG_LEAF && !simple_leaf( kind( n ) ) ) return false;
            if( op == CSG_RUN ) { for( int m = 1; m <= n; m++ ) if( !simple_leaf( kind( ( *prog )[ pc + m ] >> 4 ) ) ) return false; pc += n; }
            if( op == CSG_ENV ) pc++;
        }
        return true;
    }

    // boolean function of a program as a literal expression over B(v) = bit v of `vars`
    std::string expression( const I4& pr ) const
    {
        std::vector<std::string> stk;
        int v = 0;
        char b[ 64 ];
        for( int pc = pr.x; pc < pr.x + pr.y; pc++ )
        {
            const int ins = ( *prog )[ pc ], op = ins & 15;
            if( op == CSG_LEAF )      { snprintf( b, sizeof( b ), "B(%d)", v++ ); stk.push_back( b ); }
            else if( op == CSG_RUN )  { snprintf( b, sizeof( b ), "B(%d)", v++ ); stk.push_back( b ); pc += ins >> 4; }
            else if( op == CSG_CLIP ) { snprintf( b, sizeof( b ), "B(%d)", v++ ); stk.back() = "(" + stk.back() + "&" + b + ")"; }
            else if( op == CSG_NEG )  stk.back() = "(" + stk.back() + "^1u)";
            else if( op == CSG_AND || op == CSG_OR )
            {
                const std::string r = stk.back(); stk.pop_back();
                stk.back() = "(" + stk.back() + ( op == CSG_AND ? "&" : "|" ) + r + ")";
            }
            else if( op == CSG_ENV ) pc++;
        }
        return stk.empty() ? std::string( "0u" ) : stk.back();
    }

    void emit_event_store( int v, int k, const char* t, const std::string& code )
    {   // crossing k of variable v, kept only when it lies before the caller's horizon
        p( "            if( %s <= t_far ) { t%d_%d = %s; k%d_%d = %s; }\n", t, k, v, t, k, v, code.c_str() );
    }

    // ---- one CSG program -> csg_spec_<root>
    void emit_program( int root )
    {
        const I4 pr = ( *prog_ref )[ root ];
        const int nv = pr.w;
        const char* VT = nv > 32 ? "unsigned long long" : "unsigned int";
        const char* DT = 2 * nv > 32 ? "unsigned long long" : "unsigned int";
        p( "// CSG program of node %d: %d words, %d variables%s\n", root, pr.y, nv, pr.z >= 0 ? ", truth table" : "" );
        p( "template <typename R, bool SH> __device__ __forceinline__ R csg_spec_%d( const SceneView<R, SH>& sv, const Ray<R>& ray, V3<R>* nor, HitCtx ctx, const R t_far )\n{\n", root );
        p( "    const R inf = Num<R>::inf();\n    %s vars = 0;\n", VT );
        // per variable: up to two crossings (t, code = leaf id | variable << 8 | crossing << 16)
        std::vector<int> ncross( nv, 2 );
        for( int v = 0; v < nv; v++ ) p( "    R t0_%d = inf, t1_%d = inf; int k0_%d = 0, k1_%d = 0;\n", v, v, v, v );
        // ---- pass 1
        int v = 0, depth = 0;
        std::vector<int> clip_var;       // CLIP variable of every open ENV
        for( int pc = pr.x; pc < pr.x + pr.y; pc++ )
        {
            const int ins = ( *prog )[ pc ], op = ins & 15, n = ins >> 4, id = pc - pr.x;
            if( op == CSG_LEAF )
            {
                const int k = kind( n );
                ncross[ v ] = max_crossings( k );
                p( "    {   // variable %d: leaf %d\n        int s0; R a0 = R( 0 ), a1 = R( 0 );\n", v, n );
                if( k == ACN_KIND_SQUAROID && ool_squaroid ) p( "        const int c = squaroid_events_ool( sv, %d, ray, &s0, &a0, &a1 );\n", n );
                else p( "        const int c = leaf_events<false>( sv, %d, %d, ray, &s0, &a0, &a1 );\n", k, n );
                p( "        vars |= ( %s )s0 << %d;\n", VT, v );
                p( "        if( c >= 1 && a0 <= t_far ) { t0_%d = a0; k0_%d = %d; }\n", v, v, id | ( v << 8 ) );
                if( ncross[ v ] == 2 ) p( "        if( c == 2 && a1 <= t_far ) { t1_%d = a1; k1_%d = %d; }\n", v, v, id | ( v << 8 ) | ( 1 << 16 ) );
                p( "    }\n" );
                v++;
            }
            else if( op == CSG_RUN )
            {
                p( "    {   // variable %d: convex run of %d members, inside-set = [ max entry, min exit ]\n", v, n );
                p( "        R lo = R( -1 ), hi = inf; int i0 = %d, i1 = %d; bool alive = true;\n", ( int )CSG_VIRTUAL, ( int )CSG_VIRTUAL );
                // consecutive members of one kind (and one sign) form a segment: a long segment is a LOOP over its member words
                // (they are in the staged program table) with the kind literal, a short one is unrolled
                for( int m0 = 1; m0 <= n; )
                {
                    const int w0 = ( *prog )[ pc + m0 ], k0 = kind( w0 >> 4 ), neg0 = ( w0 & 15 ) == CSG_MEMBER_NEG;
                    int m1 = m0;
                    while( m1 + 1 <= n && kind( ( *prog )[ pc + m1 + 1 ] >> 4 ) == k0 && ( ( ( *prog )[ pc + m1 + 1 ] & 15 ) == CSG_MEMBER_NEG ) == neg0 ) m1++;
                    const int len = m1 - m0 + 1;
                    if( len > run_unroll )
                    {
                        p( "        if( __any_sync( __activemask(), alive ) )\n        {\n" );
                        p( "            #pragma unroll 1\n            for( int m = %d; m <= %d; m++ )\n            {\n", m0, m1 );
                        p( "                const int node = sv.prog[ %d + m ] >> 4;\n", pc );
                        p( "                if( alive ) { int ms0; R a0 = R( 0 ), a1 = R( 0 ), mlo, mhi;\n" );
                        if( k0 == ACN_KIND_SQUAROID && ool_squaroid ) p( "                    const int mc = squaroid_events_ool( sv, node, ray, &ms0, &a0, &a1 );%s\n", neg0 ? " ms0 ^= 1;" : "" );
                        else p( "                    const int mc = leaf_events<false>( sv, %d, node, ray, &ms0, &a0, &a1 );%s\n", k0, neg0 ? " ms0 ^= 1;" : "" );
                        p( "                    member_interval( ms0, mc, a0, a1, &mlo, &mhi ); if( mlo > lo ) { lo = mlo; i0 = %d + m; } if( mhi < hi ) { hi = mhi; i1 = %d + m; }\n", id, id );
                        p( "                    if( !( lo < hi ) || lo > t_far ) { lo = inf; alive = false; } }\n" );
                        p( "                if( ( m & 7 ) == 0 && !__any_sync( __activemask(), alive ) ) break;\n            }\n        }\n" );
                    }
                    else
                    {
                        for( int m = m0; m <= m1; m++ )
                        {
                            const int node = ( *prog )[ pc + m ] >> 4;
                            p( "        if( alive ) { int ms0; R a0 = R( 0 ), a1 = R( 0 ), mlo, mhi; " );
                            if( k0 == ACN_KIND_SQUAROID && ool_squaroid ) p( "const int mc = squaroid_events_ool( sv, %d, ray, &ms0, &a0, &a1 );%s\n", node, neg0 ? " ms0 ^= 1;" : "" );
                            else p( "const int mc = leaf_events<false>( sv, %d, %d, ray, &ms0, &a0, &a1 );%s\n", k0, node, neg0 ? " ms0 ^= 1;" : "" );
                            p( "            member_interval( ms0, mc, a0, a1, &mlo, &mhi ); if( mlo > lo ) { lo = mlo; i0 = %d; } if( mhi < hi ) { hi = mhi; i1 = %d; }\n", id + m, id + m );
                            p( "            if( !( lo < hi ) || lo > t_far ) { lo = inf; alive = false; } }\n" );
                        }
                    }
                    m0 = m1 + 1;
                }
                p( "        if( lo < R( 0 ) ) { vars |= ( %s )1 << %d; if( hi <= t_far ) { t0_%d = hi; k0_%d = i1 | %d; } }\n", VT, v, v, v, v << 8 );
                p( "        else if( lo < hi ) { t0_%d = lo; k0_%d = i0 | %d; if( hi <= t_far ) { t1_%d = hi; k1_%d = i1 | %d; } }\n",
                   v, v, v << 8, v, v, ( v << 8 ) | ( 1 << 16 ) );
                p( "    }\n" );
                pc += n; v++;
            }
            else if( op == CSG_ENV )
            {
                const int w2 = ( *prog )[ ++pc ];
                const int vc = v + ( w2 >> 16 ) - 1;
                clip_var.push_back( vc );
                p( "    {   // envelope of node %d gates variables %d..%d; its own crossings are virtual (variable %d)\n", n, v, vc - 1, vc );
                p( "        int s0; R a0 = R( 0 ), a1 = R( 0 ); const R4<R> e = sv.env[ %d ];\n", n );
                p( "        const int c = sphere_events( xyz( e ), e.w, ray, &s0, &a0, &a1 );\n" );
                p( "        if( c > 0 ) {\n" );
                p( "        vars |= ( %s )s0 << %d;\n", VT, vc );
                p( "        if( a0 <= t_far ) { t0_%d = a0; k0_%d = %d; }\n", vc, vc, ( int )CSG_VIRTUAL | ( vc << 8 ) );
                p( "        if( c == 2 && a1 <= t_far ) { t1_%d = a1; k1_%d = %d; }\n", vc, vc, ( int )CSG_VIRTUAL | ( vc << 8 ) | ( 1 << 16 ) );
                depth++;
            }
            else if( op == CSG_CLIP )
            {
                p( "        }\n    }\n" );
                depth--; clip_var.pop_back(); v++;
            }
        }
        // ---- boolean function
        if( pr.z >= 0 ) p( "    #define F( x ) ( ( ( unsigned int )sv.prog[ %d + ( int )( ( x ) >> 5 ) ] >> ( ( unsigned int )( x ) & 31u ) ) & 1u )\n", pr.z );
        else
        {
            p( "    #define B( i ) ( ( unsigned int )( x_ >> ( i ) ) & 1u )\n" );
            p( "    auto F = [ & ]( %s x_ ) -> unsigned int { return %s; };\n", VT, expression( pr ).c_str() );
        }
        // ---- sweep: the first crossing, in order of t, at which the function flips at a real (non-virtual) crossing
        p( "    unsigned int s = F( vars );\n    %s done = 0;\n", DT );
        p( "    for( ;; )\n    {\n        R tmin = inf; int code = -1;\n" );
        for( int u = 0; u < nv; u++ )
            for( int k = 0; k < ncross[ u ]; k++ )
                p( "        if( !( done & ( ( %s )1 << %d ) ) && t%d_%d < tmin ) { tmin = t%d_%d; code = k%d_%d; }\n", DT, 2 * u + k, k, u, k, u, k, u );
        p( "        if( code < 0 ) break;\n" );
        p( "        const int var = ( code >> 8 ) & 255;\n" );
        p( "        done |= ( %s )1 << ( 2 * var + ( code >> 16 ) );\n", DT );
        p( "        vars ^= ( %s )1 << var;\n", VT );
        p( "        const unsigned int s2 = F( vars );\n" );
        p( "        if( s2 != s && ( code & 255 ) != %d ) return csg_report_hit<false>( sv, %d, %d, ray, tmin, code & 255, nor, ctx );\n", ( int )CSG_VIRTUAL, root, pr.x );
        p( "        s = s2;\n    }\n" );
        p( pr.z >= 0 ? "    #undef F\n" : "    #undef B\n" );
        p( "    return inf;\n}\n\n" );
        n_programs++;
    }

    // ---- element tests of one root compound, nested compounds opened up (compound.c:215-299)
    bool count_elements( int c, int guard )
    {
        if( guard > 16 ) return false;
        const acn_flat_node& cn = fs->nodes[ c ];
        for( int i = 0; i < cn.child1; i++ )
        {
            const int e = fs->children[ cn.child0 + i ];
            if( kind( e ) == ACN_KIND_COMPOUND ) { if( !count_elements( e, guard + 1 ) ) return false; }
            else n_elements++;
        }
        return n_elements <= MAX_ELEMENTS;
    }

    void emit_element( int e, bool nested, const char* ind )
    {
        const int k = kind( e );
        p( "%s{   // node %d\n", ind, e );
        p( "%s    const R hor = ( want_trans ? r_min( r_min( min_a + sv.eps, %s ), far0 ) : r_min( min_a, far0 ) ) + slack;\n", ind, nested ? "el_a" : "inf" );
        if( has_env( e ) ) { p( "%s    if( !found && ", ind ); p( env_test( e ), e ); p( " )\n" ); }
        else               p( "%s    if( !found )\n", ind );
        p( "%s    {\n%s        V3<R> n = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); R a;\n", ind, ind );
        bool roughen_here = rough( e );
        if( k == ACN_KIND_SQUAROID && ool_squaroid ) p( "%s        { const HitN<R> h = squaroid_hit_ool( sv, %d, ray, want_trans ); a = h.a; n = h.n; }\n", ind, e );
        else if( simple_leaf( k ) ) p( "%s        a = prim_hit( sv, %d, %d, ray, want_trans ? &n : ( V3<R>* )nullptr );\n", ind, k, e );
        else if( k >= ACN_KIND_PAIR_INSIDE && program_ok( e ) ) p( "%s        a = csg_spec_%d<R, SH>( sv, ray, want_trans ? &n : ( V3<R>* )nullptr, ctx, hor );\n", ind, e );
        else { p( "%s        a = elem_hit<R, MARCH>( sv, sv.link[ %d ], %d, ray, want_trans ? &n : ( V3<R>* )nullptr, ctx, cm, hor );\n", ind, e, e ); roughen_here = false; }
        if( roughen_here ) p( "%s        if( want_trans && a < inf ) roughen( sv, %d, ray, a, &n, ctx );\n", ind, e );
        p( "%s        if( !want_trans ) { if( a < min_a ) { min_a = a; if( a <= t_far ) found = true; } }\n", ind );
        if( nested ) p( "%s        else if( a < el_a ) { el_a = a; el_n = n; el_obj = %d; }\n", ind, e );
        else         p( "%s        else trans_commit( sv, ray, a, n, %d, &min_a, &tl );\n", ind, e );
        p( "%s    }\n%s}\n", ind, ind );
    }

    // children of a nested compound: all levels below a top-level compound element share its el_* (one element of the root)
    void emit_nested( int c, const char* ind )
    {
        const acn_flat_node& cn = fs->nodes[ c ];
        for( int i = 0; i < cn.child1; i++ )
        {
            const int e = fs->children[ cn.child0 + i ];
            if( kind( e ) == ACN_KIND_COMPOUND )
            {
                p( "%s{   // compound %d\n", ind, e );
                p( "%s    const R hor = ( want_trans ? r_min( r_min( min_a + sv.eps, el_a ), far0 ) : r_min( min_a, far0 ) ) + slack;\n", ind );
                if( has_env( e ) ) { p( "%s    if( !found && ", ind ); p( env_test( e ), e ); p( " )\n" ); }
                else               p( "%s    if( !found )\n", ind );
                p( "%s    {\n", ind );
                std::string in2 = std::string( ind ) + "        ";
                emit_nested( e, in2.c_str() );
                p( "%s    }\n%s}\n", ind, ind );
            }
            else emit_element( e, true, ind );
        }
    }

    void emit_root( int root, const char* flag )
    {
        const acn_flat_node& rn = fs->nodes[ root ];
        p( "    {   // root compound %d\n", root );
        p( "        bool act = !found && ( flags & %s ) != 0;\n", flag );
        p( "        const R far0 = r_min( t_far, best );\n" );
        if( has_env( root ) ) p( "        if( act && !envelope_hits_before( sv.env[ %d ], ray, far0 + slack ) ) act = false;\n", root );
        p( "        R min_a = inf;\n        Trans<R> tl; tl.exit_obj = tl.enter_obj = -1; tl.exit_nor = v3<R>( R( 0 ), R( 0 ), R( 0 ) );\n" );
        p( "        if( act )\n        {\n" );
        for( int i = 0; i < rn.child1; i++ )
        {
            const int e = fs->children[ rn.child0 + i ];
            if( kind( e ) == ACN_KIND_COMPOUND )
            {
                p( "            {   // nested compound %d: one element of the root, its hit = the closest hit of its contents\n", e );
                p( "                const R hor = ( want_trans ? r_min( min_a + sv.eps, far0 ) : r_min( min_a, far0 ) ) + slack;\n" );
                if( has_env( e ) ) { p( "                if( !found && " ); p( env_test( e ), e ); p( " )\n" ); }
                else               p( "                if( !found )\n" );
                p( "                {\n                    R el_a = inf; V3<R> el_n = v3<R>( R( 0 ), R( 0 ), R( 0 ) ); int el_obj = -1;\n" );
                emit_nested( e, "                    " );
                p( "                    if( want_trans ) trans_commit( sv, ray, el_a, el_n, el_obj, &min_a, &tl );\n" );
                p( "                }\n            }\n" );
            }
            else emit_element( e, false, "            " );
        }
        p( "        }\n" );
        p( "        if( min_a < best ) { best = min_a; if( want_trans ) *trans = tl; }\n    }\n" );
    }

    // returns false when the scene does not qualify (too many elements): the generic kernels are used then
    bool generate()
    {
        out.clear(); n_elements = 0; n_programs = 0;
        read_env();
        if( !count_elements( fs->light_root, 0 ) || !count_elements( fs->matter_root, 0 ) ) return false;
        p( "// generated by SpecGen (acn_spec.h): structure of one scene, %d nodes, %d element tests\n", fs->n_nodes, n_elements );
        p( "#define ACN_SPEC_SCENE 1\n\n" );
        // programs of every element (any nesting level)
        std::vector<int> todo;
        std::function<void( int )> walk = [ & ]( int c )
        {
            const acn_flat_node& cn = fs->nodes[ c ];
            for( int i = 0; i < cn.child1; i++ )
            {
                const int e = fs->children[ cn.child0 + i ];
                if( kind( e ) == ACN_KIND_COMPOUND ) walk( e );
                else if( kind( e ) >= ACN_KIND_PAIR_INSIDE && program_ok( e ) ) todo.push_back( e );
            }
        };
        walk( fs->light_root ); walk( fs->matter_root );
        std::sort( todo.begin(), todo.end() ); todo.erase( std::unique( todo.begin(), todo.end() ), todo.end() );
        for( int e : todo ) emit_program( e );
        p( "// scene_s_trans_hit / compound_s_ray_trans_hit / compound_s_ray_hit (scene.c:362-382, compound.c:215-299) with the element\n"
           "// lists of this scene unrolled; semantics as scene_query (acn_isect.cuh)\n" );
        p( "template <typename R, int MARCH, bool SH> __device__ __forceinline__ R spec_scene_query( const SceneView<R, SH>& sv, const Ray<R>& ray, const int flags, const R t_far,\n"
           "                                                                                 Trans<R>* trans, HitCtx ctx, const CsgMem<R>& cm )\n{\n" );
        p( "    const R inf = Num<R>::inf();\n    const bool want_trans = ( flags & Q_TRANS ) != 0;\n    const R slack = R( 2 ) * sv.eps;\n" );
        p( "    R best = inf;\n    bool found = false;\n" );
        emit_root( fs->light_root, "Q_LIGHT" );
        emit_root( fs->matter_root, "Q_MATTER" );
        p( "    return best;\n}\n" );
        return true;
    }
};

} // namespace acn
