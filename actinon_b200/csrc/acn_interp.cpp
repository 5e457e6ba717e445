// acn_interp.cpp — front-end for the Actinon scene language (.acn).
//
// Host-side, runs once per script.  It evaluates a script with the semantics of the reference
// interpreter (src/interpreter.c: tokens :207-511, expression scheme :1412-1730, statements and
// flow control :1734-1850, closures :1896-1923, built-ins :1945-2015; src/closures.c) on top of
// the scene model in acn_model.h, and RECORDS every scene.create_image( file ) call instead of
// rendering it — rendering is the tracer's job.
//
// The evaluation scheme is the reference's hand-rolled one, not a precedence table:
//   * postfix  ( )  [ ]  .  bind tightest;
//   * unary + - ! (&) (|) (:) (@) apply to the postfix-complete operand;
//   * * / % and the comparisons chain left to right;
//   * + -  and  & | ^  evaluate their whole right-hand side first (so a & b & c = a & (b & c));
//   * ':' builds / extends lists left to right;
//   * '/' multiplies by the inverse (always floating point);
//   * values are cloned on def / assignment / push, a bare name yields a reference.
#include "acn_model.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdexcept>
#include <sys/stat.h>

namespace acnh {

namespace {

enum Tk
{
    T_END = 0, T_INT, T_NUM, T_STR, T_BOOL, T_NAME, T_TYPE, T_BLOCK,
    T_COMMA, T_SEMI, T_LPAR, T_RPAR, T_LBRK, T_RBRK, T_DEF, T_FSIG, T_DYNARR,
    // operators
    T_DOT, T_QUERY, T_DQUERY, T_MUL, T_DIV, T_MOD, T_ADD, T_SUB,
    T_ASSIGN, T_MUL_ASSIGN, T_ADD_ASSIGN, T_SUB_ASSIGN, T_DIV_ASSIGN, T_MOD_ASSIGN,
    T_EQUAL, T_SMALLER, T_UNEQUAL, T_SMALLER_EQUAL, T_LARGER, T_LARGER_EQUAL,
    T_NOT, T_AND, T_OR, T_XOR, T_CAT, T_INSIDE_CPS, T_OUTSIDE_CPS, T_COMPOUND, T_ENVELOPE,
    // flow
    T_IF, T_WHILE, T_ELSE, T_FOR, T_IN
};

bool is_operator( int t ) { return t >= T_DOT && t <= T_ENVELOPE; }
bool is_assign_op( int t ) { return t >= T_ASSIGN && t <= T_MOD_ASSIGN; }

struct Code;
struct Token
{
    int kind = T_END;
    long long i = 0;
    double f = 0;
    std::string s;
    std::shared_ptr<Code> block;
    size_t jump = 0;          // if / while / for / else: index of the matching else or ';'
    int file = 0; size_t pos = 0;
};

struct Code { std::vector<Token> tok; std::vector<std::string> files; };

struct Error : std::runtime_error { Error( const std::string& m ) : std::runtime_error( m ) {} };

// ---------------------------------------------------------------------------------------------
// tokenizer (interpreter.c:207-511)
// ---------------------------------------------------------------------------------------------
struct Lexer
{
    std::string src, file;
    size_t p = 0;
    int file_index = 0;

    [[noreturn]] void fail( const std::string& msg ) const
    {
        size_t line = 1;
        for( size_t i = 0; i < p && i < src.size(); i++ ) if( src[ i ] == '\n' ) line++;
        throw Error( file + ":" + std::to_string( line ) + ": " + msg );
    }

    void skip_space()
    {
        for( ;; )
        {
            while( p < src.size() && ( src[ p ] == ' ' || src[ p ] == '\t' || src[ p ] == '\n' || src[ p ] == '\r' ) ) p++;
            if( p + 1 < src.size() && src[ p ] == '/' && src[ p + 1 ] == '/' ) { while( p < src.size() && src[ p ] != '\n' ) p++; continue; }
            if( p + 1 < src.size() && src[ p ] == '/' && src[ p + 1 ] == '*' )
            {
                p += 2;
                while( p + 1 < src.size() && !( src[ p ] == '*' && src[ p + 1 ] == '/' ) ) p++;
                p = p + 2 <= src.size() ? p + 2 : src.size();
                continue;
            }
            break;
        }
    }

    bool eos() const { return p >= src.size(); }
    bool at( const char* s ) const { return src.compare( p, strlen( s ), s ) == 0; }
    bool take( const char* s ) { if( at( s ) ) { p += strlen( s ); return true; } return false; }
};

std::string read_file( const std::string& path )
{
    FILE* f = fopen( path.c_str(), "rb" );
    if( !f ) throw Error( "cannot open '" + path + "'" );
    std::string s;
    char buf[ 65536 ];
    size_t n;
    while( ( n = fread( buf, 1, sizeof( buf ), f ) ) > 0 ) s.append( buf, n );
    fclose( f );
    return s;
}

void parse_into( Code& code, Lexer& lx );

std::shared_ptr<Code> parse_block( Lexer& lx, const std::vector<std::string>& files )
{
    std::shared_ptr<Code> c( new Code() );
    c->files = files;
    parse_into( *c, lx );
    return c;
}

void parse_into( Code& code, Lexer& lx )
{
    std::vector<size_t> jmp;      // pending if/while/for/else tokens awaiting their target
    lx.skip_space();
    while( !lx.eos() )
    {
        Token t;
        t.file = lx.file_index; t.pos = lx.p;
        const char c = lx.src[ lx.p ];
        bool push = true;

        if( c >= '0' && c <= '9' )                                   // number literal (:247-281)
        {
            unsigned long long vi = 0;
            while( !lx.eos() && lx.src[ lx.p ] >= '0' && lx.src[ lx.p ] <= '9' ) { vi = vi * 10 + ( unsigned )( lx.src[ lx.p ] - '0' ); lx.p++; }
            bool is_int = true;
            double vf = 0;
            long long vx = 0;
            if( !lx.eos() && lx.src[ lx.p ] == '.' )
            {
                lx.p++;
                is_int = false;
                double f = 0.1;                                      // digit-by-digit accumulation like the reference
                while( !lx.eos() && lx.src[ lx.p ] >= '0' && lx.src[ lx.p ] <= '9' ) { vf += f * ( lx.src[ lx.p ] - '0' ); f *= 0.1; lx.p++; }
            }
            if( !lx.eos() && ( lx.src[ lx.p ] == 'e' || lx.src[ lx.p ] == 'E' ) )
            {
                lx.p++;
                is_int = false;
                bool negx = false;
                if( !lx.eos() && ( lx.src[ lx.p ] == '+' || lx.src[ lx.p ] == '-' ) ) { negx = lx.src[ lx.p ] == '-'; lx.p++; }
                while( !lx.eos() && lx.src[ lx.p ] >= '0' && lx.src[ lx.p ] <= '9' ) { vx = vx * 10 + ( lx.src[ lx.p ] - '0' ); lx.p++; }
                if( negx ) vx = -vx;
            }
            if( is_int ) { t.kind = T_INT; t.i = ( long long )vi; }
            else { t.kind = T_NUM; double v = ( double )vi + vf; v *= pow( 10.0, ( double )vx ); t.f = v; }
        }
        else if( c == '"' )                                          // string literal (:282-305)
        {
            lx.p++;
            t.kind = T_STR;
            for( ;; )
            {
                if( lx.eos() ) lx.fail( "Stream ends in string literal" );
                char ch = lx.src[ lx.p++ ];
                if( ch == '"' ) break;
                if( ch == '\\' && !lx.eos() )
                {
                    char e = lx.src[ lx.p ];
                    if( e == '"' ) { t.s.push_back( '"' ); lx.p++; }
                    else if( e == 'n' ) { t.s.push_back( '\n' ); lx.p++; }
                    else if( e == 'r' ) { t.s.push_back( '\r' ); lx.p++; }
                    else if( e == 't' ) { t.s.push_back( '\t' ); lx.p++; }
                    else if( e == '0' ) { t.s.push_back( '\0' ); lx.p++; }
                    else if( e == '\\' ) { t.s.push_back( '\\' ); lx.p++; }
                    else t.s.push_back( '\\' );
                }
                else t.s.push_back( ch );
            }
        }
        else if( ( c >= 'A' && c <= 'Z' ) || ( c >= 'a' && c <= 'z' ) || c == '_' )     // names and keywords (:306-384)
        {
            std::string name;
            while( !lx.eos() )
            {
                char ch = lx.src[ lx.p ];
                if( ( ch >= 'A' && ch <= 'Z' ) || ( ch >= 'a' && ch <= 'z' ) || ch == '_' || ( ch >= '0' && ch <= '9' ) ) { name.push_back( ch ); lx.p++; }
                else break;
            }
            if( name == "true" ) { t.kind = T_BOOL; t.i = 1; }
            else if( name == "false" ) { t.kind = T_BOOL; t.i = 0; }
            else if( name == "AND" ) t.kind = T_AND;
            else if( name == "OR" ) t.kind = T_OR;
            else if( name == "XOR" ) t.kind = T_XOR;
            else if( name == "NOT" ) t.kind = T_NOT;
            else if( name == "CAT" ) t.kind = T_CAT;
            else if( name == "def" ) t.kind = T_DEF;
            else if( name == "if" || name == "while" || name == "for" )
            {
                t.kind = name == "if" ? T_IF : name == "while" ? T_WHILE : T_FOR;
                jmp.push_back( code.tok.size() );
            }
            else if( name == "in" ) t.kind = T_IN;
            else if( name == "else" )
            {
                if( jmp.empty() ) lx.fail( "'else' without 'if'" );
                code.tok[ jmp.back() ].jump = code.tok.size();
                jmp.pop_back();
                t.kind = T_ELSE;
                jmp.push_back( code.tok.size() );
            }
            else if( name == "bool" || name == "int" || name == "float" || name == "num" || name == "string" || name == "map" ||
                     name == "list" || name == "object" || name == "v3d" || name == "func" ) { t.kind = T_TYPE; t.s = name; }
            else { t.kind = T_NAME; t.s = name; }
        }
        else if( strchr( "!?.=+-*/%><&|^:", c ) )                     // operators (:386-420)
        {
            lx.p++;
            auto eq = [ & ]() { if( !lx.eos() && lx.src[ lx.p ] == '=' ) { lx.p++; return true; } return false; };
            switch( c )
            {
                case '!': t.kind = T_NOT; break;
                case '?': if( !lx.eos() && lx.src[ lx.p ] == '?' ) { lx.p++; t.kind = T_DQUERY; } else t.kind = T_QUERY; break;
                case '.': t.kind = T_DOT; break;
                case '=': t.kind = eq() ? T_EQUAL : T_ASSIGN; break;
                case '+': t.kind = eq() ? T_ADD_ASSIGN : T_ADD; break;
                case '-': t.kind = eq() ? T_SUB_ASSIGN : T_SUB; break;
                case '*': t.kind = eq() ? T_MUL_ASSIGN : T_MUL; break;
                case '/': t.kind = eq() ? T_DIV_ASSIGN : T_DIV; break;
                case '%': t.kind = eq() ? T_MOD_ASSIGN : T_MOD; break;
                case '<':
                    if( eq() ) t.kind = T_SMALLER_EQUAL;
                    else if( !lx.eos() && lx.src[ lx.p ] == '>' ) { lx.p++; t.kind = T_UNEQUAL; }
                    else if( !lx.eos() && lx.src[ lx.p ] == '-' ) { lx.p++; t.kind = T_FSIG; }
                    else t.kind = T_SMALLER;
                    break;
                case '>': t.kind = eq() ? T_LARGER_EQUAL : T_LARGER; break;
                case '&': t.kind = T_AND; break;
                case '|': t.kind = T_OR; break;
                case '^': t.kind = T_XOR; break;
                case ':': t.kind = T_CAT; break;
            }
        }
        else if( strchr( ";,()[]", c ) )                              // controls (:422-462)
        {
            lx.p++;
            switch( c )
            {
                case ';':
                    if( !jmp.empty() ) { code.tok[ jmp.back() ].jump = code.tok.size(); jmp.pop_back(); }
                    if( !jmp.empty() ) lx.fail( "Trailing jump address at end of statement." );
                    t.kind = T_SEMI;
                    break;
                case ',': t.kind = T_COMMA; break;
                case '(':
                    if( lx.take( "&)" ) ) t.kind = T_INSIDE_CPS;
                    else if( lx.take( "|)" ) ) t.kind = T_OUTSIDE_CPS;
                    else if( lx.take( ":)" ) ) t.kind = T_COMPOUND;
                    else if( lx.take( "@)" ) ) t.kind = T_ENVELOPE;
                    else t.kind = T_LPAR;
                    break;
                case ')': t.kind = T_RPAR; break;
                case '[': if( !lx.eos() && lx.src[ lx.p ] == ']' ) { lx.p++; t.kind = T_DYNARR; } else t.kind = T_LBRK; break;
                case ']': t.kind = T_RBRK; break;
            }
        }
        else if( c == '{' )                                          // nested block = one data token (:463-469)
        {
            lx.p++;
            t.kind = T_BLOCK;
            t.block = parse_block( lx, code.files );
            lx.skip_space();
            if( lx.eos() || lx.src[ lx.p ] != '}' ) lx.fail( "'}' expected" );
            lx.p++;
        }
        else if( c == '}' ) break;                                   // end of block, not consumed (:470-473)
        else if( lx.take( "#parse" ) )                               // textual include (:474-498)
        {
            lx.skip_space();
            if( lx.eos() || lx.src[ lx.p ] != '"' ) lx.fail( "File name expected." );
            lx.p++;
            std::string fn;
            while( !lx.eos() && lx.src[ lx.p ] != '"' ) fn.push_back( lx.src[ lx.p++ ] );
            if( !lx.eos() ) lx.p++;
            if( fn.empty() ) lx.fail( "File name expected." );
            if( fn[ 0 ] != '/' )
            {
                size_t idx = lx.file.rfind( '/' );
                if( idx != std::string::npos ) fn = lx.file.substr( 0, idx ) + "/" + fn;
            }
            Lexer inc;
            inc.src = read_file( fn ); inc.file = fn;
            code.files.push_back( fn );
            inc.file_index = ( int )code.files.size() - 1;
            parse_into( code, inc );
            push = false;
        }
        else if( lx.take( "#source_file_name" ) ) { t.kind = T_STR; t.s = lx.file; }
        else lx.fail( "Syntax error." );

        if( push ) code.tok.push_back( t );
        lx.skip_space();
    }
}

// ---------------------------------------------------------------------------------------------
// runtime
// ---------------------------------------------------------------------------------------------
} // namespace

struct Frame
{
    std::vector<std::pair<std::string, VP>> vars;
    std::shared_ptr<Frame> external;

    VP* get_local( const std::string& k ) { for( auto& kv : vars ) if( kv.first == k ) return &kv.second; return nullptr; }
    VP* get( const std::string& k ) { for( Frame* f = this; f; f = f->external.get() ) if( VP* p = f->get_local( k ) ) return p; return nullptr; }
    VP* set( const std::string& k, VP v ) { if( VP* p = get_local( k ) ) { *p = v; return p; } vars.push_back( std::make_pair( k, v ) ); return &vars.back().second; }
};

struct Closure
{
    std::shared_ptr<Code> code;
    std::vector<std::string> params;
    std::shared_ptr<Frame> lexical;
    bool is_signature = false;     // a bare '<-( ... )' value awaiting '* { block }'
};

namespace {

enum Builtin
{
    B_VEC, B_VECX, B_VECY, B_VECZ, B_ROTX, B_ROTY, B_ROTZ, B_COLOR, B_COLR, B_COLG, B_COLB,
    B_SQRT, B_SQR, B_EXP, B_LOG, B_TO_DEG, B_TO_RAD, B_SIN, B_COS, B_TAN, B_SIN_D, B_COS_D, B_TAN_D, B_ASIN, B_ACOS, B_ATAN,
    B_POW, B_FLOOR, B_CEILING, B_FILE_EXISTS, B_FILE_TOUCH, B_FILE_DELETE, B_FILE_RENAME,
    B_CREATE_PLANE, B_CREATE_SPHERE, B_CREATE_SQUAROID, B_CREATE_CYLINDER, B_CREATE_TORUS, B_CREATE_HYPERBOLOID1,
    B_CREATE_HYPERBOLOID2, B_CREATE_ELLIPSOID, B_CREATE_CONE, B_STRING_FA, B_STRING_TO_NUM, B_BETH_OBJECT, B_GET_TIME
};

struct BuiltinDef { const char* name; int id; int nargs; };
const BuiltinDef builtin_defs[] =
{
    { "vec", B_VEC, 3 }, { "vecx", B_VECX, 1 }, { "vecy", B_VECY, 1 }, { "vecz", B_VECZ, 1 },
    { "rotx", B_ROTX, 1 }, { "roty", B_ROTY, 1 }, { "rotz", B_ROTZ, 1 },
    { "color", B_COLOR, 3 }, { "colr", B_COLR, 1 }, { "colg", B_COLG, 1 }, { "colb", B_COLB, 1 },
    { "sqrt", B_SQRT, 1 }, { "sqr", B_SQR, 1 }, { "exp", B_EXP, 1 }, { "log", B_LOG, 1 }, { "to_deg", B_TO_DEG, 1 }, { "to_rad", B_TO_RAD, 1 },
    { "sin", B_SIN, 1 }, { "cos", B_COS, 1 }, { "tan", B_TAN, 1 }, { "sin_d", B_SIN_D, 1 }, { "cos_d", B_COS_D, 1 }, { "tan_d", B_TAN_D, 1 },
    { "asin", B_ASIN, 1 }, { "acos", B_ACOS, 1 }, { "atan", B_ATAN, 1 }, { "pow", B_POW, 2 }, { "floor", B_FLOOR, 1 }, { "ceiling", B_CEILING, 1 },
    { "file_exists", B_FILE_EXISTS, 1 }, { "file_touch", B_FILE_TOUCH, 1 }, { "file_delete", B_FILE_DELETE, 1 }, { "file_rename", B_FILE_RENAME, 2 },
    { "create_plane", B_CREATE_PLANE, 0 }, { "create_sphere", B_CREATE_SPHERE, 1 }, { "create_squaroid", B_CREATE_SQUAROID, 4 },
    { "create_cylinder", B_CREATE_CYLINDER, 2 }, { "create_torus", B_CREATE_TORUS, 2 }, { "create_hyperboloid1", B_CREATE_HYPERBOLOID1, 3 },
    { "create_hyperboloid2", B_CREATE_HYPERBOLOID2, 3 }, { "create_ellipsoid", B_CREATE_ELLIPSOID, 3 }, { "create_cone", B_CREATE_CONE, 3 },
    { "string_fa", B_STRING_FA, 2 }, { "string_to_num", B_STRING_TO_NUM, 1 }, { "beth_object", B_BETH_OBJECT, 1 }, { "get_time", B_GET_TIME, 0 },
};

const char* type_name( const Value& v )
{
    switch( v.type )
    {
        case Value::NIL: return "null"; case Value::BOOL: return "bl_t"; case Value::INT: return "s3_t"; case Value::NUM: return "f3_t";
        case Value::STR: return "st_s"; case Value::VEC: return "v3d_s"; case Value::MAT: return "m3d_s"; case Value::OBJ: return "spect_obj";
        case Value::CMP: return "compound_s"; case Value::LIST: return "arr_s"; case Value::MAP: return "map_s"; case Value::FUNC: return "mclosure_s";
        case Value::SCENE: return "scene_s"; case Value::BUILTIN: return "closure";
    }
    return "?";
}

bool is_null( const VP& v ) { return !v || v->type == Value::NIL; }

struct Interp
{
    Scene& scene;
    std::shared_ptr<Code> code;        // code being executed
    std::shared_ptr<Frame> frame;
    size_t index = 0;
    int depth = 0;

    Interp( Scene& s ) : scene( s ) {}

    const Token& peek() const { static Token end; return index < code->tok.size() ? code->tok[ index ] : end; }
    int peek_kind() const { return peek().kind; }
    const Token& get() { static Token end; return index < code->tok.size() ? code->tok[ index++ ] : end; }
    bool try_kind( int k ) { if( peek_kind() == k ) { index++; return true; } return false; }

    [[noreturn]] void fail( const std::string& msg ) const
    {
        std::string where;
        size_t i = index < code->tok.size() ? index : ( code->tok.empty() ? 0 : code->tok.size() - 1 );
        if( i < code->tok.size() )
        {
            const Token& t = code->tok[ i ];
            std::string fn = t.file < ( int )code->files.size() ? code->files[ t.file ] : "?";
            size_t line = 1;
            try { std::string s = read_file( fn ); for( size_t k = 0; k < t.pos && k < s.size(); k++ ) if( s[ k ] == '\n' ) line++; } catch( ... ) {}
            where = fn + ":" + std::to_string( line ) + ": ";
        }
        throw Error( where + msg );
    }
    void expect( int k, const char* sym ) { if( !try_kind( k ) ) fail( std::string( "'" ) + sym + "' expected." ); }

    // ---- numeric helpers
    double to_f3( const VP& v )
    {
        if( !v ) fail( "Scalar expected." );
        switch( v->type ) { case Value::INT: return ( double )v->i; case Value::NUM: return v->f; case Value::BOOL: return v->b ? 1.0 : 0.0; default: fail( "Scalar expected." ); }
    }
    bool is_num( const VP& v ) { return v && ( v->type == Value::INT || v->type == Value::NUM || v->type == Value::BOOL ); }
    V3d to_v3d( const VP& v ) { if( !v || v->type != Value::VEC ) fail( "Vector expected." ); return v->v; }

    // ---- operators (interpreter.c:651-1231)
    VP op_mul( VP a, VP b )
    {
        if( is_null( a ) || is_null( b ) ) fail( "Cannot multiply an empty object." );
        const Value::Type t1 = a->type, t2 = b->type;
        if( t1 == Value::INT && t2 == Value::INT ) return make_int( a->i * b->i );
        if( t1 == Value::INT && t2 == Value::BOOL ) return make_int( a->i * ( b->b ? 1 : 0 ) );
        if( t1 == Value::BOOL && t2 == Value::INT ) return make_int( ( a->b ? 1 : 0 ) * b->i );
        if( t1 == Value::BOOL && t2 == Value::BOOL ) return make_bool( a->b && b->b );
        if( is_num( a ) && is_num( b ) ) return make_num( to_f3( a ) * to_f3( b ) );
        if( is_num( a ) && t2 == Value::VEC ) return make_vec( b->v * to_f3( a ) );
        if( t1 == Value::VEC && is_num( b ) ) return make_vec( a->v * to_f3( b ) );
        if( t1 == Value::VEC && t2 == Value::VEC ) return make_num( acn::dot( a->v, b->v ) );
        if( t1 == Value::MAT && ( t2 == Value::INT || t2 == Value::NUM ) )
        {
            double f = to_f3( b ); M3d m = a->m; m.x = m.x * f; m.y = m.y * f; m.z = m.z * f; return make_mat( m );
        }
        if( t1 == Value::MAT && t2 == Value::VEC ) return make_vec( acn::mlv( a->m, b->v ) );
        if( t1 == Value::MAT && t2 == Value::MAT ) return make_mat( mat_mlm( a->m, b->m ) );
        if( t1 == Value::LIST || t1 == Value::MAP || t1 == Value::CMP || t1 == Value::OBJ )
        {
            if( t2 == Value::INT || t2 == Value::NUM ) { VP r = a->clone(); value_scale( *r, to_f3( b ) ); return r; }
            if( t2 == Value::MAT ) { VP r = a->clone(); value_rotate( *r, b->m ); return r; }
            if( t1 == Value::OBJ && t2 == Value::VEC ) return make_obj( make_scale( *a->obj, b->v ) );
        }
        if( t1 == Value::FUNC && a->func && a->func->is_signature && t2 == Value::FUNC && b->func && !b->func->is_signature )
        {
            VP r( new Value() ); r->type = Value::FUNC;
            r->func.reset( new Closure() );
            r->func->code = b->func->code; r->func->lexical = b->func->lexical; r->func->params = a->func->params;
            return r;
        }
        fail( std::string( "Cannot evaluate '" ) + type_name( *a ) + "' * '" + type_name( *b ) + "'" );
    }

    VP op_mod( VP a, VP b )
    {
        if( a && b && a->type == Value::INT && b->type == Value::INT && b->i != 0 ) return make_int( a->i % b->i );
        fail( "Cannot evaluate '%' on these operands" );
    }

    static std::string fmt_f3( double v ) { char buf[ 64 ]; snprintf( buf, sizeof( buf ), "%g", v ); return buf; }

    VP op_add( VP a, VP b )
    {
        if( is_null( a ) || is_null( b ) ) fail( "Cannot add an empty object." );
        const Value::Type t1 = a->type, t2 = b->type;
        if( t1 == Value::INT && t2 == Value::INT ) return make_int( a->i + b->i );
        if( ( t1 == Value::INT || t1 == Value::BOOL ) && ( t2 == Value::INT || t2 == Value::BOOL ) ) return make_int( ( long long )to_f3( a ) + ( long long )to_f3( b ) );
        if( is_num( a ) && is_num( b ) ) return make_num( to_f3( a ) + to_f3( b ) );
        if( t1 == Value::INT && t2 == Value::STR ) return make_str( std::to_string( a->i ) + b->s );
        if( t1 == Value::NUM && t2 == Value::STR ) return make_str( fmt_f3( a->f ) + b->s );
        if( t1 == Value::VEC && t2 == Value::VEC ) return make_vec( a->v + b->v );
        if( t1 == Value::STR && t2 == Value::STR ) return make_str( a->s + b->s );
        if( t1 == Value::STR && t2 == Value::INT ) return make_str( a->s + std::to_string( b->i ) );
        if( t1 == Value::STR ) return make_str( a->s );
        if( ( t1 == Value::LIST || t1 == Value::MAP || t1 == Value::CMP || t1 == Value::OBJ ) && t2 == Value::VEC )
        {
            VP r = a->clone(); value_move( *r, b->v ); return r;
        }
        fail( std::string( "Cannot evaluate '" ) + type_name( *a ) + "' + '" + type_name( *b ) + "'" );
    }

    int op_cmp( VP a, VP b )
    {
        if( !is_num( a ) || !is_num( b ) ) fail( "Cannot compare these operands" );
        double x = to_f3( a ), y = to_f3( b );
        return x < y ? 1 : x > y ? -1 : 0;
    }

    VP op_inverse( VP a )
    {
        if( !a || ( a->type != Value::INT && a->type != Value::NUM ) ) fail( "Cannot invert this operand" );
        double v = to_f3( a );
        return make_num( v != 0 ? 1.0 / v : INFINITY );
    }

    VP op_and( VP a, VP b )
    {
        if( a && b && a->type == Value::BOOL && b->type == Value::BOOL ) return make_bool( a->b && b->b );
        if( a && b && a->type == Value::OBJ && b->type == Value::OBJ ) return make_obj( make_pair_inside( *a->obj, *b->obj ) );
        fail( "Cannot evaluate AND on these operands" );
    }
    VP op_or( VP a, VP b )
    {
        if( a && b && a->type == Value::BOOL && b->type == Value::BOOL ) return make_bool( a->b || b->b );
        if( a && b && a->type == Value::OBJ && b->type == Value::OBJ ) return make_obj( make_pair_outside( *a->obj, *b->obj ) );
        fail( "Cannot evaluate OR on these operands" );
    }
    VP op_xor( VP a, VP b )
    {
        if( a && b && a->type == Value::BOOL && b->type == Value::BOOL ) return make_bool( a->b != b->b );
        fail( "Cannot evaluate XOR on these operands" );
    }
    VP op_not( VP a )
    {
        if( a && a->type == Value::BOOL ) return make_bool( !a->b );
        if( a && a->type == Value::OBJ ) return make_obj( make_neg( *a->obj ) );
        fail( "Cannot evaluate NOT on this operand" );
    }
    VP op_cat( VP a, VP b )
    {
        if( is_null( a ) ) fail( "Cannot catenate an empty object." );
        if( a->type == Value::LIST )
        {
            VP r = a->clone();
            if( b && b->type == Value::LIST ) { for( const VP& e : b->list ) r->list.push_back( e ? e->clone() : VP() ); }
            else r->list.push_back( b ? b->clone() : VP() );
            return r;
        }
        VP r = make_list();
        r->list.push_back( a->clone() );
        r->list.push_back( b ? b->clone() : VP() );
        return r;
    }
    VP list_to_compound( const Value& l )
    {
        std::unique_ptr<Compound> c( new Compound() );
        std::string err;
        for( const VP& e : l.list ) if( e && !compound_push_value( *c, *e, &err ) ) fail( err );
        return make_cmp( std::move( c ) );
    }
    VP op_auto_envelope( VP a )       // '(@)' (interpreter.c:1172-1200)
    {
        if( a && a->type == Value::LIST ) { VP r = list_to_compound( *a ); r->cmp->set_auto_envelope(); return r; }
        if( a && a->type == Value::CMP ) { VP r = a->clone(); r->cmp->set_auto_envelope(); return r; }
        if( a && a->type == Value::OBJ ) { VP r = a->clone(); r->obj->set_auto_envelope(); return r; }
        fail( "Cannot compute envelope for this operand" );
    }

    // typed copy into an existing value (bcore_inst_t_copy_typed, interpreter.c:1477)
    void assign_into( Value& dst, const Value& src )
    {
        if( dst.type == Value::INT && src.type == Value::NUM ) { dst.i = ( long long )src.f; return; }
        if( dst.type == Value::NUM && src.type == Value::INT ) { dst.f = ( double )src.i; return; }
        if( dst.type == Value::INT && src.type == Value::BOOL ) { dst.i = src.b; return; }
        VP c = src.clone();
        dst.type = c->type; dst.b = c->b; dst.i = c->i; dst.f = c->f; dst.s = c->s; dst.v = c->v; dst.m = c->m;
        dst.obj = std::move( c->obj ); dst.cmp = std::move( c->cmp ); dst.list = std::move( c->list ); dst.map = std::move( c->map );
        dst.func = c->func; dst.builtin = c->builtin;
    }

    // ---- calls
    VP call_closure( const VP& fn, std::vector<VP>& args )
    {
        Closure& c = *fn->func;
        if( c.is_signature ) fail( "A bare signature is no function." );
        if( ++depth > 200 ) fail( "Call depth exceeds 200." );
        std::shared_ptr<Frame> local( new Frame() );
        local->external = c.lexical;
        for( size_t i = 0; i < c.params.size(); i++ ) local->set( c.params[ i ], i < args.size() ? args[ i ] : VP() );
        std::shared_ptr<Code> save_code = code; std::shared_ptr<Frame> save_frame = frame; size_t save_index = index;
        code = c.code; frame = local; index = 0;
        VP ret;
        try { ret = execute(); }
        catch( ... ) { throw; }
        code = save_code; frame = save_frame; index = save_index;
        depth--;
        return ret;
    }

    VP eval_call( const VP& fn )     // interpreter.c:1374-1407
    {
        if( !fn || ( fn->type != Value::FUNC && fn->type != Value::BUILTIN ) ) fail( std::string( "'" ) + ( fn ? type_name( *fn ) : "null" ) + "' is no function." );
        expect( T_LPAR, "(" );
        size_t n = fn->type == Value::FUNC ? fn->func->params.size() : ( size_t )builtin_defs[ fn->builtin ].nargs;
        std::vector<VP> args;
        for( size_t i = 0; i < n; i++ )
        {
            if( i > 0 ) expect( T_COMMA, "," );
            args.push_back( eval( VP() ) );
        }
        VP ret = fn->type == Value::FUNC ? call_closure( fn, args ) : call_builtin( builtin_defs[ fn->builtin ].id, args );
        expect( T_RPAR, ")" );
        return ret;
    }

    static bool file_exists( const std::string& f ) { struct stat st; return stat( f.c_str(), &st ) == 0; }

    std::string string_fa( const std::string& fmt, const VP& arg )
    {
        // the subset of beth's format language used by the scripts: #pl<N>'<c>'{ ... }, #<s3_t*>, #<f3_t*>, #<sc_t>
        std::string out;
        for( size_t i = 0; i < fmt.size(); )
        {
            if( fmt[ i ] != '#' ) { out.push_back( fmt[ i++ ] ); continue; }
            if( fmt.compare( i, 3, "#pl" ) == 0 )
            {
                size_t j = i + 3; size_t width = 0;
                while( j < fmt.size() && fmt[ j ] >= '0' && fmt[ j ] <= '9' ) width = width * 10 + ( fmt[ j++ ] - '0' );
                char pad = ' ';
                if( j + 2 < fmt.size() && fmt[ j ] == '\'' ) { pad = fmt[ j + 1 ]; j += 3; }
                if( j < fmt.size() && fmt[ j ] == '{' )
                {
                    size_t e = fmt.find( '}', j );
                    std::string inner = string_fa( fmt.substr( j + 1, e == std::string::npos ? std::string::npos : e - j - 1 ), arg );
                    while( inner.size() < width ) inner.insert( inner.begin(), pad );
                    out += inner;
                    i = e == std::string::npos ? fmt.size() : e + 1;
                    continue;
                }
                i = j; continue;
            }
            if( fmt.compare( i, 2, "#<" ) == 0 )
            {
                size_t e = fmt.find( '>', i );
                if( arg && arg->type == Value::INT ) out += std::to_string( arg->i );
                else if( arg && arg->type == Value::NUM ) out += fmt_f3( arg->f );
                else if( arg && arg->type == Value::STR ) out += arg->s;
                i = e == std::string::npos ? fmt.size() : e + 1;
                continue;
            }
            out.push_back( fmt[ i++ ] );
        }
        return out;
    }

    VP call_builtin( int id, std::vector<VP>& a )
    {
        const double PI = 3.14159265358979323846;
        auto f = [ & ]( size_t i ) { return to_f3( a[ i ] ); };
        auto str = [ & ]( size_t i ) -> const std::string& { if( !a[ i ] || a[ i ]->type != Value::STR ) fail( "String expected." ); return a[ i ]->s; };
        switch( id )
        {
            case B_VEC: case B_COLOR: return make_vec( vec3( f( 0 ), f( 1 ), f( 2 ) ) );
            case B_VECX: case B_COLR: return make_vec( vec3( f( 0 ), 0, 0 ) );
            case B_VECY: case B_COLG: return make_vec( vec3( 0, f( 0 ), 0 ) );
            case B_VECZ: case B_COLB: return make_vec( vec3( 0, 0, f( 0 ) ) );
            case B_ROTX: return make_mat( mat_rot_x( ( PI / 180.0 ) * f( 0 ) ) );
            case B_ROTY: return make_mat( mat_rot_y( ( PI / 180.0 ) * f( 0 ) ) );
            case B_ROTZ: return make_mat( mat_rot_z( ( PI / 180.0 ) * f( 0 ) ) );
            case B_SQRT: return make_num( sqrt( f( 0 ) ) );
            case B_SQR: { double v = f( 0 ); return make_num( v * v ); }
            case B_EXP: return make_num( exp( f( 0 ) ) );
            case B_LOG: return make_num( log( f( 0 ) ) );
            case B_TO_DEG: return make_num( f( 0 ) * 180.0 / PI );
            case B_TO_RAD: return make_num( f( 0 ) * PI / 180.0 );
            case B_SIN: return make_num( sin( f( 0 ) ) );
            case B_COS: return make_num( cos( f( 0 ) ) );
            case B_TAN: return make_num( tan( f( 0 ) ) );
            case B_SIN_D: return make_num( sin( PI * f( 0 ) / 180.0 ) );
            case B_COS_D: return make_num( cos( PI * f( 0 ) / 180.0 ) );
            case B_TAN_D: return make_num( tan( PI * f( 0 ) / 180.0 ) );
            case B_ASIN: return make_num( asin( f( 0 ) ) );
            case B_ACOS: return make_num( acos( f( 0 ) ) );
            case B_ATAN: return make_num( atan( f( 0 ) ) );
            case B_POW: return make_num( pow( f( 0 ), f( 1 ) ) );
            case B_FLOOR: return make_num( floor( f( 0 ) ) );
            case B_CEILING: return make_num( ceil( f( 0 ) ) );
            case B_FILE_EXISTS: return make_bool( file_exists( str( 0 ) ) );
            // the front-end only records images: the lock files of the video scripts are not touched
            case B_FILE_TOUCH: str( 0 ); return make_bool( true );
            case B_FILE_DELETE: str( 0 ); return make_bool( true );
            case B_FILE_RENAME: str( 0 ); str( 1 ); return make_bool( true );
            case B_CREATE_PLANE: return make_obj( make_plane() );
            case B_CREATE_SPHERE: return make_obj( make_sphere( f( 0 ) ) );
            case B_CREATE_SQUAROID: return make_obj( make_squaroid( f( 0 ), f( 1 ), f( 2 ), f( 3 ) ) );
            case B_CREATE_CYLINDER: return make_obj( make_cylinder( f( 0 ), f( 1 ) ) );
            case B_CREATE_TORUS: { double r1 = f( 0 ); if( r1 == 0 ) fail( "create_torus: radius1 is 0" ); return make_obj( make_torus( r1, f( 1 ) ) ); }
            case B_CREATE_HYPERBOLOID1: return make_obj( make_hyperboloid1( f( 0 ), f( 1 ), f( 2 ) ) );
            case B_CREATE_HYPERBOLOID2: return make_obj( make_hyperboloid2( f( 0 ), f( 1 ), f( 2 ) ) );
            case B_CREATE_ELLIPSOID: return make_obj( make_ellipsoid( f( 0 ), f( 1 ), f( 2 ) ) );
            case B_CREATE_CONE: return make_obj( make_cone( f( 0 ), f( 1 ), f( 2 ) ) );
            case B_STRING_FA: return make_str( string_fa( str( 0 ), a[ 1 ] ) );
            case B_STRING_TO_NUM:
            {
                const std::string& s = str( 0 );
                bool is_float = s.find_first_of( ".eE" ) != std::string::npos;
                return is_float ? make_num( atof( s.c_str() ) ) : make_int( atoll( s.c_str() ) );
            }
            case B_BETH_OBJECT:
            {
                const std::string& s = str( 0 );
                if( s == "obj_sphere_s" ) return make_obj( make_sphere( 1.0 ) );
                if( s == "obj_plane_s" ) return make_obj( make_plane() );
                if( s == "obj_squaroid_s" ) return make_obj( make_squaroid( 1, 1, 1, -1 ) );
                if( s == "map_s" ) return make_map();
                if( s == "arr_s" ) return make_list();
                if( s == "compound_s" ) return make_cmp( std::unique_ptr<Compound>( new Compound() ) );
                fail( "beth_object: unsupported type '" + s + "'" );
            }
            case B_GET_TIME: return make_num( 0.0 );
        }
        fail( "unknown built-in" );
    }

    // ---- member access
    VP scene_member( const std::string& key, bool assign )
    {
        acn_flat_params& p = scene.params;
        struct IM { const char* n; int32_t* p; };
        struct DM { const char* n; double* p; };
        struct VM { const char* n; double* p; };
        IM im[] = { { "threads", &p.threads }, { "image_width", &p.image_width }, { "image_height", &p.image_height },
                    { "gradient_samples", &p.gradient_samples }, { "gradient_cycles", &p.gradient_cycles }, { "trace_depth", &p.trace_depth },
                    { "direct_samples", &p.direct_samples }, { "path_samples", &p.path_samples } };
        DM dm[] = { { "gamma", &p.gamma }, { "gradient_threshold", &p.gradient_threshold }, { "camera_focal_length", &p.camera_focal_length },
                    { "trace_min_intensity", &p.trace_min_intensity }, { "max_path_length", &p.max_path_length } };
        VM vm[] = { { "background_color", p.background_color }, { "camera_position", p.camera_position },
                    { "camera_view_direction", p.camera_view_direction }, { "camera_top_direction", p.camera_top_direction } };
        for( IM& m : im ) if( key == m.n ) { if( assign ) { *m.p = ( int32_t )to_f3( eval( VP() ) ); } return make_int( *m.p ); }
        for( DM& m : dm ) if( key == m.n ) { if( assign ) { *m.p = to_f3( eval( VP() ) ); } return make_num( *m.p ); }
        for( VM& m : vm ) if( key == m.n )
        {
            if( assign ) { V3d v = to_v3d( eval( VP() ) ); m.p[ 0 ] = v.x; m.p[ 1 ] = v.y; m.p[ 2 ] = v.z; }
            return make_vec( vec3( m.p[ 0 ], m.p[ 1 ], m.p[ 2 ] ) );
        }
        if( key == "experimental_level" ) { if( assign ) eval( VP() ); return make_int( 0 ); }
        return VP();
    }

    Envelope eval_envelope_arg()
    {
        VP v = eval( VP() );
        if( !v || v->type != Value::OBJ || v->obj->kind != ACN_KIND_SPHERE ) fail( "Object cannot be used as envelope (use a sphere)." );
        Envelope e{ v->obj->prp.pos, v->obj->tail[ 0 ] };
        return e;
    }

    VP member( VP front )     // interpreter.c:1481-1523 + the per-type *_meval_key tables
    {
        if( peek_kind() != T_NAME ) fail( "Name expected after '.'" );
        const std::string key = get().s;
        Value& o = *front;
        switch( o.type )
        {
            case Value::SCENE:
            {
                const bool assign = peek_kind() == T_ASSIGN;
                if( assign ) index++;
                VP r = scene_member( key, assign );
                if( r ) return assign ? front : r;
                if( assign ) fail( "scene_s has no member '" + key + "'." );
                if( key == "clear" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); scene.clear(); return VP(); }
                if( key == "push" )
                {
                    expect( T_LPAR, "(" ); VP x = eval( VP() ); expect( T_RPAR, ")" );
                    std::string err;
                    if( x && !scene.push( *x, &err ) ) fail( err );
                    return VP();
                }
                if( key == "create_image" )
                {
                    expect( T_LPAR, "(" ); VP x = eval( VP() ); expect( T_RPAR, ")" );
                    if( !x || x->type != Value::STR ) fail( "String expected." );
                    RecordedImage im;
                    im.name = x->s; im.params = scene.params; im.light = scene.light.clone(); im.matter = scene.matter.clone();
                    scene.images.push_back( std::move( im ) );
                    return VP();
                }
                fail( "scene_s has no member '" + key + "'." );
            }
            case Value::VEC:
            {
                double* c = key == "x" ? &o.v.x : key == "y" ? &o.v.y : key == "z" ? &o.v.z : nullptr;
                if( !c ) fail( "v3d_s has no element named '" + key + "'." );
                if( try_kind( T_ASSIGN ) ) { *c = to_f3( eval( VP() ) ); return front; }
                return make_num( *c );
            }
            case Value::MAP:
            {
                if( VP v = o.map_get( key ) ) return v;
                if( try_kind( T_ASSIGN ) ) { VP v = eval( VP() ); VP c = v ? v->clone() : make_nil(); o.map_set( key, c ); return c; }
                if( key == "move" ) { expect( T_LPAR, "(" ); V3d v = to_v3d( eval( VP() ) ); expect( T_RPAR, ")" ); value_move( o, v ); return VP(); }
                if( key == "rotate" ) { expect( T_LPAR, "(" ); VP m = eval( VP() ); expect( T_RPAR, ")" ); if( !m || m->type != Value::MAT ) fail( "Rotation expected." ); value_rotate( o, m->m ); return VP(); }
                if( key == "scale" ) { expect( T_LPAR, "(" ); double f = to_f3( eval( VP() ) ); expect( T_RPAR, ")" ); value_scale( o, f ); return VP(); }
                if( key == "has" ) { expect( T_LPAR, "(" ); if( peek_kind() != T_NAME ) fail( "Name expected." ); std::string k = get().s; expect( T_RPAR, ")" ); return make_bool( ( bool )o.map_get( k ) ); }
                fail( "Map has no element of name " + key + "." );
            }
            case Value::LIST:
            {
                if( key == "push" ) { expect( T_LPAR, "(" ); VP v = eval( VP() ); expect( T_RPAR, ")" ); VP c = v ? v->clone() : VP(); o.list.push_back( c ); return c; }
                if( key == "move" ) { expect( T_LPAR, "(" ); V3d v = to_v3d( eval( VP() ) ); expect( T_RPAR, ")" ); value_move( o, v ); return VP(); }
                if( key == "rotate" ) { expect( T_LPAR, "(" ); VP m = eval( VP() ); expect( T_RPAR, ")" ); if( !m || m->type != Value::MAT ) fail( "Rotation expected." ); value_rotate( o, m->m ); return VP(); }
                if( key == "scale" ) { expect( T_LPAR, "(" ); double f = to_f3( eval( VP() ) ); expect( T_RPAR, ")" ); value_scale( o, f ); return VP(); }
                if( key == "size" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); return make_int( ( long long )o.list.size() ); }
                if( key == "clear" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); o.list.clear(); return VP(); }
                if( key == "create_inside_composite" || key == "create_outside_composite" )
                {
                    expect( T_LPAR, "(" ); expect( T_RPAR, ")" );
                    std::string err;
                    auto r = key == "create_inside_composite" ? list_inside_composite( o.list, 0, o.list.size(), &err ) : list_outside_composite( o.list, 0, o.list.size(), &err );
                    if( !r ) fail( err );
                    return make_obj( std::move( r ) );
                }
                if( key == "create_compound" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); return list_to_compound( o ); }
                fail( "arr_s has no element of name " + key + "." );
            }
            case Value::CMP:
            {
                Compound& c = *o.cmp;
                if( key == "push" )
                {
                    expect( T_LPAR, "(" ); VP v = eval( VP() ); expect( T_RPAR, ")" );
                    if( !v || ( v->type != Value::OBJ && v->type != Value::CMP ) ) fail( "Cannot push this value to compound_s." );
                    std::string err; compound_push_value( c, *v, &err );
                    return VP();
                }
                if( key == "move" ) { expect( T_LPAR, "(" ); V3d v = to_v3d( eval( VP() ) ); expect( T_RPAR, ")" ); c.move( v ); return VP(); }
                if( key == "rotate" ) { expect( T_LPAR, "(" ); VP m = eval( VP() ); expect( T_RPAR, ")" ); if( !m || m->type != Value::MAT ) fail( "Rotation expected." ); c.rotate( m->m ); return VP(); }
                if( key == "scale" ) { expect( T_LPAR, "(" ); double f = to_f3( eval( VP() ) ); expect( T_RPAR, ")" ); c.scale( f ); return VP(); }
                if( key == "set_envelope" ) { expect( T_LPAR, "(" ); Envelope e = eval_envelope_arg(); expect( T_RPAR, ")" ); c.has_envelope = true; c.envelope = e; return VP(); }
                if( key == "set_auto_envelope" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); c.set_auto_envelope(); return VP(); }
                fail( "Compound has no element of name " + key + "." );
            }
            case Value::OBJ:
            {
                Obj& b = *o.obj;
                auto arg_f3 = [ & ]() { expect( T_LPAR, "(" ); double f = to_f3( eval( VP() ) ); expect( T_RPAR, ")" ); return f; };
                auto arg_v3 = [ & ]() { expect( T_LPAR, "(" ); V3d v = to_v3d( eval( VP() ) ); expect( T_RPAR, ")" ); return v; };
                if( key == "move" ) { b.move( arg_v3() ); return VP(); }
                if( key == "rotate" ) { expect( T_LPAR, "(" ); VP m = eval( VP() ); expect( T_RPAR, ")" ); if( !m || m->type != Value::MAT ) fail( "Rotation expected." ); b.rotate( m->m ); return VP(); }
                if( key == "scale" ) { b.scale( arg_f3() ); return VP(); }
                if( key == "set_color" ) { b.prp.color = arg_v3(); return VP(); }
                if( key == "set_transparency" ) { b.prp.transparency = arg_v3(); return VP(); }
                if( key == "set_refractive_index" ) { b.set_refractive_index( arg_f3() ); return VP(); }
                if( key == "set_radiance" ) { b.prp.radiance = arg_f3(); return VP(); }
                if( key == "set_fresnel_reflectivity" ) { b.prp.fresnel_reflectivity = arg_f3(); return VP(); }
                if( key == "set_chromatic_reflectivity" ) { b.prp.chromatic_reflectivity = arg_f3(); return VP(); }
                if( key == "set_diffuse_reflectivity" ) { b.prp.diffuse_reflectivity = arg_f3(); return VP(); }
                if( key == "set_sigma" ) { b.prp.sigma = arg_f3(); return VP(); }
                if( key == "set_surface_roughness" ) { b.prp.surface_roughness = arg_f3(); return VP(); }
                if( key == "set_envelope" ) { expect( T_LPAR, "(" ); Envelope e = eval_envelope_arg(); expect( T_RPAR, ")" ); b.prp.has_envelope = true; b.prp.envelope = e; return VP(); }
                if( key == "set_auto_envelope" ) { expect( T_LPAR, "(" ); expect( T_RPAR, ")" ); b.set_auto_envelope(); return VP(); }
                if( key == "set_material" )
                {
                    expect( T_LPAR, "(" ); VP s = eval( VP() ); expect( T_RPAR, ")" );
                    if( !s || s->type != Value::STR ) fail( "set_surface: string-argument expected." );
                    if( !b.set_material( s->s ) ) fail( "set_surface: Unknown material specification '" + s->s + "'." );
                    return VP();
                }
                if( key == "radius" && b.kind == ACN_KIND_SPHERE )
                {
                    if( try_kind( T_ASSIGN ) ) { b.tail[ 0 ] = to_f3( eval( VP() ) ); return front; }
                    return make_num( b.tail[ 0 ] );
                }
                fail( "Object has no member or function '" + key + "'." );
            }
            default: fail( std::string( "Object '" ) + type_name( o ) + "' has no element named '" + key + "'." );
        }
    }

    // ---- expression evaluation (interpreter.c:1412-1730)
    VP eval( VP front )
    {
        int opr = 0;
        if( front )
        {
            const int k = peek_kind();
            if( is_operator( k ) ) { opr = k; index++; }
            else if( k == T_LPAR ) return eval_call( front );
            else if( k == T_LBRK )
            {
                index++;
                if( front->type != Value::LIST ) fail( std::string( "Cannot index '" ) + type_name( *front ) + "'." );
                VP idx = eval( VP() );
                expect( T_RBRK, "]" );
                if( !is_num( idx ) ) fail( "Numeric index expected." );
                long long i = ( long long )to_f3( idx );
                if( i < 0 ) fail( "Index is negative." );
                if( i > 1000000000ll ) fail( "Attempting to allocate a huge array seems unintended." );
                if( ( size_t )i >= front->list.size() ) front->list.resize( ( size_t )i + 1 );
                VP& slot = front->list[ ( size_t )i ];
                if( !slot && peek_kind() == T_ASSIGN ) { index++; VP v = eval( VP() ); slot = v ? v->clone() : VP(); }
                return slot;
            }
            else return front;

            if( is_assign_op( opr ) )
            {
                VP obj = eval( VP() );
                if( is_null( obj ) ) fail( "Assignment from empty object." );
                switch( opr )
                {
                    case T_ADD_ASSIGN: obj = op_add( front, obj ); break;
                    case T_SUB_ASSIGN: obj = op_add( front, op_mul( make_num( -1 ), obj ) ); break;
                    case T_MUL_ASSIGN: obj = op_mul( front, obj ); break;
                    case T_DIV_ASSIGN: obj = op_mul( front, op_inverse( obj ) ); break;
                    case T_MOD_ASSIGN: obj = op_mod( front, obj ); break;
                    default: break;
                }
                assign_into( *front, *obj );
                return front;
            }
            if( opr == T_DOT ) return member( front );
        }
        else
        {
            const int k = peek_kind();
            if( k == T_QUERY || k == T_DQUERY )
            {
                index++;
                VP v = eval( VP() );
                if( v && v->type == Value::STR ) { fputs( v->s.c_str(), stdout ); fputc( '\n', stdout ); }
                else if( v && v->type == Value::INT ) printf( "%lld\n", v->i );
                else if( v && v->type == Value::NUM ) printf( "%g\n", v->f );
                else if( v && v->type == Value::BOOL ) printf( "%s\n", v->b ? "true" : "false" );
                else if( v && v->type == Value::VEC ) printf( "(%g, %g, %g)\n", v->v.x, v->v.y, v->v.z );
                else if( v ) printf( "<%s>\n", type_name( *v ) );
                return VP();
            }
        }

        int unary = 0;
        switch( peek_kind() )
        {
            case T_ADD: case T_SUB: case T_NOT: case T_INSIDE_CPS: case T_OUTSIDE_CPS: case T_COMPOUND: case T_ENVELOPE:
                unary = get().kind; break;
            default: break;
        }

        VP obj;
        const int k = peek_kind();
        if( k == T_INT ) obj = make_int( get().i );
        else if( k == T_NUM ) obj = make_num( get().f );
        else if( k == T_STR ) obj = make_str( get().s );
        else if( k == T_BOOL ) obj = make_bool( get().i != 0 );
        else if( k == T_TYPE ) fail( "Unexpected type name." );
        else if( k == T_BLOCK )
        {
            const Token& t = get();
            obj.reset( new Value() ); obj->type = Value::FUNC;
            obj->func.reset( new Closure() );
            obj->func->code = t.block; obj->func->lexical = frame;
        }
        else if( k == T_NAME )
        {
            const std::string name = get().s;
            VP* p = frame->get( name );
            const int pk = peek_kind();
            if( is_assign_op( pk ) )
            {
                if( !p ) fail( "'" + name + "' was not defined. Use 'def " + name + "' to define it." );
                if( is_null( *p ) )
                {
                    expect( T_ASSIGN, "=" );
                    VP v = eval( VP() );
                    *p = v ? v->clone() : VP();
                }
                else obj = eval( *p );          // consume the assignment in a nested cycle
            }
            else if( !p ) fail( "Unknown name '" + name + "'" );
            else obj = *p;
        }
        else if( k == T_DYNARR ) { index++; obj = make_list(); }
        else if( k == T_FSIG )
        {
            index++;
            obj.reset( new Value() ); obj->type = Value::FUNC;
            obj->func.reset( new Closure() ); obj->func->is_signature = true;
            expect( T_LPAR, "(" );
            while( !try_kind( T_RPAR ) )
            {
                if( peek_kind() == T_TYPE ) index++;
                if( peek_kind() != T_NAME ) fail( "Argument name expected." );
                obj->func->params.push_back( get().s );
                if( peek_kind() != T_RPAR ) expect( T_COMMA, "," );
            }
        }
        else if( k == T_LPAR ) { index++; obj = eval( VP() ); expect( T_RPAR, ")" ); }
        else if( k == T_DEF )
        {
            index++;
            if( peek_kind() != T_NAME ) fail( "Name expected after 'def'." );
            const std::string name = get().s;
            if( frame->get_local( name ) ) fail( "'" + name + "' is already defined." );
            if( try_kind( T_ASSIGN ) )
            {
                VP v = eval( VP() );
                obj = *frame->set( name, v ? v->clone() : VP() );
            }
            else frame->set( name, VP() );
        }

        // operations on the object taking priority over standard operators
        if( obj )
        {
            int c = peek_kind();
            while( obj && ( c == T_LPAR || c == T_LBRK || c == T_DOT ) ) { obj = eval( obj ); c = peek_kind(); }
        }

        if( obj )
        {
            switch( unary )
            {
                case T_SUB: obj = op_mul( make_int( -1 ), obj ); break;
                case T_NOT: obj = op_not( obj ); break;
                case T_INSIDE_CPS:
                case T_OUTSIDE_CPS:
                {
                    if( obj->type != Value::LIST ) fail( "Cannot create a composite of this operand" );
                    std::string err;
                    auto r = unary == T_INSIDE_CPS ? list_inside_composite( obj->list, 0, obj->list.size(), &err ) : list_outside_composite( obj->list, 0, obj->list.size(), &err );
                    if( !r ) fail( err );
                    obj = make_obj( std::move( r ) );
                }
                break;
                case T_COMPOUND: if( obj->type != Value::LIST ) fail( "Cannot create compound of this operand" ); obj = list_to_compound( *obj ); break;
                case T_ENVELOPE: obj = op_auto_envelope( obj ); break;
                default: break;
            }

            if( opr )
            {
                switch( opr )
                {
                    case T_MUL: return eval( op_mul( front, obj ) );
                    case T_DIV: return eval( op_mul( front, op_inverse( obj ) ) );
                    case T_MOD: return eval( op_mod( front, obj ) );
                    case T_EQUAL:         return eval( make_bool( op_cmp( front, obj ) == 0 ) );
                    case T_UNEQUAL:       return eval( make_bool( op_cmp( front, obj ) != 0 ) );
                    case T_SMALLER:       return eval( make_bool( op_cmp( front, obj ) >  0 ) );
                    case T_SMALLER_EQUAL: return eval( make_bool( op_cmp( front, obj ) >= 0 ) );
                    case T_LARGER:        return eval( make_bool( op_cmp( front, obj ) <  0 ) );
                    case T_LARGER_EQUAL:  return eval( make_bool( op_cmp( front, obj ) <= 0 ) );
                    case T_ADD: return op_add( front, eval( obj ) );
                    case T_SUB: return op_add( front, eval( op_mul( make_int( -1 ), obj ) ) );
                    case T_AND: return op_and( front, eval( obj ) );
                    case T_OR:  return op_or( front, eval( obj ) );
                    case T_XOR: return op_xor( front, eval( obj ) );
                    case T_CAT: return eval( op_cat( front, obj ) );
                    default: fail( "Invalid operator." );
                }
            }
            else obj = eval( obj );       // operations subordinate to standard operators
        }
        else if( opr ) fail( "Expression does not yield an operand for the operator." );
        return obj;
    }

    // ---- statements (interpreter.c:1734-1850)
    VP execute()
    {
        VP ret;
        while( index < code->tok.size() )
        {
            VP obj;
            const int k = peek_kind();
            if( k == T_IF )
            {
                const size_t target = get().jump;
                expect( T_LPAR, "(" ); VP cond = eval( VP() ); expect( T_RPAR, ")" );
                if( !cond || cond->type != Value::BOOL ) fail( "Expression does not evaluate to boolean." );
                const bool flag = cond->b;
                if( flag ) obj = eval( VP() ); else index = target;
                if( peek_kind() == T_ELSE )
                {
                    const size_t t2 = get().jump;
                    if( flag ) index = t2; else obj = eval( VP() );
                }
            }
            else if( k == T_WHILE )
            {
                const size_t end_while = get().jump;
                const size_t begin_while = index;
                int guard = 0;
                for( ;; )
                {
                    expect( T_LPAR, "(" ); VP cond = eval( VP() ); expect( T_RPAR, ")" );
                    if( !cond || cond->type != Value::BOOL ) fail( "Expression does not evaluate to boolean." );
                    if( cond->b ) { obj = eval( VP() ); index = begin_while; }
                    else { index = end_while; break; }
                    if( ++guard > 100000000 ) fail( "while: iteration limit" );
                }
            }
            else if( k == T_FOR )
            {
                const size_t end_for = get().jump;
                std::shared_ptr<Frame> for_frame( new Frame() );
                for_frame->external = frame;
                std::shared_ptr<Frame> save = frame;
                frame = for_frame;
                if( peek_kind() != T_NAME ) fail( "Name expected after 'for'." );
                const std::string name = get().s;
                for_frame->set( name, VP() );
                expect( T_LPAR, "(" );
                expect( T_IN, "in" );
                VP arr = eval( VP() );
                if( !arr || arr->type != Value::LIST ) fail( "Expected: for '" + name + "' in 'list-expression'." );
                expect( T_RPAR, ")" );
                const size_t begin_loop = index;
                for( size_t i = 0; i < arr->list.size(); i++ )
                {
                    if( !arr->list[ i ] ) continue;
                    for_frame->set( name, arr->list[ i ] );       // the element itself, not a copy
                    eval( VP() );
                    index = begin_loop;
                }
                index = end_for;
                frame = save;
            }
            else obj = eval( VP() );

            expect( T_SEMI, ";" );
            ret = obj;
        }
        return ret;
    }
};

} // namespace

int interpret_file( Scene& scene, const std::string& path, const std::vector<std::string>& args, std::string* err )
{
    try
    {
        Lexer lx;
        lx.src = read_file( path ); lx.file = path;
        // interpreter selector "<mclosure_s></>" (e.g. src_acn/primitives.acn:18), after leading comments
        // (a '#!' line and comments may precede it: everything before the selector is skipped)
        {
            size_t h = lx.src.find( "<mclosure_s>" );
            if( h != std::string::npos )
            {
                lx.p = h + strlen( "<mclosure_s>" );
                lx.skip_space();
                if( !lx.take( "</>" ) ) lx.fail( "'</>' expected after '<mclosure_s>'" );
            }
            else if( lx.at( "#!" ) ) { while( !lx.eos() && lx.src[ lx.p ] != '\n' ) lx.p++; }
        }
        std::shared_ptr<Code> code = parse_block( lx, std::vector<std::string>{ path } );
        lx.skip_space();
        if( !lx.eos() ) lx.fail( "Syntax error (unbalanced '}')." );

        std::shared_ptr<Frame> root( new Frame() );
        for( size_t i = 0; i < sizeof( builtin_defs ) / sizeof( builtin_defs[ 0 ] ); i++ )
        {
            VP b( new Value() ); b->type = Value::BUILTIN; b->builtin = ( int )i;
            root->set( builtin_defs[ i ].name, b );
        }
        { VP s( new Value() ); s->type = Value::SCENE; root->set( "scene_s", s ); }
        root->set( "obj_sphere_s", make_obj( make_sphere( 1.0 ) ) );
        root->set( "obj_plane_s", make_obj( make_plane() ) );
        root->set( "arr_s", make_list() );
        root->set( "map_s", make_map() );
        { VP a = make_list(); for( const std::string& s : args ) a->list.push_back( make_str( s ) ); root->set( "program_args", a ); }

        Interp in( scene );
        in.code = code;
        in.frame.reset( new Frame() );
        in.frame->external = root;
        in.index = 0;
        in.execute();
        return ACN_OK;
    }
    catch( const std::exception& e )
    {
        if( err ) *err = e.what();
        return ACN_ERR_PARSE;
    }
}

} // namespace acnh
