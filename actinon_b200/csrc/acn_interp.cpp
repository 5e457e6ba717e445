// acn_interp.cpp — .acn front-end (placeholder until the interpreter lands)
#include "acn_model.h"
namespace acnh {
int interpret_file( Scene&, const std::string& path, const std::vector<std::string>&, std::string* err )
{
    if( err ) *err = "the .acn front-end is not built yet: " + path;
    return ACN_ERR_UNSUPPORTED;
}
}
