#include "bcore_std.h"
