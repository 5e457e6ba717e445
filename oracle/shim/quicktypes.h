/* shadows the reference's quicktypes.h (beth type hashes are irrelevant for the leaf math) */
#ifndef ACN_SHIM_QUICKTYPES_H
#define ACN_SHIM_QUICKTYPES_H
#endif
